// Shared definitions of the host-side drop-in classes (mirror of the reference's
// include/KaminoHeader.cuh:26-64 as far as the solver path needs it). No CUDA headers:
// the host classes talk to the device only through include/kamino_b200.h.
#pragma once

#include <cmath>
#include <cstddef>
#include <iostream>
#include <string>
#include <vector>

#include "kamino_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define M_2PI 6.28318530717958647692
#define M_hPI 1.57079632679489661923

// stagger offsets in units of the grid spacing (include/KaminoHeader.cuh:30-41)
#define centeredPhiOffset 0.0
#define centeredThetaOffset 0.5
#define vPhiPhiOffset -0.5
#define vPhiThetaOffset 0.5
#define vThetaPhiOffset 0.0
#define vThetaThetaOffset 1.0

typedef float fReal;

// The reference's error convention (cuda_util_headers/helper_cuda.h:981-995): print the
// failing call and exit(EXIT_FAILURE). `ctx` may be NULL for creation errors.
void kaminoCheck(int code, kamino_ctx* ctx, const char* what, const char* file, int line);
#define KAMINO_CHECK(ctx, call) kaminoCheck((call), (ctx), #call, __FILE__, __LINE__)
