// Shared device/host definitions for the kamino_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kb {

// pi and 2*pi as the reference spells them (include/KaminoHeader.cuh:26-27): doubles.
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
// The fp32 neighbours of those constants from above. For an fp32 x:
//   (double)x > kPi   <=>  x >= kPiF      (kPi lies strictly between two fp32 values)
//   0 <= x < kTwoPiF   =>  floorf((float)((double)x / kTwoPi)) == 0
// (see DESIGN.md "exactness notes").
constexpr float kPiF = 3.14159274101257324f;      // 0x40490FDB
constexpr float kTwoPiF = 6.28318548202514648f;   // 0x40C90FDB

enum SampleKind { kVPhi = 0, kVTheta = 1, kCentered = 2 };

// Per-context constants, passed to every kernel by value (lands in the constant bank).
struct GridParams {
    int nTheta;        // rows of u_phi / density / pressure; u_theta has nTheta - 1
    int nPhi;          // 2 * nTheta, power of two
    int log2NPhi;
    float radius;      // radiusGlobal
    float dt;          // timeStepGlobal
    float h;           // gridLenGlobal = (float)(pi / nTheta)
    float invH;        // (float)(1.0 / (double)h)   (kernel/KaminoCore.cu:42)
    float halfH;       // 0.5f * h (exact)
    float cofTheta;    // dt / radius                (kernel/KaminoCore.cu:204)
    size_t cells;      // nTheta * nPhi : per-simulation stride of every field buffer
    long numParticles; // per simulation
    // theta band this launch works on (rows rowBegin .. rowBegin + rowCount - 1 of the GLOBAL grid;
    // the whole grid unless a kamino_band_* entry point narrowed it). Arrays and indices stay
    // global: a band-decomposed run keeps full-size buffers and only computes its rows.
    int rowBegin, rowCount;
};

// Field buffers of the whole batch; simulation b starts at ptr + b * cells.
struct FieldSet {
    const float* velPhi;
    const float* velTheta;
    const float* density;
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The five kernels of a step form a chain in which every kernel consumes what the previous one
// wrote. Inside a captured step graph they are launched with the programmatic-stream-
// serialization attribute: the launch of kernel k+1 is set up while kernel k drains (its blocks
// become schedulable when every block of kernel k has exited or is exiting), and every kernel
// calls pdlWait() before its first dependent global access (read OR write), which blocks until
// kernel k has completed and its writes are visible, so the chain stays transitively ordered.
// Stream capture turns the attribute into programmatic graph edges.
// Measured (r01j A/B): triggering the dependents EARLY (griddepcontrol.launch_dependents at
// kernel entry) is a large loss -- the waiting blocks of kernel k+1 take SM slots from the
// later waves of kernel k -- so no kernel triggers explicitly. Outside graph capture (phase
// entry points, which may follow a memcpy) launches are plain.
#ifdef __CUDACC__
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdlEnabled();
void pdlSetCapturing(bool capturing);   // set by the context around graph capture

// Launch of a kernel of the step chain (with the programmatic edge while a step graph is being captured).
// Measured and rejected (r02a A/B): letting the blocks of a kernel's last wave release the dependents at
// their entry -- 62.2 instead of 55.1 us/step at 512 x 1024, 462 instead of 449 us at 2048 x 4096.
template <typename... KArgs, typename... Args>
cudaError_t launchChained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdlEnabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#define KB_CUDA_OK(expr)                                                     \
    do {                                                                     \
        cudaError_t kb_err__ = (expr);                                       \
        if (kb_err__ != cudaSuccess) return (int)kb_err__;                   \
    } while (0)

} // namespace kb
