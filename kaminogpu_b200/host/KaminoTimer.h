// KaminoTimer (reference: include/KaminoTimer.cuh, kernel/KaminoTimer.cu:3-22): start/stop
// timer returning milliseconds. The reference brackets cudaEvents on the default stream and
// synchronises; here the solver's context is synchronised and the host clock is read.
#pragma once

#include <chrono>

#include "KaminoHeader.h"

class KaminoTimer
{
private:
    kamino_ctx* ctx;
    std::chrono::steady_clock::time_point start;
    float timeElapsed;
public:
    explicit KaminoTimer(kamino_ctx* context = nullptr);
    ~KaminoTimer();

    void startTimer();
    float stopTimer();
};
