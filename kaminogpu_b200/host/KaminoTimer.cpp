#include "KaminoTimer.h"

KaminoTimer::KaminoTimer(kamino_ctx* context) : ctx(context), timeElapsed(0.0f) {}
KaminoTimer::~KaminoTimer() {}

void KaminoTimer::startTimer()
{
    if (ctx) KAMINO_CHECK(ctx, kamino_sync(ctx));
    start = std::chrono::steady_clock::now();
}

float KaminoTimer::stopTimer()
{
    if (ctx) KAMINO_CHECK(ctx, kamino_sync(ctx));
    timeElapsed = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - start).count();
    return timeElapsed;
}
