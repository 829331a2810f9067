// CPU check of the index logic of kaminogpu_b200/csrc/fft_core.cuh: the "threads" of one
// transform are executed in a loop, pass by pass, and the result is compared with a direct
// O(N^2) DFT in double precision. Built and run by tests/test_fft_host.py (nvcc, host only).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../kaminogpu_b200/csrc/fft_core.cuh"

using namespace kb::fft;

template <int SIGN>
static void runFFT(std::vector<float2>& data, int N, int log2N, const std::vector<float2>& tw)
{
    const int T = N / 16;
    std::vector<float2> a(paddedSize(N)), b(paddedSize(N));
    for (int i = 0; i < N; ++i) a[pad(i)] = data[i];
    int Ns = 1;
    const int lead = log2N & 3;
    bool first = true;
    const float2* twPass = tw.data();
    while (Ns < N) {
        int R = 16;
        if (first && lead) R = 1 << lead;
        for (int t = 0; t < T; ++t) {
            float2 v[16];
            passGather(v, a.data(), t, T);
            if (R == 16) { passCompute<SIGN, 16>(v, t, T, N, Ns, twPass); passScatter<16>(v, b.data(), t, T, Ns); }
            else if (R == 8) { passCompute<SIGN, 8>(v, t, T, N, Ns, twPass); passScatter<8>(v, b.data(), t, T, Ns); }
            else if (R == 4) { passCompute<SIGN, 4>(v, t, T, N, Ns, twPass); passScatter<4>(v, b.data(), t, T, Ns); }
            else { passCompute<SIGN, 2>(v, t, T, N, Ns, twPass); passScatter<2>(v, b.data(), t, T, Ns); }
        }
        a.swap(b);
        if (!first) twPass += 15 * Ns;
        Ns *= R;
        first = false;
    }
    for (int i = 0; i < N; ++i) data[i] = a[pad(i)];
}

int main()
{
    int failures = 0;
    for (int log2N = 4; log2N <= 14; ++log2N) {
        const int N = 1 << log2N;
        // packed per-pass tables [m-1][k] (fft_core.cuh), built as buildTablesKernel does
        std::vector<float2> tw(N);
        {
            int off = 0;
            for (int Ns = firstRadix(log2N); Ns < N; Ns <<= 4) {
                for (int m = 1; m < 16; ++m)
                    for (int k = 0; k < Ns; ++k) {
                        const double ang = -2.0 * M_PI * (double)(k * m) / (double)(16 * Ns);
                        tw[off + (m - 1) * Ns + k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
                    }
                off += 15 * Ns;
            }
            if (off != twiddleTableSize(N, log2N)) { std::printf("table size mismatch\n"); return 1; }
        }
        for (int sign = -1; sign <= 1; sign += 2) {
            std::vector<float2> x(N);
            srand(17 + log2N);
            for (auto& v : x) v = make_float2((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
            std::vector<float2> y = x;
            if (sign < 0) runFFT<-1>(y, N, log2N, tw); else runFFT<+1>(y, N, log2N, tw);
            // direct DFT on a subset of bins (all bins for small N)
            double maxErr = 0.0, maxRef = 0.0;
            const int stride = N <= 1024 ? 1 : N / 257;
            for (int k = 0; k < N; k += stride) {
                double re = 0.0, im = 0.0;
                for (int n = 0; n < N; ++n) {
                    const double ang = sign * 2.0 * M_PI * (double)((long)k * n % N) / N;
                    const double c = std::cos(ang), s = std::sin(ang);
                    re += x[n].x * c - x[n].y * s;
                    im += x[n].x * s + x[n].y * c;
                }
                maxErr = std::fmax(maxErr, std::hypot(re - y[k].x, im - y[k].y));
                maxRef = std::fmax(maxRef, std::hypot(re, im));
            }
            const bool ok = maxErr <= 2e-6 * maxRef * log2N;
            std::printf("N=%5d sign=%+d maxErr/maxRef=%.2e %s\n", N, sign, maxErr / maxRef, ok ? "ok" : "FAIL");
            if (!ok) ++failures;
        }
    }
    return failures ? 1 : 0;
}
