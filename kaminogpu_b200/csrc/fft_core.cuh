// Register-resident Stockham FFT building blocks (complex, length N = 2^m, m >= 4).
//
// One transform is executed by T = N/16 threads; every thread owns 16 complex values per
// pass. A pass of radix R (2, 4, 8 or 16) reads in[t + e*T] (e = 0..15), applies the
// inter-pass twiddles from a read-only table, performs 16/R R-point DFTs in registers and
// scatters the results in Stockham (autosort) order, so the final pass leaves the spectrum
// in natural order. Passes: one leading pass of radix 2^(m mod 4) (if m mod 4 != 0; it
// needs no twiddles) followed by m/4 radix-16 passes: N = 1024 is 4 x 16 x 16 (3 shared-
// memory round trips instead of 5 radix-4 ones), N = 4096 is 16 x 16 x 16.
//
// Shared-memory layout: element idx lives at idx + (idx >> 4) (one float2 of padding per
// 16), which makes both the strided Stockham scatter and the unit-stride gather
// conflict-free for 64-bit accesses.
//
// The functions are __host__ __device__ so that the index logic is unit-tested on the CPU
// (tests/fft_host_check.cu) with the "threads" of a transform executed in a loop.
#pragma once

#include <cuda_runtime.h>

namespace kb {
namespace fft {

#define KB_HD __host__ __device__ __forceinline__

KB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
KB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
KB_HD float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// multiply by -i (SIGN < 0, forward e^{-i...}) or +i (SIGN > 0)
template <int SIGN> KB_HD float2 mulI(float2 a)
{
    return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

KB_HD int pad(int idx) { return idx + (idx >> 4); }
KB_HD int paddedSize(int n) { return n + (n >> 4) + 1; }

// 2-point DFT
KB_HD void dft2(float2& a, float2& b)
{
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// 4-point DFT, natural order in and out: X[k] = sum_n x[n] w^(nk), w = exp(SIGN*2*pi*i/4)
template <int SIGN> KB_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3)
{
    const float2 a0 = cadd(x0, x2), a1 = csub(x0, x2), a2 = cadd(x1, x3);
    const float2 a3 = mulI<SIGN>(csub(x1, x3));
    x0 = cadd(a0, a2);
    x1 = cadd(a1, a3);
    x2 = csub(a0, a2);
    x3 = csub(a1, a3);
}

// multiply by exp(SIGN * 2*pi*i * m / 16) for compile-time m
template <int SIGN, int M> KB_HD float2 mulW16(float2 a)
{
    constexpr float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f;   // cos, sin(pi/8)
    constexpr float r = 0.70710678118654752440f;
    constexpr int m = M & 15;
    if (m == 0) return a;
    if (m == 4) return mulI<SIGN>(a);
    if (m == 8) return make_float2(-a.x, -a.y);
    if (m == 12) { const float2 t = mulI<SIGN>(a); return make_float2(-t.x, -t.y); }
    float wr, wi;     // w = wr + i * wi for the + sign; conj for the - sign
    if (m == 1) { wr = c1; wi = s1; }
    else if (m == 2) { wr = r; wi = r; }
    else if (m == 3) { wr = s1; wi = c1; }
    else if (m == 5) { wr = -s1; wi = c1; }
    else if (m == 6) { wr = -r; wi = r; }
    else if (m == 7) { wr = -c1; wi = s1; }
    else if (m == 9) { wr = -c1; wi = -s1; }
    else if (m == 10) { wr = -r; wi = -r; }
    else if (m == 11) { wr = -s1; wi = -c1; }
    else if (m == 13) { wr = s1; wi = -c1; }
    else if (m == 14) { wr = r; wi = -r; }
    else { wr = c1; wi = -s1; }
    if (SIGN < 0) wi = -wi;
    return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}

// 8-point DFT on v[0], v[S], ..., v[7S]; natural order in and out
template <int SIGN, int S> KB_HD void dft8(float2* v)
{
    // n = 2*n1 + n2 (n1 < 4, n2 < 2), k = k1 + 4*k2
    dft4<SIGN>(v[0 * S], v[2 * S], v[4 * S], v[6 * S]);      // n2 = 0 : Y0[k1]
    dft4<SIGN>(v[1 * S], v[3 * S], v[5 * S], v[7 * S]);      // n2 = 1 : Y1[k1]
    // twiddle Y1[k1] *= w8^k1 = w16^(2 k1)
    v[3 * S] = mulW16<SIGN, 2>(v[3 * S]);
    v[5 * S] = mulW16<SIGN, 4>(v[5 * S]);
    v[7 * S] = mulW16<SIGN, 6>(v[7 * S]);
    // X[k1 + 4 k2] = Y0[k1] + (-1)^k2 Y1[k1]; Y0[k1] sits in v[2 k1], Y1[k1] in v[2 k1 + 1]
    float2 y[8];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        y[k1] = cadd(v[(2 * k1) * S], v[(2 * k1 + 1) * S]);
        y[k1 + 4] = csub(v[(2 * k1) * S], v[(2 * k1 + 1) * S]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k * S] = y[k];
}

// 16-point DFT on v[0..15], natural order in and out
template <int SIGN> KB_HD void dft16(float2* v)
{
    // n = 4*n1 + n2, k = k1 + 4*k2
    // step 1: for every n2 a 4-point DFT over n1 (inputs v[n2], v[4+n2], v[8+n2], v[12+n2]);
    //         afterwards v[4*k1 + n2] = Y[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<SIGN>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
    // step 2: twiddle Y[n2][k1] *= w16^(n2*k1)
    v[4 * 1 + 1] = mulW16<SIGN, 1>(v[4 * 1 + 1]);
    v[4 * 1 + 2] = mulW16<SIGN, 2>(v[4 * 1 + 2]);
    v[4 * 1 + 3] = mulW16<SIGN, 3>(v[4 * 1 + 3]);
    v[4 * 2 + 1] = mulW16<SIGN, 2>(v[4 * 2 + 1]);
    v[4 * 2 + 2] = mulW16<SIGN, 4>(v[4 * 2 + 2]);
    v[4 * 2 + 3] = mulW16<SIGN, 6>(v[4 * 2 + 3]);
    v[4 * 3 + 1] = mulW16<SIGN, 3>(v[4 * 3 + 1]);
    v[4 * 3 + 2] = mulW16<SIGN, 6>(v[4 * 3 + 2]);
    v[4 * 3 + 3] = mulW16<SIGN, 9>(v[4 * 3 + 3]);
    // step 3: for every k1 a 4-point DFT over n2 (inputs v[4*k1 + 0..3]); output k2 lands in
    //         v[4*k1 + k2] = X[k1 + 4*k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<SIGN>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // step 4: transpose to natural order X[k] -> v[k], k = k1 + 4*k2
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = k1 + 1; k2 < 4; ++k2) {
            const float2 t = v[4 * k1 + k2];
            v[4 * k1 + k2] = v[4 * k2 + k1];
            v[4 * k2 + k1] = t;
        }
}

// Twiddle tables. Only the radix-16 passes after the first pass need inter-pass twiddles: the
// pass that starts from Ns already-transformed points multiplies element m (1..15) of butterfly
// k = tb mod Ns by exp(-2*pi*i * k*m / (16*Ns)). One table per such pass, laid out [m-1][k]
// (k fastest): the threads of a warp hold consecutive k for the same m, so a warp reads
// consecutive float2 (conflict-free from shared memory; with the single N-entry table indexed
// k*m*N/(16*Ns) the 16 distinct k of the second pass all fell on the same bank: 69 % of the
// shared-memory wavefronts of the r01f inverse FFT were bank-conflict replays). The tables of
// the successive passes are packed back to back: 15*Ns entries each, N - firstRadix in total.
// Forward sign stored; conjugated for SIGN > 0. (`tw` may point to global or shared memory.)
KB_HD int firstRadix(int log2N) { return (log2N & 3) ? (1 << (log2N & 3)) : 16; }
KB_HD int twiddleTableSize(int N, int log2N) { return N - firstRadix(log2N); }

template <int SIGN> KB_HD float2 twiddleAt(const float2* tw, int p)
{
    float2 w = tw[p];
    if (SIGN > 0) w.y = -w.y;
    return w;
}

// One Stockham pass of radix R over the 16 values of thread t (T = N/16 threads per transform).
//   v[e] holds in[t + e*T] on entry; the function twiddles, transforms and scatters to `out`
//   (padded shared memory, or any array indexed through pad()).
//   Ns = product of the radices of the previous passes; twPass = this pass's [m-1][k] table
//   (unused by the first pass, Ns == 1, the only one that may have R < 16).
template <int SIGN, int R>
KB_HD void passCompute(float2* v, int t, int T, int N, int Ns, const float2* twPass)
{
    if (R == 16 && Ns > 1) {
        const int k = t & (Ns - 1);
#pragma unroll
        for (int m = 1; m < 16; ++m) v[m] = cmul(v[m], twiddleAt<SIGN>(twPass, (m - 1) * Ns + k));
    }
    if (R == 16) dft16<SIGN>(v);
    else if (R == 8) { dft8<SIGN, 2>(v); dft8<SIGN, 2>(v + 1); }
    else if (R == 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dft4<SIGN>(v[q], v[q + 4], v[q + 8], v[q + 12]);
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) dft2(v[q], v[q + 8]);
    }
}

template <int R>
KB_HD void passScatter(const float2* v, float2* out, int t, int T, int Ns)
{
    constexpr int Q = 16 / R;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int tb = t + q * T;
        const int k = tb & (Ns - 1);
        const int base = (tb - k) * R + k;
#pragma unroll
        for (int m = 0; m < R; ++m) out[pad(base + m * Ns)] = v[q + m * Q];
    }
}

KB_HD void passGather(float2* v, const float2* in, int t, int T)
{
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = in[pad(t + e * T)];
}

#undef KB_HD

} // namespace fft
} // namespace kb
