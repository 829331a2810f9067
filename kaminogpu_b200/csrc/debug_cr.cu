// Parity instrumentation, not part of the step: the theta solve of the projection in the REFERENCE's
// operation order -- cyclic reduction on (a, b, c, d) in fp32, once for the real and once for the imaginary
// right-hand side (kernel/tdm.cu:3-96, launched per wavenumber by kernel/KaminoCore.cu:779-792).
//
// The product solves the same systems with an LU factorisation (tridiag.cu) whose result is ~36x closer to
// an fp64 solve than the reference's own CR. The distance between the two builds' pressure is therefore
// the reference's CR rounding error; this kernel exists so that a TEST can show that: with the solve done
// in CR order (same elimination order, same fused multiply-adds as the reference's SASS: b - c*t1 - a*t2
// contracts to two FFMAs), the pressure falls
// within the north-star 1e-5 of the reference's dump. kamino_debug_project_cr (kamino_ctx.cu) runs
// divergence+FFT -> this kernel -> inverse FFT+gradient; it is never captured into a step graph.
#include "kamino_kernels.cuh"

namespace kb {

namespace {

// one block per wavenumber slot, nTheta / 2 threads, dynamic smem 7 * nTheta floats
__global__ void cyclicReductionKernel(GridParams g, SpectralTables t, float2* __restrict__ spectrumAll)
{
    extern __shared__ float sh[];
    const int n = g.nTheta, half = g.nPhi >> 1;
    float* a = sh;
    float* b = a + n;
    float* c = b + n;
    float* dRe = c + n;
    float* dIm = dRe + n;
    float* xRe = dIm + n;
    float* xIm = xRe + n;
    const int slot = blockIdx.x;
    const int wave = (slot == 0) ? half : slot;              // Nyquist mode in slot 0
    const float nSq = (float)(wave * wave);
    float2* spectrum = spectrumAll + (size_t)blockIdx.y * (g.cells >> 1) + slot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        // precomputeABCKernel, kernel/KaminoSolver.cu:128-153 (as tridiag.cu builds them)
        float va = t.triA[i], vc = t.triC[i];
        float vb = (float)(t.minusTwoOverH2 - (double)__fdiv_rn(nSq, t.sinSq[i]));
        if (i == 0) { vb = __fadd_rn(vb, va); va = 0.0f; }
        if (i == n - 1) { vb = __fadd_rn(vb, vc); vc = 0.0f; }
        a[i] = va; b[i] = vb; c[i] = vc;
        const float2 d = spectrum[(size_t)i * half];
        dRe[i] = d.x; dIm[i] = d.y;
    }
    int levels = 0;
    while ((2 << levels) < n) ++levels;                      // log2(n / 2), kernel/tdm.cu:12
    int stride = 1, active = n / 2;
    // forward elimination, kernel/tdm.cu:43-63
    for (int lvl = 0; lvl < levels; ++lvl) {
        __syncthreads();
        stride *= 2;
        const int delta = stride / 2;
        if ((int)threadIdx.x < active) {
            const int i = stride * threadIdx.x + stride - 1;
            const int iLeft = i - delta;
            int iRight = i + delta;
            if (iRight >= n) iRight = n - 1;
            const float tmp1 = __fdiv_rn(a[i], b[iLeft]);
            const float tmp2 = __fdiv_rn(c[i], b[iRight]);
            const float bi = __fmaf_rn(-a[iRight], tmp2, __fmaf_rn(-c[iLeft], tmp1, b[i]));
            const float dr = __fmaf_rn(-dRe[iRight], tmp2, __fmaf_rn(-dRe[iLeft], tmp1, dRe[i]));
            const float di = __fmaf_rn(-dIm[iRight], tmp2, __fmaf_rn(-dIm[iLeft], tmp1, dIm[i]));
            const float ai = __fmul_rn(-a[iLeft], tmp1);
            const float ci = __fmul_rn(-c[iRight], tmp2);
            // the reference updates in place without a barrier between its reads and writes: rows
            // i of this level are only read by their owner, so the order within the level is free
            b[i] = bi; dRe[i] = dr; dIm[i] = di; a[i] = ai; c[i] = ci;
        }
        active /= 2;
    }
    __syncthreads();
    // 2 x 2 system, kernel/tdm.cu:65-72
    if (threadIdx.x == 0) {
        const int p = stride - 1, q = 2 * stride - 1;
        const float det = __fmaf_rn(b[q], b[p], -__fmul_rn(c[p], a[q]));
        xRe[p] = __fdiv_rn(__fmaf_rn(b[q], dRe[p], -__fmul_rn(c[p], dRe[q])), det);
        xRe[q] = __fdiv_rn(__fmaf_rn(dRe[q], b[p], -__fmul_rn(dRe[p], a[q])), det);
        xIm[p] = __fdiv_rn(__fmaf_rn(b[q], dIm[p], -__fmul_rn(c[p], dIm[q])), det);
        xIm[q] = __fdiv_rn(__fmaf_rn(dIm[q], b[p], -__fmul_rn(dIm[p], a[q])), det);
    }
    // back substitution, kernel/tdm.cu:75-90
    active = 2;
    for (int lvl = 0; lvl < levels; ++lvl) {
        const int delta = stride / 2;
        __syncthreads();
        if ((int)threadIdx.x < active) {
            const int i = stride * threadIdx.x + stride / 2 - 1;
            if (i == delta - 1) {
                xRe[i] = __fdiv_rn(__fmaf_rn(-c[i], xRe[i + delta], dRe[i]), b[i]);
                xIm[i] = __fdiv_rn(__fmaf_rn(-c[i], xIm[i + delta], dIm[i]), b[i]);
            } else {
                xRe[i] = __fdiv_rn(__fmaf_rn(-c[i], xRe[i + delta], __fmaf_rn(-a[i], xRe[i - delta], dRe[i])), b[i]);
                xIm[i] = __fdiv_rn(__fmaf_rn(-c[i], xIm[i + delta], __fmaf_rn(-a[i], xIm[i - delta], dIm[i])), b[i]);
            }
        }
        stride /= 2;
        active *= 2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) spectrum[(size_t)i * half] = make_float2(xRe[i], xIm[i]);
}

} // namespace

// nTheta <= 2048: the block and shared-memory limits of the reference's own launch (kernel/KaminoCore.cu:779-784)
cudaError_t launchCyclicReductionDebug(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch, cudaStream_t stream)
{
    if (g.nTheta > 2048) return cudaErrorInvalidValue;
    const size_t smem = 7 * (size_t)g.nTheta * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cyclicReductionKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(g.nPhi / 2, batch);
    cyclicReductionKernel<<<grid, g.nTheta / 2, smem, stream>>>(g, t, spectrum);
    return cudaGetLastError();
}

} // namespace kb
