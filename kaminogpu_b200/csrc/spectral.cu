// Pressure projection: divergence + FFT along phi, per-wavenumber tridiagonal solves along
// theta, inverse FFT + gradient subtraction. Three launches.
//
// Replaces fillDivergenceKernel, cufftExecC2C (x2), shiftFKernel, crKernel (x2),
// copy2UFourier, cacheZeroComponents, shiftUKernel, applyPressureTheta, applyPressurePhi
// and KaminoSolver::projection (kernel/KaminoCore.cu:587-842, kernel/tdm.cu), i.e. eleven
// launches, two cuFFT executions and ten device syncs.
//
// What the reference computes (verified against its dumps):
//   F_n(theta_j) = 1/N sum_m div[j][m] exp(-i n phi_m),  n = -N/2 .. N/2-1
//   tridiagonal solve in theta for every n != 0 (n = 0 is an identity "solve" whose result
//   is subtracted again, KaminoSolver.cu:154-159 + KaminoCore.cu:692-700)
//   p[j][i] = Re sum_{n != 0} U_n(theta_j) exp(+i n phi_i)
// The divergence is real, so only n = 1 .. N/2 is needed (U_-n = conj U_n); the half
// spectrum of a row is N/2 complex values, stored with the Nyquist mode in slot 0.
//
// FFT strategy: every transform is ONE complex FFT of length N per block, staged in shared
// memory (Stockham autosort, radix 4, twiddles from a read-only table):
//   forward : two divergence rows are packed as z = div_j + i div_{j+1}; their spectra are
//             separated with the Hermitian symmetry.
//   inverse : row j needs p_j (for the phi gradient) and p_{j+1} - p_j (for the theta
//             gradient). By linearity both come from one transform of X + iY with
//             X = U_j, Y = U_{j+1} - U_j, so no block ever needs another block's output.
#include "kamino_kernels.cuh"

namespace kb {

namespace {

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// In-place-semantics complex FFT of length N held in shared memory, executed by N/4
// threads (tid = 0 .. N/4-1). `a` holds the input, `b` is scratch of the same size; returns
// the buffer that holds the result (natural order). SIGN = -1: exp(-2 pi i k m / N).
// Callers must __syncthreads() after filling `a`; the result is synchronised on return.
template <int SIGN>
__device__ __forceinline__ float2* fftShared(float2* a, float2* b, int N, int log2N, int tid,
                                             const float2* __restrict__ twiddle)
{
    const int quarter = N >> 2;
    int Ns = 1;
    if (log2N & 1) {
        // leading radix-2 stage: two butterflies per thread
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = tid + h * quarter;
            const float2 v0 = a[j], v1 = a[j + (N >> 1)];
            b[2 * j] = cadd(v0, v1);
            b[2 * j + 1] = csub(v0, v1);
        }
        float2* t = a; a = b; b = t;
        Ns = 2;
        __syncthreads();
    }
    while (Ns < N) {
        const int k = tid & (Ns - 1);
        const int base = k * (N / (4 * Ns));
        float2 v0 = a[tid], v1 = a[tid + quarter], v2 = a[tid + 2 * quarter], v3 = a[tid + 3 * quarter];
        if (Ns > 1) {
            float2 w1 = __ldg(twiddle + base), w2 = __ldg(twiddle + 2 * base), w3 = __ldg(twiddle + 3 * base);
            if (SIGN > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
        }
        const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3);
        const float2 d = csub(v1, v3);
        const float2 a3 = (SIGN < 0) ? make_float2(d.y, -d.x) : make_float2(-d.y, d.x);
        const int idx = ((tid - k) << 2) + k;
        b[idx] = cadd(a0, a2);
        b[idx + Ns] = cadd(a1, a3);
        b[idx + 2 * Ns] = csub(a0, a2);
        b[idx + 3 * Ns] = csub(a1, a3);
        float2* t = a; a = b; b = t;
        Ns <<= 2;
        __syncthreads();
    }
    return a;
}

// ---- tables ---------------------------------------------------------------------------

__global__ void buildTablesKernel(GridParams g, SpectralTables t)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < g.nPhi) {
        double s, c;
        sincospi(-2.0 * (double)k / (double)g.nPhi, &s, &c);
        t.twiddle[k] = make_float2((float)c, (float)s);
    }
    if (k < g.nTheta) {
        const float h = g.h;
        // theta of row k: (k + centeredThetaOffset) * gridLen, an exact fp64 product rounded once
        const float theta = __fmul_rn(__fadd_rn((float)k, 0.5f), h);
        const float sinT = sinf(theta), cosT = cosf(theta);
        // fillDivergenceKernel row constants, kernel/KaminoCore.cu:596-628
        const float halfStep = __fmul_rn(0.5f, h);
        t.sinSouth[k] = sinf(__fadd_rn(theta, halfStep));
        t.sinNorth[k] = sinf(__fsub_rn(theta, halfStep));
        t.divFactor[k] = __fdiv_rn(__fdiv_rn(1.0f, sinT), h);
        // applyPressurePhi, kernel/KaminoCore.cu:743-744
        t.gradPhiDenom[k] = __fmul_rn(-h, sinT);
        // precomputeABCKernel, kernel/KaminoSolver.cu:128-138
        const double h2 = (double)__fmul_rn(h, h);
        const double cot = (double)cosT / 2.0 / (double)h / (double)sinT;
        t.triA[k] = (float)(1.0 / h2 - cot);
        t.triC[k] = (float)(1.0 / h2 + cot);
        t.sinSq[k] = __fmul_rn(sinT, sinT);
    }
}

// ---- K4: divergence + forward FFT ------------------------------------------------------

// divergence of cell (j, i), kernel/KaminoCore.cu:598-632
__device__ __forceinline__ float divergenceAt(const GridParams& g, const SpectralTables& t,
                                              const float* __restrict__ velPhi,
                                              const float* __restrict__ velTheta, int j, int i)
{
    const int N = g.nPhi;
    const float uWest = __ldg(velPhi + (size_t)j * N + i);
    const float uEast = __ldg(velPhi + (size_t)j * N + ((i + 1) & (N - 1)));
    float vNorth = 0.0f, vSouth = 0.0f;
    if (j != 0) vNorth = __ldg(velTheta + (size_t)(j - 1) * N + i);
    if (j != g.nTheta - 1) vSouth = __ldg(velTheta + (size_t)j * N + i);
    const float factor = __ldg(t.divFactor + j);
    const float termTheta = __fmul_rn(factor, __fmaf_rn(vSouth, __ldg(t.sinSouth + j),
                                                        -__fmul_rn(vNorth, __ldg(t.sinNorth + j))));
    return __fmaf_rn(factor, __fsub_rn(uEast, uWest), termTheta);
}

// grid (nTheta/2, batch), block N/4 threads, dynamic smem 2 * N * sizeof(float2)
__global__ void divergenceFFTKernel(GridParams g, SpectralTables t, const float* __restrict__ velPhiAll,
                                    const float* __restrict__ velThetaAll, float2* __restrict__ spectrumAll)
{
    extern __shared__ float2 smem[];
    const int N = g.nPhi, half = N >> 1;
    float2* bufA = smem;
    float2* bufB = smem + N;
    const int sim = blockIdx.y;
    const float* velPhi = velPhiAll + (size_t)sim * g.cells;
    const float* velTheta = velThetaAll + (size_t)sim * g.cells;
    float2* spectrum = spectrumAll + (size_t)sim * (g.cells >> 1);
    const int j = 2 * blockIdx.x;
    const int tid = threadIdx.x;

#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = tid + r * (N >> 2);
        bufA[i] = make_float2(divergenceAt(g, t, velPhi, velTheta, j, i),
                              divergenceAt(g, t, velPhi, velTheta, j + 1, i));
    }
    __syncthreads();
    const float2* Z = fftShared<-1>(bufA, bufB, N, g.log2NPhi, tid, t.twiddle);

    // separate the two real rows and scale by 1/N (shiftFKernel, kernel/KaminoCore.cu:652-653)
    const float scale = 0.5f / (float)N;
    float2* rowA = spectrum + (size_t)j * half;
    float2* rowB = spectrum + (size_t)(j + 1) * half;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int k = tid + r * (N >> 2);          // 0 .. N/2-1
        if (k == 0) {
            const float2 zn = Z[half];               // Nyquist: both rows real
            rowA[0] = make_float2(2.0f * scale * zn.x, 0.0f);
            rowB[0] = make_float2(2.0f * scale * zn.y, 0.0f);
        } else {
            const float2 zk = Z[k], zc = Z[N - k];
            // A_k = (Z_k + conj Z_{N-k}) / 2,  B_k = (Z_k - conj Z_{N-k}) / (2i)
            rowA[k] = make_float2(scale * (zk.x + zc.x), scale * (zk.y - zc.y));
            rowB[k] = make_float2(scale * (zk.y + zc.y), scale * (zc.x - zk.x));
        }
    }
}

// ---- K5: tridiagonal solves (cyclic reduction in the reference's elimination order) ----

__device__ __forceinline__ int padIdx(int i) { return i + (i >> 5); }

// grid (N/2 / W, batch), block W * nTheta/2 threads. Each block solves W wavenumber slots,
// real and imaginary right-hand sides together (the reference runs crKernel twice and
// reloads a, b, c, kernel/KaminoCore.cu:779-792). Coefficients are generated in the kernel
// from per-row tables (precomputeABCKernel, kernel/KaminoSolver.cu:117-163).
__global__ void tridiagonalKernel(GridParams g, SpectralTables t, float2* __restrict__ spectrumAll, int W)
{
    extern __shared__ float smemF[];
    const int nT = g.nTheta, N = g.nPhi, half = N >> 1;
    const int L = nT + (nT >> 5) + 1;               // padded length of one array
    const int tid = threadIdx.x;
    const int w = tid % W;                          // which system of this block
    const int th = tid / W;                         // 0 .. nT/2-1
    float* a = smemF + (size_t)w * 7 * L;
    float* b = a + L;
    float* c = b + L;
    float* dr = c + L;
    float* di = dr + L;
    float* xr = di + L;
    float* xi = xr + L;

    float2* spectrum = spectrumAll + (size_t)blockIdx.y * (g.cells >> 1);
    const int slot = blockIdx.x * W + w;
    const int n = (slot == 0) ? half : slot;        // wavenumber of this slot (never 0)
    const float nSq = (float)(n * n);

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = th + r * (nT >> 1);
        float valA = __ldg(t.triA + i), valC = __ldg(t.triC + i);
        float valB = (float)(t.minusTwoOverH2 - (double)__fdiv_rn(nSq, __ldg(t.sinSq + i)));
        if (i == 0) { valB = __fadd_rn(valB, valA); valA = 0.0f; }
        if (i == nT - 1) { valB = __fadd_rn(valB, valC); valC = 0.0f; }
        const float2 f = spectrum[(size_t)i * half + slot];
        const int p = padIdx(i);
        a[p] = valA; b[p] = valB; c[p] = valC; dr[p] = f.x; di[p] = f.y;
    }

    // forward elimination, kernel/tdm.cu:43-63
    int stride = 1;
    int numThreads = nT >> 1;
    int iteration = 0;
    while ((2 << iteration) < nT) ++iteration;      // log2(nT / 2)
    for (int lvl = 0; lvl < iteration; ++lvl) {
        __syncthreads();
        stride <<= 1;
        const int delta = stride >> 1;
        if (th < numThreads) {
            const int i = stride * th + stride - 1;
            const int iLeft = i - delta;
            int iRight = i + delta;
            if (iRight >= nT) iRight = nT - 1;
            const int pi = padIdx(i), pl = padIdx(iLeft), pr = padIdx(iRight);
            const float tmp1 = __fdiv_rn(a[pi], b[pl]);
            const float tmp2 = __fdiv_rn(c[pi], b[pr]);
            const float bi = __fmaf_rn(a[pr], -tmp2, __fmaf_rn(c[pl], -tmp1, b[pi]));
            const float dri = __fmaf_rn(dr[pr], -tmp2, __fmaf_rn(dr[pl], -tmp1, dr[pi]));
            const float dii = __fmaf_rn(di[pr], -tmp2, __fmaf_rn(di[pl], -tmp1, di[pi]));
            const float ai = __fmul_rn(a[pl], -tmp1);
            const float ci = __fmul_rn(c[pr], -tmp2);
            b[pi] = bi; dr[pi] = dri; di[pi] = dii; a[pi] = ai; c[pi] = ci;
        }
        numThreads >>= 1;
    }
    __syncthreads();
    // 2 x 2 system, kernel/tdm.cu:65-72
    if (th < 2) {
        const int p1 = padIdx(stride - 1), p2 = padIdx(2 * stride - 1);
        const float det = __fmaf_rn(b[p2], b[p1], -__fmul_rn(c[p1], a[p2]));
        if (th == 0) {
            xr[p1] = __fdiv_rn(__fmaf_rn(b[p2], dr[p1], -__fmul_rn(c[p1], dr[p2])), det);
            xi[p1] = __fdiv_rn(__fmaf_rn(b[p2], di[p1], -__fmul_rn(c[p1], di[p2])), det);
        } else {
            xr[p2] = __fdiv_rn(__fmaf_rn(dr[p2], b[p1], -__fmul_rn(dr[p1], a[p2])), det);
            xi[p2] = __fdiv_rn(__fmaf_rn(di[p2], b[p1], -__fmul_rn(di[p1], a[p2])), det);
        }
    }
    // back substitution, kernel/tdm.cu:75-90
    numThreads = 2;
    for (int lvl = 0; lvl < iteration; ++lvl) {
        const int delta = stride >> 1;
        __syncthreads();
        if (th < numThreads) {
            const int i = stride * th + (stride >> 1) - 1;
            const int pi = padIdx(i), pp = padIdx(i + delta);
            if (i == delta - 1) {
                xr[pi] = __fdiv_rn(__fmaf_rn(-c[pi], xr[pp], dr[pi]), b[pi]);
                xi[pi] = __fdiv_rn(__fmaf_rn(-c[pi], xi[pp], di[pi]), b[pi]);
            } else {
                const int pm = padIdx(i - delta);
                xr[pi] = __fdiv_rn(__fmaf_rn(-c[pi], xr[pp], __fmaf_rn(-a[pi], xr[pm], dr[pi])), b[pi]);
                xi[pi] = __fdiv_rn(__fmaf_rn(-c[pi], xi[pp], __fmaf_rn(-a[pi], xi[pm], di[pi])), b[pi]);
            }
        }
        stride >>= 1;
        numThreads <<= 1;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = th + r * (nT >> 1);
        const int p = padIdx(i);
        spectrum[(size_t)i * half + slot] = make_float2(xr[p], xi[p]);
    }
}

// ---- K6: inverse FFT + gradient subtraction ---------------------------------------------

// grid (nTheta, batch), block N/4 threads, dynamic smem 2 * N * sizeof(float2)
__global__ void inverseFFTGradientKernel(GridParams g, SpectralTables t, const float2* __restrict__ spectrumAll,
                                         float* __restrict__ velPhiAll, float* __restrict__ velThetaAll,
                                         float* __restrict__ pressureAll)
{
    extern __shared__ float2 smem[];
    const int N = g.nPhi, half = N >> 1, nT = g.nTheta;
    float2* bufA = smem;
    float2* bufB = smem + N;
    const int sim = blockIdx.y;
    const float2* spectrum = spectrumAll + (size_t)sim * (g.cells >> 1);
    float* velPhi = velPhiAll + (size_t)sim * g.cells;
    float* velTheta = velThetaAll + (size_t)sim * g.cells;
    const int j = blockIdx.x;
    const int tid = threadIdx.x;
    const bool hasSouth = (j < nT - 1);
    const float2* rowU = spectrum + (size_t)j * half;
    const float2* rowS = spectrum + (size_t)(hasSouth ? j + 1 : j) * half;

    // W_k = X_k + i Y_k with X = U_j, Y = U_{j+1} - U_j (Hermitian completions), X_0 = Y_0 = 0
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int k = tid + r * (N >> 2);
        const float2 x = rowU[k];
        float2 y = make_float2(0.0f, 0.0f);
        if (hasSouth) { const float2 s = rowS[k]; y = make_float2(s.x - x.x, s.y - x.y); }
        if (k == 0) {
            bufA[0] = make_float2(0.0f, 0.0f);
            bufA[half] = make_float2(x.x, y.x);            // Nyquist: real parts only
        } else {
            bufA[k] = make_float2(x.x - y.y, x.y + y.x);
            bufA[N - k] = make_float2(x.x + y.y, y.x - x.y);
        }
    }
    __syncthreads();
    const float2* z = fftShared<+1>(bufA, bufB, N, g.log2NPhi, tid, t.twiddle);

    const float denomPhi = __ldg(t.gradPhiDenom + j);
    const float negH = -g.h;
    float* pressure = pressureAll ? pressureAll + (size_t)sim * g.cells : nullptr;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = tid + r * (N >> 2);
        const float2 zi = z[i];
        const float pWest = z[(i - 1) & (N - 1)].x;
        // applyPressurePhi, kernel/KaminoCore.cu:740-746
        const size_t at = (size_t)j * N + i;
        velPhi[at] = __fadd_rn(velPhi[at], __fdiv_rn(__fsub_rn(zi.x, pWest), denomPhi));
        // applyPressureTheta, kernel/KaminoCore.cu:716-721 (zi.y = p[j+1][i] - p[j][i])
        if (hasSouth) velTheta[at] = __fadd_rn(velTheta[at], __fdiv_rn(zi.y, negH));
        if (pressure) pressure[at] = zi.x;
    }
}

int tridiagonalWidth(const GridParams& g)
{
    // W systems per block: W * nTheta/2 threads <= 1024, W <= 4 (32-byte row segments)
    int W = 2048 / g.nTheta;
    if (W > 4) W = 4;
    if (W < 1) W = 1;
    while ((g.nPhi / 2) % W) W >>= 1;
    return W;
}

size_t tridiagonalSmem(const GridParams& g, int W)
{
    const int L = g.nTheta + (g.nTheta >> 5) + 1;
    return (size_t)W * 7 * L * sizeof(float);
}

} // namespace

size_t spectralTableBytes(const GridParams& g)
{
    return sizeof(float2) * g.nPhi + sizeof(float) * 7 * g.nTheta + 256 * 8;
}

cudaError_t launchBuildTables(const GridParams& g, SpectralTables t, cudaStream_t stream)
{
    const int threads = 256;
    const int blocks = (g.nPhi + threads - 1) / threads;
    buildTablesKernel<<<blocks, threads, 0, stream>>>(g, t);
    return cudaGetLastError();
}

cudaError_t configureKernels(const GridParams& g)
{
    cudaError_t e;
    const size_t fftSmem = 2 * (size_t)g.nPhi * sizeof(float2);
    e = cudaFuncSetAttribute(divergenceFFTKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fftSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(inverseFFTGradientKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fftSmem);
    if (e != cudaSuccess) return e;
    const int W = tridiagonalWidth(g);
    e = cudaFuncSetAttribute(tridiagonalKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tridiagonalSmem(g, W));
    return e;
}

cudaError_t launchDivergenceFFT(const GridParams& g, const SpectralTables& t, const float* velPhi,
                                const float* velTheta, float2* spectrum, int batch, cudaStream_t stream)
{
    dim3 grid(g.nTheta / 2, batch);
    divergenceFFTKernel<<<grid, g.nPhi / 4, 2 * (size_t)g.nPhi * sizeof(float2), stream>>>(
        g, t, velPhi, velTheta, spectrum);
    return cudaGetLastError();
}

cudaError_t launchTridiagonal(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch,
                              cudaStream_t stream)
{
    const int W = tridiagonalWidth(g);
    dim3 grid((g.nPhi / 2) / W, batch);
    tridiagonalKernel<<<grid, W * (g.nTheta / 2), tridiagonalSmem(g, W), stream>>>(g, t, spectrum, W);
    return cudaGetLastError();
}

cudaError_t launchInverseFFTGradient(const GridParams& g, const SpectralTables& t, const float2* spectrum,
                                     float* velPhi, float* velTheta, float* pressure, int batch,
                                     cudaStream_t stream)
{
    dim3 grid(g.nTheta, batch);
    inverseFFTGradientKernel<<<grid, g.nPhi / 4, 2 * (size_t)g.nPhi * sizeof(float2), stream>>>(
        g, t, spectrum, velPhi, velTheta, pressure);
    return cudaGetLastError();
}

} // namespace kb
