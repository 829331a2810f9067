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
// FFT strategy: every transform is ONE complex FFT of length N executed by N/16 threads, 16
// values per thread in registers, radix-16 Stockham passes through padded shared memory
// (fft_core.cuh: N = 1024 is 4 x 16 x 16, N = 4096 is 16 x 16 x 16), twiddles from a table:
//   forward : two divergence rows are packed as z = div_j + i div_{j+1}; their spectra are
//             separated with the Hermitian symmetry.
//   inverse : row j needs p_j (for the phi gradient) and p_{j+1} - p_j (for the theta
//             gradient). By linearity both come from one transform of X + iY with
//             X = U_j, Y = U_{j+1} - U_j, so no block ever needs another block's output.
#include <cstdlib>

#include "kamino_kernels.cuh"
#include "fft_core.cuh"
#include "tma_bulk.cuh"

namespace kb {

namespace {

// Complex FFT of length N by the T = N/16 threads of one transform (t = 0 .. T-1), entirely
// from registers: on entry v[e] = x[t + e*T]; on exit v[m] = X[t + m*T] and the whole
// spectrum also sits in `buf` (padded shared memory, natural order, synchronised).
// Every thread of the block must call this (it contains __syncthreads()).
// `tw` is the twiddle table (global, or its shared-memory copy once `twReady` has completed:
// the first pass needs no twiddles, so the wait sits after it).
// LOG2N > 0: the transform length is a compile-time constant (the BASELINE sizes), so T, the pass sizes Ns and the
// padded shared-memory offsets of the scatters / gathers fold into immediates: in the runtime-N build 34 % of the
// executed instructions of the FFT kernels were IADD3 / LEA / SHF address arithmetic (r02j SASS profile).
// LOG2N == 0: any power of two >= 32, sizes read from the arguments. Same arithmetic either way.
template <int SIGN, int LOG2N>
__device__ __forceinline__ void fftFromRegisters(float2* v, float2* buf, int t, int T, int N, int log2N,
                                                 const float2* tw, uint64_t* twReady)
{
    using namespace fft;
    if constexpr (LOG2N != 0) {
        constexpr int kN = 1 << LOG2N, kT = kN >> 4;
        constexpr int kFirst = (LOG2N & 3) ? (1 << (LOG2N & 3)) : 16;
        passCompute<SIGN, kFirst>(v, t, kT, kN, 1, tw);
        passScatter<kFirst>(v, buf, t, kT, 1);
        if (twReady) tma::mbarWait(twReady, 0);
        __syncthreads();
        const float2* twPass = tw;                 // packed per-pass tables, fft_core.cuh
#pragma unroll
        for (int Ns = kFirst; Ns < kN; Ns <<= 4) {
            passGather(v, buf, t, kT);
            __syncthreads();                       // everyone has read before anyone overwrites
            passCompute<SIGN, 16>(v, t, kT, kN, Ns, twPass);
            passScatter<16>(v, buf, t, kT, Ns);
            twPass += 15 * Ns;
            __syncthreads();
        }
    } else {
        int Ns;
        switch (log2N & 3) {
        case 1: passCompute<SIGN, 2>(v, t, T, N, 1, tw); passScatter<2>(v, buf, t, T, 1); Ns = 2; break;
        case 2: passCompute<SIGN, 4>(v, t, T, N, 1, tw); passScatter<4>(v, buf, t, T, 1); Ns = 4; break;
        case 3: passCompute<SIGN, 8>(v, t, T, N, 1, tw); passScatter<8>(v, buf, t, T, 1); Ns = 8; break;
        default: passCompute<SIGN, 16>(v, t, T, N, 1, tw); passScatter<16>(v, buf, t, T, 1); Ns = 16; break;
        }
        if (twReady) tma::mbarWait(twReady, 0);
        __syncthreads();
        const float2* twPass = tw;                 // packed per-pass tables, fft_core.cuh
        while (Ns < N) {
            passGather(v, buf, t, T);
            __syncthreads();                       // everyone has read before anyone overwrites
            passCompute<SIGN, 16>(v, t, T, N, Ns, twPass);
            passScatter<16>(v, buf, t, T, Ns);
            twPass += 15 * Ns;
            Ns <<= 4;
            __syncthreads();
        }
    }
}

// ---- tables ---------------------------------------------------------------------------

__global__ void buildTablesKernel(GridParams g, SpectralTables t)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < fft::twiddleTableSize(g.nPhi, g.log2NPhi)) {
        // entry k of the packed per-pass tables: pass with Ns points done, element m, butterfly kk
        int Ns = fft::firstRadix(g.log2NPhi), rel = k;
        while (rel >= 15 * Ns) { rel -= 15 * Ns; Ns <<= 4; }
        const int m = rel / Ns + 1, kk = rel - (m - 1) * Ns;
        double s, c;
        sincospi(-2.0 * (double)(kk * m) / (double)(16 * Ns), &s, &c);
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
        // geometricFillKernel, kernel/KaminoCore.cu:494
        t.geoG[k] = __fdiv_rn(__fmul_rn(g.dt, cosT), __fmul_rn(g.radius, sinT));
        // advection kernels, kernel/KaminoCore.cu:196,202-203 / 241,247-248 / 286,292-293
        t.cofPhiCentred[k] = __fdiv_rn(g.dt, __fmul_rn(g.radius, sinT));
        const float thetaV = __fmul_rn(__fadd_rn((float)k, 1.0f), h);
        t.cofPhiTheta[k] = __fdiv_rn(g.dt, __fmul_rn(g.radius, sinf(thetaV)));
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

// Shared-memory twiddle staging: one thread starts a TMA bulk copy of the N-entry table
// (N * 8 bytes, L2-resident) while everybody loads its inputs; the table is first needed by
// the second pass. Returns the barrier to wait on (NULL when the table stays in global memory).
template <bool STAGE>
__device__ __forceinline__ uint64_t* stageTwiddles(const float2* __restrict__ twGlobal, float2* twShared,
                                                   uint64_t* bar, int twEntries)
{
    if (!STAGE) return nullptr;
    if (threadIdx.x == 0) { tma::mbarInit(bar, 1); tma::fenceBarrierInit(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbarExpectTx(bar, (uint32_t)(twEntries * sizeof(float2)));
        tma::bulkLoad(twShared, twGlobal, (uint32_t)(twEntries * sizeof(float2)), bar);
    }
    return bar;
}

// Each transform handles a pair of theta rows (z = div_j + i div_{j+1}) with T = N/16 threads;
// a block of BLOCK threads holds BLOCK/T transforms. grid (ceil(nTheta/2 / (BLOCK/T)), batch),
// dynamic smem: [twiddles N float2 if STAGE] + (BLOCK/T) * paddedSize(N) float2.
template <int BLOCK, bool STAGE, int LOG2N, bool PACKED>
__global__ void __launch_bounds__(BLOCK)
divergenceFFTKernel(GridParams g, SpectralTables t, const float* __restrict__ velPhiAll,
                    const float* __restrict__ velThetaAll, float2* __restrict__ spectrumAll, SpectrumLayout lay)
{
    extern __shared__ __align__(16) float2 smem[];
    __shared__ __align__(8) uint64_t twBar;
    const int log2N = LOG2N ? LOG2N : g.log2NPhi;
    const int N = 1 << log2N, half = N >> 1, log2T = log2N - 4, T = 1 << log2T;
    const int local = threadIdx.x >> log2T, tt = threadIdx.x & (T - 1);
    const int pairs = (g.rowBegin + g.rowCount) >> 1;             // one past the last pair of the band
    const int pairRaw = (g.rowBegin >> 1) + blockIdx.x * (BLOCK >> log2T) + local;
    const bool valid = pairRaw < pairs;
    const int j = 2 * (valid ? pairRaw : pairs - 1);
    float2* twShared = smem;
    float2* buf = smem + (STAGE ? N : 0) + (size_t)local * fft::paddedSize(N);
    const int sim = blockIdx.y;
    const float* velPhi = velPhiAll + (size_t)sim * g.cells;
    const float* velTheta = velThetaAll + (size_t)sim * g.cells;
    float2* spectrum = spectrumAll + (size_t)sim * (g.cells >> 1);

    uint64_t* twReady = stageTwiddles<STAGE>(t.twiddle, twShared, &twBar, fft::twiddleTableSize(N, log2N));
    const float2* tw = STAGE ? twShared : t.twiddle;
    pdlWait();                                   // the velocity comes from the previous kernel

    // divergence of rows j and j + 1 (fillDivergenceKernel, kernel/KaminoCore.cu:598-632) at the
    // 16 columns i = tt + e*T of this thread; all loads of a half are issued before they are used
    const float* uA = velPhi + (size_t)j * N;                  // u_phi rows j, j+1
    const float* uB = uA + N;
    const float* vN = velTheta + (size_t)(j > 0 ? j - 1 : 0) * N;             // u_theta rows j-1, j, j+1
    const float* vM = velTheta + (size_t)j * N;
    const float* vS = velTheta + (size_t)(j + 1 < g.nTheta - 1 ? j + 1 : j) * N;
    const float nMask = j > 0 ? 1.0f : 0.0f;                    // v_N = 0 on the first row
    const float sMask = (j + 1 < g.nTheta - 1) ? 1.0f : 0.0f;   // v_S = 0 on the last row
    const float facA = __ldg(t.divFactor + j), facB = __ldg(t.divFactor + j + 1);
    const float snA = __ldg(t.sinNorth + j), ssA = __ldg(t.sinSouth + j);
    const float snB = __ldg(t.sinNorth + j + 1), ssB = __ldg(t.sinSouth + j + 1);
    float2 v[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float a0[8], a1[8], b0[8], b1[8], c0[8], c1[8], c2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = tt + ((h * 8 + e) << log2T);
            const int ie = (i + 1) & (N - 1);
            a0[e] = __ldg(uA + i); a1[e] = __ldg(uA + ie);
            b0[e] = __ldg(uB + i); b1[e] = __ldg(uB + ie);
            c0[e] = __ldg(vN + i); c1[e] = __ldg(vM + i); c2[e] = __ldg(vS + i);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float vn = c0[e] * nMask, vs = c2[e] * sMask;
            const float termA = __fmul_rn(facA, __fmaf_rn(c1[e], ssA, -__fmul_rn(vn, snA)));
            const float termB = __fmul_rn(facB, __fmaf_rn(vs, ssB, -__fmul_rn(c1[e], snB)));
            v[h * 8 + e] = make_float2(__fmaf_rn(facA, __fsub_rn(a1[e], a0[e]), termA),
                                       __fmaf_rn(facB, __fsub_rn(b1[e], b0[e]), termB));
        }
    }
    fftFromRegisters<-1, LOG2N>(v, buf, tt, T, N, log2N, tw, twReady);

    // v[m] = Z[tt + m*T]. Separate the two real rows and scale by 1/N (shiftFKernel,
    // kernel/KaminoCore.cu:652-653); only k = 1 .. N/2 is kept, the Nyquist mode in slot 0.
    if (!valid) return;
    const float scale = 0.5f / (float)N;
    // dense: [row][slot]; PACKED (theta bands, dist.cu): the send layout of the transpose, see SpectrumLayout
    const bool peers = PACKED && lay.peerTable != nullptr;
    float2* rowA = PACKED ? spectrum + (size_t)(j - lay.rowBase) * lay.rowPitch : spectrum + (size_t)j * half;
    float2* rowB = rowA + (PACKED ? lay.rowPitch : half);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int k = tt + (m << log2T);          // 0 .. N/2-1
        size_t at = PACKED ? (size_t)(k >> lay.log2Block) * lay.blockPitch + (k & ((1 << lay.log2Block) - 1)) : (size_t)k;
        if (peers) {
            // peer-memory transpose: block k >> log2Block goes straight to its owner, at [global row j][slot in block]
            rowA = lay.peerTable[k >> lay.log2Block] + (size_t)j * lay.rowPitch;
            rowB = rowA + lay.rowPitch;
            at = (size_t)(k & ((1 << lay.log2Block) - 1));
        }
        if (k == 0) {
            const float2 zn = v[8];                // Z[N/2] (tt = 0): both rows real
            rowA[at] = make_float2(2.0f * scale * zn.x, 0.0f);
            rowB[at] = make_float2(2.0f * scale * zn.y, 0.0f);
        } else {
            const float2 zk = v[m], zc = buf[fft::pad(N - k)];
            // A_k = (Z_k + conj Z_{N-k}) / 2,  B_k = (Z_k - conj Z_{N-k}) / (2i)
            rowA[at] = make_float2(scale * (zk.x + zc.x), scale * (zk.y - zc.y));
            rowB[at] = make_float2(scale * (zk.y + zc.y), scale * (zc.x - zk.x));
        }
    }
}

// ---- K6: inverse FFT + gradient subtraction ---------------------------------------------

// One transform per theta row with T = N/16 threads, BLOCK/T rows per block.
// grid (ceil(nTheta / (BLOCK/T)), batch), dynamic smem as for the forward kernel.
template <int BLOCK, bool STAGE, int LOG2N, bool PACKED>
__global__ void __launch_bounds__(BLOCK)
inverseFFTGradientKernel(GridParams g, SpectralTables t, const float2* __restrict__ spectrumAll,
                         float* __restrict__ velPhiAll, float* __restrict__ velThetaAll,
                         float* __restrict__ pressureAll, SpectrumLayout lay)
{
    extern __shared__ __align__(16) float2 smem[];
    __shared__ __align__(8) uint64_t twBar;
    const int log2N = LOG2N ? LOG2N : g.log2NPhi;
    const int N = 1 << log2N, half = N >> 1, nT = g.nTheta, log2T = log2N - 4, T = 1 << log2T;
    const int local = threadIdx.x >> log2T, tt = threadIdx.x & (T - 1);
    const int rowEnd = g.rowBegin + g.rowCount;
    const int rowRaw = g.rowBegin + blockIdx.x * (BLOCK >> log2T) + local;
    const bool valid = rowRaw < rowEnd;
    const int j = valid ? rowRaw : rowEnd - 1;
    float2* twShared = smem;
    float2* buf = smem + (STAGE ? N : 0) + (size_t)local * fft::paddedSize(N);
    const int sim = blockIdx.y;
    const float2* spectrum = spectrumAll + (size_t)sim * (g.cells >> 1);
    float* velPhi = velPhiAll + (size_t)sim * g.cells + (size_t)j * N;
    float* velTheta = velThetaAll + (size_t)sim * g.cells + (size_t)j * N;
    const bool hasSouth = (j < nT - 1);
    const float2* rowU = PACKED ? spectrum + (size_t)(j - lay.rowBase) * lay.rowPitch : spectrum + (size_t)j * half;
    const float2* rowS = hasSouth ? rowU + (PACKED ? lay.rowPitch : half) : rowU;     // !hasSouth: Y = 0

    uint64_t* twReady = stageTwiddles<STAGE>(t.twiddle, twShared, &twBar, fft::twiddleTableSize(N, log2N));
    const float2* tw = STAGE ? twShared : t.twiddle;
    pdlWait();                                   // the spectrum comes from the previous kernel

    // W_k = X_k + i Y_k with X = U_j, Y = U_{j+1} - U_j (Hermitian completions), X_0 = Y_0 = 0.
    // Thread tt needs W at idx = tt + e*T: for idx < N/2 from slot idx, for idx > N/2 from the
    // mirrored slot N - idx, for idx = N/2 from slot 0 (the Nyquist mode, real parts only).
    float2 v[16];
    // N = 4096 (256-thread blocks, 80 registers, three blocks per SM): all 32 loads of the two rows are issued before
    // any is used, and so are the 32 loads of the old velocity in the epilogue -- one exposed memory latency each
    // instead of two; together with the twiddles read through L1 instead of a 32 KB shared-memory copy per block
    // (r02m A/B at 2048 x 4096: 48.1 us against 55.2; either change alone is a loss: 57.3 / 53.3 us; the forward
    // kernel keeps its staged twiddles: 45.0 us without them against 37.9). The 1024-thread blocks of N = 16384 have no
    // registers to spare (972 vs 965 us), the 64-thread blocks of N <= 1024 are launch-latency-bound either way.
    constexpr bool kOneLatency = (BLOCK == 256);
    if constexpr (kOneLatency) {
        float2 x[16], sth[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int idx = tt + (e << log2T);
            const int slot = (e < 8) ? idx : ((N - idx) & (half - 1));     // idx = N/2 -> 0
            const size_t at = PACKED ? (size_t)(slot >> lay.log2Block) * lay.blockPitch + (slot & ((1 << lay.log2Block) - 1)) : (size_t)slot;
            x[e] = __ldg(rowU + at);
            sth[e] = __ldg(rowS + at);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int idx = tt + (e << log2T);
            const float2 y = make_float2(sth[e].x - x[e].x, sth[e].y - x[e].y);
            if (e < 8) v[e] = (idx == 0) ? make_float2(0.0f, 0.0f) : make_float2(x[e].x - y.y, x[e].y + y.x);
            else v[e] = (idx == half) ? make_float2(x[e].x, y.x) : make_float2(x[e].x + y.y, y.x - x[e].y);
        }
    } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float2 x[8], sth[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tt + ((h * 8 + e) << log2T);
            const int slot = (h == 0) ? idx : ((N - idx) & (half - 1));     // idx = N/2 -> 0
            const size_t at = PACKED ? (size_t)(slot >> lay.log2Block) * lay.blockPitch + (slot & ((1 << lay.log2Block) - 1)) : (size_t)slot;
            x[e] = __ldg(rowU + at);
            sth[e] = __ldg(rowS + at);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tt + ((h * 8 + e) << log2T);
            const float2 y = make_float2(sth[e].x - x[e].x, sth[e].y - x[e].y);
            if (h == 0) v[e] = (idx == 0) ? make_float2(0.0f, 0.0f) : make_float2(x[e].x - y.y, x[e].y + y.x);
            else v[8 + e] = (idx == half) ? make_float2(x[e].x, y.x) : make_float2(x[e].x + y.y, y.x - x[e].y);
        }
    }
    }
    fftFromRegisters<+1, LOG2N>(v, buf, tt, T, N, log2N, tw, twReady);

    // v[m] = z[i], i = tt + m*T: z.x = p[j][i], z.y = p[j+1][i] - p[j][i]
    if (!valid) return;
    // gradient subtraction (applyPressurePhi / applyPressureTheta, kernel/KaminoCore.cu:716-746);
    // the projection is compared with the reference to fp32 round-off, not bit for bit, so the
    // two divisions by row constants become multiplications by their reciprocals
    const float invDenomPhi = 1.0f / __ldg(t.gradPhiDenom + j);
    const float invNegH = -1.0f / g.h;
    float* pressure = pressureAll ? pressureAll + (size_t)sim * g.cells + (size_t)j * N : nullptr;
    if constexpr (kOneLatency) {
        float uOld[16], vOld[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int i = tt + (m << log2T);
            uOld[m] = velPhi[i];
            vOld[m] = hasSouth ? velTheta[i] : 0.0f;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int i = tt + (m << log2T);
            const float2 zi = v[m];
            const float pWest = buf[fft::pad((i - 1) & (N - 1))].x;
            velPhi[i] = __fmaf_rn(__fsub_rn(zi.x, pWest), invDenomPhi, uOld[m]);
            if (hasSouth) velTheta[i] = __fmaf_rn(zi.y, invNegH, vOld[m]);
            if (pressure) pressure[i] = zi.x;
        }
    } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float uOld[8], vOld[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = tt + ((h * 8 + m) << log2T);
            uOld[m] = velPhi[i];
            vOld[m] = hasSouth ? velTheta[i] : 0.0f;
        }
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int i = tt + ((h * 8 + m) << log2T);
            const float2 zi = v[h * 8 + m];
            const float pWest = buf[fft::pad((i - 1) & (N - 1))].x;
            velPhi[i] = __fmaf_rn(__fsub_rn(zi.x, pWest), invDenomPhi, uOld[m]);
            if (hasSouth) velTheta[i] = __fmaf_rn(zi.y, invNegH, vOld[m]);
            if (pressure) pressure[i] = zi.x;
        }
    }
    }
}

} // namespace

size_t spectralTableBytes(const GridParams& g)
{
    return sizeof(float2) * g.nPhi + sizeof(float) * 10 * g.nTheta + sizeof(float) * solveTableFloats(g) + 256 * 18;
}

cudaError_t launchBuildTables(const GridParams& g, SpectralTables t, cudaStream_t stream)
{
    const int threads = 256;
    const int blocks = (g.nPhi + threads - 1) / threads;
    buildTablesKernel<<<blocks, threads, 0, stream>>>(g, t);
    return cudaGetLastError();
}

namespace {

// FFT launch geometry: T = N/16 threads per transform, blocks of max(T, 64) threads; the
// twiddle table is staged in shared memory up to N = 4096 (32 KB)
struct FftLaunch { int block; int perBlock; bool stage; size_t smem; };
FftLaunch fftLaunch(const GridParams& g)
{
    const int T = g.nPhi >> 4;
    FftLaunch l;
    l.block = T < 64 ? 64 : T;
    l.perBlock = l.block / T;
    l.stage = g.nPhi <= 4096 && g.nPhi >= 64;
    l.smem = ((size_t)l.perBlock * fft::paddedSize(g.nPhi) + (l.stage ? g.nPhi : 0)) * sizeof(float2);
    return l;
}

template <int BLOCK, bool STAGE, int LOG2N, bool PACKED>
cudaError_t fftDispatchLayout(int which, const GridParams& g, const SpectralTables& t, const FftLaunch& l,
                              const float* velPhiIn, const float* velThetaIn, float2* spectrum,
                              float* velPhi, float* velTheta, float* pressure, int batch, cudaStream_t stream, const SpectrumLayout& lay)
{
    // the inverse kernel of N = 4096 reads its twiddles through L1 (see the kernel)
    constexpr bool kStageInverse = STAGE && BLOCK != 256;
    const size_t smemInverse = l.smem - ((STAGE && !kStageInverse) ? sizeof(float2) * (size_t)g.nPhi : 0);
    if (which == 0) {           // configure
        cudaError_t e = cudaFuncSetAttribute(divergenceFFTKernel<BLOCK, STAGE, LOG2N, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(inverseFFTGradientKernel<BLOCK, kStageInverse, LOG2N, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemInverse);
    }
    if (which == 1) {
        const int pairs = g.rowCount / 2;
        dim3 grid((pairs + l.perBlock - 1) / l.perBlock, batch);
        return launchChained(divergenceFFTKernel<BLOCK, STAGE, LOG2N, PACKED>, grid, dim3(BLOCK), l.smem, stream, g, t, velPhiIn, velThetaIn, spectrum, lay);
    } else {
        dim3 grid((g.rowCount + l.perBlock - 1) / l.perBlock, batch);
        return launchChained(inverseFFTGradientKernel<BLOCK, kStageInverse, LOG2N, PACKED>, grid, dim3(BLOCK), smemInverse, stream, g, t,
                             (const float2*)spectrum, velPhi, velTheta, pressure, lay);
    }
}

template <int BLOCK, bool STAGE, int LOG2N>
cudaError_t fftDispatch(int which, const GridParams& g, const SpectralTables& t, const FftLaunch& l,
                        const float* velPhiIn, const float* velThetaIn, float2* spectrum,
                        float* velPhi, float* velTheta, float* pressure, int batch, cudaStream_t stream, const SpectrumLayout* lay)
{
    if (which == 0) {           // configure both addressing variants
        cudaError_t e = fftDispatchLayout<BLOCK, STAGE, LOG2N, false>(0, g, t, l, velPhiIn, velThetaIn, spectrum, velPhi, velTheta, pressure, batch, stream, SpectrumLayout{});
        if (e != cudaSuccess) return e;
        return fftDispatchLayout<BLOCK, STAGE, LOG2N, true>(0, g, t, l, velPhiIn, velThetaIn, spectrum, velPhi, velTheta, pressure, batch, stream, SpectrumLayout{});
    }
    if (lay) return fftDispatchLayout<BLOCK, STAGE, LOG2N, true>(which, g, t, l, velPhiIn, velThetaIn, spectrum, velPhi, velTheta, pressure, batch, stream, *lay);
    return fftDispatchLayout<BLOCK, STAGE, LOG2N, false>(which, g, t, l, velPhiIn, velThetaIn, spectrum, velPhi, velTheta, pressure, batch, stream, SpectrumLayout{});
}

cudaError_t fftSelect(int which, const GridParams& g, const SpectralTables& t, const float* velPhiIn,
                      const float* velThetaIn, float2* spectrum, float* velPhi, float* velTheta,
                      float* pressure, int batch, cudaStream_t stream, const SpectrumLayout* lay = nullptr)
{
    const FftLaunch l = fftLaunch(g);
#define KB_FFT(B, S, L2N) return fftDispatch<B, S, L2N>(which, g, t, l, velPhiIn, velThetaIn, spectrum, velPhi, velTheta, pressure, batch, stream, lay)
    // blocks of 128 threads and more hold one transform, so the block size fixes N; 64-thread blocks serve every
    // N <= 1024, of which 512 (C4) and 1024 (C2) get compile-time builds
    switch (l.block) {
    case 64:
        if (g.log2NPhi == 10) KB_FFT(64, true, 10);
        if (g.log2NPhi == 9) KB_FFT(64, true, 9);
        if (l.stage) KB_FFT(64, true, 0); else KB_FFT(64, false, 0);
    case 128: KB_FFT(128, true, 11);
    // (r02a A/B: capping the 256-thread kernels at 80 registers for three blocks per SM is a loss at
    // 2048 x 4096: forward 54.9 vs 47.5 us, inverse 64.3 vs 59.1 us)
    case 256: KB_FFT(256, true, 12);
    case 512: KB_FFT(512, false, 13);
    case 1024: KB_FFT(1024, false, 14);
    default: return cudaErrorInvalidValue;
    }
#undef KB_FFT
}

} // namespace

cudaError_t configureKernels(const GridParams& g, int batch)
{
    SpectralTables none{};
    cudaError_t e = fftSelect(0, g, none, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1, nullptr);
    if (e != cudaSuccess) return e;
    return configureTridiagonal(g, batch);
}

cudaError_t launchDivergenceFFT(const GridParams& g, const SpectralTables& t, const float* velPhi,
                                const float* velTheta, float2* spectrum, int batch, cudaStream_t stream,
                                const SpectrumLayout* packed)
{
    return fftSelect(1, g, t, velPhi, velTheta, spectrum, nullptr, nullptr, nullptr, batch, stream, packed);
}

cudaError_t launchInverseFFTGradient(const GridParams& g, const SpectralTables& t, const float2* spectrum,
                                     float* velPhi, float* velTheta, float* pressure, int batch,
                                     cudaStream_t stream, const SpectrumLayout* packed)
{
    return fftSelect(2, g, t, nullptr, nullptr, const_cast<float2*>(spectrum), velPhi, velTheta, pressure, batch, stream, packed);
}

} // namespace kb
