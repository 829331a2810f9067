// Per-wavenumber tridiagonal solves along theta (K5 of the projection).
//
// Replaces precomputeABCKernel (kernel/KaminoSolver.cu:117-163), the two crKernel launches
// per step (kernel/tdm.cu:3-96, kernel/KaminoCore.cu:779-792) and the two transposes around
// them (shiftFKernel's transposed store, copy2UFourier: kernel/KaminoCore.cu:640-670).
//
// The reference runs cyclic reduction (CR) on (a, b, c, d) every step, once for the real and
// once for the imaginary right-hand side. CR in fp32 is the least accurate part of its step
// (at 512 x 1024 its pressure is 1.7e-5 away, in relative L2, from an fp64 solve of the same
// fp32 coefficients; 1.3e-4 at 2048 x 4096), and its log2(nTheta) barrier-
// separated levels with halving parallelism are slow. This file solves the same systems --
// same a, b, c, bit for bit the reference's fp32 coefficients -- with an LU (Thomas)
// factorisation computed ONCE at context creation in fp64 and applied per step as two
// first-order linear recurrences,
//     forward :  y_i = d_i - l_i y_{i-1}
//     backward:  x_i = y_i / b'_i - h_i x_{i+1},          l_i = a_i / b'_{i-1},  h_i = c_i / b'_i,
// which are parallelised over P chunks of L rows with precomputed multiplier products
// (beta_i = prod -l, delta_i = prod -h inside a chunk): each chunk runs its recurrence from a
// zero carry, one thread per system chains the P chunk carries, and a final pass adds
// carry x product. Dependent chain: 2 L + 2 P steps instead of 2 nTheta (Thomas) or
// 2 log2(nTheta) block-wide barriers with strided shared-memory traffic (CR). In fp32 the
// result is 4.7e-7 from the fp64 solve at 512 x 1024 (36 x closer than the reference's own
// CR), so the pressure differs from the reference's by the reference's rounding error; the
// parity tests bound it through the fp64 operator.
//
// Tried and rejected (r01n A/B): dropping the l and h tables and rebuilding the multipliers in the
// kernel from per-row off-diagonals and the 1/b' table (8 of the 20 table bytes per cell less HBM
// traffic) -- 53.0 us instead of 50.8 us at 2048 x 4096: with one 512-thread block per SM the kernel
// is bound by the latency of its load phases, not by bandwidth, and the extra dependent loads cost more.
//
// Layout: the half spectrum is solved in place in its [theta][slot] layout (no transposes).
// A block owns W consecutive wavenumber slots and P * W threads (thread = chunk p, system w).
// Tables are [slot group][row][w] so that every access of a warp is a run of W floats.
#include "kamino_kernels.cuh"

namespace kb {

namespace {

// ---- launch geometry ----------------------------------------------------------------------------

struct TriLaunch { int W; int L; int P; size_t smem; };

TriLaunch triLaunch(const GridParams& g, int batch)
{
    const int nT = g.nTheta, half = g.nPhi / 2;
    TriLaunch l;
    l.L = nT <= 64 ? 4 : nT <= 256 ? 8 : nT <= 1024 ? 16 : nT <= 4096 ? 32 : 64;
    // (r02a A/B of the chunk length: 512 rows 10.4 / 10.0 / 13.4 us for L = 8 / 16 / 32; 2048 rows 63.3 / 51.0 /
    // 75.4 us for L = 16 / 32 / 64)
    l.P = nT / l.L;
    // W: as wide as possible (64-byte runs) while the grid still has about one block per SM and
    // the right-hand sides (nTheta * W float2) fit in shared memory (r01o, r02a A/B: W = 4 at 512 x 1024
    // (W = 2: 11.4 vs 10.0 us), W = 8 at 2048 x 4096 (W = 4: 63.7 vs 51.0 us))
    l.W = 8;
    while (l.W > 2 && ((long)(half / l.W) * batch < 128 || (size_t)nT * l.W * sizeof(float2) > 160 * 1024)) l.W >>= 1;
    while (l.P * l.W > (l.L >= 64 ? 512 : 1024)) l.W >>= 1;
    l.smem = ((size_t)l.P * (l.L * l.W + l.W) + 2 * (size_t)(l.P + 1) * l.W) * sizeof(float2)
           + 2 * (size_t)l.P * l.W * sizeof(float);
    return l;
}

// ---- setup: LU factors and chunk products, one thread per wavenumber slot, fp64 -------------------

// (slotBegin, slotCount): the slots this table set covers -- all of them, or the wavenumber band of one rank of
// a theta-band run (dist.cu), whose tables are indexed by slot - slotBegin.
__global__ void buildSolveTablesKernel(GridParams g, SpectralTables t, int W, int L, int slotBegin, int slotCount)
{
    const int nT = g.nTheta, half = g.nPhi >> 1;
    const int local = blockIdx.x * blockDim.x + threadIdx.x;
    if (local >= slotCount) return;
    const int wave = slotBegin + local;
    const int n = (wave == 0) ? half : wave;        // wavenumber of this slot (never 0)
    const int slot = local;                         // table index
    const size_t base = (size_t)(slot / W) * nT * W + (slot % W);
    const float nSq = (float)(n * n);
    const int P = nT / L;

    // forward: l_i, 1/b'_i, h_i and beta_i = prod_{m = chunk start .. i} (-l_m) (from the ROUNDED l)
    double bPrev = 1.0, cPrev = 0.0, beta = 1.0;
    for (int i = 0; i < nT; ++i) {
        // coefficients exactly as precomputeABCKernel builds them, kernel/KaminoSolver.cu:128-153
        float a = t.triA[i], c = t.triC[i];
        float b = (float)(t.minusTwoOverH2 - (double)__fdiv_rn(nSq, t.sinSq[i]));
        if (i == 0) { b = __fadd_rn(b, a); a = 0.0f; }
        if (i == nT - 1) { b = __fadd_rn(b, c); c = 0.0f; }
        const double l = (i == 0) ? 0.0 : (double)a / bPrev;
        const double bp = (double)b - l * cPrev;
        const float l32 = (float)l;
        const double invb = 1.0 / bp;
        if (i % L == 0) beta = 1.0;
        beta *= -(double)l32;
        const size_t e = base + (size_t)i * W;
        t.thL[e] = l32;
        t.thInvB[e] = (float)invb;
        t.thBetaInv[e] = (float)(beta * invb);
        t.thH[e] = (float)((double)c * invb);
        if (i % L == L - 1) t.thBetaEnd[((size_t)(slot / W) * P + i / L) * W + (slot % W)] = (float)beta;
        bPrev = bp; cPrev = (double)c;
    }
    // backward: delta_i = prod_{m = i .. chunk end} (-h_m) (from the ROUNDED h)
    double delta = 1.0;
    for (int i = nT - 1; i >= 0; --i) {
        if (i % L == L - 1) delta = 1.0;
        const size_t e = base + (size_t)i * W;
        delta *= -(double)t.thH[e];
        t.thDelta[e] = (float)delta;
    }
}

// The same factorisation for ONE theta band of a band-decomposed run (reduced-interface / SPIKE solve,
// banded.py): rows [rowBegin, rowBegin + rows) of every wavenumber's system, cut loose from the
// neighbouring bands (the sub-diagonal of the first band row and the super-diagonal of the last one
// are left out; the driver re-introduces them through the spike vectors). Global rows 0 and
// nTheta - 1 keep the reference's Neumann fold. Tables are laid out as for a grid of `rows` rows.
__global__ void buildBandSolveTablesKernel(GridParams g, SpectralTables t, SpectralTables band, int rowBegin, int rows, int W, int L)
{
    const int nT = g.nTheta, half = g.nPhi >> 1;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= half) return;
    const int n = (slot == 0) ? half : slot;
    const size_t base = (size_t)(slot / W) * rows * W + (slot % W);
    const float nSq = (float)(n * n);
    const int P = rows / L;
    double bPrev = 1.0, cPrev = 0.0, beta = 1.0;
    for (int i = 0; i < rows; ++i) {
        const int gi = rowBegin + i;
        float a = t.triA[gi], c = t.triC[gi];
        float b = (float)(t.minusTwoOverH2 - (double)__fdiv_rn(nSq, t.sinSq[gi]));
        if (gi == 0) { b = __fadd_rn(b, a); a = 0.0f; }
        if (gi == nT - 1) { b = __fadd_rn(b, c); c = 0.0f; }
        if (i == 0) a = 0.0f;                       // coupling to the band above: handled by the spikes
        if (i == rows - 1) c = 0.0f;                // coupling to the band below
        const double l = (i == 0) ? 0.0 : (double)a / bPrev;
        const double bp = (double)b - l * cPrev;
        const float l32 = (float)l;
        const double invb = 1.0 / bp;
        if (i % L == 0) beta = 1.0;
        beta *= -(double)l32;
        const size_t e = base + (size_t)i * W;
        band.thL[e] = l32;
        band.thInvB[e] = (float)invb;
        band.thBetaInv[e] = (float)(beta * invb);
        band.thH[e] = (float)((double)c * invb);
        if (i % L == L - 1) band.thBetaEnd[((size_t)(slot / W) * P + i / L) * W + (slot % W)] = (float)beta;
        bPrev = bp; cPrev = (double)c;
    }
    double delta = 1.0;
    for (int i = rows - 1; i >= 0; --i) {
        if (i % L == L - 1) delta = 1.0;
        const size_t e = base + (size_t)i * W;
        delta *= -(double)band.thH[e];
        band.thDelta[e] = (float)delta;
    }
}

// ---- per step -------------------------------------------------------------------------------------
// grid ((N/2) / W, batch), P * W threads, dynamic smem (P (L W + W) + 2 (P + 1) W) float2 + 2 P W floats.
// Shared-memory rows of the right-hand sides are padded by one W-run per chunk (pitch L*W + W):
// the 64-bit accesses of a half-warp (16 / W chunks x W systems) then fall in distinct banks.
// W, L and P are compile-time so that every loop is fully unrolled: all global loads of a phase
// are in flight together and the only dependent chains are the FMA recurrences themselves.
// SCATTER (theta-band runs with peer-memory transposes, dist.cu): the solution is not written back in place but straight
// into the owning ranks' buffers. A compile-time variant: the in-place kernel keeps exactly the code (and SASS) it had
// before -- a run-time branch around the store made the pointer-table stores alias-block the table loads of the unrolled
// loop (r02r ncu at 2048 x 4096: 68 instead of 48 us).
template <int W, int L, int P, bool SCATTER>
__global__ void __launch_bounds__(P * W)
tridiagonalKernel(GridParams g, SpectralTables t, float2* __restrict__ spectrumAll, int pitch, int groupOffset, PeerScatter scatter)
{
    extern __shared__ __align__(16) float2 sm[];
    constexpr int nT = P * L;
    constexpr int kChunkPitch = L * W + W;              // float2 elements per chunk in smem
    constexpr int kThreads = P * W;
    const int half = pitch;                             // row pitch of the spectrum buffer (float2)
    const int group = blockIdx.x + groupOffset;         // slot group: selects the table slice
    float2* d = sm;                                     // P chunks of kChunkPitch
    float2* carryY = sm + P * kChunkPitch;              // (P + 1) x W : y at the end of chunk p-1
    float2* carryX = carryY + (P + 1) * W;              // (P + 1) x W : x at the start of chunk p
    float* betaEnd = reinterpret_cast<float*>(carryX + (P + 1) * W);   // P x W
    float* deltaStart = betaEnd + P * W;                               // P x W
    const int tid = threadIdx.x;
    const int p = tid / W, w = tid % W;
    float2* spectrum = spectrumAll + (size_t)blockIdx.y * (g.cells >> 1) + (size_t)blockIdx.x * W;
    const size_t tabBase = (size_t)group * nT * W;
    const float* tabL = t.thL + tabBase;
    const float* tabInvB = t.thInvB + tabBase;
    const float* tabBetaInv = t.thBetaInv + tabBase;
    const float* tabH = t.thH + tabBase;
    const float* tabDelta = t.thDelta + tabBase;

    // tables needed first (independent of the previous kernel): my chunk's forward multipliers
    // and the carry multipliers of the whole block
    // (multipliers are register-prefetched ahead of the barriers for chunks of up to 16 rows;
    // longer chunks read them inside the recurrences, 8 rows ahead, to stay within the register file)
    constexpr bool kPrefetch = (L <= 16);
    constexpr int kPre = kPrefetch ? L : 1;
    constexpr int kBatch = L < 16 ? L : 16;             // items per thread per load / store batch
    const int row0 = p * L;
    float lReg[kPre];
    if (kPrefetch) {
#pragma unroll
        for (int ii = 0; ii < L; ++ii) lReg[ii] = __ldg(tabL + (row0 + ii) * W + w);
    }
    betaEnd[tid] = __ldg(t.thBetaEnd + (size_t)group * P * W + tid);
    deltaStart[tid] = __ldg(tabDelta + row0 * W + w);

    pdlWait();                                   // the right-hand sides come from the previous kernel
    // load: item = (row i, system w), runs of W float2 per row; L items per thread
#pragma unroll 1
    for (int k0 = 0; k0 < L; k0 += kBatch) {
        float2 v[kBatch];
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int item = tid + (k0 + k) * kThreads;
            v[k] = spectrum[(size_t)(item / W) * half + (item % W)];
        }
#pragma unroll
        for (int k = 0; k < kBatch; ++k) {
            const int item = tid + (k0 + k) * kThreads;
            const int i = item / W, ww = item % W;
            d[(i / L) * kChunkPitch + (i % L) * W + ww] = v[k];
        }
    }
    __syncthreads();

    float2* mine = d + p * kChunkPitch + w;
    // forward recurrence inside the chunk from a zero carry: alpha_i = d_i - l_i alpha_{i-1}
    {
        float2 prev = make_float2(0.0f, 0.0f);
#pragma unroll(kPrefetch ? L : 8)
        for (int ii = 0; ii < L; ++ii) {
            const float l = kPrefetch ? lReg[ii] : __ldg(tabL + (row0 + ii) * W + w);
            float2 v = mine[ii * W];
            v.x = __fmaf_rn(-l, prev.x, v.x);
            v.y = __fmaf_rn(-l, prev.y, v.y);
            mine[ii * W] = v;
            prev = v;
        }
    }
    // my chunk's backward multipliers: issued now, consumed after the carry chain
    float invB[kPre], betaInv[kPre], hReg[kPre];
    if (kPrefetch) {
#pragma unroll
        for (int ii = 0; ii < L; ++ii) {
            const int e = (row0 + ii) * W + w;
            invB[ii] = __ldg(tabInvB + e);
            betaInv[ii] = __ldg(tabBetaInv + e);
            hReg[ii] = __ldg(tabH + e);
        }
    }
    __syncthreads();
    // chain the chunk carries: Y_p = alpha_end(p) + beta_end(p) Y_{p-1}; carryY[p] = Y_{p-1}
    if (tid < W) {
        float2 y = make_float2(0.0f, 0.0f);
        carryY[tid] = y;
#pragma unroll 8
        for (int q = 0; q < P; ++q) {
            const float2 a = d[q * kChunkPitch + (L - 1) * W + tid];
            const float be = betaEnd[q * W + tid];
            y.x = __fmaf_rn(be, y.x, a.x);
            y.y = __fmaf_rn(be, y.y, a.y);
            carryY[(q + 1) * W + tid] = y;
        }
    }
    __syncthreads();
    // backward recurrence inside the chunk from a zero carry:
    //   t_i = (alpha_i + beta_i Y_{p-1}) / b'_i ;  gamma_i = t_i - h_i gamma_{i+1}
    {
        const float2 yPrev = carryY[p * W + w];
        float2 next = make_float2(0.0f, 0.0f);
#pragma unroll(kPrefetch ? L : 8)
        for (int ii = L - 1; ii >= 0; --ii) {
            const int e = (row0 + ii) * W + w;
            const float ib = kPrefetch ? invB[ii] : __ldg(tabInvB + e);
            const float bi = kPrefetch ? betaInv[ii] : __ldg(tabBetaInv + e);
            const float hh = kPrefetch ? hReg[ii] : __ldg(tabH + e);
            float2 v = mine[ii * W];
            v.x = __fmaf_rn(bi, yPrev.x, __fmul_rn(v.x, ib));
            v.y = __fmaf_rn(bi, yPrev.y, __fmul_rn(v.y, ib));
            v.x = __fmaf_rn(-hh, next.x, v.x);
            v.y = __fmaf_rn(-hh, next.y, v.y);
            mine[ii * W] = v;
            next = v;
        }
    }
    // delta of my output items (load mapping): issued now, consumed after the carry chain
    float de[kPre];
    if (kPrefetch) {
#pragma unroll
        for (int k = 0; k < L; ++k) de[k] = __ldg(tabDelta + tid + k * kThreads);
    }
    __syncthreads();
    // chain the carries upwards: X_p = gamma_start(p) + delta_start(p) X_{p+1}; carryX[p] = X_p
    if (tid < W) {
        float2 x = make_float2(0.0f, 0.0f);
        carryX[P * W + tid] = x;
#pragma unroll 8
        for (int q = P - 1; q >= 0; --q) {
            const float2 gm = d[q * kChunkPitch + tid];
            const float ds = deltaStart[q * W + tid];
            x.x = __fmaf_rn(ds, x.x, gm.x);
            x.y = __fmaf_rn(ds, x.y, gm.y);
            carryX[q * W + tid] = x;
        }
    }
    __syncthreads();
    // x_i = gamma_i + delta_i X_{p+1}, written back in the load mapping (coalesced)
    if constexpr (!SCATTER) {
#pragma unroll(kPrefetch ? L : 8)
        for (int k = 0; k < L; ++k) {
            const int item = tid + k * kThreads;
            const int i = item / W, ww = item % W;
            const int q = i / L;
            const float dl = kPrefetch ? de[k] : __ldg(tabDelta + item);
            const float2 gm = d[q * kChunkPitch + (i % L) * W + ww];
            const float2 xn = carryX[(q + 1) * W + ww];
            spectrum[(size_t)i * half + ww] = make_float2(__fmaf_rn(dl, xn.x, gm.x), __fmaf_rn(dl, xn.y, gm.y));
        }
    } else {
        // peer-memory transpose: row i goes straight to its owner, in the layout the owner's inverse FFT reads
        // ([source rank][rows + 1][K]); the first row of a band is also the extra row of the band above it. The values
        // of a batch are computed first (all table loads in flight together), then stored: the stores go through
        // pointers the compiler cannot prove distinct from the tables.
        const int rows = 1 << scatter.log2Rows;
        const size_t blockPitch = (size_t)(rows + 1) * scatter.kper;
        constexpr int kOut = L < 8 ? L : 8;
#pragma unroll 1
        for (int k0 = 0; k0 < L; k0 += kOut) {
            float2 x[kOut];
#pragma unroll
            for (int u = 0; u < kOut; ++u) {
                const int item = tid + (k0 + u) * kThreads;
                const int i = item / W, ww = item % W;
                const int q = i / L;
                const float dl = __ldg(tabDelta + item);
                const float2 gm = d[q * kChunkPitch + (i % L) * W + ww];
                const float2 xn = carryX[(q + 1) * W + ww];
                x[u] = make_float2(__fmaf_rn(dl, xn.x, gm.x), __fmaf_rn(dl, xn.y, gm.y));
            }
#pragma unroll
            for (int u = 0; u < kOut; ++u) {
                const int item = tid + (k0 + u) * kThreads;
                const int i = item / W, ww = item % W;
                const int dest = i >> scatter.log2Rows, local = i & (rows - 1);
                const size_t at = (size_t)scatter.myRank * blockPitch + (size_t)blockIdx.x * W + ww;
                scatter.table[dest][at + (size_t)local * scatter.kper] = x[u];
                if (local == 0 && dest > 0) scatter.table[dest - 1][at + (size_t)rows * scatter.kper] = x[u];
            }
        }
    }
}

template <int W, int L, int P>
cudaError_t launchTri(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch, const TriLaunch& l,
                      cudaStream_t stream, bool configureOnly, int pitch, int slotBegin, int slotCount, const PeerScatter& scatter)
{
    if (configureOnly) {
        cudaError_t e = cudaFuncSetAttribute(tridiagonalKernel<W, L, P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(tridiagonalKernel<W, L, P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem);
    }
    if (slotBegin % W != 0 || slotCount % W != 0) return cudaErrorInvalidValue;
    dim3 grid(slotCount / W, batch);
    if (scatter.table)
        return launchChained(tridiagonalKernel<W, L, P, true>, grid, dim3(P * W), l.smem, stream, g, t, spectrum, pitch, slotBegin / W, scatter);
    return launchChained(tridiagonalKernel<W, L, P, false>, grid, dim3(P * W), l.smem, stream, g, t, spectrum, pitch, slotBegin / W, scatter);
}

cudaError_t dispatchTri(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch,
                        cudaStream_t stream, bool configureOnly, int tableBatch, int pitch, int slotBegin, int slotCount,
                        const PeerScatter& scatter = PeerScatter{})
{
    const TriLaunch l = triLaunch(g, tableBatch);      // the tables were laid out for this W at creation
#define KB_TRI(WW, LL, PP) if (l.W == WW && l.L == LL && l.P == PP) return launchTri<WW, LL, PP>(g, t, spectrum, batch, l, stream, configureOnly, pitch, slotBegin, slotCount, scatter)
#define KB_TRI_W(LL, PP) KB_TRI(2, LL, PP); KB_TRI(4, LL, PP); KB_TRI(8, LL, PP)
    KB_TRI_W(4, 4); KB_TRI_W(4, 8); KB_TRI_W(4, 16);          // nTheta = 16, 32, 64
    KB_TRI_W(8, 16); KB_TRI_W(8, 32);                         // 128, 256
    KB_TRI_W(16, 32); KB_TRI_W(16, 64);                       // 512, 1024
    KB_TRI_W(32, 64);                                         // 2048
    KB_TRI(2, 32, 128); KB_TRI(4, 32, 128); KB_TRI(8, 32, 128);   // 4096
    KB_TRI(2, 64, 128); KB_TRI(4, 64, 128);                   // 8192
#undef KB_TRI_W
#undef KB_TRI
    return cudaErrorInvalidValue;
}

} // namespace

size_t solveTableFloats(const GridParams& g)
{
    // l, 1/b', beta/b', h, delta: 5 x nTheta x N/2 floats, + beta at the chunk ends (<= nTheta/4 x N/2)
    return (size_t)g.nTheta * (g.nPhi / 2) * 5 + (size_t)(g.nTheta / 4) * (g.nPhi / 2);
}

cudaError_t configureTridiagonal(const GridParams& g, int batch)
{
    SpectralTables none{};
    return dispatchTri(g, none, nullptr, batch, nullptr, true, batch, g.nPhi / 2, 0, g.nPhi / 2);
}

cudaError_t launchBuildSolveTables(const GridParams& g, SpectralTables t, int batch, cudaStream_t stream,
                                   int slotBegin, int slotCount)
{
    const TriLaunch l = triLaunch(g, batch);
    if (slotCount < 0) slotCount = g.nPhi / 2;
    if (slotBegin % l.W || slotCount % l.W) return cudaErrorInvalidValue;
    const int threads = 64;
    buildSolveTablesKernel<<<(slotCount + threads - 1) / threads, threads, 0, stream>>>(g, t, l.W, l.L, slotBegin, slotCount);
    return cudaGetLastError();
}

cudaError_t launchTridiagonal(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch,
                              cudaStream_t stream)
{
    return dispatchTri(g, t, spectrum, batch, stream, false, batch, g.nPhi / 2, 0, g.nPhi / 2);
}

cudaError_t launchTridiagonalBand(const GridParams& g, const SpectralTables& t, float2* packed, int pitch,
                                  int slotBegin, int slotCount, int tableBatch, cudaStream_t stream, const PeerScatter* scatter)
{
    return dispatchTri(g, t, packed, 1, stream, false, tableBatch, pitch, slotBegin, slotCount, scatter ? *scatter : PeerScatter{});
}

// ---- band-local solve (reduced-interface / SPIKE mode of the band-decomposed run) -------------------

namespace {
GridParams bandSolveParams(const GridParams& g, int rows)
{
    GridParams b = g;
    b.nTheta = rows;        // selects L, P and the kernel instantiation; nPhi (the slot count) is unchanged
    return b;
}
} // namespace

size_t bandSolveTableFloats(const GridParams& g, int rows)
{
    return (size_t)rows * (g.nPhi / 2) * 5 + (size_t)(rows / 4 + 1) * (g.nPhi / 2);
}

cudaError_t launchBuildBandSolveTables(const GridParams& g, const SpectralTables& t, const SpectralTables& band,
                                       int rowBegin, int rows, cudaStream_t stream)
{
    const GridParams b = bandSolveParams(g, rows);
    cudaError_t e = dispatchTri(b, band, nullptr, 1, nullptr, true, 1, g.nPhi / 2, 0, g.nPhi / 2);    // opt-in smem
    if (e != cudaSuccess) return e;
    const TriLaunch l = triLaunch(b, 1);
    const int half = g.nPhi / 2, threads = 64;
    buildBandSolveTablesKernel<<<(half + threads - 1) / threads, threads, 0, stream>>>(g, t, band, rowBegin, rows, l.W, l.L);
    return cudaGetLastError();
}

// in place on spectrum rows [rowBegin, rowBegin + rows), all wavenumber slots
cudaError_t launchBandLocalSolve(const GridParams& g, const SpectralTables& band, float2* spectrum, int rowBegin, int rows,
                                 cudaStream_t stream)
{
    const GridParams b = bandSolveParams(g, rows);
    const int half = g.nPhi / 2;
    return dispatchTri(b, band, spectrum + (size_t)rowBegin * half, 1, stream, false, 1, half, 0, half);
}

} // namespace kb
