// Per-wavenumber tridiagonal solves along theta (K5 of the projection).
//
// Replaces precomputeABCKernel (kernel/KaminoSolver.cu:117-163), the two crKernel launches
// per step (kernel/tdm.cu:3-96, kernel/KaminoCore.cu:779-792) and the two transposes around
// them (shiftFKernel's transposed store, copy2UFourier: kernel/KaminoCore.cu:640-670).
//
// The reference runs cyclic reduction (CR) on (a, b, c, d) every step, once for the real
// and once for the imaginary right-hand side. The (a, b, c) part of that recursion does not
// depend on the right-hand side, so it is executed ONCE, at context creation, by
// buildCrTablesKernel -- in the reference's elimination order and with its fp32 operations,
// so every factor has the bits crKernel would produce -- and stored:
//   crFwd[e]   = (tmp1, tmp2)   elimination factors of forward level l, element idx,
//                               e = nTheta - (nTheta >> l) + idx          (tdm.cu:55-56)
//   crA/B/C[e]                  a, b, c of the rows as the back substitution sees them, in
//                               the order it visits them (contiguous per level)
// The per-step kernel then only streams the right-hand sides: 2 FMAs per element and level
// going down, one FMA chain and one IEEE division per element coming up, real and imaginary
// parts together, reading and writing the half spectrum in place in its [theta][slot]
// layout (no transposes). A block owns W consecutive wavenumber slots; the tables are laid
// out per block, [slot group][row][w], so that the slice a block needs is one contiguous
// range: for grids up to nTheta = 1024 it is brought into shared memory by four TMA bulk
// copies issued by one thread while the others load the right-hand sides (one exposed
// global-memory latency per block instead of one per CR level); for larger grids the levels
// read it directly with fully coalesced loads.
//
// Shared memory holds the right-hand sides of the W systems as float2 d[sw(i) * W + w]; sw()
// XOR-swizzles the low bits of the row index so that the power-of-two strides of CR are
// bank-conflict-free without padding.
#include "kamino_kernels.cuh"
#include "tma_bulk.cuh"

namespace kb {

namespace {

// log2(16 / W) low bits of i are XORed with the fold of all higher bit groups
template <int W>
__device__ __forceinline__ int swizzleRow(int i)
{
    constexpr int B = (W == 2) ? 3 : (W == 4) ? 2 : 1;
    const int x = i >> B;
    int f = 0;
#pragma unroll
    for (int s = 0; s < 14; s += B) f ^= (x >> s);
    return i ^ (f & ((1 << B) - 1));
}

// ---- setup: CR on the coefficients only, one block per wavenumber slot ---------------------
// dynamic smem: 3 * nTheta floats
__global__ void buildCrTablesKernel(GridParams g, SpectralTables t, int W)
{
    extern __shared__ float sm[];
    const int nT = g.nTheta, half = g.nPhi >> 1;
    float* a = sm;
    float* b = a + nT;
    float* c = b + nT;
    const int slot = blockIdx.x;
    const int n = (slot == 0) ? half : slot;        // wavenumber of this slot (never 0)
    // table element (row r) of this slot: [(slot / W) * nT + r] * W + slot % W
    const size_t tabBase = (size_t)(slot / W) * nT * W + (slot % W);
    const float nSq = (float)(n * n);

    // precomputeABCKernel, kernel/KaminoSolver.cu:128-153
    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
        float valA = t.triA[i], valC = t.triC[i];
        float valB = (float)(t.minusTwoOverH2 - (double)__fdiv_rn(nSq, t.sinSq[i]));
        if (i == 0) { valB = __fadd_rn(valB, valA); valA = 0.0f; }
        if (i == nT - 1) { valB = __fadd_rn(valB, valC); valC = 0.0f; }
        a[i] = valA; b[i] = valB; c[i] = valC;
    }
    // forward elimination of the coefficients, kernel/tdm.cu:43-63
    int levels = 0;
    while ((2 << levels) < nT) ++levels;            // log2(nT / 2)
    int stride = 1;
    for (int lvl = 0; lvl < levels; ++lvl) {
        __syncthreads();
        stride <<= 1;
        const int delta = stride >> 1;
        const int count = nT >> (lvl + 1);
        const int entry0 = nT - (nT >> lvl);
        // two-phase (read, barrier, write) so that a block smaller than `count` stays correct
        for (int base = 0; base < count; base += blockDim.x) {
            const int idx = base + threadIdx.x;
            float ai = 0.f, bi = 0.f, ci = 0.f, tmp1 = 0.f, tmp2 = 0.f;
            int i = 0;
            if (idx < count) {
                i = stride * idx + stride - 1;
                const int iLeft = i - delta;
                int iRight = i + delta;
                if (iRight >= nT) iRight = nT - 1;
                tmp1 = __fdiv_rn(a[i], b[iLeft]);
                tmp2 = __fdiv_rn(c[i], b[iRight]);
                bi = __fmaf_rn(a[iRight], -tmp2, __fmaf_rn(c[iLeft], -tmp1, b[i]));
                ai = __fmul_rn(a[iLeft], -tmp1);
                ci = __fmul_rn(c[iRight], -tmp2);
            }
            // rows written at this level (odd multiples of delta, minus one) are never read at
            // this level (reads touch i +- delta, which are rows of the previous level), so no
            // barrier is needed between the chunks
            if (idx < count) {
                a[i] = ai; b[i] = bi; c[i] = ci;
                t.crFwd[tabBase + (size_t)(entry0 + idx) * W] = make_float2(tmp1, tmp2);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
        // stored in the order the back substitution visits the rows: row i with
        // i + 1 = (2*idx + 1) * 2^k is entry nTheta - (nTheta >> k) + idx
        const int k = __ffs(i + 1) - 1;
        const int e = nT - (nT >> k) + ((i + 1) >> (k + 1));
        t.crA[tabBase + (size_t)e * W] = a[i];
        t.crB[tabBase + (size_t)e * W] = b[i];
        t.crC[tabBase + (size_t)e * W] = c[i];
    }
}

// ---- per step ---------------------------------------------------------------------------------
// grid ((N/2) / W, batch), THREADS threads, dynamic smem nTheta * W * (8 [+ 20 if STAGE]) bytes
template <int W, int THREADS, bool STAGE>
__global__ void __launch_bounds__(THREADS)
tridiagonalKernel(GridParams g, SpectralTables t, float2* __restrict__ spectrumAll)
{
    extern __shared__ __align__(16) float2 d[];
    __shared__ __align__(8) uint64_t tabBar;
    const int nT = g.nTheta, half = g.nPhi >> 1;
    const int slot0 = blockIdx.x * W;
    float2* spectrum = spectrumAll + (size_t)blockIdx.y * (g.cells >> 1) + slot0;
    const size_t tabBase = (size_t)blockIdx.x * nT * W;
    const float2* fwd = t.crFwd + tabBase;
    const float* tabA = t.crA + tabBase;
    const float* tabB = t.crB + tabBase;
    const float* tabC = t.crC + tabBase;
    const int tid = threadIdx.x;
    constexpr int LW = (W == 2) ? 1 : (W == 4) ? 2 : 3;
    if (STAGE) {
        float2* sFwd = d + nT * W;
        float* sA = reinterpret_cast<float*>(sFwd + nT * W);
        float* sB = sA + nT * W;
        float* sC = sB + nT * W;
        if (tid == 0) { tma::mbarInit(&tabBar, 1); tma::fenceBarrierInit(); }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(nT * W * sizeof(float));
            tma::mbarExpectTx(&tabBar, 5 * bytes);
            tma::bulkLoad(sFwd, fwd, 2 * bytes, &tabBar);
            tma::bulkLoad(sA, tabA, bytes, &tabBar);
            tma::bulkLoad(sB, tabB, bytes, &tabBar);
            tma::bulkLoad(sC, tabC, bytes, &tabBar);
        }
        fwd = sFwd; tabA = sA; tabB = sB; tabC = sC;
    }

    pdlWait();                                   // the right-hand sides come from the previous kernel
    for (int item = tid; item < nT * W; item += THREADS) {
        const int i = item >> LW, w = item & (W - 1);
        d[swizzleRow<W>(i) * W + w] = spectrum[(size_t)i * half + w];
    }

    int levels = 0;
    while ((2 << levels) < nT) ++levels;
    if (STAGE) tma::mbarWait(&tabBar, 0);
    // forward elimination of the right-hand sides, kernel/tdm.cu:43-63 (d only)
    int stride = 1;
    for (int lvl = 0; lvl < levels; ++lvl) {
        __syncthreads();
        stride <<= 1;
        const int delta = stride >> 1;
        const int count = (nT >> (lvl + 1)) * W;
        const float2* f = fwd + (size_t)(nT - (nT >> lvl)) * W;
        for (int item = tid; item < count; item += THREADS) {
            const int idx = item >> LW, w = item & (W - 1);
            const int i = stride * idx + stride - 1;
            const int iLeft = i - delta;
            int iRight = i + delta;
            if (iRight >= nT) iRight = nT - 1;
            const float2 tm = f[idx * W + w];
            const int pi = swizzleRow<W>(i) * W + w;
            const float2 dl = d[swizzleRow<W>(iLeft) * W + w], dr = d[swizzleRow<W>(iRight) * W + w];
            float2 di = d[pi];
            di.x = __fmaf_rn(dr.x, -tm.y, __fmaf_rn(dl.x, -tm.x, di.x));
            di.y = __fmaf_rn(dr.y, -tm.y, __fmaf_rn(dl.y, -tm.x, di.y));
            d[pi] = di;
        }
    }
    __syncthreads();
    // 2 x 2 system, kernel/tdm.cu:65-72
    if (tid < W) {
        const int w = tid;
        const int i1 = stride - 1, i2 = 2 * stride - 1;
        const float b1 = tabB[(nT - 2) * W + w], c1 = tabC[(nT - 2) * W + w];     // row i1
        const float a2 = tabA[(nT - 1) * W + w], b2 = tabB[(nT - 1) * W + w];     // row i2
        const int p1 = swizzleRow<W>(i1) * W + w, p2 = swizzleRow<W>(i2) * W + w;
        const float2 d1 = d[p1], d2 = d[p2];
        const float det = __fmaf_rn(b2, b1, -__fmul_rn(c1, a2));
        float2 x1, x2;
        x1.x = __fdiv_rn(__fmaf_rn(b2, d1.x, -__fmul_rn(c1, d2.x)), det);
        x1.y = __fdiv_rn(__fmaf_rn(b2, d1.y, -__fmul_rn(c1, d2.y)), det);
        x2.x = __fdiv_rn(__fmaf_rn(d2.x, b1, -__fmul_rn(d1.x, a2)), det);
        x2.y = __fdiv_rn(__fmaf_rn(d2.y, b1, -__fmul_rn(d1.y, a2)), det);
        d[p1] = x1;
        d[p2] = x2;
    }
    // back substitution, kernel/tdm.cu:75-90; the solution overwrites the right-hand side
    int rows = 2;
    for (int lvl = 0; lvl < levels; ++lvl) {
        const int delta = stride >> 1;
        const int e0 = (nT - (nT >> (levels - 1 - lvl))) * W;      // first table entry of this level
        __syncthreads();
        for (int item = tid; item < rows * W; item += THREADS) {
            const int idx = item >> LW, w = item & (W - 1);
            const int i = stride * idx + delta - 1;
            const float ci = tabC[e0 + item], bi = tabB[e0 + item];
            const int pi = swizzleRow<W>(i) * W + w;
            const float2 di = d[pi], xp = d[swizzleRow<W>(i + delta) * W + w];
            float2 x;
            if (i == delta - 1) {
                x.x = __fdiv_rn(__fmaf_rn(-ci, xp.x, di.x), bi);
                x.y = __fdiv_rn(__fmaf_rn(-ci, xp.y, di.y), bi);
            } else {
                const float ai = tabA[e0 + item];
                const float2 xm = d[swizzleRow<W>(i - delta) * W + w];
                x.x = __fdiv_rn(__fmaf_rn(-ci, xp.x, __fmaf_rn(-ai, xm.x, di.x)), bi);
                x.y = __fdiv_rn(__fmaf_rn(-ci, xp.y, __fmaf_rn(-ai, xm.y, di.y)), bi);
            }
            d[pi] = x;
        }
        stride >>= 1;
        rows <<= 1;
    }
    __syncthreads();
    for (int item = tid; item < nT * W; item += THREADS) {
        const int i = item >> LW, w = item & (W - 1);
        spectrum[(size_t)i * half + w] = d[swizzleRow<W>(i) * W + w];
    }
}

struct TriLaunch { int W; int threads; bool stage; size_t smem; };

TriLaunch triLaunch(const GridParams& g)
{
    const int half = g.nPhi / 2;
    TriLaunch l;
    l.stage = g.nTheta <= 1024;
    if (l.stage) {
        // latency regime: as many blocks as possible, tables staged in shared memory
        l.W = (half / 4 >= 4 * 148) ? 4 : 2;
    } else {
        // throughput regime: widest slot group that still gives every SM a couple of blocks
        l.W = (half / 8 >= 2 * 148) ? 8 : 4;
        while (l.W > 2 && (size_t)g.nTheta * l.W * sizeof(float2) > 100 * 1024) l.W >>= 1;
    }
    const int items = g.nTheta / 2 * l.W;
    l.threads = items >= 1024 ? 512 : (items >= 256 ? 256 : 64);
    l.smem = (size_t)g.nTheta * l.W * (sizeof(float2) + (l.stage ? 5 * sizeof(float) : 0));
    return l;
}

template <int W, int THREADS, bool STAGE>
cudaError_t launchTri(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch, size_t smem,
                      cudaStream_t stream, bool configureOnly)
{
    if (configureOnly)
        return cudaFuncSetAttribute(tridiagonalKernel<W, THREADS, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((g.nPhi / 2) / W, batch);
    return launchChained(tridiagonalKernel<W, THREADS, STAGE>, grid, dim3(THREADS), smem, stream, g, t, spectrum);
}

cudaError_t dispatchTri(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch,
                        cudaStream_t stream, bool configureOnly)
{
    const TriLaunch l = triLaunch(g);
#define KB_TRI(WW, TT, SS) if (l.W == WW && l.threads == TT && l.stage == SS) return launchTri<WW, TT, SS>(g, t, spectrum, batch, l.smem, stream, configureOnly)
    KB_TRI(2, 64, true); KB_TRI(2, 256, true); KB_TRI(2, 512, true);
    KB_TRI(4, 64, true); KB_TRI(4, 256, true); KB_TRI(4, 512, true);
    KB_TRI(2, 512, false); KB_TRI(4, 512, false); KB_TRI(8, 512, false);
#undef KB_TRI
    return cudaErrorInvalidValue;
}

} // namespace

size_t crTableFloats(const GridParams& g)
{
    // fwd: nTheta x N/2 float2, bwd a/b/c: 3 x nTheta x N/2 floats
    return (size_t)g.nTheta * (g.nPhi / 2) * 5;
}

cudaError_t configureTridiagonal(const GridParams& g)
{
    SpectralTables none{};
    return dispatchTri(g, none, nullptr, 1, nullptr, true);
}

cudaError_t launchBuildCrTables(const GridParams& g, SpectralTables t, cudaStream_t stream)
{
    const size_t smem = 3 * sizeof(float) * (size_t)g.nTheta;
    cudaError_t e = cudaFuncSetAttribute(buildCrTablesKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int threads = g.nTheta / 2 < 256 ? (g.nTheta / 2 < 32 ? 32 : g.nTheta / 2) : 256;
    buildCrTablesKernel<<<g.nPhi / 2, threads, smem, stream>>>(g, t, triLaunch(g).W);
    return cudaGetLastError();
}

cudaError_t launchTridiagonal(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch,
                              cudaStream_t stream)
{
    return dispatchTri(g, t, spectrum, batch, stream, false);
}

} // namespace kb
