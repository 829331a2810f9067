// Semi-Lagrangian advection of u_phi, u_theta, density and the tracer particles.
//
// Replaces advectionVPhiKernel / advectionVThetaKernel / advectionCentered /
// advectionParticles and KaminoSolver::advection (kernel/KaminoCore.cu:186-384): four
// launches and two device syncs become ONE launch whose 1-D grid is partitioned into
// two block ranges (tile blocks: u_phi, u_theta and density of 8 x 32 cells | particle blocks).
// Everything reads the pre-advection velocity, as in the reference.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "kamino_kernels.cuh"
#include "sampler.cuh"

namespace kb {

namespace { thread_local bool g_capturing = false; }

void pdlSetCapturing(bool capturing) { g_capturing = capturing; }
bool pdlEnabled() { return g_capturing; }

namespace {

constexpr int kAdvectThreads = 256;

// One RK2 backtrace from the node of `KIND` cell (j, i); samples `src` at the foot point. The
// two velocity samples of each stage are issued together. kernel/KaminoCore.cu:195-228
// (and :240-273, :285-318); arithmetic notes below.
template <int KIND, bool SAFE>
__device__ __forceinline__ float backtrace(const SamplerRegs& g, const SamplerConsts* __restrict__ sc, float cofTheta,
                                           const float* __restrict__ velPhi, const float* __restrict__ velTheta,
                                           const float* __restrict__ src, int i, int j, float cofPhi,
                                           unsigned tilePhi, unsigned tileTheta, unsigned tileSrc)
{
    const float offPhi = (KIND == kVPhi) ? -0.5f : 0.0f;
    const float offTheta = (KIND == kVTheta) ? 1.0f : 0.5f;
    const float gPhi = __fmul_rn(__fadd_rn((float)i, offPhi), g.h);
    const float gTheta = __fmul_rn(__fadd_rn((float)j, offTheta), g.h);
    PendingSample pu = sampleIssueTiled<kVPhi, SAFE>(g, sc, velPhi, gPhi, gTheta, tilePhi);
    PendingSample pv = sampleIssueTiled<kVTheta, SAFE>(g, sc, velTheta, gPhi, gTheta, tileTheta);
    const float guPhi = sampleFinish(pu);
    const float guTheta = sampleFinish(pv);
    const float deltaPhi = __fmul_rn(guPhi, cofPhi);
    const float deltaTheta = __fmul_rn(guTheta, cofTheta);
    const float midPhi = __fmaf_rn(-0.5f, deltaPhi, gPhi);
    const float midTheta = __fmaf_rn(-0.5f, deltaTheta, gTheta);
    pu = sampleIssueTiled<kVPhi, SAFE>(g, sc, velPhi, midPhi, midTheta, tilePhi);
    pv = sampleIssueTiled<kVTheta, SAFE>(g, sc, velTheta, midPhi, midTheta, tileTheta);
    const float muPhi = sampleFinish(pu);
    const float muTheta = sampleFinish(pv);
    const float averuPhi = __fmul_rn(0.5f, __fadd_rn(muPhi, guPhi));
    const float averuTheta = __fmul_rn(0.5f, __fadd_rn(muTheta, guTheta));
    const float pPhi = __fmaf_rn(-averuPhi, cofPhi, gPhi);
    const float pTheta = __fmaf_rn(-averuTheta, cofTheta, gTheta);
    return sampleFinish(sampleIssueTiled<KIND, SAFE>(g, sc, src, pPhi, pTheta, tileSrc));
}

// Arithmetic notes for backtrace():
//  * node coordinate ((fReal)i + offset) * gridLen is an exact fp64 sum and product
//    rounded once, which equals the fp32 product of the exact fp32 sum;
//  * g - 0.5*delta (fp64 in the reference) is an exact product and a sum of two fp32
//    values, i.e. fmaf(-0.5, delta, g);
//  * 0.5 * (mu + gu) is an fp32 add followed by an exact halving;
//  * g - aver*cof is one FFMA in the reference's SASS;
//  * cofPhi = dt / (R sinf(gTheta)) depends on the row only: tabulated (same device functions).
// Tried and rejected (r01g/r01h A/B): advancing the three backtraces of a cell stage by stage so
// that six samples are in flight per round -- 80-100 registers, 10 % slower than this form.

// backtrace() of an interior block without a branch: see sampleIssueFast. `bad` comes back true for
// the lanes whose result must be discarded.
template <int KIND>
__device__ __forceinline__ float backtraceFast(const SamplerRegs& g, float cofTheta, int i, int j, float cofPhi,
                                               unsigned tilePhi, unsigned tileTheta, unsigned tileSrc, bool& bad)
{
    const float offPhi = (KIND == kVPhi) ? -0.5f : 0.0f;
    const float offTheta = (KIND == kVTheta) ? 1.0f : 0.5f;
    const float gPhi = __fmul_rn(__fadd_rn((float)i, offPhi), g.h);
    const float gTheta = __fmul_rn(__fadd_rn((float)j, offTheta), g.h);
    bad = false;
#ifdef KB_STAGE1_CHECK
    constexpr bool kNodeCheck = true;
#else
    constexpr bool kNodeCheck = false;
#endif
    PendingSample pu = sampleIssueFast<kVPhi, kNodeCheck>(g, gPhi, gTheta, tilePhi, bad);       // at the cell's own node
    PendingSample pv = sampleIssueFast<kVTheta, kNodeCheck>(g, gPhi, gTheta, tileTheta, bad);
    const float guPhi = sampleFinish(pu);
    const float guTheta = sampleFinish(pv);
    const float deltaPhi = __fmul_rn(guPhi, cofPhi);
    const float deltaTheta = __fmul_rn(guTheta, cofTheta);
    const float midPhi = __fmaf_rn(-0.5f, deltaPhi, gPhi);
    const float midTheta = __fmaf_rn(-0.5f, deltaTheta, gTheta);
    pu = sampleIssueFast<kVPhi>(g, midPhi, midTheta, tilePhi, bad);
    pv = sampleIssueFast<kVTheta>(g, midPhi, midTheta, tileTheta, bad);
    const float muPhi = sampleFinish(pu);
    const float muTheta = sampleFinish(pv);
    const float averuPhi = __fmul_rn(0.5f, __fadd_rn(muPhi, guPhi));
    const float averuTheta = __fmul_rn(0.5f, __fadd_rn(muTheta, guTheta));
    const float pPhi = __fmaf_rn(-averuPhi, cofPhi, gPhi);
    const float pTheta = __fmaf_rn(-averuTheta, cofTheta, gTheta);
    return sampleFinish(sampleIssueFast<KIND>(g, pPhi, pTheta, tileSrc, bad));
}

// kernel/KaminoCore.cu:321-342, in two halves so that a thread can have the gathers of several
// particles in flight: the two velocity samples are issued, then finished and applied.
struct PendingParticle { PendingSample u, v; };

__device__ __forceinline__ PendingParticle particleIssue(const SamplerRegs& g, const SamplerConsts* __restrict__ sc,
                                                         const float* __restrict__ velPhi,
                                                         const float* __restrict__ velTheta, float2 pos)
{
    PendingParticle p;
    p.u = sampleIssue<kVPhi>(g, sc, velPhi, pos.x, pos.y);
    p.v = sampleIssue<kVTheta>(g, sc, velTheta, pos.x, pos.y);
    return p;
}

__device__ __forceinline__ float2 particleFinish(float radius, float dt, float cofTheta, float2 pos, const PendingParticle& p)
{
    const float posPhi = pos.x, posTheta = pos.y;
    const float uPhi = sampleFinish(p.u);
    const float uTheta = sampleFinish(p.v);
    const float latRadius = __fmul_rn(radius, sinf(posTheta));
    const float cofPhi = __fdiv_rn(dt, latRadius);
    float updatedTheta = __fmaf_rn(uTheta, cofTheta, posTheta);
    float updatedPhi = posPhi;
    if (latRadius > 1e-7f) updatedPhi = __fmaf_rn(uPhi, cofPhi, posPhi);
    if (!(__float_as_uint(updatedTheta) < 0x40490FDBu && __float_as_uint(updatedPhi) < 0x40C90FDBu)) {
        const Validated v = validateCoord(updatedPhi, updatedTheta);
        updatedPhi = v.phi; updatedTheta = v.theta;
    }
    return make_float2(updatedPhi, updatedTheta);
}

__device__ __forceinline__ float2 pushParticle(const SamplerRegs& g, const SamplerConsts* __restrict__ sc,
                                               float radius, float dt, float cofTheta,
                                               const float* __restrict__ velPhi,
                                               const float* __restrict__ velTheta, float2 pos)
{
    return particleFinish(radius, dt, cofTheta, pos, particleIssue(g, sc, velPhi, velTheta, pos));
}

// Grid: [tile blocks | particle blocks] x batch. A tile block owns kTileRows x 32 cells (one warp
// per row segment, so stores stay 128-byte coalesced), stages the 13 x 40 neighbourhood of
// u_phi, u_theta and density in shared memory and runs the three backtraces of its cells one
// after the other: nearly all of their 15 samples per cell are served from the tiles, the rest
// (polar rows, where the phi displacement exceeds the halo) gather from L1/L2, and every field
// is streamed from HBM once per step. The particle blocks (one
// thread per particle) come last in the grid and fill the tail of the tile blocks.
constexpr int kTileRows = kAdvectThreads / 32;

// Register budget: 6 blocks per SM (r01k/r01n A/B at 2048 x 4096: 217.6 / 209.1 / 198.6 / 195.8 us for
// 4 / 5 / 6 / 7 blocks per SM; 7 costs 3 us at 512 x 1024).
__global__ void __launch_bounds__(kAdvectThreads, 6)
advectKernel(GridParams g, AdvectArgs a)
{
    const int sim = blockIdx.y;
    const float* velPhi = pinPointer(a.velPhi + (size_t)sim * g.cells);
    const float* velTheta = pinPointer(a.velTheta + (size_t)sim * g.cells);
    const SamplerRegs sr(a.consts);            // read-only table: independent of the previous kernel
    const int block = blockIdx.x;
    const bool isTile = block < a.tileBlocks;
    const int particleBlock = block - a.tileBlocks;
    pdlWait();

    if (isTile) {
        __shared__ __align__(16) float tiles[3][kTileH * kTileStride];
        const float* density = pinPointer(a.density + (size_t)sim * g.cells);
        const int log2TilesX = g.log2NPhi - 5;
        const int i0 = (block & ((1 << log2TilesX) - 1)) << 5;
        const int j0 = g.rowBegin + (block >> log2TilesX) * kTileRows;
        SamplerRegs tr = sr;
        tr.tileRow0 = j0 - 2;
        tr.tileCol0 = i0 - 4;
        // stage the neighbourhood of the block's cells as float4 (tileCol0 is a multiple of 4, so a
        // group of four columns never straddles the seam; rows clamped into the arrays: clamped rows
        // are never addressed by an interior sample; columns wrap around the seam)
        constexpr int kVecPerRow = kTileW / 4, kVecPerTile = kTileH * kVecPerRow;
        for (int e = threadIdx.x; e < 3 * kVecPerTile; e += kAdvectThreads) {
            const int f = e / kVecPerTile, rem = e - f * kVecPerTile;
            const int r = rem / kVecPerRow, q = rem - r * kVecPerRow;
            const int col = (tr.tileCol0 + 4 * q) & sr.mask;
            const int rowC = min(max(tr.tileRow0 + r, 0), sr.nTheta - 1);
            const int row = (f == 1) ? min(rowC, sr.nTheta - 2) : rowC;
            const float* src = (f == 0) ? velPhi : (f == 1) ? velTheta : density;
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (row * sr.N + col)));
            *reinterpret_cast<float4*>(&tiles[f][r * kTileStride + 4 * q]) = v;
        }
        __syncthreads();
        unsigned tileBase = (unsigned)__cvta_generic_to_shared(&tiles[0][0]);
        asm volatile("" : "+r"(tileBase));       // opaque: kept in a register instead of re-derived per sample
        const unsigned tPhi = tileBase, tTheta = tileBase + kTileBytes, tRho = tileBase + 2 * kTileBytes;
        const int i = i0 + (threadIdx.x & 31);
        const int j = j0 + (threadIdx.x >> 5);
        const size_t cell = (size_t)sim * g.cells + (size_t)j * g.nPhi + i;
        const float cofCentred = __ldg(a.cofPhiCentred + j);
        // blocks whose tile lies inside rows [2, nTheta-4] and columns [2, N-2]: see sampleIssueTiled
        const bool safe = tr.tileRow0 >= 2 && tr.tileRow0 + (kTileH - 1) <= sr.nTheta - 4
                       && tr.tileCol0 >= 2 && tr.tileCol0 + (kTileW - 1) <= sr.N - 2;
        const float cofTheta = __ldg(a.cofPhiTheta + j);
        // interior blocks: the three backtraces without a branch; a lane whose samples left the tile
        // (or a block that is not interior) redoes the backtrace below with the branching samplers
        bool redoU = true, redoV = true, redoRho = true;
        if (safe) {
            const float u = backtraceFast<kVPhi>(tr, g.cofTheta, i, j, cofCentred, tPhi, tTheta, tPhi, redoU);
            const float v = backtraceFast<kVTheta>(tr, g.cofTheta, i, j, cofTheta, tPhi, tTheta, tTheta, redoV);
            const float rho = backtraceFast<kCentered>(tr, g.cofTheta, i, j, cofCentred, tPhi, tTheta, tRho, redoRho);
            if (!redoU) a.velPhiOut[cell] = u;
            if (!redoV) a.velThetaOut[cell] = v;           // interior blocks never hold row nTheta-1
            if (!redoRho) a.densityOut[cell] = rho;
        }
        if (redoU)
            a.velPhiOut[cell] = backtrace<kVPhi, false>(tr, a.consts, g.cofTheta, velPhi, velTheta, velPhi, i, j, cofCentred,
                                                        tPhi, tTheta, tPhi);
        if (redoV && j < g.nTheta - 1)
            a.velThetaOut[cell] = backtrace<kVTheta, false>(tr, a.consts, g.cofTheta, velPhi, velTheta, velTheta, i, j,
                                                            cofTheta, tPhi, tTheta, tTheta);
        if (redoRho)
            a.densityOut[cell] = backtrace<kCentered, false>(tr, a.consts, g.cofTheta, velPhi, velTheta, density, i, j,
                                                             cofCentred, tPhi, tTheta, tRho);
        return;
    }
    unsigned k;                                     // particle counts are far below 2^32
    if (a.latticeInner > 0) {
        // compact patch of the particle lattice per warp / block (see AdvectArgs); pb / blocksInner by a multiply
        // with ceil(2^32 / blocksInner) (exact: pb * blocksInner < 2^32; magic 0 stands for a divisor of 1)
        const unsigned pb = (unsigned)particleBlock;
        const unsigned blockO = a.innerMagic ? __umulhi(pb, a.innerMagic) : pb, blockI = pb - blockO * a.blocksInner;
        const unsigned log2BI = a.log2Inner < 4 ? 4 : a.log2Inner;          // block patch: 2^log2BI rows
        const unsigned sh = log2BI - a.log2Inner;
        const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const unsigned wi = warp & ((1u << sh) - 1), wo = warp >> sh;
        const unsigned inner = (blockI << log2BI) + (wi << a.log2Inner) + (lane & ((1u << a.log2Inner) - 1));
        const unsigned outer = blockO * (kAdvectThreads >> log2BI) + wo * (32u >> a.log2Inner) + (lane >> a.log2Inner);
        k = (inner < (unsigned)a.latticeInner && outer < (unsigned)a.latticeOuter) ? outer * a.latticeInner + inner : 0xffffffffu;
    } else {
        k = (unsigned)particleBlock * kAdvectThreads + threadIdx.x;
    }
    if (k < (unsigned long)g.numParticles) {        // the reference has no tail guard (:323)
        const float2* in = reinterpret_cast<const float2*>(a.particles) + (size_t)sim * g.numParticles;
        float2* out = reinterpret_cast<float2*>(a.particlesOut) + (size_t)sim * g.numParticles;
        out[k] = pushParticle(sr, a.consts, g.radius, g.dt, g.cofTheta, velPhi, velTheta, in[k]);
    }
}

template <int KIND>
__global__ void locateKernel(const SamplerConsts* __restrict__ consts, long n, const float* __restrict__ phiRaw,
                             const float* __restrict__ thetaRaw, int* phiIndex, int* thetaIndex,
                             float* alphaPhi, float* alphaTheta, float* phiOut, float* thetaOut, int* flags)
{
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const SamplerRegs sr(consts);
    Location loc = locate<KIND>(sr, phiRaw[k], thetaRaw[k]);
    phiIndex[k] = loc.phiIndex;
    thetaIndex[k] = loc.thetaIndex;
    alphaPhi[k] = loc.alphaPhi;
    alphaTheta[k] = loc.alphaTheta;
    phiOut[k] = loc.phi;
    thetaOut[k] = loc.theta;
    flags[k] = (loc.flipped ? 1 : 0) | (poleBranch<KIND>(sr, loc) ? 2 : 0);
}

} // namespace

void fillSamplerConsts(const GridParams& g, void* hostBlock64, int validLo, int validHi, int* haloViolation)
{
    SamplerConsts c{};
    c.validLo = validLo; c.validHi = validHi < 0 ? g.nTheta : validHi; c.haloViolation = haloViolation;
    c.h = g.h; c.halfH = g.halfH; c.invH = g.invH;
    c.N = g.nPhi; c.mask = g.nPhi - 1; c.halfN = g.nPhi >> 1; c.nTheta = g.nTheta;
    auto bits = [](float x) { unsigned u; memcpy(&u, &x, 4); return u; };
    // interior fast path of sample(): see sampler.cuh. lastRow = nTheta-1 (u_phi, centred) / nTheta-2 (u_theta)
    const float lo = 1.5f * g.h;
    const float hiCentred = ((float)(g.nTheta - 1) - 0.5f) * g.h;
    const float hiVTheta = ((float)(g.nTheta - 2) - 0.5f) * g.h;
    c.thetaLoBits = bits(lo);
    c.thetaSpanCentred = bits(hiCentred) - bits(lo);
    c.thetaSpanVTheta = bits(hiVTheta) - bits(lo);
    c.phiLoBits = bits(lo);
    c.phiSpan = bits(kTwoPiF) - bits(lo);
    static_assert(sizeof(SamplerConsts) == 64, "SamplerConsts is a 64-byte block");
    memcpy(hostBlock64, &c, sizeof(c));
}
cudaError_t launchAdvect(const GridParams& g, AdvectArgs a, int batch, cudaStream_t stream)
{
    a.tileBlocks = (g.nPhi / 32) * (g.rowCount / kTileRows);      // rowBegin, rowCount: multiples of kTileRows
    int blocksParticles = (a.particles && g.numParticles > 0)
        ? (int)((g.numParticles + kAdvectThreads - 1) / kAdvectThreads) : 0;
    a.latticeInner = a.latticeOuter = a.log2Inner = a.blocksInner = 0; a.innerMagic = 0;
    if (blocksParticles > 0) {
        // numOfParticles = numTheta * (2 numTheta) for a seeded set (kernel/KaminoParticles.cu:22-25)
        const long m = (long)(sqrt((double)g.numParticles / 2.0) + 0.5);
        if (m >= 16 && 2 * m * m == g.numParticles) {
            // warp patch height: about three grid cells of lattice rows (spacing nTheta / m cells)
            int log2Inner = 2;
            while (log2Inner < 5 && (2L << log2Inner) * g.nTheta <= 3 * m) ++log2Inner;
            const int log2BI = log2Inner < 4 ? 4 : log2Inner;
            if (log2Inner < 5) {       // a 32-row patch is the linear mapping (dense particle sets, r01g A/B at C1)
                a.latticeInner = (int)m; a.latticeOuter = (int)(2 * m); a.log2Inner = log2Inner;
                a.blocksInner = (int)((m + (1 << log2BI) - 1) >> log2BI);
                a.innerMagic = a.blocksInner > 1 ? (unsigned)(((1ull << 32) + a.blocksInner - 1) / a.blocksInner) : 0u;   // 0: divisor 1
                const int rowsOuter = kAdvectThreads >> log2BI;
                blocksParticles = a.blocksInner * (int)((2 * m + rowsOuter - 1) / rowsOuter);
            }
        }
    }
    // (r01h A/B: interleaving tile and particle blocks along the grid is a loss, 35.3 -> 41.4 us at 512 x 1024:
    // tile blocks first, the particle blocks fill their tail waves)
    dim3 grid(a.tileBlocks + blocksParticles, batch);
    return launchChained(advectKernel, grid, dim3(kAdvectThreads), 0, stream, g, a);
}

cudaError_t launchLocate(const SamplerConsts* consts, int kind, long n, const float* phiRaw, const float* thetaRaw,
                         int* phiIndex, int* thetaIndex, float* alphaPhi, float* alphaTheta,
                         float* phiOut, float* thetaOut, int* flags, cudaStream_t stream)
{
    const int threads = 256;
    const int blocks = (int)((n + threads - 1) / threads);
    if (blocks == 0) return cudaSuccess;
    switch (kind) {
    case kVPhi:
        locateKernel<kVPhi><<<blocks, threads, 0, stream>>>(consts, n, phiRaw, thetaRaw, phiIndex, thetaIndex,
                                                            alphaPhi, alphaTheta, phiOut, thetaOut, flags);
        break;
    case kVTheta:
        locateKernel<kVTheta><<<blocks, threads, 0, stream>>>(consts, n, phiRaw, thetaRaw, phiIndex, thetaIndex,
                                                              alphaPhi, alphaTheta, phiOut, thetaOut, flags);
        break;
    default:
        locateKernel<kCentered><<<blocks, threads, 0, stream>>>(consts, n, phiRaw, thetaRaw, phiIndex, thetaIndex,
                                                                alphaPhi, alphaTheta, phiOut, thetaOut, flags);
        break;
    }
    return cudaGetLastError();
}

} // namespace kb
