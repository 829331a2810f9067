// Bilinear samplers on the staggered spherical grid (device code).
//
// Behavioural contract: sampleVPhi / sampleVTheta / sampleCentered and validateCoord of
// the reference (kernel/KaminoCore.cu:11-184), including its quirks (the
// sampled value is not sign-flipped across a pole, sampleVPhi's pole branch does not halve
// the theta weight, the belt order in the flipped north-pole case).
//
// The reference evaluates these in a mix of fp32 and fp64 because its constants are double
// literals. This implementation produces the same bits with almost no fp64:
//  * (float)(double-op of two fp32 values) == the fp32 op, for + - * / (double rounding is
//    innocuous when the wide format has >= 2p+2 bits); this covers the stagger shifts, the
//    node coordinates and 1.0/gridLen.
//  * (double)x > M_PI  <=>  x >= kPiF, and 0 <= x < kTwoPiF  =>  the 2*pi wrap is the
//    identity; other inputs take the reference's own fp64 expressions (rare: seam and
//    pole-crossing lanes).
//  * kaminoLerp = (float)fma(1.0 - a, from, (double)(a*to)). When 1-a is exact in fp32
//    (a is a multiple of 2^-24, true whenever the cell index is >= 1) this equals
//    fmaf(1-a, from, a*to) except for a ~2^-29 chance of a double-rounding tie; the other
//    lanes evaluate the fp64 form.
#pragma once

#include "kamino_common.cuh"

namespace kb {

struct Location {
    int phiIndex;      // before the modulo
    int thetaIndex;
    float alphaPhi;
    float alphaTheta;  // before the optional halving
    float phi;         // validated coordinates
    float theta;
    bool flipped;      // validateCoord returned -1
};

// x - (int)floorf(x / M_2PI) * M_2PI   (kernel/KaminoCore.cu:14,27)
__device__ __forceinline__ float wrapTwoPi(float x)
{
    if (x >= 0.0f && x < kTwoPiF) return x;
    double xd = (double)x;
    int k = (int)floorf((float)(xd / kTwoPi));
    return (float)(xd - (double)k * kTwoPi);
}

// kernel/KaminoCore.cu:11-29; `flipped` is true when the reference returns -1.
struct Validated { float phi, theta; bool flipped; };

__device__ __forceinline__ Validated validateCoordInline(float phi, float theta)
{
    bool flipped = false;
    theta = wrapTwoPi(theta);
    if (theta >= kPiF) {                       // (double)theta > M_PI
        theta = (float)(kTwoPi - (double)theta);
        phi = (float)((double)phi + kPi);
        flipped = !flipped;
    }
    if (theta < 0.0f) {
        theta = -theta;
        phi = (float)((double)phi + kPi);
        flipped = !flipped;
    }
    phi = wrapTwoPi(phi);
    return Validated{phi, theta, flipped};
}

// Out of line (values in registers both ways): only lanes that cross a pole or the seam come
// here, the callers test the in-range case first.
__device__ __noinline__ Validated validateCoord(float phi, float theta)
{
    return validateCoordInline(phi, theta);
}

// kernel/KaminoCore.cu:31-34
__device__ __forceinline__ float lerpWide(float from, float to, float alpha)
{
    float at = __fmul_rn(alpha, to);
    return (float)fma(1.0 - (double)alpha, (double)from, (double)at);
}

__device__ __forceinline__ float lerpFast(float from, float to, float alpha, float oneMinusAlpha)
{
    return __fmaf_rn(oneMinusAlpha, from, __fmul_rn(alpha, to));
}

// Grid constants held in registers for the whole kernel. They are LOADED from a small global
// table (SamplerConsts, built by the host at context creation) rather than read from the kernel
// parameters: ptxas re-materialises parameter reads from the constant bank at every use (6 extra
// LDC/LDCU per sample in the r01e profile of this issue-bound kernel), but it never re-issues a
// global load.
struct SamplerConsts {               // 16 words, 16-byte aligned
    float h, halfH, invH;
    int validLo;                     // first field row held in memory (0 unless the context is a theta band)
    int N, mask, halfN, nTheta;
    // interior fast path: shifted theta in [thetaLo, thetaHi[kind]) and shifted phi in [phiLo, 2 pi)
    // as unsigned ranges on the bit patterns: (bits - lo) < span
    unsigned thetaLoBits, thetaSpanCentred, thetaSpanVTheta;
    int validHi;                     // one past the last row held in memory (nTheta unless a theta band)
    unsigned phiLoBits, phiSpan;
    // Band-decomposed runs (dist.cu) hold rows [validLo, validHi) only. A gather of the general path that
    // would leave them (a backtrace longer than the halo: theta-CFL too large for the halo width) is
    // clamped into memory and recorded here; the host checks the flag at the next synchronisation.
    int* haloViolation;              // NULL when the whole grid is resident
};

struct SamplerRegs {
    float h, halfH, invH;
    int N, mask, halfN, nTheta;
    unsigned thetaLoBits, thetaSpanCentred, thetaSpanVTheta, phiLoBits, phiSpan;
    int tileRow0 = 0, tileCol0 = 0;     // grid row / column of element (0, 0) of the block's smem tiles
    __device__ __forceinline__ explicit SamplerRegs(const SamplerConsts* __restrict__ c)
    {
        const float4 a = __ldg(reinterpret_cast<const float4*>(c));
        const int4 b = __ldg(reinterpret_cast<const int4*>(c) + 1);
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(c) + 2);
        const uint2 e = __ldg(reinterpret_cast<const uint2*>(c) + 6);
        h = a.x; halfH = a.y; invH = a.z;
        N = b.x; mask = b.y; halfN = b.z; nTheta = b.w;
        thetaLoBits = d.x; thetaSpanCentred = d.y; thetaSpanVTheta = d.z;
        phiLoBits = e.x; phiSpan = e.y;
    }
};

// keep a base pointer in a register pair so that element `idx` is one IMAD.WIDE away
__device__ __forceinline__ const float* pinPointer(const float* p)
{
    asm volatile("" : "+l"(p));
    return p;
}

// Index / weight / predicate part of the samplers (kernel/KaminoCore.cu:36-53, 86-103, 136-153).
template <int KIND>
__device__ __forceinline__ Location locate(const SamplerRegs& g, float phiRaw, float thetaRaw)
{
    Location loc;
    // stagger shift (kernel/KaminoCore.cu:38-39, 88-89, 138-139)
    float phi = (KIND == kVPhi) ? __fadd_rn(phiRaw, g.halfH) : phiRaw;
    float theta = (KIND == kVTheta) ? __fsub_rn(thetaRaw, g.h) : __fsub_rn(thetaRaw, g.halfH);
    // in range (0 <= theta < pi and 0 <= phi < 2 pi) validateCoord is the identity
    loc.flipped = false;
    if (!(__float_as_uint(theta) < 0x40490FDBu && __float_as_uint(phi) < 0x40C90FDBu)) {
        const Validated v = validateCoord(phi, theta);
        phi = v.phi; theta = v.theta; loc.flipped = v.flipped;
    }
    const float normedPhi = __fmul_rn(phi, g.invH);
    const float normedTheta = __fmul_rn(theta, g.invH);
    loc.phiIndex = (int)floorf(normedPhi);
    loc.thetaIndex = (int)floorf(normedTheta);       // >= 0: theta >= 0 after validateCoord
    loc.alphaPhi = __fsub_rn(normedPhi, (float)loc.phiIndex);
    loc.alphaTheta = __fsub_rn(normedTheta, (float)loc.thetaIndex);
    loc.phi = phi;
    loc.theta = theta;
    return loc;
}

// pole branch of the reference (:52-53, :102-103, :152-153)
template <int KIND>
__device__ __forceinline__ bool poleBranch(const SamplerRegs& g, const Location& loc)
{
    const int lastRow = (KIND == kVTheta) ? g.nTheta - 2 : g.nTheta - 1;
    return (loc.thetaIndex == lastRow) || (loc.thetaIndex == 0 && loc.flipped);
}

// One bilinear sample of `field` (rows x nPhi, dense) at the raw coordinate: the general path
// (pole and seam crossings, first row / column, last rows). All four gathers are addressed with
// 32-bit element offsets from one base pointer; the pole / out-of-range row handling is two
// selects; the in-range case of validateCoord is detected with two unsigned compares on the
// bit patterns (-0.0f and NaN fall through to validateCoord, which treats them as the
// reference does).
template <int KIND>
__device__ __forceinline__ float sampleGeneralInline(const SamplerRegs& g, const SamplerConsts* __restrict__ consts,
                                                     const float* __restrict__ field, float phiRaw, float thetaRaw)
{
    const Location loc = locate<KIND>(g, phiRaw, thetaRaw);
    const int phiIndex = loc.phiIndex, thetaIndex = loc.thetaIndex;
    const float alphaPhi = loc.alphaPhi;
    float alphaTheta = loc.alphaTheta;
    const int N = g.N, mask = g.mask;
    const int lastRow = (KIND == kVTheta) ? g.nTheta - 2 : g.nTheta - 1;
    const bool pole = poleBranch<KIND>(g, loc);
    // rows past the array are clamped to the last one: the reference reads out of bounds
    // there (undefined; only reachable when the theta-CFL exceeds 1 at the south pole)
    const bool oneRow = pole || thetaIndex > lastRow;
    int rowLo = min(thetaIndex, lastRow);
    if (consts->haloViolation) {
        // theta band: rows [validLo, validHi) are resident (cold path: one extra load of the constants block)
        const int lo = consts->validLo, hi = consts->validHi - (oneRow ? 1 : 2);
        if (rowLo < lo || rowLo > hi) {
            atomicOr(consts->haloViolation, 1);
            rowLo = max(lo, min(rowLo, hi));
        }
    }
    const int colShift = pole ? g.halfN : 0;             // second belt = same row at phi + pi
    const int rowStep = oneRow ? 0 : N;
    const int c0 = phiIndex & mask;                         // size_t % nPhi, nPhi = 2^k
    const int c1 = (c0 + 1) & mask;
    const int cA = (c0 + colShift) & mask;
    const int cB = (cA + 1) & mask;
    const int lo = rowLo * N;
    const int hi = lo + rowStep;
    const float v00 = __ldg(field + (lo + c0));
    const float v01 = __ldg(field + (lo + c1));
    const float v10 = __ldg(field + (hi + cA));
    const float v11 = __ldg(field + (hi + cB));

    if (KIND != kVPhi) alphaTheta = pole ? __fmul_rn(0.5f, alphaTheta) : alphaTheta;   // :115, :165

    // alpha is a multiple of 2^-23 (2^-24 after the halving) whenever the index is >= 1, which
    // makes 1 - alpha exact in fp32 (see the file header)
    if (min(phiIndex, thetaIndex) >= 1) {
        const float omPhi = __fsub_rn(1.0f, alphaPhi);
        const float omTheta = __fsub_rn(1.0f, alphaTheta);
        const float lowerBelt = lerpFast(v00, v01, alphaPhi, omPhi);
        const float higherBelt = lerpFast(v10, v11, alphaPhi, omPhi);
        return lerpFast(lowerBelt, higherBelt, alphaTheta, omTheta);
    } else {
        const float lowerBelt = lerpWide(v00, v01, alphaPhi);
        const float higherBelt = lerpWide(v10, v11, alphaPhi);
        return lerpWide(lowerBelt, higherBelt, alphaTheta);
    }
}

// Out-of-line instance of the general path; it re-loads the constants it needs (cold code).
template <int KIND>
__device__ __noinline__ float sampleGeneral(const SamplerConsts* __restrict__ consts, const float* __restrict__ field,
                                            float phiRaw, float thetaRaw)
{
    const SamplerRegs g(consts);
    return sampleGeneralInline<KIND>(g, consts, field, phiRaw, thetaRaw);
}

// The sample as the kernels call it, split in two so that the gathers of several independent
// samples are in flight together (the kernel is otherwise bound by one exposed memory latency
// per sample): sampleIssue() computes the cell and issues the four loads, sampleFinish() does
// the three lerps.
//
// Interior fast path: when the shifted coordinate satisfies
//   1.5 h <= theta < (lastRow - 0.5) h   and   1.5 h <= phi < 2 pi
// then validateCoord is the identity, 1 <= thetaIndex < lastRow (no pole branch, no clamping),
// phiIndex >= 1, and both weights are multiples of 2^-23, so 1 - alpha is exact in fp32 and
// the three lerps are FMUL + FFMA (file header). Everything else takes the general path, out of
// line, and its finished value v travels through sampleFinish() as the degenerate bilinear
// form (v, v, v, v; alpha = 0), which returns v unchanged (1*v + 0*v, exact for finite v).
// Both paths evaluate the reference's expressions, so which one a lane takes never changes a bit.
//
// Shared-memory tiles (grid-cell blocks of the advection kernel): the block stages the
// kTileH x kTileW neighbourhood of its cells of every sampled field in shared memory once
// (row tileRow0 + r, column (tileCol0 + c) mod N at tile[r * kTileW + c]); an interior sample
// whose cell and its +1 neighbours lie inside the tile reads the four values from there (same
// values, ~30 cycles, no L1 tag traffic), everything else gathers from global memory.
struct PendingSample { float v00, v01, v10, v11, alphaPhi, alphaTheta; };

constexpr int kTileH = 13;      // 8 rows of cells + 2 above + 3 below
constexpr int kTileW = 40;      // 32 columns + 4 left + 4 right
// Row stride of the tiles in shared memory (floats): dense. Measured and rejected (r02a A/B): rows padded
// to 64 floats, which removes the bank conflicts between the rows of a warp's samples (27 % of the
// shared-memory wavefronts in the r01j ncu capture) -- 33.9 instead of 33.5 us at 512 x 1024, 199.0 instead
// of 197.7 us at 2048 x 4096: the kernel is bound by instruction issue, not by the shared-memory pipe.
constexpr int kTileStride = kTileW;
constexpr int kTileBytes = kTileH * kTileStride * 4;

// The four corners of a cell from a shared-memory tile addressed by its 32-bit shared-window
// address: explicit ld.shared with immediate offsets, so the tile base lives in one register for
// the whole kernel (with generic pointers ptxas re-materialised the shared-window base -- S2UR,
// UMOV, ULEA -- at every one of the 15 samples of a cell, r01f SASS profile).
__device__ __forceinline__ void loadTileCell(unsigned addr, PendingSample& p)
{
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p.v00) : "r"(addr));
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(p.v01) : "r"(addr));
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(p.v10) : "r"(addr), "n"(kTileStride * 4));
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(p.v11) : "r"(addr), "n"(kTileStride * 4 + 4));
}

// Sample from the block's tile (tile blocks of the advection kernel). `tile` is the shared-window
// address of the field's kTileH x kTileW tile. Everything that is not served by the tile goes to
// the general path out of line (polar rows, whose phi displacement exceeds the halo).
//   SAFE = false: the interior test on the coordinates plus the tile test on the indices.
//   SAFE = true : blocks whose tile lies inside rows [2, nTheta-4] and columns [2, N-2] (all but the
//     first / last tile row and column of the grid). There the tile test alone implies the interior
//     test: thetaIndex in [2, nTheta-4] means RN(theta * invH) in [2, nTheta-3), hence
//     1.5 h <= theta < (lastRow - 0.5) h for either lastRow, and likewise 1.5 h <= phi < 2 pi
//     (NaN / inf convert to indices that fail the test). Lanes that pass run the same fast-path
//     arithmetic as before, all others the general path: no result changes, five instructions per
//     sample disappear.
template <int KIND, bool SAFE>
__device__ __forceinline__ PendingSample sampleIssueTiled(const SamplerRegs& g, const SamplerConsts* __restrict__ consts,
                                                          const float* __restrict__ field, float phiRaw, float thetaRaw,
                                                          unsigned tile)
{
    const float phi = (KIND == kVPhi) ? __fadd_rn(phiRaw, g.halfH) : phiRaw;
    const float theta = (KIND == kVTheta) ? __fsub_rn(thetaRaw, g.h) : __fsub_rn(thetaRaw, g.halfH);
    const float normedPhi = __fmul_rn(phi, g.invH);
    const float normedTheta = __fmul_rn(theta, g.invH);
    const int phiIndex = (int)floorf(normedPhi);
    const int thetaIndex = (int)floorf(normedTheta);
    const int tr = thetaIndex - g.tileRow0;
    int tc;
    bool ok;
    if (SAFE) {
        tc = phiIndex - g.tileCol0;
        ok = (unsigned)tr < (unsigned)(kTileH - 1) && (unsigned)tc < (unsigned)(kTileW - 1);
    } else {
        const unsigned thetaSpan = (KIND == kVTheta) ? g.thetaSpanVTheta : g.thetaSpanCentred;
        const bool interior = (__float_as_uint(theta) - g.thetaLoBits) < thetaSpan
                           && (__float_as_uint(phi) - g.phiLoBits) < g.phiSpan;
        tc = (phiIndex - g.tileCol0) & g.mask;
        ok = interior && (unsigned)tr < (unsigned)(kTileH - 1) && (unsigned)tc < (unsigned)(kTileW - 1);
    }
    PendingSample p;
    if (ok) {
        p.alphaPhi = __fsub_rn(normedPhi, (float)phiIndex);
        p.alphaTheta = __fsub_rn(normedTheta, (float)thetaIndex);
        loadTileCell(tile + 4u * (unsigned)(tr * kTileStride + tc), p);
    } else {
        const float v = sampleGeneral<KIND>(consts, field, phiRaw, thetaRaw);
        p.v00 = p.v01 = p.v10 = p.v11 = v;
        p.alphaPhi = p.alphaTheta = 0.0f;
    }
    return p;
}

// Branch-free form of the SAFE tile sample (interior blocks of the advection kernel): the fast-path
// arithmetic is executed unconditionally, the tile address of a lane that fails the tile test is
// clamped to a valid one, and the failure is recorded in the sticky flag `bad`. The caller discards
// everything a flagged lane computed and redoes that backtrace with sampleIssueTiled<KIND, false>,
// so the values of unflagged lanes are exactly those of the branching form; what goes away is
// the divergence bookkeeping around every sample (two branches, BSSY, BSYNC) and, with the call
// to the out-of-line general path gone from the hot loop, the register spills around it.
// CHECK = false: the sample sits at a node of one of the block's own cells (first stage of a backtrace). Its
// cell index is within one of the node's own (the coordinate is a product of an integer or half-integer and h,
// rounded a few times: |normed - exact| < 0.01 up to index 16384), i.e. at least one row and three columns inside
// the tile's halo, so neither the test nor the clamp is needed.
template <int KIND, bool CHECK = true>
__device__ __forceinline__ PendingSample sampleIssueFast(const SamplerRegs& g, float phiRaw, float thetaRaw,
                                                         unsigned tile, bool& bad)
{
    const float phi = (KIND == kVPhi) ? __fadd_rn(phiRaw, g.halfH) : phiRaw;
    const float theta = (KIND == kVTheta) ? __fsub_rn(thetaRaw, g.h) : __fsub_rn(thetaRaw, g.halfH);
    const float normedPhi = __fmul_rn(phi, g.invH);
    const float normedTheta = __fmul_rn(theta, g.invH);
    const int phiIndex = (int)floorf(normedPhi);
    const int thetaIndex = (int)floorf(normedTheta);
    const int tr = thetaIndex - g.tileRow0;
    const int tc = phiIndex - g.tileCol0;
    if (CHECK) bad = bad || (unsigned)tr >= (unsigned)(kTileH - 1) || (unsigned)tc >= (unsigned)(kTileW - 1);
    // valid lanes: tr * kTileStride + tc <= (kTileH - 2) * kTileStride + kTileW - 2, never altered by the clamp
    const unsigned cellIndex = CHECK ? min((unsigned)(tr * kTileStride + tc), (unsigned)((kTileH - 2) * kTileStride + kTileW - 2))
                                     : (unsigned)(tr * kTileStride + tc);
    PendingSample p;
    p.alphaPhi = __fsub_rn(normedPhi, (float)phiIndex);
    p.alphaTheta = __fsub_rn(normedTheta, (float)thetaIndex);
    loadTileCell(tile + 4u * cellIndex, p);
    return p;
}

template <int KIND>
__device__ __forceinline__ PendingSample sampleIssue(const SamplerRegs& g, const SamplerConsts* __restrict__ consts,
                                                     const float* __restrict__ field, float phiRaw, float thetaRaw)
{
    const float phi = (KIND == kVPhi) ? __fadd_rn(phiRaw, g.halfH) : phiRaw;
    const float theta = (KIND == kVTheta) ? __fsub_rn(thetaRaw, g.h) : __fsub_rn(thetaRaw, g.halfH);
    const unsigned thetaSpan = (KIND == kVTheta) ? g.thetaSpanVTheta : g.thetaSpanCentred;
    const bool interior = (__float_as_uint(theta) - g.thetaLoBits) < thetaSpan
                       && (__float_as_uint(phi) - g.phiLoBits) < g.phiSpan;
    PendingSample p;
    if (interior) {
        const float normedPhi = __fmul_rn(phi, g.invH);
        const float normedTheta = __fmul_rn(theta, g.invH);
        const int phiIndex = (int)floorf(normedPhi);
        const int thetaIndex = (int)floorf(normedTheta);
        p.alphaPhi = __fsub_rn(normedPhi, (float)phiIndex);
        p.alphaTheta = __fsub_rn(normedTheta, (float)thetaIndex);
        const int c0 = phiIndex & g.mask;
        const int c1 = (c0 + 1) & g.mask;
        const int lo0 = thetaIndex * g.N + c0;
        const int lo1 = thetaIndex * g.N + c1;
        p.v00 = __ldg(field + lo0);
        p.v01 = __ldg(field + lo1);
        p.v10 = __ldg(field + (lo0 + g.N));
        p.v11 = __ldg(field + (lo1 + g.N));
    } else {
        const float v = sampleGeneral<KIND>(consts, field, phiRaw, thetaRaw);
        p.v00 = p.v01 = p.v10 = p.v11 = v;
        p.alphaPhi = p.alphaTheta = 0.0f;
    }
    return p;
}

__device__ __forceinline__ float sampleFinish(const PendingSample& p)
{
    const float omPhi = __fsub_rn(1.0f, p.alphaPhi);
    const float omTheta = __fsub_rn(1.0f, p.alphaTheta);
    const float lowerBelt = lerpFast(p.v00, p.v01, p.alphaPhi, omPhi);
    const float higherBelt = lerpFast(p.v10, p.v11, p.alphaPhi, omPhi);
    return lerpFast(lowerBelt, higherBelt, p.alphaTheta, omTheta);
}

template <int KIND>
__device__ __forceinline__ float sample(const SamplerRegs& g, const SamplerConsts* __restrict__ consts,
                                        const float* __restrict__ field, float phiRaw, float thetaRaw)
{
    return sampleFinish(sampleIssue<KIND>(g, consts, field, phiRaw, thetaRaw));
}

} // namespace kb
