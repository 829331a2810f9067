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

// kernel/KaminoCore.cu:11-29; returns true when the reference returns -1.
__device__ __forceinline__ bool validateCoord(float& phi, float& theta)
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
    return flipped;
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

template <int KIND>
__device__ __forceinline__ Location locate(const GridParams& g, float phiRaw, float thetaRaw)
{
    Location loc;
    // stagger shift (kernel/KaminoCore.cu:38-39, 88-89, 138-139)
    float phi = (KIND == kVPhi) ? __fadd_rn(phiRaw, g.halfH) : phiRaw;
    float theta = (KIND == kVTheta) ? __fsub_rn(thetaRaw, g.h) : __fsub_rn(thetaRaw, g.halfH);
    loc.flipped = validateCoord(phi, theta);
    float normedPhi = __fmul_rn(phi, g.invH);
    float normedTheta = __fmul_rn(theta, g.invH);
    loc.phiIndex = (int)floorf(normedPhi);
    loc.thetaIndex = (int)floorf(normedTheta);
    loc.alphaPhi = __fsub_rn(normedPhi, (float)loc.phiIndex);
    loc.alphaTheta = __fsub_rn(normedTheta, (float)loc.thetaIndex);
    loc.phi = phi;
    loc.theta = theta;
    return loc;
}

template <int KIND>
__device__ __forceinline__ bool poleBranch(const GridParams& g, const Location& loc)
{
    const int poleRow = (KIND == kVTheta) ? g.nTheta - 2 : g.nTheta - 1;
    return (loc.thetaIndex == 0 && loc.flipped) || loc.thetaIndex == poleRow;
}

// One bilinear sample of `field` (rows x nPhi, dense) at the raw coordinate.
template <int KIND>
__device__ __forceinline__ float sample(const GridParams& g, const float* __restrict__ field,
                                        float phiRaw, float thetaRaw)
{
    const Location loc = locate<KIND>(g, phiRaw, thetaRaw);
    const int N = g.nPhi;
    const int rows = (KIND == kVTheta) ? g.nTheta - 1 : g.nTheta;
    const int phiLower = loc.phiIndex & (N - 1);          // size_t % nPhi, nPhi = 2^k
    const int phiHigher = (phiLower + 1) & (N - 1);
    const bool pole = poleBranch<KIND>(g, loc);

    // Rows outside the array are clamped: the reference reads out of bounds there
    // (undefined; only reachable when the theta-CFL exceeds 1 at the south pole).
    const int rowLo = min(max(loc.thetaIndex, 0), rows - 1);
    int colA, colB, rowHi;
    if (pole) {
        // single-row branch: second belt is the same row at phi + pi
        rowHi = rowLo;
        colA = (phiLower + (N >> 1)) & (N - 1);
        colB = (colA + 1) & (N - 1);
    } else {
        rowHi = min(rowLo + 1, rows - 1);
        colA = phiLower;
        colB = phiHigher;
    }
    const float* lo = field + (size_t)rowLo * N;
    const float* hi = field + (size_t)rowHi * N;
    const float v00 = __ldg(lo + phiLower);
    const float v01 = __ldg(lo + phiHigher);
    const float v10 = __ldg(hi + colA);
    const float v11 = __ldg(hi + colB);

    float alphaTheta = loc.alphaTheta;
    if (pole && KIND != kVPhi) alphaTheta = __fmul_rn(0.5f, alphaTheta);   // :115, :165

    const float omPhi = __fsub_rn(1.0f, loc.alphaPhi);
    const float omTheta = __fsub_rn(1.0f, alphaTheta);
    const bool exact = (__fsub_rn(1.0f, omPhi) == loc.alphaPhi) && (__fsub_rn(1.0f, omTheta) == alphaTheta);
    if (exact) {
        const float lowerBelt = lerpFast(v00, v01, loc.alphaPhi, omPhi);
        const float higherBelt = lerpFast(v10, v11, loc.alphaPhi, omPhi);
        return lerpFast(lowerBelt, higherBelt, alphaTheta, omTheta);
    } else {
        const float lowerBelt = lerpWide(v00, v01, loc.alphaPhi);
        const float higherBelt = lerpWide(v10, v11, loc.alphaPhi);
        return lerpWide(lowerBelt, higherBelt, alphaTheta);
    }
}

} // namespace kb
