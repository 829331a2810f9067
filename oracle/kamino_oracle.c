/*
 * kamino_oracle.c -- TEST INFRASTRUCTURE ONLY. Not part of the product.
 *
 * A CPU restatement of the per-timestep path of KaminoGPU (KaminoSolver::stepForward,
 * /root/reference/KaminoGPU/kernel/KaminoSolver.cu:197-221) used as (1) the checker in
 * tests/, __graft_entry__.smoke() and (2) the "port" CPU baseline in bench.py. Nothing
 * under kaminogpu_b200/ may link, import or call this file.
 *
 * Parity status: PINNED. The restatement is checked in tests/test_oracle_golden.py
 * against raw state dumps produced by the reference's own CUDA build
 * (oracle/_ref/kamino_ref, built from the sources under /root/reference by
 * oracle/ref_harness/Makefile and run on a B200); the dumps are committed under
 * tests/golden/ together with the script that made them.
 *
 * Arithmetic contract. The reference stores fp32 but evaluates many sub-expressions in
 * fp64 because its constants are double literals. Every
 * expression below is written with explicit casts so that the evaluation type of each
 * operation is visible; the file must be compiled with -ffp-contract=off. Where nvcc's
 * default -fmad=true contracts an fp32 a*b+c in the reference, fmaf() is used here.
 * sinf/cosf/acosf come from the host libm and differ from CUDA's by an ulp or so, so
 * field values agree with the GPU to rounding level, not bit for bit; the index / pole
 * predicate path (ko_locate) uses IEEE operations only and is bit-exact.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference/KaminoGPU/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define KO_PI  3.14159265358979323846   /* include/KaminoHeader.cuh:26 */
#define KO_2PI 6.28318530717958647692   /* include/KaminoHeader.cuh:27 */

/* sampler kinds: stagger offsets from include/KaminoHeader.cuh:30-41 */
enum { KO_VPHI = 0, KO_VTHETA = 1, KO_CENTERED = 2 };

typedef struct {
    int nTheta;      /* rows of u_phi / density / pressure; u_theta has nTheta-1 */
    int nPhi;        /* 2 * nTheta */
    float radius;
    float dt;        /* timeStepGlobal: the kernels ignore stepForward's argument */
    float gridLen;   /* (float)(pi / nTheta), kernel/KaminoCore.cu:849 */
} ko_params;

typedef struct {
    int phiIndex;      /* before the % nPhi */
    int thetaIndex;
    float alphaPhi;
    float alphaTheta;  /* before the optional halving */
    float phi;         /* validated coordinates */
    float theta;
    int flipped;       /* 1 when validateCoord returned -1 */
    int poleBranch;    /* 1 when the single-row (pole) interpolation is taken */
} ko_location;

static const double ko_off_phi[3]   = { -0.5, 0.0, 0.0 };
static const double ko_off_theta[3] = {  0.5, 1.0, 0.5 };

int ko_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* kernel/KaminoCore.cu:11-29 */
static float ko_validate_coord(float *phi, float *theta)
{
    float ret = 1.0f;
    float th = *theta, ph = *phi;
    int k = (int)floorf((float)((double)th / KO_2PI));
    th = (float)((double)th - (double)k * KO_2PI);
    if ((double)th > KO_PI) {
        th = (float)(KO_2PI - (double)th);
        ph = (float)((double)ph + KO_PI);
        ret = -ret;
    }
    if (th < 0.0f) {
        th = -th;
        ph = (float)((double)ph + KO_PI);
        ret = -ret;
    }
    k = (int)floorf((float)((double)ph / KO_2PI));
    ph = (float)((double)ph - (double)k * KO_2PI);
    *phi = ph;
    *theta = th;
    return ret;
}

/* kernel/KaminoCore.cu:31-34 */
static float ko_lerp(float from, float to, float alpha)
{
    float at = alpha * to;
    return (float)fma(1.0 - (double)alpha, (double)from, (double)at);
}

/* index / predicate part of the three samplers, kernel/KaminoCore.cu:36-53, 86-103, 136-153 */
void ko_locate(const ko_params *p, int kind, float phiRaw, float thetaRaw, ko_location *loc)
{
    float h = p->gridLen;
    float phi = (float)((double)phiRaw - (double)h * ko_off_phi[kind]);
    float theta = (float)((double)thetaRaw - (double)h * ko_off_theta[kind]);
    float invGridLen = (float)(1.0 / (double)h);
    float flip = ko_validate_coord(&phi, &theta);
    float normedPhi = phi * invGridLen;
    float normedTheta = theta * invGridLen;
    int phiIndex = (int)floorf(normedPhi);
    int thetaIndex = (int)floorf(normedTheta);
    int poleRow = (kind == KO_VTHETA) ? p->nTheta - 2 : p->nTheta - 1;
    loc->phiIndex = phiIndex;
    loc->thetaIndex = thetaIndex;
    loc->alphaPhi = normedPhi - (float)phiIndex;
    loc->alphaTheta = normedTheta - (float)thetaIndex;
    loc->phi = phi;
    loc->theta = theta;
    loc->flipped = (flip == -1.0f);
    loc->poleBranch = ((thetaIndex == 0 && flip == -1.0f) || thetaIndex == poleRow);
}

/* rows are clamped into the array: the reference reads out of bounds there
 * (kernel/KaminoCore.cu:121-131 when theta-CFL > 1 near the south pole; undefined). */
static inline int ko_clamp_row(int r, int rows)
{
    return r < 0 ? 0 : (r >= rows ? rows - 1 : r);
}

/* kernel/KaminoCore.cu:36-84 (sampleVPhi), 86-134 (sampleVTheta), 136-184 (sampleCentered) */
float ko_sample(const ko_params *p, int kind, const float *field, float phiRaw, float thetaRaw)
{
    ko_location loc;
    ko_locate(p, kind, phiRaw, thetaRaw, &loc);
    const int N = p->nPhi;
    const int rows = (kind == KO_VTHETA) ? p->nTheta - 1 : p->nTheta;
    /* size_t modulo of a non-negative int; nPhi is a power of two */
    int phiLower = (int)(((uint64_t)(int64_t)loc.phiIndex) % (uint64_t)N);
    int phiHigher = (phiLower + 1) % N;
    if (loc.poleBranch) {
        int row = ko_clamp_row(loc.thetaIndex, rows);
        const float *r = field + (size_t)row * N;
        float lowerBelt = ko_lerp(r[phiLower], r[phiHigher], loc.alphaPhi);
        int oppLower = (phiLower + N / 2) % N;
        int oppHigher = (oppLower + 1) % N;
        float higherBelt = ko_lerp(r[oppLower], r[oppHigher], loc.alphaPhi);
        float alphaTheta = loc.alphaTheta;
        if (kind != KO_VPHI)                       /* :115, :165 -- not in sampleVPhi (:66) */
            alphaTheta = (float)(0.5 * (double)alphaTheta);
        return ko_lerp(lowerBelt, higherBelt, alphaTheta);
    } else {
        int r0 = ko_clamp_row(loc.thetaIndex, rows);
        int r1 = ko_clamp_row(loc.thetaIndex + 1, rows);
        const float *lo = field + (size_t)r0 * N;
        const float *hi = field + (size_t)r1 * N;
        float lowerBelt = ko_lerp(lo[phiLower], lo[phiHigher], loc.alphaPhi);
        float higherBelt = ko_lerp(hi[phiLower], hi[phiHigher], loc.alphaPhi);
        return ko_lerp(lowerBelt, higherBelt, loc.alphaTheta);
    }
}

/* kernel/KaminoCore.cu:186-229, 231-274, 276-319: one semi-Lagrangian backtrace.
 * kind selects the node position; `src` is the field sampled at the backtraced point. */
static float ko_backtrace(const ko_params *p, int kind, const float *velPhi, const float *velTheta,
                          const float *src, int i, int j)
{
    float h = p->gridLen;
    float gPhi = (float)(((double)(float)i + ko_off_phi[kind]) * (double)h);
    float gTheta = (float)(((double)(float)j + ko_off_theta[kind]) * (double)h);

    float guPhi = ko_sample(p, KO_VPHI, velPhi, gPhi, gTheta);
    float guTheta = ko_sample(p, KO_VTHETA, velTheta, gPhi, gTheta);

    float latRadius = p->radius * sinf(gTheta);
    float cofPhi = p->dt / latRadius;
    float cofTheta = p->dt / p->radius;

    float deltaPhi = guPhi * cofPhi;
    float deltaTheta = guTheta * cofTheta;

    /* RUNGE_KUTTA is defined, include/KaminoHeader.cuh:63 */
    float midPhi = (float)((double)gPhi - 0.5 * (double)deltaPhi);
    float midTheta = (float)((double)gTheta - 0.5 * (double)deltaTheta);
    float muPhi = ko_sample(p, KO_VPHI, velPhi, midPhi, midTheta);
    float muTheta = ko_sample(p, KO_VTHETA, velTheta, midPhi, midTheta);
    float averuPhi = (float)(0.5 * (double)(muPhi + guPhi));
    float averuTheta = (float)(0.5 * (double)(muTheta + guTheta));

    /* gPhi - averuPhi * cofPhi: contracted to one FFMA by nvcc's default -fmad=true */
    float pPhi = fmaf(-averuPhi, cofPhi, gPhi);
    float pTheta = fmaf(-averuTheta, cofTheta, gTheta);

    return ko_sample(p, kind, src, pPhi, pTheta);
}

/* kernel/KaminoCore.cu:321-342 */
static void ko_advect_particle(const ko_params *p, const float *velPhi, const float *velTheta,
                               const float *in, float *out)
{
    float posPhi = in[0], posTheta = in[1];
    float uPhi = ko_sample(p, KO_VPHI, velPhi, posPhi, posTheta);
    float uTheta = ko_sample(p, KO_VTHETA, velTheta, posPhi, posTheta);
    float latRadius = p->radius * sinf(posTheta);
    float cofPhi = p->dt / latRadius;
    float cofTheta = p->dt / p->radius;
    float updatedTheta = fmaf(uTheta, cofTheta, posTheta);
    float updatedPhi = posPhi;
    if (latRadius > 1e-7f)
        updatedPhi = fmaf(uPhi, cofPhi, posPhi);
    ko_validate_coord(&updatedPhi, &updatedTheta);
    out[0] = updatedPhi;
    out[1] = updatedTheta;
}

/* kernel/KaminoCore.cu:344-384. Outputs are separate arrays; all four advections read the
 * pre-advection velocity. */
void ko_advection(const ko_params *p, const float *velPhi, const float *velTheta, const float *density,
                  const float *particles, long numParticles,
                  float *velPhiOut, float *velThetaOut, float *densityOut, float *particlesOut)
{
    const int N = p->nPhi, nT = p->nTheta;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nT; ++j) {
        for (int i = 0; i < N; ++i)
            velPhiOut[(size_t)j * N + i] = ko_backtrace(p, KO_VPHI, velPhi, velTheta, velPhi, i, j);
        if (j < nT - 1)
            for (int i = 0; i < N; ++i)
                velThetaOut[(size_t)j * N + i] = ko_backtrace(p, KO_VTHETA, velPhi, velTheta, velTheta, i, j);
        if (density)
            for (int i = 0; i < N; ++i)
                densityOut[(size_t)j * N + i] = ko_backtrace(p, KO_CENTERED, velPhi, velTheta, density, i, j);
    }
    if (particles) {
#pragma omp parallel for schedule(static)
        for (long k = 0; k < numParticles; ++k)
            ko_advect_particle(p, velPhi, velTheta, particles + 2 * k, particlesOut + 2 * k);
    }
}

/* kernel/KaminoCore.cu:386-407. The reference loops forever on inf/NaN; this one is bounded. */
static float ko_root3_pos(float x)
{
    float s = 1.0f;
    int guard = 0;
    while ((double)x < 1.0 && guard++ < 200) { x = (float)((double)x * 8.0); s = (float)((double)s * 0.5); }
    guard = 0;
    while ((double)x > 8.0 && guard++ < 200) { x = (float)((double)x * 0.125); s = (float)((double)s * 2.0); }
    float r = 1.5f;
    for (int it = 0; it < 6; ++it) {
        float t = r - x / (r * r);
        r = (float)((double)r - (1.0 / 3.0) * (double)t);
    }
    return r * s;
}

/* kernel/KaminoCore.cu:409-417 */
static float ko_root3(double x)
{
    if (x > 0) return ko_root3_pos((float)x);
    else if (x < 0) return -ko_root3_pos((float)(-x));
    else return 0.0f;
}

/* kernel/KaminoCore.cu:421-454 with eps = 1e-7f (:419) */
static float ko_solve_cubic(float a, float b, float c)
{
    float a2 = a * a;
    float q = (float)((double)fmaf(-3.0f, b, a2) / 9.0);
    float r = (float)(((double)a * (2.0 * (double)a2 - 9.0 * (double)b) + 27.0 * (double)c) / 54.0);
    float r2 = r * r;
    float q3 = q * q * q;
    if (r2 <= (q3 + 1e-7f)) {
        double t = (double)(r / sqrtf(q3));
        if (t < -1) t = -1;
        if (t > 1) t = 1;
        t = (double)acosf((float)t);
        a = (float)((double)a / 3.0);
        q = (float)(-2.0 * (double)sqrtf(q));
        return fmaf(q, cosf((float)(t / 3.0)), -a);
    } else {
        float A = -ko_root3((double)(fabsf(r) + sqrtf(r2 - q3)));
        if (r < 0) A = -A;
        float B = (A == 0) ? 0.0f : q / A;
        a = (float)((double)a / 3.0);
        return (A + B) - a;
    }
}

/* kernel/KaminoCore.cu:457-549: fill (cell centres) then re-average to the faces */
void ko_geometric(const ko_params *p, const float *velPhi, const float *velTheta,
                  float *velPhiOut, float *velThetaOut, float *scratchU, float *scratchV)
{
    const int N = p->nPhi, nT = p->nTheta;
    const float h = p->gridLen;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nT; ++j) {
        float gTheta = (float)(((double)(float)j + 0.5) * (double)h);
        float G = p->dt * cosf(gTheta) / (p->radius * sinf(gTheta));
        for (int i = 0; i < N; ++i) {
            int right = (i + 1) % N;
            float uPrev = (float)(0.5 * (double)(velPhi[(size_t)j * N + i] + velPhi[(size_t)j * N + right]));
            float vPrev;
            if (j == 0) {
                int opp = (i + N / 2) % N;
                vPrev = (float)(0.75 * (double)velTheta[i] + 0.25 * (double)velTheta[opp]);
            } else if (j == nT - 1) {
                int opp = (i + N / 2) % N;
                vPrev = (float)(0.75 * (double)velTheta[(size_t)(j - 1) * N + i]
                              + 0.25 * (double)velTheta[(size_t)(j - 1) * N + opp]);
            } else {
                vPrev = (float)(0.5 * (double)(velTheta[(size_t)(j - 1) * N + i] + velTheta[(size_t)j * N + i]));
            }
            float uNext;
            if (fabsf(G) > 1e-7f) {
                float cof = G * G;
                float B = (float)(((double)(G * vPrev) + 1.0) / (double)cof);
                float C = -uPrev / cof;
                uNext = ko_solve_cubic(0.0f, B, C);
            } else {
                uNext = uPrev;
            }
            float vNext = fmaf(G * uNext, uNext, vPrev);
            scratchU[(size_t)j * N + i] = uNext;
            scratchV[(size_t)j * N + i] = vNext;
        }
    }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nT; ++j) {
        for (int i = 0; i < N; ++i) {
            int left = (i == 0) ? N - 1 : i - 1;
            velPhiOut[(size_t)j * N + i] =
                (float)(0.5 * (double)(scratchU[(size_t)j * N + left] + scratchU[(size_t)j * N + i]));
            if (j < nT - 1)
                velThetaOut[(size_t)j * N + i] =
                    (float)(0.5 * (double)(scratchV[(size_t)j * N + i] + scratchV[(size_t)(j + 1) * N + i]));
        }
    }
}

/* kernel/KaminoSolver.cu:117-163: tridiagonal coefficients of wavenumber n = nIdx - nPhi/2 */
void ko_abc_row(const ko_params *p, int n, float *a, float *b, float *c)
{
    const int nT = p->nTheta;
    const float h = p->gridLen;
    for (int i = 0; i < nT; ++i) {
        float thetaI = (float)(((double)i + 0.5) * (double)h);
        float cosT = cosf(thetaI), sinT = sinf(thetaI);
        float valB = (float)(-2.0 / (double)(h * h) - (double)((float)(n * n) / (sinT * sinT)));
        float valA = (float)(1.0 / (double)(h * h) - (double)cosT / 2.0 / (double)h / (double)sinT);
        float valC = (float)(1.0 / (double)(h * h) + (double)cosT / 2.0 / (double)h / (double)sinT);
        if (n != 0) {
            if (i == 0) { valB += valA; valA = 0.0f; }
            if (i == nT - 1) { valB += valC; valC = 0.0f; }
        } else {
            valA = 0.0f; valB = 1.0f; valC = 0.0f;
        }
        a[i] = valA; b[i] = valB; c[i] = valC;
    }
}

/* kernel/tdm.cu:3-96: cyclic reduction in the reference's elimination order, fp32.
 * a, b, c, d are overwritten (they are shared-memory copies in the reference). */
void ko_cyclic_reduction(int n, float *a, float *b, float *c, float *d, float *x)
{
    int iteration = 0;
    while ((1 << (iteration + 1)) < n) ++iteration;      /* log2(n / 2) */
    int stride = 1;
    int numThreads = n / 2;
    for (int lvl = 0; lvl < iteration; ++lvl) {
        stride *= 2;
        int delta = stride / 2;
        for (int t = 0; t < numThreads; ++t) {
            int i = stride * t + stride - 1;
            int iLeft = i - delta;
            int iRight = i + delta;
            if (iRight >= n) iRight = n - 1;
            float tmp1 = a[i] / b[iLeft];
            float tmp2 = c[i] / b[iRight];
            float bi = fmaf(-a[iRight], tmp2, fmaf(-c[iLeft], tmp1, b[i]));
            float di = fmaf(-d[iRight], tmp2, fmaf(-d[iLeft], tmp1, d[i]));
            float ai = -a[iLeft] * tmp1;
            float ci = -c[iRight] * tmp2;
            b[i] = bi; d[i] = di; a[i] = ai; c[i] = ci;
        }
        numThreads /= 2;
    }
    {
        int addr1 = stride - 1, addr2 = 2 * stride - 1;
        float tmp3 = fmaf(b[addr2], b[addr1], -(c[addr1] * a[addr2]));
        x[addr1] = fmaf(b[addr2], d[addr1], -(c[addr1] * d[addr2])) / tmp3;
        x[addr2] = fmaf(d[addr2], b[addr1], -(d[addr1] * a[addr2])) / tmp3;
    }
    numThreads = 2;
    for (int lvl = 0; lvl < iteration; ++lvl) {
        int delta = stride / 2;
        for (int t = 0; t < numThreads; ++t) {
            int i = stride * t + stride / 2 - 1;
            if (i == delta - 1)
                x[i] = fmaf(-c[i], x[i + delta], d[i]) / b[i];
            else
                x[i] = fmaf(-c[i], x[i + delta], fmaf(-a[i], x[i - delta], d[i])) / b[i];
        }
        stride /= 2;
        numThreads *= 2;
    }
}

/* in-place iterative radix-2 complex FFT in fp64; sign = -1 forward, +1 inverse, unnormalised.
 * Stands in for cuFFT (closed source; kernel/KaminoSolver.cu:61-66, KaminoCore.cu:763,812):
 * the transform is the mathematically defined DFT, evaluated here more accurately than
 * cuFFT's fp32. */
static void ko_fft(double *re, double *im, int n, int sign)
{
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        double ang = sign * KO_2PI / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < len / 2; ++k) {
                double wr = cos(ang * k), wi = sin(ang * k);
                double ur = re[i + k], ui = im[i + k];
                double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                re[i + k] = ur + vr; im[i + k] = ui + vi;
                re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
            }
        }
    }
}

/* kernel/KaminoCore.cu:587-638 */
void ko_divergence(const ko_params *p, const float *velPhi, const float *velTheta, float *div)
{
    const int N = p->nPhi, nT = p->nTheta;
    const float h = p->gridLen;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nT; ++j) {
        float coord = (float)(((double)(float)j + 0.5) * (double)h);
        float halfStep = (float)(0.5 * (double)h);
        float thetaSouth = coord + halfStep;
        float thetaNorth = coord - halfStep;
        float invGridSine = (float)(1.0 / (double)sinf(coord));
        float sinNorth = sinf(thetaNorth);
        float sinSouth = sinf(thetaSouth);
        float factor = invGridSine / h;
        for (int i = 0; i < N; ++i) {
            int east = (i + 1) % N;
            float uWest = velPhi[(size_t)j * N + i];
            float uEast = velPhi[(size_t)j * N + east];
            float vNorth = 0.0f, vSouth = 0.0f;
            if (j != 0) vNorth = velTheta[(size_t)(j - 1) * N + i];
            if (j != nT - 1) vSouth = velTheta[(size_t)j * N + i];
            float termTheta = factor * fmaf(vSouth, sinSouth, -(vNorth * sinNorth));
            div[(size_t)j * N + i] = fmaf(factor, uEast - uWest, termTheta);
        }
    }
}

/* kernel/KaminoCore.cu:749-842 (projection) with its kernels :587-747 and tdm.cu.
 * pressure (nT x N) receives the reference's pressure->gpuThisStep. Velocity is updated in place. */
void ko_projection(const ko_params *p, float *velPhi, float *velTheta, float *pressure)
{
    const int N = p->nPhi, nT = p->nTheta;
    const float h = p->gridLen;
    float *div = (float *)malloc(sizeof(float) * (size_t)N * nT);
    /* [nIdx][theta] like gpuFReal / gpuFImag / gpuUReal / gpuUImag */
    float *FRe = (float *)malloc(sizeof(float) * (size_t)N * nT);
    float *FIm = (float *)malloc(sizeof(float) * (size_t)N * nT);
    float *URe = (float *)malloc(sizeof(float) * (size_t)N * nT);
    float *UIm = (float *)malloc(sizeof(float) * (size_t)N * nT);

    ko_divergence(p, velPhi, velTheta, div);

    /* cuFFT inverse (unnormalised) + shiftFKernel (:640-656) */
#pragma omp parallel
    {
        double *re = (double *)malloc(sizeof(double) * N), *im = (double *)malloc(sizeof(double) * N);
#pragma omp for schedule(static)
        for (int j = 0; j < nT; ++j) {
            for (int i = 0; i < N; ++i) { re[i] = (double)div[(size_t)j * N + i]; im[i] = 0.0; }
            ko_fft(re, im, N, +1);
            for (int nIdx = 0; nIdx < N; ++nIdx) {
                int fftIndex = N / 2 - nIdx;
                if (fftIndex < 0) fftIndex += N;
                FRe[(size_t)nIdx * nT + j] = (float)re[fftIndex] / (float)N;
                FIm[(size_t)nIdx * nT + j] = (float)im[fftIndex] / (float)N;
            }
        }
        free(re); free(im);
    }

    /* crKernel twice (KaminoCore.cu:779-792), coefficients from precomputeABCKernel */
#pragma omp parallel
    {
        float *a = (float *)malloc(sizeof(float) * nT * 8);
        float *b = a + nT, *c = b + nT, *d = c + nT, *x = d + nT;
        float *a0 = x + nT, *b0 = a0 + nT, *c0 = b0 + nT;
#pragma omp for schedule(dynamic, 4)
        for (int nIdx = 0; nIdx < N; ++nIdx) {
            ko_abc_row(p, nIdx - N / 2, a0, b0, c0);
            for (int part = 0; part < 2; ++part) {
                const float *F = part ? FIm : FRe;
                float *U = part ? UIm : URe;
                memcpy(a, a0, sizeof(float) * nT); memcpy(b, b0, sizeof(float) * nT);
                memcpy(c, c0, sizeof(float) * nT);
                memcpy(d, F + (size_t)nIdx * nT, sizeof(float) * nT);
                ko_cyclic_reduction(nT, a, b, c, d, x);
                memcpy(U + (size_t)nIdx * nT, x, sizeof(float) * nT);
            }
        }
        free(a);
    }

    /* copy2UFourier, cacheZeroComponents, cuFFT forward, shiftUKernel (:658-703) */
#pragma omp parallel
    {
        double *re = (double *)malloc(sizeof(double) * N), *im = (double *)malloc(sizeof(double) * N);
#pragma omp for schedule(static)
        for (int j = 0; j < nT; ++j) {
            for (int nIdx = 0; nIdx < N; ++nIdx) {
                re[nIdx] = (double)URe[(size_t)nIdx * nT + j];
                im[nIdx] = (double)UIm[(size_t)nIdx * nT + j];
            }
            float zero = URe[(size_t)(N / 2) * nT + j];
            ko_fft(re, im, N, -1);
            for (int i = 0; i < N; ++i) {
                int fftIndex = (i != 0) ? N - i : 0;
                float y = (float)re[fftIndex];
                pressure[(size_t)j * N + i] = (i % 2 == 0) ? (y - zero) : (-y - zero);
            }
        }
        free(re); free(im);
    }

    /* applyPressureTheta (:705-722), applyPressurePhi (:724-747) */
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nT; ++j) {
        float thetaBelt = (float)(((double)j + 0.5) * (double)h);
        float denomPhi = -h * sinf(thetaBelt);
        for (int i = 0; i < N; ++i) {
            if (j < nT - 1) {
                float dV = (pressure[(size_t)(j + 1) * N + i] - pressure[(size_t)j * N + i]) / (-h);
                velTheta[(size_t)j * N + i] = velTheta[(size_t)j * N + i] + dV;
            }
            int west = (i == 0) ? N - 1 : i - 1;
            float dU = (pressure[(size_t)j * N + i] - pressure[(size_t)j * N + west]) / denomPhi;
            velPhi[(size_t)j * N + i] = velPhi[(size_t)j * N + i] + dU;
        }
    }
    free(div); free(FRe); free(FIm); free(URe); free(UIm);
}

/* kernel/KaminoSolver.cu:197-221: one step = advection -> geometric -> projection.
 * Fields are updated in place; `phase` (1..3) stops after that phase (0 = full step). */
void ko_step(const ko_params *p, float *velPhi, float *velTheta, float *density, float *pressure,
             float *particles, long numParticles, int phase)
{
    const size_t cells = (size_t)p->nPhi * p->nTheta;
    const size_t cellsT = (size_t)p->nPhi * (p->nTheta - 1);
    float *uN = (float *)malloc(sizeof(float) * cells);
    float *vN = (float *)malloc(sizeof(float) * cells);
    float *rN = density ? (float *)malloc(sizeof(float) * cells) : NULL;
    float *pN = particles ? (float *)malloc(sizeof(float) * 2 * (size_t)numParticles) : NULL;
    float *sU = (float *)malloc(sizeof(float) * cells);
    float *sV = (float *)malloc(sizeof(float) * cells);

    ko_advection(p, velPhi, velTheta, density, particles, numParticles, uN, vN, rN, pN);
    memcpy(velPhi, uN, sizeof(float) * cells);
    memcpy(velTheta, vN, sizeof(float) * cellsT);
    if (density) memcpy(density, rN, sizeof(float) * cells);
    if (particles) memcpy(particles, pN, sizeof(float) * 2 * (size_t)numParticles);
    if (phase != 1) {
        ko_geometric(p, velPhi, velTheta, uN, vN, sU, sV);
        memcpy(velPhi, uN, sizeof(float) * cells);
        memcpy(velTheta, vN, sizeof(float) * cellsT);
        if (phase != 2)
            ko_projection(p, velPhi, velTheta, pressure);
    }
    free(uN); free(vN); free(rN); free(pN); free(sU); free(sV);
}

/* ---- host-side initialisers of the reference (needed for identical initial fields) ---- */

/* kernel/KaminoInitializer.cu:127-134 (vec2 holds doubles, include/vectorUtil.cuh:13) */
static float ko_hash_rand(double ax, double ay)
{
    float dotProd = (float)(ax * 12.9898 + ay * 4.1414);
    float val = (float)sin((double)dotProd * 43758.5453);
    return val - floorf(val);
}

/* kernel/KaminoInitializer.cu:104-107 */
static float ko_lerp_host(float from, float to, float alpha)
{
    float at = alpha * to;
    return (float)((1.0 - (double)alpha) * (double)from + (double)at);
}

/* kernel/KaminoInitializer.cu:109-125 */
static float ko_interp_noise(float x, float y)
{
    float intX = floorf(x), fractX = x - intX;
    float intY = floorf(y), fractY = y - intY;
    float v1 = ko_hash_rand((double)intX, (double)intY);
    float v2 = ko_hash_rand((double)(intX + 1), (double)intY);
    float v3 = ko_hash_rand((double)intX, (double)(intY + 1));
    float v4 = ko_hash_rand((double)(intX + 1), (double)(intY + 1));
    float i1 = ko_lerp_host(v1, v2, fractX);
    float i2 = ko_lerp_host(v3, v4, fractX);
    return ko_lerp_host(i1, i2, fractY);
}

/* kernel/KaminoInitializer.cu:87-102 */
static float ko_fbm(float x, float y)
{
    float total = 0.0f;
    const float resolutionX = 0.15f, resolutionY = 0.5f, persistance = 0.5f;
    for (int i = 0; i < 4; ++i) {
        float freq = (float)pow(2.0, (double)i);
        float amp = (float)pow((double)persistance, (double)i);
        total += amp * ko_interp_noise(x * freq / resolutionX, y * freq / resolutionY);
    }
    float a = 1 - persistance;
    return a * total / 2.0f;
}

/* kernel/KaminoInitializer.cu:3-85. The solver's own gridLen is (float)(2pi / nPhi),
 * kernel/KaminoSolver.cu:14. */
void ko_init_velocity(int nTheta, float radius, float *velPhi, float *velTheta)
{
    const int N = 2 * nTheta;
    const float gridLen = (float)(KO_2PI / (double)N);
    const float gain = (float)(4096.0 / (double)N);
    const float rg = radius * gridLen;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < nTheta; ++j) {
        for (int i = 0; i < N; ++i) {
            float ur_x, ul_x;
            if (i == 0) {
                ur_x = gridLen / 2;
                ul_x = (float)(2 * KO_PI - (double)(gridLen / 2));
            } else {
                ur_x = (float)i * gridLen + gridLen / 2;
                ul_x = (float)i * gridLen - gridLen / 2;
            }
            float up_y = (float)(j + 1) * gridLen;
            float lo_y = (float)j * gridLen;
            float noise_ur = ko_fbm(ur_x, up_y), noise_lr = ko_fbm(ur_x, lo_y);
            float noise_ul = ko_fbm(ul_x, up_y), noise_ll = ko_fbm(ul_x, lo_y);
            float dyl = (noise_ur - noise_lr) / rg;
            float dyr = (noise_ul - noise_ll) / rg;
            float avg = (float)((double)(dyl + dyr) / 2.0);
            velPhi[(size_t)j * N + i] = avg * gain;
        }
    }
#pragma omp parallel for schedule(static)
    for (int j = 1; j < nTheta; ++j) {
        for (int i = 0; i < N; ++i) {
            float r_x = (float)(i + 1) * gridLen;
            float l_x = (float)i * gridLen;
            float up_y = (float)j * gridLen + gridLen / 2;
            float lo_y = (float)j * gridLen - gridLen / 2;
            float noise_ur = ko_fbm(r_x, up_y), noise_lr = ko_fbm(r_x, lo_y);
            float noise_ul = ko_fbm(l_x, up_y);
            float noise_ll = ko_fbm(l_x, up_y);          /* the reference's ll_y equals ul_y (:69) */
            float dyu = -1.0f * (noise_ur - noise_ul) / rg;
            float dyd = -1.0f * (noise_lr - noise_ll) / rg;
            float avg = (float)((double)(dyu + dyd) / 2.0);
            velTheta[(size_t)(j - 1) * N + i] = avg * gain;
        }
    }
}

/* kernel/KaminoParticles.cu:20-26 */
long ko_particle_count(int nTheta, float particleDensity)
{
    float linearDensity = sqrtf(particleDensity);
    unsigned int numTheta = (unsigned int)(linearDensity * (float)nTheta);
    unsigned int numPhi = 2 * numTheta;
    return (long)numTheta * (long)numPhi;
}

/* kernel/KaminoParticles.cu:20-62: jittered lattice driven by libc rand() in the
 * reference's call order. Reseeds with srand(1) (= the never-seeded state). */
void ko_seed_particles(int nTheta, float particleDensity, float *coords)
{
    float linearDensity = sqrtf(particleDensity);
    float delta = (float)(KO_PI / (double)nTheta / (double)linearDensity);
    float halfDelta = (float)((double)delta / 2.0);
    unsigned int numTheta = (unsigned int)(linearDensity * (float)nTheta);
    unsigned int numPhi = 2 * numTheta;
    srand(1);
    for (unsigned int i = 0; i < numPhi; ++i) {
        for (unsigned int j = 0; j < numTheta; ++j) {
            float signPhi = (float)rand() / (float)RAND_MAX;
            signPhi = ((double)signPhi >= 0.5) ? 1.0f : -1.0f;
            float signTheta = (float)rand() / (float)RAND_MAX;
            signTheta = ((double)signTheta >= 0.5) ? 1.0f : -1.0f;
            float randPhi = signPhi * halfDelta * (float)rand() / (float)RAND_MAX;
            float randTheta = signTheta * halfDelta * (float)rand() / (float)RAND_MAX;
            float phi = (float)i * delta + randPhi;
            float theta = (float)j * delta + randTheta;
            if (phi < 0.0f) phi = 0.0f;
            if (theta < 0.0f) theta = 0.0f;
            size_t index = (size_t)i * numTheta + j;
            coords[2 * index] = phi;
            coords[2 * index + 1] = theta;
        }
    }
}

/* Deterministic synthetic density of SURVEY.md section 8d (the reference leaves density
 * uninitialised without an image); same formula as oracle/ref_harness/ref_harness.cu. */
void ko_synthetic_density(int nTheta, float *density)
{
    const int N = 2 * nTheta;
    const float gridLen = (float)(KO_PI / (double)nTheta);
    for (int j = 0; j < nTheta; ++j)
        for (int i = 0; i < N; ++i) {
            double phi = (double)i * (double)gridLen;
            double theta = ((double)j + 0.5) * (double)gridLen;
            double st = sin(theta);
            density[(size_t)j * N + i] = (float)(0.5 + 0.5 * sin(4.0 * phi) * st * st);
        }
}
