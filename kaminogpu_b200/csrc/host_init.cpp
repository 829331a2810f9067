// Host-side initialisers of the drop-in boundary: the FBM "curl noise" initial velocity and
// the rand()-driven particle lattice. They run once, on the CPU, exactly as in the reference
// (KaminoSolver::initialize_velocity, kernel/KaminoInitializer.cu:3-134; KaminoParticles
// constructor, kernel/KaminoParticles.cu:20-62) so that a simulation started through this
// library begins from bit-identical fields. Compile with -ffp-contract=off: the reference's
// host code is plain SSE2 arithmetic without fused multiply-adds.
#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "../../include/kamino_b200.h"

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;

// value noise lattice hash in [0, 1) (kernel/KaminoInitializer.cu:127-134)
float latticeHash(double x, double y)
{
    const float dotProd = (float)(x * 12.9898 + y * 4.1414);
    const float val = (float)std::sin((double)dotProd * 43758.5453);
    return val - std::floor(val);
}

// (1.0 - t) * a + t * b with the reference's mixed types (kernel/KaminoInitializer.cu:104-107)
float mixWide(float a, float b, float t)
{
    const float tb = t * b;
    return (float)((1.0 - (double)t) * (double)a + (double)tb);
}

// bilinear value noise (kernel/KaminoInitializer.cu:109-125)
float valueNoise(float x, float y)
{
    const float x0 = std::floor(x), fx = x - x0;
    const float y0 = std::floor(y), fy = y - y0;
    const float n00 = latticeHash(x0, y0);
    const float n10 = latticeHash(x0 + 1, y0);
    const float n01 = latticeHash(x0, y0 + 1);
    const float n11 = latticeHash(x0 + 1, y0 + 1);
    return mixWide(mixWide(n00, n10, fx), mixWide(n01, n11, fx), fy);
}

// four octaves, persistence 0.5, anisotropic base resolution (kernel/KaminoInitializer.cu:87-102)
float fbm(float x, float y)
{
    const float resX = 0.15f, resY = 0.5f, persistence = 0.5f;
    // freq = (fReal)pow(2.0, octave) and amp = (fReal)pow(persistence, octave) are exact powers of two
    static const float freqs[4] = {1.0f, 2.0f, 4.0f, 8.0f}, amps[4] = {1.0f, 0.5f, 0.25f, 0.125f};
    float total = 0.0f;
    for (int octave = 0; octave < 4; ++octave) {
        const float freq = freqs[octave];
        const float amp = amps[octave];
        total += amp * valueNoise(x * freq / resX, y * freq / resY);
    }
    const float norm = 1 - persistence;
    return norm * total / 2.0f;
}

} // namespace

extern "C" int kamino_init_velocity_host(int nTheta, float radius, float* velPhi, float* velTheta)
{
    return kamino_init_velocity_host_rows(nTheta, radius, 0, nTheta, velPhi, velTheta);
}

extern "C" int kamino_init_velocity_host_rows(int nTheta, float radius, int rowBegin, int rowCount, float* velPhi, float* velTheta)
{
    if (nTheta < 2 || !velPhi || !velTheta || rowBegin < 0 || rowCount < 1 || rowBegin + rowCount > nTheta) return KAMINO_ERR_INVALID;
    const int rowEnd = rowBegin + rowCount;
    velPhi -= (size_t)rowBegin * (2 * (size_t)nTheta);          // global row indexing below
    velTheta -= (size_t)rowBegin * (2 * (size_t)nTheta);
    const int nPhi = 2 * nTheta;
    const float h = (float)(kTwoPi / (double)nPhi);           // the solver's gridLen, KaminoSolver.cu:14
    const float gain = (float)(4096.0 / (double)nPhi);        // KaminoInitializer.cu:9
    const float scale = radius * h;

    // u_phi: theta-difference of the noise, averaged over the two phi sides of the face
    // (KaminoInitializer.cu:11-55; the i = 0 column wraps to 2*pi - h/2). Every cell is a pure
    // function of its position, so the rows are spread over the host cores (the reference's serial
    // double loop takes minutes at 8192 x 16384; the values do not depend on the schedule).
#pragma omp parallel for schedule(dynamic, 8)
    for (int j = rowBegin; j < rowEnd; ++j) {
        const float yUp = (float)(j + 1) * h, yLo = (float)j * h;
        for (int i = 0; i < nPhi; ++i) {
            const float xR = (i == 0) ? h / 2 : (float)i * h + h / 2;
            const float xL = (i == 0) ? (float)(2 * kPi - (double)(h / 2)) : (float)i * h - h / 2;
            const float dR = (fbm(xR, yUp) - fbm(xR, yLo)) / scale;
            const float dL = (fbm(xL, yUp) - fbm(xL, yLo)) / scale;
            velPhi[(size_t)j * nPhi + i] = (float)((double)(dR + dL) / 2.0) * gain;
        }
    }
    // u_theta: minus the phi-difference; the reference's lower-left sample reuses the upper
    // row (KaminoInitializer.cu:69), which is kept
#pragma omp parallel for schedule(dynamic, 8)
    for (int j = rowBegin + 1; j < (rowEnd + 1 < nTheta ? rowEnd + 1 : nTheta); ++j) {   // u_theta row j - 1
        const float yUp = (float)j * h + h / 2, yLo = (float)j * h - h / 2;
        for (int i = 0; i < nPhi; ++i) {
            const float xR = (float)(i + 1) * h, xL = (float)i * h;
            const float upperLeft = fbm(xL, yUp);
            const float dU = -1.0f * (fbm(xR, yUp) - upperLeft) / scale;
            const float dD = -1.0f * (fbm(xR, yLo) - upperLeft) / scale;
            velTheta[(size_t)(j - 1) * nPhi + i] = (float)((double)(dU + dD) / 2.0) * gain;
        }
    }
    return 0;
}

namespace {
// glibc's rand() (random_r, TYPE_3: x_k = x_{k-3} + x_{k-31} mod 2^32, output x_k >> 1) restated locally, in its
// never-seeded (= srand(1)) state: the reference's particle seeding draws from it (kernel/KaminoParticles.cu:39-46),
// and re-seeding or advancing the process-global generator from inside a library would be a side effect on the host
// program. tests/test_capi_host.py checks the sequence against libc's.
struct GlibcRand {
    uint32_t r[34];
    int at = 0;            // ring position of x_{k-34}
    GlibcRand()
    {
        int32_t seedTable[34];
        seedTable[0] = 1;
        for (int i = 1; i < 31; ++i) {
            int64_t v = (16807LL * seedTable[i - 1]) % 2147483647LL;
            if (v < 0) v += 2147483647LL;
            seedTable[i] = (int32_t)v;
        }
        for (int i = 31; i < 34; ++i) seedTable[i] = seedTable[i - 31];
        for (int i = 0; i < 34; ++i) r[i] = (uint32_t)seedTable[i];
        for (int i = 34; i < 344; ++i) step();       // glibc discards the first 310 values
    }
    uint32_t step()
    {
        // r[] is a ring of the last 34 values, r[at] the oldest (x_{k-34}); x_k = x_{k-31} + x_{k-3}
        const uint32_t v = r[(at + 3) % 34] + r[(at + 31) % 34];
        r[at] = v;
        at = (at + 1) % 34;
        return v;
    }
    int next() { return (int)(step() >> 1); }
};

struct LatticeShape { float spacing; unsigned nTheta, nPhi; };

LatticeShape latticeShape(int nTheta, float particleDensity)
{
    const float linear = std::sqrt(particleDensity);                       // KaminoParticles.cu:20
    LatticeShape s;
    s.spacing = (float)(kPi / (double)nTheta / (double)linear);           // :21
    s.nTheta = (unsigned)(linear * (float)nTheta);                        // :24
    s.nPhi = 2 * s.nTheta;                                                // :25
    return s;
}
} // namespace

extern "C" long kamino_particle_count(int nTheta, float particleDensity)
{
    if (nTheta < 1 || !(particleDensity >= 0.f)) return 0;
    const LatticeShape s = latticeShape(nTheta, particleDensity);
    return (long)s.nTheta * (long)s.nPhi;
}

extern "C" int kamino_seed_particles_host(int nTheta, float particleDensity, float* coords)
{
    if (nTheta < 1 || !(particleDensity >= 0.f) || !coords) return KAMINO_ERR_INVALID;
    const LatticeShape s = latticeShape(nTheta, particleDensity);
    const float half = (float)((double)s.spacing / 2.0);
    const float randMax = 2147483647.0f;          // (float)RAND_MAX of glibc
    GlibcRand rng;       // the reference never seeds: glibc's rand() starts in the srand(1) state
    for (unsigned i = 0; i < s.nPhi; ++i) {
        for (unsigned j = 0; j < s.nTheta; ++j) {
            // four draws per particle in this order: sign phi, sign theta, jitter phi, jitter theta (:39-46)
            const float sp = ((double)((float)rng.next() / randMax) >= 0.5) ? 1.0f : -1.0f;
            const float st = ((double)((float)rng.next() / randMax) >= 0.5) ? 1.0f : -1.0f;
            const float jitterPhi = sp * half * (float)rng.next() / randMax;
            const float jitterTheta = st * half * (float)rng.next() / randMax;
            float phi = (float)i * s.spacing + jitterPhi;
            float theta = (float)j * s.spacing + jitterTheta;
            if (phi < 0.0f) phi = 0.0f;
            if (theta < 0.0f) theta = 0.0f;
            const size_t at = (size_t)i * s.nTheta + j;
            coords[2 * at] = phi;
            coords[2 * at + 1] = theta;
        }
    }
    return 0;
}
