// GPU-side initialisers: the FBM "curl noise" initial velocity evaluated on the device, and a
// counter-based particle seeding option.
//
// The drop-in path keeps the HOST initialisers (host_init.cpp), which are bit-identical to the reference's serial
// host loops (kernel/KaminoInitializer.cu:3-134, kernel/KaminoParticles.cu:20-62). These kernels exist for start-up
// at sizes where that loop is the bottleneck (two core-minutes at 8192 x 16384 against milliseconds here):
//  * velocity: the same expressions, operation for operation (explicit round-to-nearest intrinsics so that nothing
//    contracts into an FMA, as in the reference's host build). The one thing a device cannot reproduce bit for bit is
//    the lattice hash's sin(): glibc's double sin is correctly rounded in practice, CUDA's is within 2 ulp; the hash
//    keeps (float)sin(...), so a value differs only when the double results straddle an fp32 rounding boundary
//    (~ 1e-8 of the evaluations). The GPU suite measures the identical fraction against the host initialiser.
//  * particles: the reference draws from libc rand(), which is inherently serial. The option here keeps the lattice
//    (counts, spacing, +-half-spacing jitter, clamp at 0, index order) and takes the four uniforms of particle n from a
//    counter-based generator (splitmix64 of (seed, n)): reproducible for a given seed on any device, in any launch
//    shape -- but NOT the reference's sequence.
#include "kamino_kernels.cuh"

namespace kb {

namespace {

// value noise lattice hash in [0, 1) (kernel/KaminoInitializer.cu:127-134)
__device__ float latticeHash(double x, double y)
{
    const float dotProd = (float)__dadd_rn(__dmul_rn(x, 12.9898), __dmul_rn(y, 4.1414));
    const float val = (float)sin(__dmul_rn((double)dotProd, 43758.5453));
    return __fsub_rn(val, floorf(val));
}

// (1.0 - t) * a + t * b with the reference's mixed types (kernel/KaminoInitializer.cu:104-107)
__device__ float mixWide(float a, float b, float t)
{
    const float tb = __fmul_rn(t, b);
    return (float)__dadd_rn(__dmul_rn(__dsub_rn(1.0, (double)t), (double)a), (double)tb);
}

// bilinear value noise (kernel/KaminoInitializer.cu:109-125)
__device__ float valueNoise(float x, float y)
{
    const float x0 = floorf(x), fx = __fsub_rn(x, x0);
    const float y0 = floorf(y), fy = __fsub_rn(y, y0);
    const float n00 = latticeHash(x0, y0);
    const float n10 = latticeHash(__fadd_rn(x0, 1.0f), y0);
    const float n01 = latticeHash(x0, __fadd_rn(y0, 1.0f));
    const float n11 = latticeHash(__fadd_rn(x0, 1.0f), __fadd_rn(y0, 1.0f));
    return mixWide(mixWide(n00, n10, fx), mixWide(n01, n11, fx), fy);
}

// four octaves, persistence 0.5, anisotropic base resolution (kernel/KaminoInitializer.cu:87-102)
__device__ float fbm(float x, float y)
{
    const float resX = 0.15f, resY = 0.5f;
    float total = 0.0f, freq = 1.0f, amp = 1.0f;
#pragma unroll
    for (int octave = 0; octave < 4; ++octave) {
        const float vn = valueNoise(__fdiv_rn(__fmul_rn(x, freq), resX), __fdiv_rn(__fmul_rn(y, freq), resY));
        total = __fadd_rn(total, __fmul_rn(amp, vn));
        freq *= 2.0f; amp *= 0.5f;                   // exact powers of two
    }
    return __fdiv_rn(__fmul_rn(0.5f, total), 2.0f);   // norm * total / 2.0f, norm = 1 - persistence
}

// one thread per cell of rows [rowBegin, rowBegin + rowCount): u_phi(j, i) and u_theta(j, i) (row j of u_theta is the
// node below row j of cells, kernel/KaminoInitializer.cu:11-83)
__global__ void initVelocityKernel(GridParams g, float* __restrict__ velPhi, float* __restrict__ velTheta, int rowBegin, int rowCount)
{
    const long cell = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (long)rowCount * g.nPhi) return;
    const int j = rowBegin + (int)(cell >> g.log2NPhi), i = (int)(cell & (g.nPhi - 1));
    const float h = (float)(kTwoPi / (double)g.nPhi);             // the solver's gridLen, KaminoSolver.cu:14
    const float gain = (float)(4096.0 / (double)g.nPhi);          // KaminoInitializer.cu:9
    const float scale = __fmul_rn(g.radius, h);
    const float halfH = __fdiv_rn(h, 2.0f);
    {
        const float yUp = __fmul_rn((float)(j + 1), h), yLo = __fmul_rn((float)j, h);
        const float xR = (i == 0) ? halfH : __fadd_rn(__fmul_rn((float)i, h), halfH);
        const float xL = (i == 0) ? (float)__dsub_rn(2.0 * kPi, (double)halfH) : __fsub_rn(__fmul_rn((float)i, h), halfH);
        const float dR = __fdiv_rn(__fsub_rn(fbm(xR, yUp), fbm(xR, yLo)), scale);
        const float dL = __fdiv_rn(__fsub_rn(fbm(xL, yUp), fbm(xL, yLo)), scale);
        velPhi[(size_t)j * g.nPhi + i] = __fmul_rn((float)((double)__fadd_rn(dR, dL) / 2.0), gain);
    }
    if (j + 1 < g.nTheta) {                                       // u_theta row j = the reference's loop index j + 1
        const int jj = j + 1;
        const float yUp = __fadd_rn(__fmul_rn((float)jj, h), halfH), yLo = __fsub_rn(__fmul_rn((float)jj, h), halfH);
        const float xR = __fmul_rn((float)(i + 1), h), xL = __fmul_rn((float)i, h);
        const float upperLeft = fbm(xL, yUp);
        const float dU = __fdiv_rn(__fmul_rn(-1.0f, __fsub_rn(fbm(xR, yUp), upperLeft)), scale);
        const float dD = __fdiv_rn(__fmul_rn(-1.0f, __fsub_rn(fbm(xR, yLo), upperLeft)), scale);
        velTheta[(size_t)j * g.nPhi + i] = __fmul_rn((float)((double)__fadd_rn(dU, dD) / 2.0), gain);
    }
}

__device__ unsigned long long splitmix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// uniform in [0, 1] with the granularity of rand() / RAND_MAX (31 bits)
__device__ float uniform01(unsigned long long seed, unsigned long long counter)
{
    const unsigned r = (unsigned)(splitmix64(seed ^ splitmix64(counter)) >> 33);      // 31 bits
    return __fdiv_rn((float)r, 2147483647.0f);
}

// the lattice of kernel/KaminoParticles.cu:20-62, index i * numTheta + j, four uniforms per particle
__global__ void seedParticlesKernel(unsigned numTheta, unsigned numPhi, float spacing, unsigned long long seed, float2* __restrict__ coords)
{
    const unsigned long long n = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= (unsigned long long)numTheta * numPhi) return;
    const unsigned i = (unsigned)(n / numTheta), j = (unsigned)(n - (unsigned long long)i * numTheta);
    const float half = (float)((double)spacing / 2.0);
    const float sp = ((double)uniform01(seed, 4 * n) >= 0.5) ? 1.0f : -1.0f;
    const float st = ((double)uniform01(seed, 4 * n + 1) >= 0.5) ? 1.0f : -1.0f;
    const float jitterPhi = __fmul_rn(__fmul_rn(sp, half), uniform01(seed, 4 * n + 2));
    const float jitterTheta = __fmul_rn(__fmul_rn(st, half), uniform01(seed, 4 * n + 3));
    float phi = __fadd_rn(__fmul_rn((float)i, spacing), jitterPhi);
    float theta = __fadd_rn(__fmul_rn((float)j, spacing), jitterTheta);
    if (phi < 0.0f) phi = 0.0f;
    if (theta < 0.0f) theta = 0.0f;
    coords[n] = make_float2(phi, theta);
}

} // namespace

// velPhi / velTheta: pointers addressed with GLOBAL row indices (a band context passes its virtual bases)
cudaError_t launchInitVelocity(const GridParams& g, float* velPhi, float* velTheta, int rowBegin, int rowCount, cudaStream_t stream)
{
    const long cells = (long)rowCount * g.nPhi;
    const int threads = 128;
    initVelocityKernel<<<(unsigned)((cells + threads - 1) / threads), threads, 0, stream>>>(g, velPhi, velTheta, rowBegin, rowCount);
    return cudaGetLastError();
}

cudaError_t launchSeedParticles(int nTheta, float particleDensity, unsigned long long seed, float* coords, long expected, cudaStream_t stream)
{
    const float linear = sqrtf(particleDensity);                                   // KaminoParticles.cu:20
    const float spacing = (float)(kPi / (double)nTheta / (double)linear);           // :21
    const unsigned numTheta = (unsigned)(linear * (float)nTheta), numPhi = 2 * numTheta;      // :24-25
    const long n = (long)numTheta * numPhi;
    if (n != expected) return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    const int threads = 256;
    seedParticlesKernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(numTheta, numPhi, spacing, seed, reinterpret_cast<float2*>(coords));
    return cudaGetLastError();
}

} // namespace kb
