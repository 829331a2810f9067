// Launch interfaces of the kamino_b200 kernels. One step = five launches:
//   advect -> geometric -> divergence+forward FFT -> tridiagonal -> inverse FFT+gradient
#pragma once

#include "kamino_common.cuh"

namespace kb {

struct SamplerConsts;

struct AdvectArgs {
    const float* velPhi;       // this-step buffers (whole batch)
    const float* velTheta;
    const float* density;      // may be NULL
    const float* particles;    // may be NULL; interleaved (phi, theta)
    float* velPhiOut;          // next-step buffers
    float* velThetaOut;
    float* densityOut;
    float* particlesOut;
    // per-row dt / (R sinf(theta_node)) of the u_phi / density nodes and of the u_theta nodes
    const float* cofPhiCentred;
    const float* cofPhiTheta;
    const SamplerConsts* consts;  // device copy of the sampler constants (sampler.cuh)
    int tileBlocks;                              // filled by launchAdvect
    // thread -> particle mapping of the particle blocks (filled by launchAdvect). A seeded particle
    // set is a jittered numPhi x numTheta lattice stored phi-major (kernel/KaminoParticles.cu:39-62,
    // index i * numTheta + j), so 32 consecutive particles lie on 32 different theta rows and every
    // gather of a warp touches 32 cache lines. When the count has the lattice form 2 m^2 the warps
    // are laid over compact patches of the lattice instead ((32 >> log2Inner) columns x
    // (1 << log2Inner) rows per warp, 256 / blockInner x blockInner per block); any other count
    // keeps the linear mapping (latticeInner = 0). The mapping never changes a result.
    int latticeInner, latticeOuter;              // numTheta, numPhi of the lattice (0: linear mapping)
    int log2Inner;                               // log2 of the warp patch height (rows of the lattice)
    int blocksInner;                             // particle blocks along the inner (theta) dimension
    unsigned innerMagic;                         // ceil(2^32 / blocksInner): division by a multiply
};

cudaError_t launchAdvect(const GridParams& g, AdvectArgs a, int batch, cudaStream_t stream);

struct SamplerConsts;
cudaError_t launchLocate(const SamplerConsts* consts, int kind, long n, const float* phiRaw, const float* thetaRaw,
                         int* phiIndex, int* thetaIndex, float* alphaPhi, float* alphaTheta,
                         float* phiOut, float* thetaOut, int* flags, cudaStream_t stream);

// Per-context read-only per-row tables (built once by launchBuildTables with the same device
// functions the reference's kernels call per thread, so every entry is bit-identical).
struct SpectralTables {
    float2* twiddle;      // packed per-pass FFT twiddle tables (fft_core.cuh), < nPhi entries
    float* divFactor;     // nTheta: invGridSine / gridLen        (kernel/KaminoCore.cu:625,628)
    float* sinNorth;      // nTheta: sinf(theta_j - h/2)          (:626)
    float* sinSouth;      // nTheta: sinf(theta_j + h/2)          (:627)
    float* gradPhiDenom;  // nTheta: -gridLen * sinf(theta_j)     (:743-744)
    float* triA;          // nTheta: sub-diagonal before the Neumann fold (KaminoSolver.cu:135-136)
    float* triC;          // nTheta: super-diagonal               (:137-138)
    float* sinSq;         // nTheta: sinf(theta_j)^2              (:134)
    float* geoG;          // nTheta: dt*cosf(theta_j)/(R*sinf(theta_j))   (kernel/KaminoCore.cu:494)
    float* cofPhiCentred; // nTheta: dt/(R*sinf((j+1/2)h))  cofPhi of u_phi / density nodes (:202-203, :292-293)
    float* cofPhiTheta;   // nTheta: dt/(R*sinf((j+1)h))    cofPhi of u_theta nodes (:247-248)
    SamplerConsts* samplerConsts;  // 64 bytes, filled by fillSamplerConsts + a host-to-device copy
    double minusTwoOverH2;  // -2.0 / (h*h)                        (:133)
    // LU (Thomas) factors of every wavenumber slot and their chunk products (tridiag.cu),
    // layout [slot group][row][w]
    float* thL;           // nTheta x N/2: l_i = a_i / b'_{i-1}
    float* thInvB;        // 1 / b'_i
    float* thBetaInv;     // beta_i / b'_i, beta_i = prod_{chunk start..i} (-l)
    float* thH;           // h_i = c_i / b'_i
    float* thDelta;       // delta_i = prod_{i..chunk end} (-h)
    float* thBetaEnd;     // beta at the last row of every chunk, [slot group][chunk][w]
};

// geometric phase: velPhi/velTheta (this) -> velPhiOut/velThetaOut (next)
cudaError_t launchGeometric(const GridParams& g, const SpectralTables& t, const float* velPhi, const float* velTheta,
                            float* velPhiOut, float* velThetaOut, int batch, cudaStream_t stream);

size_t spectralTableBytes(const GridParams& g);
size_t solveTableFloats(const GridParams& g);
cudaError_t configureTridiagonal(const GridParams& g, int batch);
// runs after launchBuildTables (same stream): LU factorisation of every wavenumber's system, once
// (slotBegin, slotCount: build the tables of a wavenumber band only, indexed from 0; default: every slot)
cudaError_t launchBuildSolveTables(const GridParams& g, SpectralTables t, int batch, cudaStream_t stream,
                                   int slotBegin = 0, int slotCount = -1);
cudaError_t launchBuildTables(const GridParams& g, SpectralTables t, cudaStream_t stream);

// spectrum layout: S[sim][j][k], k = 0 .. nPhi/2-1, float2; slot k holds wavenumber
// n = k for k >= 1 and the Nyquist mode n = nPhi/2 in slot 0 (the n = 0 mode is never
// projected by the reference, kernel/KaminoSolver.cu:154-159 + KaminoCore.cu:692-700).
// Theta-band runs (dist.cu) keep the half spectrum in the layouts of the transposes around the theta solve
// instead of dense [row][slot]: slots are grouped into blocks of 2^log2Block (one block per destination /
// source rank), element (row, k) sits at (k >> log2Block) * blockPitch + (row - rowBase) * rowPitch +
// (k & (2^log2Block - 1)). The forward FFT stores straight into the send buffer of the all-to-all and the
// inverse FFT reads straight from its receive buffer: no pack / unpack passes.
// peerTable != NULL (peer-memory transposes): block b of slots is not stored locally but straight into rank b's buffer
// peerTable[b] (a device pointer valid on this GPU: the rank's own memory, or a peer's mapped through CUDA IPC / the same
// process), at [global row][slot within the block] with row pitch rowPitch -- the receive layout of the transpose.
struct SpectrumLayout { int rowBase, rowPitch, log2Block; size_t blockPitch; float2* const* peerTable; };
// Where the theta solve of a band-decomposed run puts its solution when the transposes go through peer memory: row i of
// my K slots belongs to rank i >> log2Rows, whose buffer table[rank] holds [source rank][rows + 1][K]; the first row of a
// band is also the extra row of the band above it. table == NULL: in place (single GPU, or NCCL transposes).
struct PeerScatter { float2* const* table; int log2Rows, myRank, kper; };
cudaError_t launchDivergenceFFT(const GridParams& g, const SpectralTables& t, const float* velPhi,
                                const float* velTheta, float2* spectrum, int batch, cudaStream_t stream,
                                const SpectrumLayout* packed = nullptr);
cudaError_t launchTridiagonal(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch,
                              cudaStream_t stream);
// band-decomposed runs: solve slots [slotBegin, slotBegin + slotCount) whose right-hand sides sit in
// `packed` as [theta][slot - slotBegin] with row pitch `pitch` (float2 elements), one simulation
cudaError_t launchTridiagonalBand(const GridParams& g, const SpectralTables& t, float2* packed, int pitch,
                                  int slotBegin, int slotCount, int tableBatch, cudaStream_t stream,
                                  const PeerScatter* scatter = nullptr);
// band-local theta solve of the reduced-interface (SPIKE) mode: `band` holds th* tables built for rows
// [rowBegin, rowBegin + rows) cut loose from their neighbours; only its th* members are used
size_t bandSolveTableFloats(const GridParams& g, int rows);
cudaError_t launchBuildBandSolveTables(const GridParams& g, const SpectralTables& t, const SpectralTables& band,
                                       int rowBegin, int rows, cudaStream_t stream);
cudaError_t launchBandLocalSolve(const GridParams& g, const SpectralTables& band, float2* spectrum, int rowBegin, int rows,
                                 cudaStream_t stream);
// velPhi / velTheta updated in place; pressure (may be NULL) receives p.
cudaError_t launchInverseFFTGradient(const GridParams& g, const SpectralTables& t, const float2* spectrum,
                                     float* velPhi, float* velTheta, float* pressure, int batch,
                                     cudaStream_t stream, const SpectrumLayout* packed = nullptr);

// parity instrumentation (debug_cr.cu): the theta solve in the reference's cyclic-reduction order, in place
cudaError_t launchCyclicReductionDebug(const GridParams& g, const SpectralTables& t, float2* spectrum, int batch, cudaStream_t stream);

// GPU-side initialisers (device_init.cu): the FBM initial velocity of rows [rowBegin, rowBegin + rowCount) into buffers
// addressed with global row indices, and the counter-based particle lattice (expected = the allocated particle count)
cudaError_t launchInitVelocity(const GridParams& g, float* velPhi, float* velTheta, int rowBegin, int rowCount, cudaStream_t stream);
cudaError_t launchSeedParticles(int nTheta, float particleDensity, unsigned long long seed, float* coords, long expected, cudaStream_t stream);

// host: the constants block of the samplers for this grid
// (validLo, validHi, haloViolation: the resident row range and the device flag of a theta-band context, dist.cu)
void fillSamplerConsts(const GridParams& g, void* hostBlock64, int validLo = 0, int validHi = -1, int* haloViolation = nullptr);

// One-time setup of kernel attributes (opt-in shared memory); called at context creation.
cudaError_t configureKernels(const GridParams& g, int batch);

} // namespace kb
