/*
 * kamino_b200.h -- C ABI of the B200-native (sm_100a) implementation of KaminoGPU's
 * per-timestep solver path (advection -> geometric -> projection, plus tracer particles).
 *
 * The reference has no FFI: its boundary for this path is the C++ class surface
 * KaminoSolver / KaminoQuantity / KaminoParticles / Kamino (paths below are relative to
 * /root/reference/KaminoGPU/). The C++ host classes in kaminogpu_b200/host/ keep that
 * surface and call ONLY the functions declared here; Python (ctypes) binds the same
 * symbols for tests and bench.py. Plain pointers and sizes only -- no CUDA or torch
 * types cross this boundary (streams are passed as void*).
 *
 * Conventions: every function returns 0 on success, otherwise a non-zero code
 * (a cudaError_t value, or one of KAMINO_ERR_*); kamino_last_error() returns the text.
 * No exceptions, no process-global solver state, no cudaDeviceReset: several contexts
 * may coexist in one process (unlike the reference, include/KaminoSolver.cuh:6-112,
 * whose kernels read file-static __constant__ parameters, kernel/KaminoCore.cu:5-9).
 * A context is not thread-safe; use one context per host thread.
 *
 * Grid: nPhi = 2 * nTheta (kernel/KaminoCore.cu:849), nTheta a power of two >= 16.
 * Field layouts (row-major, row = theta index j, column = phi index i, dense pitch nPhi):
 *   KAMINO_VEL_PHI    nTheta     x nPhi   node (phi=(i-1/2)h, theta=(j+1/2)h)
 *   KAMINO_VEL_THETA  (nTheta-1) x nPhi   node (phi=i h,      theta=(j+1)h)
 *   KAMINO_DENSITY    nTheta     x nPhi   node (phi=i h,      theta=(j+1/2)h)
 *   KAMINO_PRESSURE   nTheta     x nPhi   same nodes as density
 * Particles: numParticles x (phi, theta) interleaved fp32 (kernel/KaminoParticles.cu:59-62).
 * A context may hold `batch` independent simulations of the same shape (ensembles);
 * `sim` selects one.
 */
#ifndef KAMINO_B200_H
#define KAMINO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kamino_ctx kamino_ctx;

enum {
    KAMINO_VEL_PHI = 0,
    KAMINO_VEL_THETA = 1,
    KAMINO_DENSITY = 2,
    KAMINO_PRESSURE = 3
};

enum {
    KAMINO_ERR_INVALID = 10001,   /* bad argument */
    KAMINO_ERR_NO_DEVICE = 10002, /* no CUDA device / not an sm_100 part */
    KAMINO_ERR_STATE = 10003      /* call not valid in the current state */
};

/* sampler kinds for kamino_debug_locate: sampleVPhi / sampleVTheta / sampleCentered,
 * kernel/KaminoCore.cu:36-184 */
enum { KAMINO_SAMPLE_VPHI = 0, KAMINO_SAMPLE_VTHETA = 1, KAMINO_SAMPLE_CENTERED = 2 };

/* ---- lifetime -------------------------------------------------------------------- */

/* Replaces the KaminoSolver constructor's device work (kernel/KaminoSolver.cu:12-67:
 * cudaSetDevice, ten cudaMalloc, precomputeABCCoef, four KaminoQuantity, cufftPlanMany)
 * and the constant uploads of Kamino::run (kernel/KaminoCore.cu:868-872). `dt` is the
 * value the reference uploads as timeStepGlobal (the kernels ignore stepForward's
 * argument, kernel/KaminoSolver.cu:199). particlesPerSim may be 0. */
int kamino_create(kamino_ctx** out, int device, int nTheta, float radius, float dt,
                  int batch, long particlesPerSim);

/* KaminoSolver::initParticlesfromPic (kernel/KaminoSolver.cu:279-282) creates the particle
 * set after the solver exists: (re)allocate the particle double buffer for
 * particlesPerSim particles per simulation (contents zeroed; upload afterwards). */
int kamino_alloc_particles(kamino_ctx* ctx, long particlesPerSim);

/* Replaces ~KaminoSolver (kernel/KaminoSolver.cu:69-104) without the cudaDeviceReset. */
int kamino_destroy(kamino_ctx* ctx);

/* Text of the last error of this context (or of the last failed kamino_create when
 * ctx is NULL). Never NULL. */
const char* kamino_last_error(const kamino_ctx* ctx);

/* Run all work of this context on an existing CUDA stream (a cudaStream_t passed as
 * void*; NULL restores the context's own stream). Invalidates captured step graphs. */
int kamino_set_stream(kamino_ctx* ctx, void* cudaStream);

/* Shape queries. */
int kamino_get_shape(const kamino_ctx* ctx, int* nTheta, int* nPhi, int* batch, long* particlesPerSim);

/* ---- state transfer ---------------------------------------------------------------- */

/* KaminoQuantity::copyToGPU / copyBackToCPU (kernel/KaminoQuantity.cu:3-18): dense host
 * array (rows x nPhi floats) <-> the "this step" device buffer of `field` of simulation
 * `sim`. Synchronous with respect to the host buffer. */
int kamino_upload_field(kamino_ctx* ctx, int field, int sim, const float* host);
int kamino_download_field(kamino_ctx* ctx, int field, int sim, float* host);

/* KaminoParticles::copy2GPU / copyBack2CPU (kernel/KaminoParticles.cu:94-110). */
int kamino_upload_particles(kamino_ctx* ctx, int sim, const float* hostPhiTheta);
int kamino_download_particles(kamino_ctx* ctx, int sim, float* hostPhiTheta);

/* Asynchronous variants for pinned host memory (used by the end-to-end frame loop);
 * completion is observed with kamino_sync. */
int kamino_download_field_async(kamino_ctx* ctx, int field, int sim, float* pinnedHost);
int kamino_download_particles_async(kamino_ctx* ctx, int sim, float* pinnedHost);
int kamino_upload_field_async(kamino_ctx* ctx, int field, int sim, const float* pinnedHost);
int kamino_upload_particles_async(kamino_ctx* ctx, int sim, const float* pinnedHost);

/* KaminoQuantity::getGPUThisStep / getGPUNextStep / get*PitchInElements
 * (include/KaminoQuantity.cuh:60-66): raw device pointer of the current this/next buffer
 * (which = 0 / 1) and its pitch in elements. The roles move after every phase as the
 * reference's swaps do (re-query the pointers after a phase rather than swapping a cached pair). */
int kamino_field_device_ptr(kamino_ctx* ctx, int field, int sim, int which,
                            void** devicePtr, size_t* pitchInElements);
/* KaminoParticles::coordGPUThisStep / coordGPUNextStep (include/KaminoParticles.cuh:18-19). */
int kamino_particles_device_ptr(kamino_ctx* ctx, int sim, int which, void** devicePtr);

/* ---- the hot path ------------------------------------------------------------------- */

/* KaminoSolver::advection (kernel/KaminoCore.cu:344-384): u_phi, u_theta, density and the
 * particles are advected with the pre-advection velocity; buffers swap. */
int kamino_advect(kamino_ctx* ctx);
/* KaminoSolver::geometric (kernel/KaminoCore.cu:551-583). */
int kamino_geometric(kamino_ctx* ctx);
/* KaminoSolver::projection (kernel/KaminoCore.cu:749-842); leaves the pressure in
 * KAMINO_PRESSURE. */
int kamino_project(kamino_ctx* ctx);

/* ---- theta-band entry points -----------------------------------------------------------
 * The reference is single-GPU (device 0 hard-coded, kernel/KaminoSolver.cu:20); these have no
 * counterpart there. A band-decomposed run (kaminogpu_b200/banded.py: one process per GPU,
 * NCCL halo exchange + all-to-all around the theta solve) keeps full-size buffers on every
 * rank and asks each kernel of the step for a range of GLOBAL theta rows only. Same kernels,
 * same arithmetic as kamino_advect / kamino_geometric / kamino_project, so a banded run is
 * bit-identical to the single-GPU run. Asynchronous on the context's stream; single-
 * simulation contexts without particles. Row ranges: multiples of 8 rows for advect and
 * geometric, of 2 for divergence_fft; buffers swap exactly as in the phase calls. */
int kamino_band_advect(kamino_ctx* ctx, int rowBegin, int rowCount);
int kamino_band_geometric(kamino_ctx* ctx, int rowBegin, int rowCount);
/* divergence + forward FFT of rows [rowBegin, rowBegin+rowCount) into the spectrum buffer
 * ([theta][slot], nPhi/2 float2 per row; slot k = wavenumber k, slot 0 = Nyquist). */
int kamino_band_divergence_fft(kamino_ctx* ctx, int rowBegin, int rowCount);
/* theta solve of slots [slotBegin, slotBegin+slotCount) (multiples of 8) for ALL rows, in place
 * on a packed device buffer [nTheta][pitch] of float2 (what the all-to-all delivers). */
int kamino_band_tridiagonal(kamino_ctx* ctx, void* packedSpectrum, int pitch, int slotBegin, int slotCount);
/* Reduced-interface ("SPIKE") alternative to the all-to-all around the theta solve (SURVEY.md 8f-3;
 * no reference counterpart). prepare builds, once, the LU factors of the rank's band
 * [rowBegin, rowBegin+rowCount) (a power of two >= 16 rows) cut loose from the neighbouring bands;
 * local_solve applies them in place to the spectrum rows of the band, all wavenumbers, with the
 * same kernel as the full solve. The driver (kaminogpu_b200/banded.py) derives the spike vectors
 * from unit right-hand sides, exchanges two interface values per wavenumber and rank, and corrects
 * the band. tridiagonal_coefficients returns the unfolded sub- and super-diagonal of a row
 * (kernel/KaminoSolver.cu:135-138), i.e. the couplings across a band boundary. */
int kamino_band_solver_prepare(kamino_ctx* ctx, int rowBegin, int rowCount);
int kamino_band_local_solve(kamino_ctx* ctx);
int kamino_tridiagonal_coefficients(kamino_ctx* ctx, int row, float* subDiagonal, float* superDiagonal);
/* inverse FFT + gradient subtraction of rows [rowBegin, rowBegin+rowCount); reads spectrum rows
 * rowBegin .. rowBegin+rowCount (one past the band, unless it is the last row of the grid). */
int kamino_band_inverse_fft_gradient(kamino_ctx* ctx, int rowBegin, int rowCount);
/* device pointer of the half-spectrum buffer of simulation `sim` (nTheta x nPhi/2 float2). */
int kamino_spectrum_device_ptr(kamino_ctx* ctx, int sim, void** devicePtr);

/* KaminoSolver::stepForward (kernel/KaminoSolver.cu:197-221) nSteps times, launched as
 * CUDA graphs (10-, 2- and 1-step graphs, captured, instantiated and uploaded at context creation /
 * particle allocation) on the context's stream. Asynchronous: returns once the work is queued. */
int kamino_step(kamino_ctx* ctx, int nSteps);

/* Block the host until all queued work of this context is complete. */
int kamino_sync(kamino_ctx* ctx);

/* Kamino::run's frame loop (kernel/KaminoCore.cu:886-904) with output to host memory
 * instead of .bgeo files: for each of nFrames frames, stepsPerFrame steps followed by
 * the read-backs the reference's writers perform (velPhi, velTheta, density and particle
 * coordinates of every simulation, kernel/KaminoSolver.cu:301-303,375) into the given
 * pinned host buffers (each sized for all `batch` simulations, any may be NULL to skip).
 * Synchronous. */
int kamino_run_frames(kamino_ctx* ctx, int nFrames, int stepsPerFrame,
                      float* hostVelPhi, float* hostVelTheta, float* hostDensity,
                      float* hostParticles);

/* Per-phase accumulated device time in seconds since creation / last reset: the
 * reference's advectionTime / geometricTime / projectionTime (kernel/KaminoSolver.cu:201-218).
 * Only kamino_advect / kamino_geometric / kamino_project accumulate (graph-launched steps
 * are not split by phase). */
int kamino_phase_times(kamino_ctx* ctx, float* advection, float* geometric, float* projection, int reset);

/* Number of kernel launches one kamino_step(ctx, 1) performs (for bench.py's gpu_launches):
 * 5 (the particles are part of the advection launch). */
int kamino_launches_per_step(const kamino_ctx* ctx);

/* Run nSteps steps with every kernel launched individually and bracketed by CUDA events on
 * the context's stream; kernelSeconds[k] (k = 0 .. kamino_launches_per_step()-1: advection,
 * geometric, divergence+FFT, tridiagonal, inverse FFT+gradient) receives the summed device
 * time of kernel k. The per-kernel counterpart of the reference's per-phase KaminoTimer
 * brackets (kernel/KaminoSolver.cu:201-218). Synchronous. */
int kamino_profile_steps(kamino_ctx* ctx, int nSteps, float* kernelSeconds);

/* ---- theta-band decomposition of one simulation over P GPUs (BASELINE config 5) -------------------------
 * No reference counterpart: the reference is single-GPU (kernel/KaminoSolver.cu:20) and cannot launch its
 * theta solve above nTheta = 2048 (kernel/KaminoCore.cu:779-784). One kamino_dist per GPU (one process per
 * GPU, or one thread per GPU): rank r owns theta rows [r nTheta/P, (r+1) nTheta/P) and the same range of
 * wavenumber slots, holds band-sized buffers only, and kamino_dist_step runs the whole step loop in C++:
 * NCCL halo exchange (24 rows of u_phi, u_theta, density per neighbour), advection / geometric /
 * divergence + FFT on the band, NCCL transpose, theta solve of the rank's wavenumbers, transpose back,
 * inverse FFT + gradient (csrc/dist.cu). Same kernels and arithmetic as kamino_step: bit-identical to the
 * single-GPU run. No tracer particles (they would migrate between ranks).
 *
 * Bootstrap: rank 0 calls kamino_dist_unique_id and hands the 128 bytes to the other ranks by any means
 * (torch.distributed broadcast in bench.py, MPI_Bcast, a file); every rank then calls kamino_dist_create
 * with it (collective: ncclCommInitRank). id128 = NULL creates a rank without a communicator: P such ranks
 * on one device are stepped by kamino_dist_group_step (device copies instead of NCCL; tests).
 * world: power of two, >= 32 rows and a multiple of 8 wavenumbers per rank. */
typedef struct kamino_dist kamino_dist;
int kamino_dist_unique_id(void* id128);
int kamino_dist_create(kamino_dist** out, int device, int nTheta, float radius, float dt, int rank, int world,
                       const void* id128);
int kamino_dist_destroy(kamino_dist* d);
const char* kamino_dist_last_error(const kamino_dist* d);
/* rows [rowBegin, rowEnd) and wavenumber slots [slotBegin, slotEnd) of this rank; bytes of device memory held */
int kamino_dist_shape(const kamino_dist* d, int* rowBegin, int* rowEnd, int* slotBegin, int* slotEnd, size_t* deviceBytes);
/* the rank's OWN rows of a field (dense, rows x nPhi; u_theta: the rows below nTheta - 1) <-> host; synchronous */
int kamino_dist_upload(kamino_dist* d, int field, const float* hostRows);
int kamino_dist_download(kamino_dist* d, int field, float* hostRows);
/* nSteps steps of the band-decomposed simulation; collective over the communicator, asynchronous */
int kamino_dist_step(kamino_dist* d, int nSteps);
int kamino_dist_group_step(kamino_dist* const* ranks, int world, int nSteps);
/* Blocks until the rank's work is complete. Returns KAMINO_ERR_STATE if a backtrace left the 24-row halo
 * since the last call (theta-CFL too large for the decomposition: the run no longer equals the single-GPU one). */
int kamino_dist_sync(kamino_dist* d);
int kamino_dist_stream(kamino_dist* d, void** cudaStream);
/* Transport of the two transposes around the theta solve. usesPeerStores = 1: the FFT kernel and the theta solve store
 * straight into the owning ranks' buffers over NVLink (CUDA IPC mappings of the peers' memory, set up at creation; the
 * default whenever every rank could map every peer), the transposes reduce to a barrier each; 0: NCCL send / recv of
 * staged buffers. setPeerStores = 1 / 0 switches (collective: every rank must make the same call; -1 only queries);
 * note receives a short text saying how the peers are mapped or why they are not. Results are bit-identical either way. */
int kamino_dist_transport(kamino_dist* d, int setPeerStores, int* usesPeerStores, const char** note);
/* Communication accounting. enable = 1 / 0 switches per-step CUDA-event brackets around the NCCL calls on / off
 * (it adds one host synchronisation per step; -1 leaves the setting); the accumulated seconds and bytes SENT per
 * step by this rank are returned. */
int kamino_dist_comm_stats(kamino_dist* d, int enable, double* haloSeconds, double* transposeSeconds, long* steps,
                           size_t* haloBytesPerStep, size_t* transposeBytesPerStep);

/* ---- host-side initialisers (pure CPU; reproduce the reference's initial state) ------ */

/* KaminoSolver::initialize_velocity (kernel/KaminoInitializer.cu:3-134): FBM curl-noise
 * initial u_phi (nTheta x nPhi) and u_theta ((nTheta-1) x nPhi). */
int kamino_init_velocity_host(int nTheta, float radius, float* velPhi, float* velTheta);
/* The same field for rows [rowBegin, rowBegin + rowCount) only (a band of a decomposed run): velPhi receives
 * rowCount rows, velTheta the rows of the range below nTheta - 1. Every cell is a pure function of its position. */
int kamino_init_velocity_host_rows(int nTheta, float radius, int rowBegin, int rowCount, float* velPhi, float* velTheta);
/* KaminoParticles constructor (kernel/KaminoParticles.cu:20-62): particle count for a
 * density, and the jittered lattice driven by libc rand() in the reference's call order
 * (the generator is put into its never-seeded state first). */
long kamino_particle_count(int nTheta, float particleDensity);
int kamino_seed_particles_host(int nTheta, float particleDensity, float* coords);

/* ---- GPU-side initialisers (no reference counterpart: its initialisers are serial host loops) ---- */

/* The FBM initial velocity of kernel/KaminoInitializer.cu:3-134 evaluated on the device, operation for operation,
 * into the "this step" velocity buffers of every simulation of the context (of the rank's band for a kamino_dist).
 * Differs from kamino_init_velocity_host only where CUDA's double sin() and glibc's straddle an fp32 rounding boundary
 * inside the lattice hash (about 1e-8 of the evaluations; measured in tests/test_parity_gpu.py). For start-up at sizes
 * where the host loop costs minutes; the drop-in classes keep the bit-identical host initialiser. */
int kamino_init_velocity_device(kamino_ctx* ctx);
int kamino_dist_init_velocity_device(kamino_dist* d);
/* The particle lattice of kernel/KaminoParticles.cu:20-62 (same counts, spacing, jitter range, clamp and index order)
 * with the four uniforms of every particle taken from a counter-based generator (splitmix64 of (seed, index)) instead
 * of libc rand(): reproducible for a seed on any device, NOT the reference's sequence. The context must hold
 * kamino_particle_count(nTheta, particleDensity) particles (kamino_alloc_particles), else KAMINO_ERR_STATE. */
int kamino_seed_particles_device(kamino_ctx* ctx, float particleDensity, unsigned long long seed);

/* ---- parity instrumentation ------------------------------------------------------------ */

/* Evaluate, on the device, the index / predicate part of the reference's samplers
 * (kernel/KaminoCore.cu:36-53, 86-103, 136-153) for n raw coordinates: cell indices
 * before the modulo, interpolation weights, validated coordinates and flags
 * (bit 0 = validateCoord returned -1, bit 1 = pole branch taken). Host arrays. */
int kamino_debug_locate(kamino_ctx* ctx, int kind, long n, const float* phiRaw, const float* thetaRaw,
                        int32_t* phiIndex, int32_t* thetaIndex, float* alphaPhi, float* alphaTheta,
                        float* phiValidated, float* thetaValidated, int32_t* flags);

/* kamino_project with the theta solve performed in the reference's own operation order (cyclic reduction in
 * fp32, kernel/tdm.cu:3-96) instead of the product's LU recurrences: lets a test show that the distance
 * between the two builds' pressure is the reference's CR rounding. nTheta <= 2048 (the reference's launch
 * limit, kernel/KaminoCore.cu:779-784). Test use only; never part of kamino_step. */
int kamino_debug_project_cr(kamino_ctx* ctx);

/* Library build information: "kamino_b200 <version> sm_100a". */
const char* kamino_version(void);

/* Allocate / free page-locked host memory (for the asynchronous transfer entry points). */
int kamino_host_alloc(void** ptr, size_t bytes);
int kamino_host_free(void* ptr);

#ifdef __cplusplus
}
#endif

#endif /* KAMINO_B200_H */
