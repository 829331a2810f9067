// Context, memory layout, step graphs and the C ABI of kamino_b200 (include/kamino_b200.h).
//
// HBM layout (one arena per context, every sub-buffer 256-byte aligned, batch-major):
//   velPhi[2], velTheta[2], density[2]   batch x nTheta x nPhi fp32 each (u_theta uses
//                                        nTheta-1 rows of its slot), double-buffered
//   pressure                             batch x nTheta x nPhi fp32
//   spectrum                             batch x nTheta x nPhi/2 float2 (half spectrum)
//   particles[2]                         batch x numParticles float2, double-buffered
//   tables                               twiddles + per-row constants + the LU factors of every
//                                        wavenumber's theta system (5 floats per spectrum entry =
//                                        10 B per cell, read-only)
// = 36 B/cell of state + 10 B/cell of tables + 16 B/particle (the reference: 76 B/cell,
// SURVEY.md appendix B).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kamino_b200.h"
#include "kamino_kernels.cuh"

using namespace kb;

struct kamino_ctx {
    int device = 0;
    int batch = 1;
    GridParams g{};
    cudaStream_t ownStream = nullptr;    // graphs are captured here
    cudaStream_t copyStream = nullptr;   // frame read-backs
    cudaStream_t stream = nullptr;       // where work is launched (ownStream unless overridden)
    char* arena = nullptr;          // fields, spectrum, tables
    size_t arenaBytes = 0;
    char* particleArena = nullptr;  // particles[2] + read-back snapshot (sized by kamino_alloc_particles)

    float* velPhi[2]{};
    float* velTheta[2]{};
    float* density[2]{};
    float* pressure = nullptr;
    float2* spectrum = nullptr;
    float* particles[2]{};
    float* snapshot = nullptr;           // staging for overlapped frame read-backs
    size_t snapshotFloats = 0;
    SpectralTables tables{};
    // band-local LU tables of the reduced-interface (SPIKE) mode, built by kamino_band_solver_prepare
    char* bandArena = nullptr;
    SpectralTables bandTables{};
    int bandRowBegin = 0, bandRows = 0;

    int velIdx = 0, densityIdx = 0, particleIdx = 0;   // which buffer is "this step"
    // Velocity ping-pong, as in the reference: advect writes the other buffer, geometric writes back into
    // the first, the projection corrects that buffer in place. (r01i A/B: a third velocity buffer, so that the
    // tracer particles run as their own kernel on a parallel graph branch, is a loss -- 63.9 vs 54.5 us/step at
    // 512 x 1024: inside the fused launch the particle blocks fill the tail waves of the tile blocks.)
    static constexpr int velBuffers = 2;

    std::map<std::pair<int, int>, cudaGraphExec_t> graphs;   // (parity, steps) -> exec

    cudaEvent_t evStart = nullptr, evStop = nullptr, evSnap = nullptr, evCopied = nullptr;
    float advectionTime = 0.f, geometricTime = 0.f, projectionTime = 0.f;

    std::string lastError;
};

namespace {

constexpr int kStepKernels = 5;

int nextVel(const kamino_ctx*, int v) { return v ^ 1; }

thread_local std::string g_createError;

int fail(kamino_ctx* ctx, int code, const std::string& what)
{
    std::string msg = what;
    if (code > 0 && code < 10000) {
        msg += ": ";
        msg += cudaGetErrorString((cudaError_t)code);
    }
    if (ctx) ctx->lastError = msg; else g_createError = msg;
    return code;
}

#define KB_TRY(ctx, expr)                                                        \
    do {                                                                         \
        cudaError_t kb_e__ = (expr);                                             \
        if (kb_e__ != cudaSuccess) return fail((ctx), (int)kb_e__, #expr);       \
    } while (0)

struct DeviceGuard {
    int previous = -1;
    explicit DeviceGuard(int device) { cudaGetDevice(&previous); if (previous != device) cudaSetDevice(device); else previous = -1; }
    ~DeviceGuard() { if (previous >= 0) cudaSetDevice(previous); }
};

size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t fieldRows(const kamino_ctx* ctx, int field)
{
    return field == KAMINO_VEL_THETA ? (size_t)ctx->g.nTheta - 1 : (size_t)ctx->g.nTheta;
}

float* fieldBuffer(kamino_ctx* ctx, int field, int which)
{
    switch (field) {
    case KAMINO_VEL_PHI: return ctx->velPhi[which ? nextVel(ctx, ctx->velIdx) : ctx->velIdx];
    case KAMINO_VEL_THETA: return ctx->velTheta[which ? nextVel(ctx, ctx->velIdx) : ctx->velIdx];
    case KAMINO_DENSITY: return ctx->density[ctx->densityIdx ^ which];
    case KAMINO_PRESSURE: return ctx->pressure;
    default: return nullptr;
    }
}

// enqueue the kernels of one phase on `s`, using and updating the index state `st`
struct IndexState { int vel, density, particle; };

AdvectArgs advectArgs(kamino_ctx* ctx, const IndexState& st)
{
    AdvectArgs a{};
    a.velPhi = ctx->velPhi[st.vel];
    a.velTheta = ctx->velTheta[st.vel];
    a.density = ctx->density[st.density];
    a.particles = ctx->g.numParticles > 0 ? ctx->particles[st.particle] : nullptr;
    a.velPhiOut = ctx->velPhi[nextVel(ctx, st.vel)];
    a.velThetaOut = ctx->velTheta[nextVel(ctx, st.vel)];
    a.densityOut = ctx->density[st.density ^ 1];
    a.particlesOut = ctx->g.numParticles > 0 ? ctx->particles[st.particle ^ 1] : nullptr;
    a.cofPhiCentred = ctx->tables.cofPhiCentred;
    a.cofPhiTheta = ctx->tables.cofPhiTheta;
    a.consts = ctx->tables.samplerConsts;
    return a;
}

// cells and particles; flips the indices
cudaError_t enqueueAdvect(kamino_ctx* ctx, IndexState& st, cudaStream_t s)
{
    cudaError_t e = launchAdvect(ctx->g, advectArgs(ctx, st), ctx->batch, s);
    st.vel = nextVel(ctx, st.vel); st.density ^= 1; st.particle ^= 1;      // kernel/KaminoCore.cu:373,380,383
    return e;
}

cudaError_t enqueueGeometric(kamino_ctx* ctx, IndexState& st, cudaStream_t s)
{
    cudaError_t e = launchGeometric(ctx->g, ctx->tables, ctx->velPhi[st.vel], ctx->velTheta[st.vel],
                                    ctx->velPhi[nextVel(ctx, st.vel)], ctx->velTheta[nextVel(ctx, st.vel)], ctx->batch, s);
    st.vel = nextVel(ctx, st.vel);                        // kernel/KaminoCore.cu:582
    return e;
}

// the three kernels of the projection (part = 0, 1, 2)
cudaError_t enqueueProjectPart(kamino_ctx* ctx, IndexState& st, int part, cudaStream_t s)
{
    switch (part) {
    case 0:
        return launchDivergenceFFT(ctx->g, ctx->tables, ctx->velPhi[st.vel], ctx->velTheta[st.vel],
                                   ctx->spectrum, ctx->batch, s);
    case 1:
        return launchTridiagonal(ctx->g, ctx->tables, ctx->spectrum, ctx->batch, s);
    default:
        // the velocity is corrected in place; the reference writes the "next" buffer and swaps
        // (kernel/KaminoCore.cu:828-841), which is the same state for every caller of this ABI
        return launchInverseFFTGradient(ctx->g, ctx->tables, ctx->spectrum, ctx->velPhi[st.vel],
                                        ctx->velTheta[st.vel], ctx->pressure, ctx->batch, s);
    }
}

cudaError_t enqueueProject(kamino_ctx* ctx, IndexState& st, cudaStream_t s)
{
    for (int part = 0; part < 3; ++part) {
        cudaError_t e = enqueueProjectPart(ctx, st, part, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// kernel k of a step: 0 advect, 1 geometric, 2 divergence+FFT, 3 tridiagonal, 4 inverse FFT+gradient
cudaError_t enqueueStepKernel(kamino_ctx* ctx, IndexState& st, int k, cudaStream_t s)
{
    if (k == 0) return enqueueAdvect(ctx, st, s);
    if (k == 1) return enqueueGeometric(ctx, st, s);
    return enqueueProjectPart(ctx, st, k - 2, s);
}

// Timing instrumentation only (scripts/step_mask_timing.py): KAMINO_DEBUG_STEP_MASK leaves kernels
// out of the captured step graph (bit k = kernel k) to attribute the in-graph step time; the results of such a run are meaningless and bench.py
// refuses to run with it set.
int debugStepMask()
{
    static const int mask = [] { const char* e = getenv("KAMINO_DEBUG_STEP_MASK"); return e ? atoi(e) : 63; }();
    return mask;
}

// kernel k of a step unless the instrumentation mask leaves it out (the buffer roles still move)
cudaError_t enqueueStepKernelMasked(kamino_ctx* ctx, IndexState& st, int k, cudaStream_t s)
{
    if (debugStepMask() & (1 << k)) return enqueueStepKernel(ctx, st, k, s);
    if (k == 0) { st.vel = nextVel(ctx, st.vel); st.density ^= 1; st.particle ^= 1; }
    else if (k == 1) st.vel = nextVel(ctx, st.vel);
    return cudaSuccess;
}

// one step in stream order
cudaError_t enqueueStep(kamino_ctx* ctx, IndexState& st, cudaStream_t s)
{
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < kStepKernels && e == cudaSuccess; ++k) e = enqueueStepKernelMasked(ctx, st, k, s);
    return e;
}

void dropGraphs(kamino_ctx* ctx)
{
    for (auto& kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    ctx->graphs.clear();
}

// A graph of `steps` consecutive steps for the current buffer roles. The velocity index
// returns to its starting value after every step; density and particles flip once per step.
int getGraph(kamino_ctx* ctx, IndexState st, int steps, cudaGraphExec_t* out)
{
    const int roles = (st.vel << 2) | (st.density << 1) | st.particle;
    auto key = std::make_pair(roles, steps);
    auto it = ctx->graphs.find(key);
    if (it != ctx->graphs.end()) { *out = it->second; return 0; }
    cudaGraph_t graph = nullptr;
    KB_TRY(ctx, cudaStreamBeginCapture(ctx->ownStream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaSuccess;
    pdlSetCapturing(true);
    for (int k = 0; k < steps && e == cudaSuccess; ++k) e = enqueueStep(ctx, st, ctx->ownStream);
    pdlSetCapturing(false);
    cudaError_t e2 = cudaStreamEndCapture(ctx->ownStream, &graph);
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return fail(ctx, (int)e, "graph capture (launch)"); }
    if (e2 != cudaSuccess) return fail(ctx, (int)e2, "cudaStreamEndCapture");
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(ctx, (int)e, "cudaGraphInstantiate");
    cudaGraphUpload(exec, ctx->ownStream);          // device-side resources now, not at the first launch
    ctx->graphs[key] = exec;
    *out = exec;
    return 0;
}

int getGraph(kamino_ctx* ctx, int steps, cudaGraphExec_t* out)
{
    return getGraph(ctx, IndexState{ctx->velIdx, ctx->densityIdx, ctx->particleIdx}, steps, out);
}

constexpr int kStepChunks[] = {10, 2, 1};           // kamino_step splits nSteps into these graph sizes

// Capture, instantiate and upload every graph kamino_step can ask for while whole steps are the only
// thing that moves the buffer roles (the velocity index returns to its value after every step, density
// and particles flip together): 3 chunk sizes x 2 parities. Called at context creation and whenever the
// particle set is re-allocated, so that no kamino_step ever instantiates inside a caller's timed region
// (r01 VERDICT: the lazily built 10-step graph cost 380 us of a 1.4 ms bench window). Phase calls that
// leave other role combinations still build their graph on first use.
int prepareGraphs(kamino_ctx* ctx)
{
    for (int parity = 0; parity < 2; ++parity)
        for (int steps : kStepChunks) {
            cudaGraphExec_t exec = nullptr;
            const IndexState st{ctx->velIdx, ctx->densityIdx ^ parity, ctx->particleIdx ^ parity};
            if (int rc = getGraph(ctx, st, steps, &exec)) return rc;
        }
    KB_TRY(ctx, cudaStreamSynchronize(ctx->ownStream));
    return 0;
}

template <typename F>
int timedPhase(kamino_ctx* ctx, float& accumulator, F&& enqueue)
{
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, cudaEventRecord(ctx->evStart, ctx->stream));
    IndexState st{ctx->velIdx, ctx->densityIdx, ctx->particleIdx};
    cudaError_t e = enqueue(st);
    if (e != cudaSuccess) return fail(ctx, (int)e, "kernel launch");
    ctx->velIdx = st.vel; ctx->densityIdx = st.density; ctx->particleIdx = st.particle;
    KB_TRY(ctx, cudaEventRecord(ctx->evStop, ctx->stream));
    KB_TRY(ctx, cudaEventSynchronize(ctx->evStop));
    float ms = 0.f;
    KB_TRY(ctx, cudaEventElapsedTime(&ms, ctx->evStart, ctx->evStop));
    accumulator += ms * 0.001f;
    return 0;
}

int checkSim(kamino_ctx* ctx, int sim)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (sim < 0 || sim >= ctx->batch) return fail(ctx, KAMINO_ERR_INVALID, "simulation index out of range");
    return 0;
}

// (re)allocate the particle double buffer and the frame snapshot for n particles per simulation
int allocParticles(kamino_ctx* ctx, long n)
{
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    dropGraphs(ctx);
    if (ctx->particleArena) { cudaFree(ctx->particleArena); ctx->particleArena = nullptr; }
    GridParams& g = ctx->g;
    g.numParticles = n;
    const size_t particleBytes = alignUp(sizeof(float) * 2 * (size_t)n * ctx->batch, 256);
    ctx->snapshotFloats = (g.cells * 3 + 2 * (size_t)n) * ctx->batch;
    const size_t snapshotBytes = alignUp(sizeof(float) * ctx->snapshotFloats, 256);
    KB_TRY(ctx, cudaMalloc((void**)&ctx->particleArena, particleBytes * 2 + snapshotBytes));
    KB_TRY(ctx, cudaMemset(ctx->particleArena, 0, particleBytes * 2 + snapshotBytes));
    KB_TRY(ctx, cudaDeviceSynchronize());          // see the arena clear in kamino_create
    ctx->particles[0] = (float*)ctx->particleArena;
    ctx->particles[1] = (float*)(ctx->particleArena + particleBytes);
    ctx->snapshot = (float*)(ctx->particleArena + 2 * particleBytes);
    ctx->particleIdx = ctx->densityIdx;
    return prepareGraphs(ctx);
}

} // namespace

extern "C" {

const char* kamino_version(void) { return "kamino_b200 0.1 sm_100a"; }

const char* kamino_last_error(const kamino_ctx* ctx)
{
    return ctx ? ctx->lastError.c_str() : g_createError.c_str();
}

int kamino_create(kamino_ctx** out, int device, int nTheta, float radius, float dt, int batch, long particlesPerSim)
{
    if (!out) return fail(nullptr, KAMINO_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (nTheta < 16 || (nTheta & (nTheta - 1)) != 0 || nTheta > 8192)
        return fail(nullptr, KAMINO_ERR_INVALID, "nTheta must be a power of two in [16, 8192]");
    if (batch < 1 || particlesPerSim < 0 || !(radius > 0.f) || !(dt > 0.f))
        return fail(nullptr, KAMINO_ERR_INVALID, "batch >= 1, particles >= 0, radius > 0, dt > 0 required");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, KAMINO_ERR_NO_DEVICE, "no CUDA device available (kamino_b200 has no CPU path)");
    if (device < 0 || device >= count) return fail(nullptr, KAMINO_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    KB_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, KAMINO_ERR_NO_DEVICE, "kamino_b200 is built for sm_100a (B200) only");

    DeviceGuard guard(device);
    kamino_ctx* ctx = new kamino_ctx();
    ctx->device = device;
    ctx->batch = batch;
    GridParams& g = ctx->g;
    g.nTheta = nTheta;
    g.nPhi = 2 * nTheta;
    g.log2NPhi = 0;
    while ((1 << g.log2NPhi) < g.nPhi) ++g.log2NPhi;
    g.radius = radius;
    g.dt = dt;
    g.h = (float)(kPi / (double)nTheta);                  // kernel/KaminoCore.cu:849
    g.invH = (float)(1.0 / (double)g.h);
    g.halfH = 0.5f * g.h;
    g.cofTheta = dt / radius;
    g.cells = (size_t)g.nTheta * g.nPhi;
    g.numParticles = particlesPerSim;
    g.rowBegin = 0;
    g.rowCount = nTheta;

    const size_t fieldBytes = alignUp(sizeof(float) * g.cells * batch, 256);
    const size_t tableBytes = alignUp(spectralTableBytes(g), 256);
    ctx->arenaBytes = fieldBytes * (4 + 2 * ctx->velBuffers) + tableBytes;   // 2 x (u_phi, u_theta), 2 x density, pressure, spectrum
    e = cudaMalloc((void**)&ctx->arena, ctx->arenaBytes);
    if (e != cudaSuccess) { int rc = fail(nullptr, (int)e, "cudaMalloc(arena)"); delete ctx; return rc; }
    // cudaMemset on the legacy stream is asynchronous for device memory and does NOT order against the non-blocking
    // streams of this context: wait for it. (Found at 8192 x 16384, r02f: the 6 GB clear was still running when the
    // table builders started on ownStream and zeroed the first ~60 rows of the LU tables after they had been written;
    // the step was finite but not repeatable. This also explains the non-finite banded run of r01o.)
    cudaMemset(ctx->arena, 0, ctx->arenaBytes);
    cudaDeviceSynchronize();
    char* p = ctx->arena;
    auto take = [&p](size_t bytes) { char* r = p; p += bytes; return r; };
    for (int k = 0; k < ctx->velBuffers; ++k) ctx->velPhi[k] = (float*)take(fieldBytes);
    for (int k = 0; k < ctx->velBuffers; ++k) ctx->velTheta[k] = (float*)take(fieldBytes);
    for (int k = 0; k < 2; ++k) ctx->density[k] = (float*)take(fieldBytes);
    ctx->pressure = (float*)take(fieldBytes);
    ctx->spectrum = (float2*)take(fieldBytes);
    {
        char* t = take(tableBytes);
        auto sub = [&t](size_t bytes) { char* r = t; t += alignUp(bytes, 256); return r; };
        ctx->tables.twiddle = (float2*)sub(sizeof(float2) * g.nPhi);
        ctx->tables.divFactor = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.sinNorth = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.sinSouth = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.gradPhiDenom = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.triA = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.triC = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.sinSq = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.geoG = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.cofPhiCentred = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.cofPhiTheta = (float*)sub(sizeof(float) * nTheta);
        ctx->tables.samplerConsts = (SamplerConsts*)sub(64);
        const size_t slotRows = (size_t)nTheta * (g.nPhi / 2);
        ctx->tables.thL = (float*)sub(sizeof(float) * slotRows);
        ctx->tables.thInvB = (float*)sub(sizeof(float) * slotRows);
        ctx->tables.thBetaInv = (float*)sub(sizeof(float) * slotRows);
        ctx->tables.thH = (float*)sub(sizeof(float) * slotRows);
        ctx->tables.thDelta = (float*)sub(sizeof(float) * slotRows);
        ctx->tables.thBetaEnd = (float*)sub(sizeof(float) * (size_t)(nTheta / 4) * (g.nPhi / 2));
        ctx->tables.minusTwoOverH2 = -2.0 / (double)(g.h * g.h);
    }

    int prioLeast = 0, prioGreatest = 0;
    cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest);
    bool ok = cudaStreamCreateWithPriority(&ctx->ownStream, cudaStreamNonBlocking, prioGreatest) == cudaSuccess
        && cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking) == cudaSuccess
        && cudaEventCreate(&ctx->evStart) == cudaSuccess && cudaEventCreate(&ctx->evStop) == cudaSuccess
        && cudaEventCreateWithFlags(&ctx->evSnap, cudaEventDisableTiming) == cudaSuccess
        && cudaEventCreateWithFlags(&ctx->evCopied, cudaEventDisableTiming) == cudaSuccess;
    ctx->stream = ctx->ownStream;
    if (ok) ok = configureKernels(g, batch) == cudaSuccess;
    if (ok) {
        char block[64];
        fillSamplerConsts(g, block);
        ok = cudaMemcpyAsync(ctx->tables.samplerConsts, block, sizeof(block), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess
            && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    if (ok) ok = launchBuildTables(g, ctx->tables, ctx->stream) == cudaSuccess;
    if (ok) ok = launchBuildSolveTables(g, ctx->tables, batch, ctx->stream) == cudaSuccess;
    if (ok) ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (ok) ok = allocParticles(ctx, particlesPerSim) == 0;
    if (!ok) {
        int rc = fail(nullptr, (int)cudaGetLastError(), "context setup");
        if (rc == 0) rc = fail(nullptr, KAMINO_ERR_STATE, "context setup failed");
        kamino_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return 0;
}

int kamino_destroy(kamino_ctx* ctx)
{
    if (!ctx) return 0;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    dropGraphs(ctx);
    if (ctx->evStart) cudaEventDestroy(ctx->evStart);
    if (ctx->evStop) cudaEventDestroy(ctx->evStop);
    if (ctx->evSnap) cudaEventDestroy(ctx->evSnap);
    if (ctx->evCopied) cudaEventDestroy(ctx->evCopied);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->particleArena) cudaFree(ctx->particleArena);
    if (ctx->bandArena) cudaFree(ctx->bandArena);
    delete ctx;
    return 0;
}

int kamino_alloc_particles(kamino_ctx* ctx, long particlesPerSim)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (particlesPerSim < 0) return fail(ctx, KAMINO_ERR_INVALID, "particlesPerSim < 0");
    return allocParticles(ctx, particlesPerSim);
}

int kamino_set_stream(kamino_ctx* ctx, void* cudaStream)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cudaStream ? (cudaStream_t)cudaStream : ctx->ownStream;
    return 0;
}

int kamino_get_shape(const kamino_ctx* ctx, int* nTheta, int* nPhi, int* batch, long* particlesPerSim)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (nTheta) *nTheta = ctx->g.nTheta;
    if (nPhi) *nPhi = ctx->g.nPhi;
    if (batch) *batch = ctx->batch;
    if (particlesPerSim) *particlesPerSim = ctx->g.numParticles;
    return 0;
}

static int copyField(kamino_ctx* ctx, int field, int sim, void* host, bool toDevice, bool async)
{
    if (int rc = checkSim(ctx, sim)) return rc;
    if (!host) return fail(ctx, KAMINO_ERR_INVALID, "host pointer is NULL");
    float* base = fieldBuffer(ctx, field, 0);
    if (!base) return fail(ctx, KAMINO_ERR_INVALID, "unknown field");
    DeviceGuard guard(ctx->device);
    float* dev = base + (size_t)sim * ctx->g.cells;
    const size_t bytes = sizeof(float) * fieldRows(ctx, field) * ctx->g.nPhi;
    if (toDevice) KB_TRY(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    else KB_TRY(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (!async) KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int copyParticles(kamino_ctx* ctx, int sim, void* host, bool toDevice, bool async)
{
    if (int rc = checkSim(ctx, sim)) return rc;
    if (ctx->g.numParticles == 0) return 0;     // kernel/KaminoParticles.cu:96,105
    if (!host) return fail(ctx, KAMINO_ERR_INVALID, "host pointer is NULL");
    DeviceGuard guard(ctx->device);
    float* dev = ctx->particles[ctx->particleIdx] + (size_t)sim * 2 * ctx->g.numParticles;
    const size_t bytes = sizeof(float) * 2 * (size_t)ctx->g.numParticles;
    if (toDevice) KB_TRY(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    else KB_TRY(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (!async) KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int kamino_upload_field(kamino_ctx* ctx, int field, int sim, const float* host)
{ return copyField(ctx, field, sim, (void*)host, true, false); }
int kamino_download_field(kamino_ctx* ctx, int field, int sim, float* host)
{ return copyField(ctx, field, sim, host, false, false); }
int kamino_upload_particles(kamino_ctx* ctx, int sim, const float* host)
{ return copyParticles(ctx, sim, (void*)host, true, false); }
int kamino_download_particles(kamino_ctx* ctx, int sim, float* host)
{ return copyParticles(ctx, sim, host, false, false); }
int kamino_download_field_async(kamino_ctx* ctx, int field, int sim, float* host)
{ return copyField(ctx, field, sim, host, false, true); }
int kamino_download_particles_async(kamino_ctx* ctx, int sim, float* host)
{ return copyParticles(ctx, sim, host, false, true); }
int kamino_upload_field_async(kamino_ctx* ctx, int field, int sim, const float* host)
{ return copyField(ctx, field, sim, (void*)host, true, true); }
int kamino_upload_particles_async(kamino_ctx* ctx, int sim, const float* host)
{ return copyParticles(ctx, sim, (void*)host, true, true); }

int kamino_field_device_ptr(kamino_ctx* ctx, int field, int sim, int which, void** devicePtr, size_t* pitchInElements)
{
    if (int rc = checkSim(ctx, sim)) return rc;
    float* base = fieldBuffer(ctx, field, which ? 1 : 0);
    if (!base || !devicePtr) return fail(ctx, KAMINO_ERR_INVALID, "unknown field or NULL output");
    *devicePtr = base + (size_t)sim * ctx->g.cells;
    if (pitchInElements) *pitchInElements = (size_t)ctx->g.nPhi;
    return 0;
}

int kamino_particles_device_ptr(kamino_ctx* ctx, int sim, int which, void** devicePtr)
{
    if (int rc = checkSim(ctx, sim)) return rc;
    if (!devicePtr) return fail(ctx, KAMINO_ERR_INVALID, "NULL output");
    *devicePtr = ctx->particles[ctx->particleIdx ^ (which ? 1 : 0)] + (size_t)sim * 2 * ctx->g.numParticles;
    return 0;
}

int kamino_advect(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    return timedPhase(ctx, ctx->advectionTime, [&](IndexState& st) { return enqueueAdvect(ctx, st, ctx->stream); });
}

int kamino_geometric(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    return timedPhase(ctx, ctx->geometricTime, [&](IndexState& st) { return enqueueGeometric(ctx, st, ctx->stream); });
}

int kamino_project(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    return timedPhase(ctx, ctx->projectionTime, [&](IndexState& st) { return enqueueProject(ctx, st, ctx->stream); });
}

// GPU-side initialisers (device_init.cu): every simulation of the batch gets the same field / lattice
int kamino_init_velocity_device(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    DeviceGuard guard(ctx->device);
    for (int sim = 0; sim < ctx->batch; ++sim) {
        cudaError_t e = launchInitVelocity(ctx->g, ctx->velPhi[ctx->velIdx] + (size_t)sim * ctx->g.cells,
                                           ctx->velTheta[ctx->velIdx] + (size_t)sim * ctx->g.cells, 0, ctx->g.nTheta, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, (int)e, "kamino_init_velocity_device");
    }
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int kamino_seed_particles_device(kamino_ctx* ctx, float particleDensity, unsigned long long seed)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (!(particleDensity >= 0.f)) return fail(ctx, KAMINO_ERR_INVALID, "particleDensity < 0");
    DeviceGuard guard(ctx->device);
    for (int sim = 0; sim < ctx->batch; ++sim) {
        cudaError_t e = launchSeedParticles(ctx->g.nTheta, particleDensity, seed,
                                            ctx->particles[ctx->particleIdx] + (size_t)sim * 2 * ctx->g.numParticles, ctx->g.numParticles, ctx->stream);
        if (e != cudaSuccess)
            return fail(ctx, e == cudaErrorInvalidValue ? KAMINO_ERR_STATE : (int)e,
                        "kamino_seed_particles_device (allocate kamino_particle_count(nTheta, particleDensity) particles first)");
    }
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// Parity instrumentation: kamino_project with the theta solve done in the reference's cyclic-reduction order
// (debug_cr.cu). Never part of a step graph.
int kamino_debug_project_cr(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (ctx->g.nTheta > 2048) return fail(ctx, KAMINO_ERR_INVALID, "the reference's cyclic reduction cannot launch above nTheta = 2048");
    return timedPhase(ctx, ctx->projectionTime, [&](IndexState& st) {
        cudaError_t e = enqueueProjectPart(ctx, st, 0, ctx->stream);
        if (e == cudaSuccess) e = launchCyclicReductionDebug(ctx->g, ctx->tables, ctx->spectrum, ctx->batch, ctx->stream);
        if (e == cudaSuccess) e = enqueueProjectPart(ctx, st, 2, ctx->stream);
        return e;
    });
}

// ---- theta-band entry points (band-decomposed multi-GPU runs, kaminogpu_b200/banded.py) ----------

static int bandRange(kamino_ctx* ctx, int rowBegin, int rowCount, int multiple, GridParams* out)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (ctx->batch != 1) return fail(ctx, KAMINO_ERR_STATE, "band entry points need a single-simulation context");
    if (rowBegin < 0 || rowCount <= 0 || rowBegin + rowCount > ctx->g.nTheta || rowBegin % multiple || rowCount % multiple)
        return fail(ctx, KAMINO_ERR_INVALID, "band rows out of range or not a multiple of the kernel's row granularity");
    *out = ctx->g;
    out->rowBegin = rowBegin;
    out->rowCount = rowCount;
    out->numParticles = 0;          // particles are not band-decomposed
    return 0;
}

int kamino_band_advect(kamino_ctx* ctx, int rowBegin, int rowCount)
{
    GridParams g;
    if (int rc = bandRange(ctx, rowBegin, rowCount, 8, &g)) return rc;
    DeviceGuard guard(ctx->device);
    IndexState st{ctx->velIdx, ctx->densityIdx, ctx->particleIdx};
    AdvectArgs a = advectArgs(ctx, st);
    a.particles = nullptr; a.particlesOut = nullptr;
    KB_TRY(ctx, launchAdvect(g, a, 1, ctx->stream));
    ctx->velIdx = nextVel(ctx, ctx->velIdx); ctx->densityIdx ^= 1;
    return 0;
}

int kamino_band_geometric(kamino_ctx* ctx, int rowBegin, int rowCount)
{
    GridParams g;
    if (int rc = bandRange(ctx, rowBegin, rowCount, 8, &g)) return rc;
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, launchGeometric(g, ctx->tables, ctx->velPhi[ctx->velIdx], ctx->velTheta[ctx->velIdx],
                                ctx->velPhi[nextVel(ctx, ctx->velIdx)], ctx->velTheta[nextVel(ctx, ctx->velIdx)], 1, ctx->stream));
    ctx->velIdx = nextVel(ctx, ctx->velIdx);
    return 0;
}

int kamino_band_divergence_fft(kamino_ctx* ctx, int rowBegin, int rowCount)
{
    GridParams g;
    if (int rc = bandRange(ctx, rowBegin, rowCount, 2, &g)) return rc;
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, launchDivergenceFFT(g, ctx->tables, ctx->velPhi[ctx->velIdx], ctx->velTheta[ctx->velIdx],
                                    ctx->spectrum, 1, ctx->stream));
    return 0;
}

int kamino_band_tridiagonal(kamino_ctx* ctx, void* packedSpectrum, int pitch, int slotBegin, int slotCount)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (ctx->batch != 1) return fail(ctx, KAMINO_ERR_STATE, "band entry points need a single-simulation context");
    const int half = ctx->g.nPhi / 2;
    if (!packedSpectrum || pitch < slotCount || slotBegin < 0 || slotCount <= 0 || slotBegin + slotCount > half)
        return fail(ctx, KAMINO_ERR_INVALID, "bad packed spectrum, pitch or slot range");
    DeviceGuard guard(ctx->device);
    cudaError_t e = launchTridiagonalBand(ctx->g, ctx->tables, (float2*)packedSpectrum, pitch, slotBegin, slotCount,
                                          ctx->batch, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, (int)e, "kamino_band_tridiagonal (slot range must be a multiple of 8)");
    return 0;
}

int kamino_band_solver_prepare(kamino_ctx* ctx, int rowBegin, int rowCount)
{
    GridParams g;
    if (int rc = bandRange(ctx, rowBegin, rowCount, 16, &g)) return rc;
    if (rowCount & (rowCount - 1)) return fail(ctx, KAMINO_ERR_INVALID, "the band-local solve needs a power-of-two row count");
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->bandArena) { cudaFree(ctx->bandArena); ctx->bandArena = nullptr; ctx->bandRows = 0; }
    const size_t slotRows = (size_t)rowCount * (ctx->g.nPhi / 2);
    const size_t endFloats = (size_t)(rowCount / 4 + 1) * (ctx->g.nPhi / 2);
    const size_t bytes = alignUp(sizeof(float) * slotRows, 256) * 5 + alignUp(sizeof(float) * endFloats, 256);
    KB_TRY(ctx, cudaMalloc((void**)&ctx->bandArena, bytes));
    char* p = ctx->bandArena;
    auto sub = [&p](size_t b) { char* r = p; p += alignUp(b, 256); return r; };
    ctx->bandTables = ctx->tables;                  // per-row tables are shared; the th* members are the band's own
    ctx->bandTables.thL = (float*)sub(sizeof(float) * slotRows);
    ctx->bandTables.thInvB = (float*)sub(sizeof(float) * slotRows);
    ctx->bandTables.thBetaInv = (float*)sub(sizeof(float) * slotRows);
    ctx->bandTables.thH = (float*)sub(sizeof(float) * slotRows);
    ctx->bandTables.thDelta = (float*)sub(sizeof(float) * slotRows);
    ctx->bandTables.thBetaEnd = (float*)sub(sizeof(float) * endFloats);
    cudaError_t e = launchBuildBandSolveTables(ctx->g, ctx->tables, ctx->bandTables, rowBegin, rowCount, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, (int)e, "kamino_band_solver_prepare");
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bandRowBegin = rowBegin;
    ctx->bandRows = rowCount;
    return 0;
}

int kamino_band_local_solve(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (ctx->bandRows == 0) return fail(ctx, KAMINO_ERR_STATE, "kamino_band_solver_prepare has not been called");
    DeviceGuard guard(ctx->device);
    cudaError_t e = launchBandLocalSolve(ctx->g, ctx->bandTables, ctx->spectrum, ctx->bandRowBegin, ctx->bandRows, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, (int)e, "kamino_band_local_solve");
    return 0;
}

int kamino_tridiagonal_coefficients(kamino_ctx* ctx, int row, float* subDiagonal, float* superDiagonal)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (row < 0 || row >= ctx->g.nTheta) return fail(ctx, KAMINO_ERR_INVALID, "row out of range");
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (subDiagonal) KB_TRY(ctx, cudaMemcpy(subDiagonal, ctx->tables.triA + row, sizeof(float), cudaMemcpyDeviceToHost));
    if (superDiagonal) KB_TRY(ctx, cudaMemcpy(superDiagonal, ctx->tables.triC + row, sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

int kamino_band_inverse_fft_gradient(kamino_ctx* ctx, int rowBegin, int rowCount)
{
    GridParams g;
    if (int rc = bandRange(ctx, rowBegin, rowCount, 1, &g)) return rc;
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, launchInverseFFTGradient(g, ctx->tables, ctx->spectrum, ctx->velPhi[ctx->velIdx],
                                         ctx->velTheta[ctx->velIdx], ctx->pressure, 1, ctx->stream));
    return 0;
}

int kamino_spectrum_device_ptr(kamino_ctx* ctx, int sim, void** devicePtr)
{
    if (int rc = checkSim(ctx, sim)) return rc;
    if (!devicePtr) return fail(ctx, KAMINO_ERR_INVALID, "NULL output");
    *devicePtr = ctx->spectrum + (size_t)sim * (ctx->g.cells >> 1);
    return 0;
}

int kamino_step(kamino_ctx* ctx, int nSteps)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (nSteps < 0) return fail(ctx, KAMINO_ERR_INVALID, "nSteps < 0");
    DeviceGuard guard(ctx->device);
    int remaining = nSteps;
    while (remaining > 0) {
        int take = 1;
        for (int c : kStepChunks)
            if (c <= remaining) { take = c; break; }
        cudaGraphExec_t exec = nullptr;
        if (int rc = getGraph(ctx, take, &exec)) return rc;
        KB_TRY(ctx, cudaGraphLaunch(exec, ctx->stream));
        if (take & 1) { ctx->densityIdx ^= 1; ctx->particleIdx ^= 1; }
        remaining -= take;
    }
    return 0;
}

int kamino_sync(kamino_ctx* ctx)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    DeviceGuard guard(ctx->device);
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    KB_TRY(ctx, cudaStreamSynchronize(ctx->copyStream));
    return 0;
}

int kamino_run_frames(kamino_ctx* ctx, int nFrames, int stepsPerFrame, float* hostVelPhi, float* hostVelTheta,
                      float* hostDensity, float* hostParticles)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (nFrames < 0 || stepsPerFrame < 1) return fail(ctx, KAMINO_ERR_INVALID, "nFrames >= 0, stepsPerFrame >= 1 required");
    DeviceGuard guard(ctx->device);
    const GridParams& g = ctx->g;
    const size_t fieldFloats = g.cells * ctx->batch;
    const size_t particleFloats = 2 * (size_t)g.numParticles * ctx->batch;
    float* snapPhi = ctx->snapshot;
    float* snapTheta = snapPhi + fieldFloats;
    float* snapRho = snapTheta + fieldFloats;
    float* snapPart = snapRho + fieldFloats;
    for (int f = 0; f < nFrames; ++f) {
        if (int rc = kamino_step(ctx, stepsPerFrame)) return rc;
        // device-side snapshot so that the read-back of frame f overlaps the steps of frame f+1
        KB_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evCopied, 0));
        if (hostVelPhi) KB_TRY(ctx, cudaMemcpyAsync(snapPhi, ctx->velPhi[ctx->velIdx], sizeof(float) * fieldFloats, cudaMemcpyDeviceToDevice, ctx->stream));
        if (hostVelTheta) KB_TRY(ctx, cudaMemcpyAsync(snapTheta, ctx->velTheta[ctx->velIdx], sizeof(float) * fieldFloats, cudaMemcpyDeviceToDevice, ctx->stream));
        if (hostDensity) KB_TRY(ctx, cudaMemcpyAsync(snapRho, ctx->density[ctx->densityIdx], sizeof(float) * fieldFloats, cudaMemcpyDeviceToDevice, ctx->stream));
        if (hostParticles && particleFloats) KB_TRY(ctx, cudaMemcpyAsync(snapPart, ctx->particles[ctx->particleIdx], sizeof(float) * particleFloats, cudaMemcpyDeviceToDevice, ctx->stream));
        KB_TRY(ctx, cudaEventRecord(ctx->evSnap, ctx->stream));
        KB_TRY(ctx, cudaStreamWaitEvent(ctx->copyStream, ctx->evSnap, 0));
        // host layout: simulation-major, each field dense (u_theta keeps nTheta-1 rows per simulation)
        if (hostVelPhi) KB_TRY(ctx, cudaMemcpyAsync(hostVelPhi, snapPhi, sizeof(float) * fieldFloats, cudaMemcpyDeviceToHost, ctx->copyStream));
        if (hostVelTheta)
            KB_TRY(ctx, cudaMemcpy2DAsync(hostVelTheta, sizeof(float) * (g.cells - g.nPhi), snapTheta, sizeof(float) * g.cells,
                                          sizeof(float) * (g.cells - g.nPhi), ctx->batch, cudaMemcpyDeviceToHost, ctx->copyStream));
        if (hostDensity) KB_TRY(ctx, cudaMemcpyAsync(hostDensity, snapRho, sizeof(float) * fieldFloats, cudaMemcpyDeviceToHost, ctx->copyStream));
        if (hostParticles && particleFloats) KB_TRY(ctx, cudaMemcpyAsync(hostParticles, snapPart, sizeof(float) * particleFloats, cudaMemcpyDeviceToHost, ctx->copyStream));
        KB_TRY(ctx, cudaEventRecord(ctx->evCopied, ctx->copyStream));
    }
    KB_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    KB_TRY(ctx, cudaStreamSynchronize(ctx->copyStream));
    return 0;
}

int kamino_phase_times(kamino_ctx* ctx, float* advection, float* geometric, float* projection, int reset)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (advection) *advection = ctx->advectionTime;
    if (geometric) *geometric = ctx->geometricTime;
    if (projection) *projection = ctx->projectionTime;
    if (reset) ctx->advectionTime = ctx->geometricTime = ctx->projectionTime = 0.f;
    return 0;
}

int kamino_launches_per_step(const kamino_ctx*) { return kStepKernels; }

int kamino_profile_steps(kamino_ctx* ctx, int nSteps, float* kernelSeconds)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (nSteps < 1 || !kernelSeconds) return fail(ctx, KAMINO_ERR_INVALID, "nSteps >= 1 and an output array required");
    DeviceGuard guard(ctx->device);
    const int nK = kStepKernels;
    std::vector<cudaEvent_t> ev((size_t)nSteps * (nK + 1));
    for (auto& e : ev) KB_TRY(ctx, cudaEventCreate(&e));
    IndexState st{ctx->velIdx, ctx->densityIdx, ctx->particleIdx};
    cudaError_t err = cudaSuccess;
    for (int s = 0; s < nSteps && err == cudaSuccess; ++s) {
        cudaEvent_t* e = &ev[(size_t)s * (nK + 1)];
        err = cudaEventRecord(e[0], ctx->stream);
        for (int k = 0; k < kStepKernels && err == cudaSuccess; ++k) {
            err = enqueueStepKernel(ctx, st, k, ctx->stream);
            if (err == cudaSuccess) err = cudaEventRecord(e[k + 1], ctx->stream);
        }
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    ctx->velIdx = st.vel; ctx->densityIdx = st.density; ctx->particleIdx = st.particle;
    std::vector<double> total(nK, 0.0);
    for (int s = 0; s < nSteps && err == cudaSuccess; ++s) {
        cudaEvent_t* e = &ev[(size_t)s * (nK + 1)];
        for (int k = 0; k < nK && err == cudaSuccess; ++k) {
            float ms = 0.f;
            err = cudaEventElapsedTime(&ms, e[k], e[k + 1]);
            total[k] += ms * 1e-3;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (err != cudaSuccess) return fail(ctx, (int)err, "kamino_profile_steps");
    for (int k = 0; k < nK; ++k) kernelSeconds[k] = (float)total[k];
    return 0;
}

int kamino_debug_locate(kamino_ctx* ctx, int kind, long n, const float* phiRaw, const float* thetaRaw,
                        int32_t* phiIndex, int32_t* thetaIndex, float* alphaPhi, float* alphaTheta,
                        float* phiValidated, float* thetaValidated, int32_t* flags)
{
    if (!ctx) return fail(nullptr, KAMINO_ERR_INVALID, "null context");
    if (kind < 0 || kind > 2 || n < 0) return fail(ctx, KAMINO_ERR_INVALID, "bad kind or n");
    if (n == 0) return 0;
    DeviceGuard guard(ctx->device);
    char* scratch = nullptr;
    const size_t one = alignUp(sizeof(float) * (size_t)n, 256);
    KB_TRY(ctx, cudaMalloc((void**)&scratch, one * 9));
    float* dPhi = (float*)scratch;
    float* dTheta = (float*)(scratch + one);
    int* dPi = (int*)(scratch + 2 * one);
    int* dTi = (int*)(scratch + 3 * one);
    float* dAp = (float*)(scratch + 4 * one);
    float* dAt = (float*)(scratch + 5 * one);
    float* dPv = (float*)(scratch + 6 * one);
    float* dTv = (float*)(scratch + 7 * one);
    int* dFl = (int*)(scratch + 8 * one);
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(dPhi, phiRaw, sizeof(float) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dTheta, thetaRaw, sizeof(float) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = launchLocate(ctx->tables.samplerConsts, kind, n, dPhi, dTheta, dPi, dTi, dAp, dAt, dPv, dTv, dFl, s);
    auto back = [&](void* host, const void* dev) {
        if (e == cudaSuccess && host) e = cudaMemcpyAsync(host, dev, sizeof(float) * n, cudaMemcpyDeviceToHost, s);
    };
    back(phiIndex, dPi); back(thetaIndex, dTi); back(alphaPhi, dAp); back(alphaTheta, dAt);
    back(phiValidated, dPv); back(thetaValidated, dTv); back(flags, dFl);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(scratch);
    if (e != cudaSuccess) return fail(ctx, (int)e, "kamino_debug_locate");
    return 0;
}

int kamino_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return fail(nullptr, KAMINO_ERR_INVALID, "ptr is NULL");
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) return fail(nullptr, (int)e, "cudaMallocHost");
    return 0;
}

int kamino_host_free(void* ptr)
{
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return fail(nullptr, (int)e, "cudaFreeHost");
    return 0;
}

} // extern "C"
