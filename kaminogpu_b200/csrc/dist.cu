// Theta-band decomposition of ONE simulation over P GPUs (BASELINE config 5): the step loop,
// the halo exchange and the two transposes around the theta solve, in C++ behind the C ABI
// (include/kamino_b200.h, kamino_dist_*), with direct NCCL point-to-point calls.
//
// The reference is single-GPU (device 0 hard-coded, kernel/KaminoSolver.cu:20) and cannot launch its theta
// solve above nTheta = 2048 (kernel/KaminoCore.cu:779-784): this mode has no reference counterpart. Its
// contract is bit-identity with this library's own single-GPU step (same kernels, same arithmetic), which
// the tests check.
//
// Rank r owns theta rows [lo, hi) = [r nTheta/P, (r+1) nTheta/P) and wavenumber slots [r K, (r+1) K),
// K = (nPhi/2)/P. Memory per rank is BAND-SIZED:
//   u_phi, u_theta, density  x2   rows [lo - 24, hi + 24) clipped to the grid (24 = halo)
//   pressure                      same rows
//   specSend   [P][rows][K]       forward FFT output, already in the send layout of the transpose
//   packed     [nTheta][K]        all rows of my wavenumber band: receive buffer, solved in place, sent back
//   specBack   [P][rows+1][K]     solution of my rows (+ the first row of the next band), all wavenumbers
//   tables                        per-row constants, twiddles, and the LU factors of MY wavenumber band only
// Kernels index rows globally; every field pointer handed to them is the allocation minus (first resident
// row) x nPhi, so no kernel knows about the band. A gather that would leave the resident rows (theta-CFL too
// large for the halo) is clamped and flagged by the sampler (sampler.cuh, haloViolation); kamino_dist_sync
// reports it.
//
// One step:
//   1. halo exchange: 24 rows of u_phi, u_theta, density with each neighbour        (ncclSend / ncclRecv)
//   2. advection on [lo-16, hi+16), geometric on [lo-8, hi+8): the extra rows are recomputed instead of
//      exchanged a second time; divergence + FFT on [lo, hi) -> specSend
//   3. transpose: specSend[p] -> rank p's packed[rows of me]                          (P sends + P receives)
//   4. theta solve of my K slots over all rows, in place on `packed`
//   5. transpose back: packed[rows of p, + 1] -> rank p's specBack[me]               (P sends + P receives)
//   6. inverse FFT + gradient on [lo, hi) reading specBack
// Everything of a rank is ordered on one stream. Transport of the two transposes (3, 5):
//   * peer memory (default when the peers' memory can be mapped): the FFT kernel of step 2 stores every block of its
//     spectrum straight into the owner's `packed`, and the theta solve of step 4 stores its solution straight into the
//     owners' `specBack`, over NVLink -- compute and transfer are ONE kernel each, tile by tile; what is left of 3 and 5
//     is a barrier (a one-word NCCL all-reduce). Peer pointers: CUDA IPC mappings of the other ranks' arenas, exchanged
//     through NCCL itself at creation.
//   * NCCL send / recv of the staged buffers (fallback; kamino_dist_transport switches).
// The halos go through NCCL send / recv either way. For P virtual ranks inside one process on one device -- what the
// single-GPU parity tests drive -- kamino_dist_group_step uses sibling pointers / plain device copies instead.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kamino_b200.h"
#include "kamino_kernels.cuh"

using namespace kb;

namespace {

constexpr int kHalo = 24;          // rows exchanged with each neighbour
constexpr int kAdvectExtra = 16;   // rows beyond the band that the advection recomputes
constexpr int kGeoExtra = 8;       // rows beyond the band that the geometric phase recomputes
constexpr int kMaxPeers = 16;      // ranks whose buffers a rank can address (peer-memory transposes)

// NCCL is loaded at run time: single-GPU users of the library need no NCCL installed, and inside a process
// that already carries one (torch) the same instance is shared.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

const NcclApi* ncclApi()
{
    static const NcclApi api = [] {
        NcclApi a;
        const char* env = getenv("KAMINO_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !n[0]) continue;
            a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return a;
        bool ok = true;
        auto sym = [&](const char* name) { void* p = dlsym(a.handle, name); if (!p) ok = false; return p; };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.Send = (decltype(a.Send))sym("ncclSend");
        a.Recv = (decltype(a.Recv))sym("ncclRecv");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.GetVersion = (decltype(a.GetVersion))sym("ncclGetVersion");
        if (!ok) { dlclose(a.handle); a = NcclApi{}; }
        return a;
    }();
    return api.handle ? &api : nullptr;
}

size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

thread_local std::string g_distCreateError;

} // namespace

struct kamino_dist {
    int device = 0, rank = 0, world = 1;
    GridParams g{};
    int lo = 0, hi = 0, rows = 0, half = 0, kper = 0, log2Kper = 0;
    int memLo = 0, memHi = 0;              // resident field rows [memLo, memHi)
    cudaStream_t stream = nullptr;
    char* arena = nullptr;
    size_t arenaBytes = 0;
    // field pointers are VIRTUAL: allocation - memLo * nPhi, so that global row indices address them
    float* velPhi[2]{};
    float* velTheta[2]{};
    float* density[2]{};
    float* pressure = nullptr;
    float2* specSend = nullptr;
    float2* packed = nullptr;
    float2* specBack = nullptr;
    SpectralTables tables{};
    int* haloViolation = nullptr;          // device flag
    int velIdx = 0, densityIdx = 0;
    ncclComm_t comm = nullptr;
    // Peer-memory transposes: the FFT kernel stores every block of its spectrum straight into the owner's `packed`
    // buffer and the theta solve stores its solution straight into the owners' `specBack` buffers, over NVLink (peer
    // pointers: CUDA IPC mappings of the other ranks' arenas, or sibling ranks of the same process); what is left of the
    // two transposes is a barrier each. Tables of world pointers in device memory, read by the kernels.
    bool peerStores = false;               // transposes through peer memory (else NCCL send / recv, or copies)
    bool peersLinked = false;
    float2** peerPackedTable = nullptr;    // device: [world]
    float2** peerBackTable = nullptr;      // device: [world]
    void* ipcBase[kMaxPeers]{};            // mappings to close at destruction
    int* barrierWord = nullptr;            // device: operand of the barrier all-reduce
    std::string peerNote = "not attempted";
    cudaEvent_t ev[6]{};                   // comm timing brackets
    bool timing = false;
    double haloSeconds = 0.0, transposeSeconds = 0.0;
    long timedSteps = 0;
    std::string lastError;
};

namespace {

int fail(kamino_dist* d, int code, const std::string& what)
{
    std::string msg = what;
    if (code > 0 && code < 10000) { msg += ": "; msg += cudaGetErrorString((cudaError_t)code); }
    if (d) d->lastError = msg; else g_distCreateError = msg;
    return code;
}

int failNccl(kamino_dist* d, ncclResult_t r, const char* what)
{
    const NcclApi* n = ncclApi();
    std::string msg = std::string(what) + ": " + (n ? n->GetErrorString(r) : "NCCL unavailable");
    if (d) d->lastError = msg; else g_distCreateError = msg;
    return KAMINO_ERR_STATE;
}

#define KD_TRY(d, expr)                                                          \
    do {                                                                         \
        cudaError_t kd_e__ = (expr);                                             \
        if (kd_e__ != cudaSuccess) return fail((d), (int)kd_e__, #expr);         \
    } while (0)
#define KD_NCCL(d, expr)                                                         \
    do {                                                                         \
        ncclResult_t kd_r__ = (expr);                                            \
        if (kd_r__ != ncclSuccess) return failNccl((d), kd_r__, #expr);          \
    } while (0)

struct DeviceGuard {
    int previous = -1;
    explicit DeviceGuard(int device) { cudaGetDevice(&previous); if (previous != device) cudaSetDevice(device); else previous = -1; }
    ~DeviceGuard() { if (previous >= 0) cudaSetDevice(previous); }
};

int clipLo(const kamino_dist* d, int extra) { return d->lo - extra < 0 ? 0 : d->lo - extra; }
int clipHi(const kamino_dist* d, int extra) { return d->hi + extra > d->g.nTheta ? d->g.nTheta : d->hi + extra; }

float* fieldThis(kamino_dist* d, int field)
{
    switch (field) {
    case KAMINO_VEL_PHI: return d->velPhi[d->velIdx];
    case KAMINO_VEL_THETA: return d->velTheta[d->velIdx];
    case KAMINO_DENSITY: return d->density[d->densityIdx];
    case KAMINO_PRESSURE: return d->pressure;
    default: return nullptr;
    }
}

GridParams rowsOf(const kamino_dist* d, int begin, int end)
{
    GridParams g = d->g;
    g.rowBegin = begin;
    g.rowCount = end - begin;
    return g;
}

SpectrumLayout forwardLayout(const kamino_dist* d);

// phases 2a-2c of the header: advection, geometric, divergence + FFT into the send layout (or straight into the peers)
cudaError_t enqueueToSpectrum(kamino_dist* d)
{
    AdvectArgs a{};
    a.velPhi = d->velPhi[d->velIdx]; a.velTheta = d->velTheta[d->velIdx]; a.density = d->density[d->densityIdx];
    a.velPhiOut = d->velPhi[d->velIdx ^ 1]; a.velThetaOut = d->velTheta[d->velIdx ^ 1]; a.densityOut = d->density[d->densityIdx ^ 1];
    a.cofPhiCentred = d->tables.cofPhiCentred; a.cofPhiTheta = d->tables.cofPhiTheta; a.consts = d->tables.samplerConsts;
    GridParams g = rowsOf(d, clipLo(d, kAdvectExtra), clipHi(d, kAdvectExtra));
    g.numParticles = 0;
    cudaError_t e = launchAdvect(g, a, 1, d->stream);
    if (e != cudaSuccess) return e;
    d->velIdx ^= 1; d->densityIdx ^= 1;
    e = launchGeometric(rowsOf(d, clipLo(d, kGeoExtra), clipHi(d, kGeoExtra)), d->tables, d->velPhi[d->velIdx], d->velTheta[d->velIdx],
                        d->velPhi[d->velIdx ^ 1], d->velTheta[d->velIdx ^ 1], 1, d->stream);
    if (e != cudaSuccess) return e;
    d->velIdx ^= 1;
    const SpectrumLayout lay = forwardLayout(d);
    return launchDivergenceFFT(rowsOf(d, d->lo, d->hi), d->tables, d->velPhi[d->velIdx], d->velTheta[d->velIdx], d->specSend, 1, d->stream, &lay);
}

cudaError_t enqueueSolve(kamino_dist* d)
{
    // my wavenumber band over ALL rows; the tables hold this band only (slot index from 0)
    int log2Rows = 0;
    while ((1 << log2Rows) < d->rows) ++log2Rows;
    const PeerScatter scatter{d->peerBackTable, log2Rows, d->rank, d->kper};
    return launchTridiagonalBand(d->g, d->tables, d->packed, d->kper, 0, d->kper, 1, d->stream, d->peerStores ? &scatter : nullptr);
}

cudaError_t enqueueInverse(kamino_dist* d)
{
    const SpectrumLayout lay{d->lo, d->kper, d->log2Kper, (size_t)(d->rows + 1) * d->kper, nullptr};
    return launchInverseFFTGradient(rowsOf(d, d->lo, d->hi), d->tables, d->specBack, d->velPhi[d->velIdx], d->velTheta[d->velIdx],
                                    d->pressure, 1, d->stream, &lay);
}

// rows of `packed` that rank p needs back: its band plus the first row of the next band (the theta gradient of a
// band's last row reads the pressure spectrum of the row below it)
int backRows(const kamino_dist* d, int p) { return d->rows + (p < d->world - 1 ? 1 : 0); }

int haloExchangeNccl(kamino_dist* d)
{
    const NcclApi* n = ncclApi();
    const size_t N = (size_t)d->g.nPhi, count = (size_t)kHalo * N;
    if (d->world == 1) return 0;
    KD_NCCL(d, n->GroupStart());
    for (int f = 0; f < 3; ++f) {
        float* base = fieldThis(d, f);
        if (d->rank > 0) {
            KD_NCCL(d, n->Send(base + (size_t)d->lo * N, count, ncclFloat, d->rank - 1, d->comm, d->stream));
            KD_NCCL(d, n->Recv(base + (size_t)(d->lo - kHalo) * N, count, ncclFloat, d->rank - 1, d->comm, d->stream));
        }
        if (d->rank < d->world - 1) {
            KD_NCCL(d, n->Send(base + (size_t)(d->hi - kHalo) * N, count, ncclFloat, d->rank + 1, d->comm, d->stream));
            KD_NCCL(d, n->Recv(base + (size_t)d->hi * N, count, ncclFloat, d->rank + 1, d->comm, d->stream));
        }
    }
    KD_NCCL(d, n->GroupEnd());
    return 0;
}

int transposeForwardNccl(kamino_dist* d)
{
    const NcclApi* n = ncclApi();
    const size_t block = (size_t)d->rows * d->kper * 2;       // floats per (rank, rank) block
    KD_NCCL(d, n->GroupStart());
    for (int p = 0; p < d->world; ++p) {
        KD_NCCL(d, n->Send((const float*)d->specSend + p * block, block, ncclFloat, p, d->comm, d->stream));
        KD_NCCL(d, n->Recv((float*)d->packed + p * block, block, ncclFloat, p, d->comm, d->stream));
    }
    KD_NCCL(d, n->GroupEnd());
    return 0;
}

int transposeBackwardNccl(kamino_dist* d)
{
    const NcclApi* n = ncclApi();
    const size_t rowFloats = (size_t)d->kper * 2;
    KD_NCCL(d, n->GroupStart());
    for (int p = 0; p < d->world; ++p) {
        KD_NCCL(d, n->Send((const float*)d->packed + (size_t)p * d->rows * rowFloats, backRows(d, p) * rowFloats, ncclFloat, p, d->comm, d->stream));
        KD_NCCL(d, n->Recv((float*)d->specBack + (size_t)p * (d->rows + 1) * rowFloats, backRows(d, d->rank) * rowFloats, ncclFloat, p, d->comm, d->stream));
    }
    KD_NCCL(d, n->GroupEnd());
    return 0;
}

// all ranks have reached this point of their streams (and their earlier writes into peer memory are complete: a kernel's
// stores are visible device-wide -- and system-wide for peer mappings -- once the kernel has finished, and the all-reduce
// kernel is ordered after it on the stream)
int barrierNccl(kamino_dist* d)
{
    const NcclApi* n = ncclApi();
    KD_NCCL(d, n->AllReduce(d->barrierWord, d->barrierWord, 1, ncclInt, ncclSum, d->comm, d->stream));
    return 0;
}

// upload the two pointer tables (host arrays of `world` device pointers)
int uploadPeerTables(kamino_dist* d, float2* const* packedOf, float2* const* backOf)
{
    KD_TRY(d, cudaMemcpyAsync(d->peerPackedTable, packedOf, sizeof(float2*) * d->world, cudaMemcpyHostToDevice, d->stream));
    KD_TRY(d, cudaMemcpyAsync(d->peerBackTable, backOf, sizeof(float2*) * d->world, cudaMemcpyHostToDevice, d->stream));
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    d->peersLinked = true;
    return 0;
}

// NCCL ranks: exchange CUDA IPC handles of the arenas (through NCCL itself: no other channel is needed), map the peers,
// agree on the outcome. Any failure anywhere leaves every rank on the NCCL transposes.
struct PeerRecord { cudaIpcMemHandle_t handle; unsigned long long offPacked, offBack; };

int linkPeersOverIpc(kamino_dist* d)
{
    const NcclApi* n = ncclApi();
    if (d->world > kMaxPeers) { d->peerNote = "more ranks than peer slots"; return 0; }
    PeerRecord mine{};
    bool ok = cudaIpcGetMemHandle(&mine.handle, d->arena) == cudaSuccess;
    if (!ok) cudaGetLastError();
    mine.offPacked = (unsigned long long)((char*)d->packed - d->arena);
    mine.offBack = (unsigned long long)((char*)d->specBack - d->arena);
    PeerRecord* dev = nullptr;
    KD_TRY(d, cudaMalloc((void**)&dev, sizeof(PeerRecord) * d->world));
    KD_TRY(d, cudaMemcpyAsync(dev + d->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, d->stream));
    ncclResult_t r = n->AllGather(dev + d->rank, dev, sizeof(PeerRecord), ncclChar, d->comm, d->stream);
    std::vector<PeerRecord> all(d->world);
    if (r == ncclSuccess) {
        cudaMemcpyAsync(all.data(), dev, sizeof(PeerRecord) * d->world, cudaMemcpyDeviceToHost, d->stream);
        ok = ok && cudaStreamSynchronize(d->stream) == cudaSuccess;
    } else ok = false;
    cudaFree(dev);
    float2* packedOf[kMaxPeers]{};
    float2* backOf[kMaxPeers]{};
    for (int p = 0; p < d->world && ok; ++p) {
        if (p == d->rank) { packedOf[p] = d->packed; backOf[p] = d->specBack; continue; }
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        d->ipcBase[p] = base;
        packedOf[p] = (float2*)((char*)base + all[p].offPacked);
        backOf[p] = (float2*)((char*)base + all[p].offBack);
    }
    // unanimous?
    int flag = ok ? 1 : 0;
    KD_TRY(d, cudaMemcpyAsync(d->barrierWord, &flag, sizeof(int), cudaMemcpyHostToDevice, d->stream));
    KD_NCCL(d, n->AllReduce(d->barrierWord, d->barrierWord, 1, ncclInt, ncclMin, d->comm, d->stream));
    KD_TRY(d, cudaMemcpyAsync(&flag, d->barrierWord, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    KD_TRY(d, cudaMemsetAsync(d->barrierWord, 0, sizeof(int), d->stream));
    if (!flag) {
        d->peerNote = ok ? "a peer could not map this rank's memory (CUDA IPC)" : "CUDA IPC mapping of a peer's memory failed";
        return 0;
    }
    if (int rc = uploadPeerTables(d, packedOf, backOf)) return rc;
    d->peerStores = true;
    d->peerNote = "CUDA IPC mappings of the peers' arenas";
    return 0;
}

SpectrumLayout forwardLayout(const kamino_dist* d)
{
    if (d->peerStores) return SpectrumLayout{0, d->kper, d->log2Kper, 0, d->peerPackedTable};
    return SpectrumLayout{d->lo, d->kper, d->log2Kper, (size_t)d->rows * d->kper, nullptr};
}

int stepNccl(kamino_dist* d)
{
    const bool t = d->timing;
    if (t) KD_TRY(d, cudaEventRecord(d->ev[0], d->stream));
    if (int rc = haloExchangeNccl(d)) return rc;
    if (t) KD_TRY(d, cudaEventRecord(d->ev[1], d->stream));
    cudaError_t e = enqueueToSpectrum(d);
    if (e != cudaSuccess) return fail(d, (int)e, "advection / geometric / divergence + FFT launch");
    if (t) KD_TRY(d, cudaEventRecord(d->ev[2], d->stream));
    if (int rc = d->peerStores ? barrierNccl(d) : transposeForwardNccl(d)) return rc;
    if (t) KD_TRY(d, cudaEventRecord(d->ev[3], d->stream));
    e = enqueueSolve(d);
    if (e != cudaSuccess) return fail(d, (int)e, "theta solve launch");
    if (t) KD_TRY(d, cudaEventRecord(d->ev[4], d->stream));
    if (int rc = d->peerStores ? barrierNccl(d) : transposeBackwardNccl(d)) return rc;
    if (t) KD_TRY(d, cudaEventRecord(d->ev[5], d->stream));
    e = enqueueInverse(d);
    if (e != cudaSuccess) return fail(d, (int)e, "inverse FFT + gradient launch");
    if (t) {
        KD_TRY(d, cudaEventSynchronize(d->ev[5]));
        float a = 0.f, b = 0.f, c = 0.f;
        KD_TRY(d, cudaEventElapsedTime(&a, d->ev[0], d->ev[1]));
        KD_TRY(d, cudaEventElapsedTime(&b, d->ev[2], d->ev[3]));
        KD_TRY(d, cudaEventElapsedTime(&c, d->ev[4], d->ev[5]));
        d->haloSeconds += a * 1e-3;
        d->transposeSeconds += (b + c) * 1e-3;
        ++d->timedSteps;
    }
    return 0;
}

int checkViolation(kamino_dist* d)
{
    int flag = 0;
    KD_TRY(d, cudaMemcpyAsync(&flag, d->haloViolation, sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    if (flag) {
        KD_TRY(d, cudaMemsetAsync(d->haloViolation, 0, sizeof(int), d->stream));
        return fail(d, KAMINO_ERR_STATE, "a backtrace left the 24-row halo of a theta band (theta-CFL too large for the band "
                                        "decomposition): the result differs from the single-GPU step");
    }
    return 0;
}

} // namespace

extern "C" {

const char* kamino_dist_last_error(const kamino_dist* d) { return d ? d->lastError.c_str() : g_distCreateError.c_str(); }

int kamino_dist_unique_id(void* id128)
{
    if (!id128) return fail(nullptr, KAMINO_ERR_INVALID, "id buffer is NULL");
    const NcclApi* n = ncclApi();
    if (!n) return fail(nullptr, KAMINO_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded; set KAMINO_NCCL_LIB");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return failNccl(nullptr, r, "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int kamino_dist_create(kamino_dist** out, int device, int nTheta, float radius, float dt, int rank, int world, const void* id128)
{
    if (!out) return fail(nullptr, KAMINO_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (nTheta < 64 || (nTheta & (nTheta - 1)) != 0 || nTheta > 8192)
        return fail(nullptr, KAMINO_ERR_INVALID, "nTheta must be a power of two in [64, 8192]");
    if (world < 1 || (world & (world - 1)) != 0 || rank < 0 || rank >= world || !(radius > 0.f) || !(dt > 0.f))
        return fail(nullptr, KAMINO_ERR_INVALID, "world must be a power of two, 0 <= rank < world, radius > 0, dt > 0");
    const int rows = nTheta / world, kper = nTheta / world;     // nPhi / 2 = nTheta wavenumber slots
    if (world > 1 && (rows < 32 || kper % 8))
        return fail(nullptr, KAMINO_ERR_INVALID, "bands need at least 32 rows and a multiple of 8 wavenumbers per rank");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(nullptr, KAMINO_ERR_NO_DEVICE, "no CUDA device available (kamino_b200 has no CPU path)");
    if (device < 0 || device >= count) return fail(nullptr, KAMINO_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    KD_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(nullptr, KAMINO_ERR_NO_DEVICE, "kamino_b200 is built for sm_100a (B200) only");
    if (id128 && world > 1 && !ncclApi())
        return fail(nullptr, KAMINO_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded; set KAMINO_NCCL_LIB");

    DeviceGuard guard(device);
    kamino_dist* d = new kamino_dist();
    d->device = device; d->rank = rank; d->world = world;
    GridParams& g = d->g;
    g.nTheta = nTheta; g.nPhi = 2 * nTheta;
    g.log2NPhi = 0;
    while ((1 << g.log2NPhi) < g.nPhi) ++g.log2NPhi;
    g.radius = radius; g.dt = dt;
    g.h = (float)(kPi / (double)nTheta);
    g.invH = (float)(1.0 / (double)g.h);
    g.halfH = 0.5f * g.h;
    g.cofTheta = dt / radius;
    g.cells = (size_t)g.nTheta * g.nPhi;
    g.numParticles = 0;
    g.rowBegin = 0; g.rowCount = nTheta;
    d->rows = rows; d->half = nTheta; d->kper = kper;
    while ((1 << d->log2Kper) < kper) ++d->log2Kper;
    d->lo = rank * rows; d->hi = d->lo + rows;
    d->memLo = clipLo(d, kHalo); d->memHi = clipHi(d, kHalo);

    const size_t N = (size_t)g.nPhi;
    const size_t fieldBytes = alignUp(sizeof(float) * (size_t)(d->memHi - d->memLo) * N, 256);
    const size_t sendBytes = alignUp(sizeof(float2) * (size_t)world * rows * kper, 256);
    const size_t packedBytes = alignUp(sizeof(float2) * (size_t)nTheta * kper, 256);
    const size_t backBytes = alignUp(sizeof(float2) * (size_t)world * (rows + 1) * kper, 256);
    const size_t slotRows = (size_t)nTheta * kper;
    const size_t tableBytes = alignUp(sizeof(float2) * N, 256) + 10 * alignUp(sizeof(float) * nTheta, 256) + 256
                            + 5 * alignUp(sizeof(float) * slotRows, 256) + alignUp(sizeof(float) * (size_t)(nTheta / 4) * kper, 256) + 256;
    d->arenaBytes = 7 * fieldBytes + sendBytes + packedBytes + backBytes + tableBytes + 1024;
    cudaError_t e = cudaMalloc((void**)&d->arena, d->arenaBytes);
    if (e != cudaSuccess) { int rc = fail(nullptr, (int)e, "cudaMalloc(band arena)"); delete d; return rc; }
    // cudaMemset on the legacy stream is asynchronous for device memory and does NOT order against the non-blocking
    // stream the table builders run on: wait for it (at 8192 x 16384 the 3 GB clear otherwise overtakes the builders
    // and zeroes the first rows of the tables they have just written)
    cudaMemset(d->arena, 0, d->arenaBytes);
    cudaDeviceSynchronize();
    char* p = d->arena;
    auto take = [&p](size_t bytes) { char* r = p; p += alignUp(bytes, 256); return r; };
    const ptrdiff_t shift = (ptrdiff_t)d->memLo * (ptrdiff_t)N;      // virtual base: global row indexing
    auto field = [&]() { return (float*)take(fieldBytes) - shift; };
    for (int k = 0; k < 2; ++k) d->velPhi[k] = field();
    for (int k = 0; k < 2; ++k) d->velTheta[k] = field();
    for (int k = 0; k < 2; ++k) d->density[k] = field();
    d->pressure = field();
    d->specSend = (float2*)take(sendBytes);
    d->packed = (float2*)take(packedBytes);
    d->specBack = (float2*)take(backBytes);
    SpectralTables& t = d->tables;
    t.twiddle = (float2*)take(sizeof(float2) * N);
    t.divFactor = (float*)take(sizeof(float) * nTheta);
    t.sinNorth = (float*)take(sizeof(float) * nTheta);
    t.sinSouth = (float*)take(sizeof(float) * nTheta);
    t.gradPhiDenom = (float*)take(sizeof(float) * nTheta);
    t.triA = (float*)take(sizeof(float) * nTheta);
    t.triC = (float*)take(sizeof(float) * nTheta);
    t.sinSq = (float*)take(sizeof(float) * nTheta);
    t.geoG = (float*)take(sizeof(float) * nTheta);
    t.cofPhiCentred = (float*)take(sizeof(float) * nTheta);
    t.cofPhiTheta = (float*)take(sizeof(float) * nTheta);
    t.samplerConsts = (SamplerConsts*)take(128);
    d->haloViolation = (int*)((char*)t.samplerConsts + 64);
    t.thL = (float*)take(sizeof(float) * slotRows);
    t.thInvB = (float*)take(sizeof(float) * slotRows);
    t.thBetaInv = (float*)take(sizeof(float) * slotRows);
    t.thH = (float*)take(sizeof(float) * slotRows);
    t.thDelta = (float*)take(sizeof(float) * slotRows);
    t.thBetaEnd = (float*)take(sizeof(float) * (size_t)(nTheta / 4) * kper);
    t.minusTwoOverH2 = -2.0 / (double)(g.h * g.h);
    d->peerPackedTable = (float2**)take(sizeof(float2*) * kMaxPeers);
    d->peerBackTable = (float2**)take(sizeof(float2*) * kMaxPeers);
    d->barrierWord = (int*)take(256);

    int prioLeast = 0, prioGreatest = 0;
    cudaDeviceGetStreamPriorityRange(&prioLeast, &prioGreatest);
    bool ok = cudaStreamCreateWithPriority(&d->stream, cudaStreamNonBlocking, prioGreatest) == cudaSuccess;
    for (auto& ev : d->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
    if (ok) ok = configureKernels(g, 1) == cudaSuccess;
    if (ok) {
        char block[64];
        fillSamplerConsts(g, block, d->memLo, d->memHi, world > 1 ? d->haloViolation : nullptr);
        ok = cudaMemcpyAsync(t.samplerConsts, block, sizeof(block), cudaMemcpyHostToDevice, d->stream) == cudaSuccess
            && cudaStreamSynchronize(d->stream) == cudaSuccess;
    }
    if (ok) ok = launchBuildTables(g, t, d->stream) == cudaSuccess;
    if (ok) ok = launchBuildSolveTables(g, t, 1, d->stream, rank * kper, kper) == cudaSuccess;
    if (ok) ok = cudaStreamSynchronize(d->stream) == cudaSuccess;
    if (!ok) {
        int rc = fail(nullptr, (int)cudaGetLastError(), "band context setup");
        if (rc == 0) rc = fail(nullptr, KAMINO_ERR_STATE, "band context setup failed");
        kamino_dist_destroy(d);
        return rc;
    }
    if (id128 && world > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclResult_t r = ncclApi()->CommInitRank(&d->comm, world, id, rank);
        if (r != ncclSuccess) { int rc = failNccl(nullptr, r, "ncclCommInitRank"); kamino_dist_destroy(d); return rc; }
        if (int rc = linkPeersOverIpc(d)) { g_distCreateError = d->lastError; kamino_dist_destroy(d); return rc; }
    }
    *out = d;
    return 0;
}

int kamino_dist_destroy(kamino_dist* d)
{
    if (!d) return 0;
    DeviceGuard guard(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (void* base : d->ipcBase) if (base) cudaIpcCloseMemHandle(base);
    if (d->comm && ncclApi()) ncclApi()->CommDestroy(d->comm);
    for (auto& ev : d->ev) if (ev) cudaEventDestroy(ev);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->arena) cudaFree(d->arena);
    delete d;
    return 0;
}

int kamino_dist_shape(const kamino_dist* d, int* rowBegin, int* rowEnd, int* slotBegin, int* slotEnd, size_t* deviceBytes)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    if (rowBegin) *rowBegin = d->lo;
    if (rowEnd) *rowEnd = d->hi;
    if (slotBegin) *slotBegin = d->rank * d->kper;
    if (slotEnd) *slotEnd = (d->rank + 1) * d->kper;
    if (deviceBytes) *deviceBytes = d->arenaBytes;
    return 0;
}

// rows of `field` a rank owns: [lo, hi), u_theta clipped to its nTheta - 1 rows
static int ownedRows(const kamino_dist* d, int field)
{
    const int end = (field == KAMINO_VEL_THETA && d->hi > d->g.nTheta - 1) ? d->g.nTheta - 1 : d->hi;
    return end - d->lo;
}

int kamino_dist_upload(kamino_dist* d, int field, const float* hostRows)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    float* base = fieldThis(d, field);
    if (!base || !hostRows) return fail(d, KAMINO_ERR_INVALID, "unknown field or NULL host pointer");
    DeviceGuard guard(d->device);
    const size_t N = (size_t)d->g.nPhi;
    KD_TRY(d, cudaMemcpyAsync(base + (size_t)d->lo * N, hostRows, sizeof(float) * ownedRows(d, field) * N, cudaMemcpyHostToDevice, d->stream));
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    return 0;
}

int kamino_dist_download(kamino_dist* d, int field, float* hostRows)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    float* base = fieldThis(d, field);
    if (!base || !hostRows) return fail(d, KAMINO_ERR_INVALID, "unknown field or NULL host pointer");
    DeviceGuard guard(d->device);
    const size_t N = (size_t)d->g.nPhi;
    KD_TRY(d, cudaMemcpyAsync(hostRows, base + (size_t)d->lo * N, sizeof(float) * ownedRows(d, field) * N, cudaMemcpyDeviceToHost, d->stream));
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    return 0;
}

// the reference's FBM initial velocity for this rank's rows, evaluated on the device (device_init.cu)
int kamino_dist_init_velocity_device(kamino_dist* d)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    DeviceGuard guard(d->device);
    cudaError_t e = launchInitVelocity(d->g, d->velPhi[d->velIdx], d->velTheta[d->velIdx], d->lo, d->rows, d->stream);
    if (e != cudaSuccess) return fail(d, (int)e, "kamino_dist_init_velocity_device");
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    return 0;
}

int kamino_dist_step(kamino_dist* d, int nSteps)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    if (nSteps < 0) return fail(d, KAMINO_ERR_INVALID, "nSteps < 0");
    if (d->world > 1 && !d->comm) return fail(d, KAMINO_ERR_STATE, "this rank was created without a communicator: use kamino_dist_group_step");
    DeviceGuard guard(d->device);
    for (int s = 0; s < nSteps; ++s) {
        if (d->world > 1) {
            if (int rc = stepNccl(d)) return rc;
            continue;
        }
        // one band = the whole grid: same phases, the "transposes" are device copies
        cudaError_t e = enqueueToSpectrum(d);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d->packed, d->specSend, sizeof(float2) * (size_t)d->rows * d->kper, cudaMemcpyDeviceToDevice, d->stream);
        if (e == cudaSuccess) e = enqueueSolve(d);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d->specBack, d->packed, sizeof(float2) * (size_t)d->rows * d->kper, cudaMemcpyDeviceToDevice, d->stream);
        if (e == cudaSuccess) e = enqueueInverse(d);
        if (e != cudaSuccess) return fail(d, (int)e, "single-band step");
    }
    return 0;
}

// P virtual ranks of one process on ONE device (created with id128 = NULL): the phases of kamino_dist_step
// with device copies in place of the NCCL calls. What the single-GPU parity tests drive; also the reference
// for what every NCCL call must move.
int kamino_dist_group_step(kamino_dist* const* ranks, int world, int nSteps)
{
    if (!ranks || world < 1) return fail(nullptr, KAMINO_ERR_INVALID, "bad rank array");
    for (int r = 0; r < world; ++r)
        if (!ranks[r] || ranks[r]->world != world || ranks[r]->rank != r || ranks[r]->device != ranks[0]->device || ranks[r]->comm)
            return fail(ranks[0], KAMINO_ERR_INVALID, "group members must be ranks 0 .. world-1 of one device, created without a communicator");
    kamino_dist* d0 = ranks[0];
    DeviceGuard guard(d0->device);
    const size_t N = (size_t)d0->g.nPhi;
    // peer-memory transposes between the virtual ranks: their buffers are plain pointers of this process
    const bool peerStores = d0->peerStores;
    for (int r = 0; r < world; ++r)
        if (ranks[r]->peerStores != peerStores) return fail(d0, KAMINO_ERR_STATE, "group members disagree on the transpose transport");
    if (peerStores && world <= kMaxPeers) {
        float2* packedOf[kMaxPeers]{};
        float2* backOf[kMaxPeers]{};
        for (int r = 0; r < world; ++r) { packedOf[r] = ranks[r]->packed; backOf[r] = ranks[r]->specBack; }
        for (int r = 0; r < world; ++r)
            if (!ranks[r]->peersLinked)
                if (int rc = uploadPeerTables(ranks[r], packedOf, backOf)) return rc;
    }
    auto barrier = [&]() -> cudaError_t {
        for (int r = 0; r < world; ++r) {
            cudaError_t e = cudaStreamSynchronize(ranks[r]->stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    for (int s = 0; s < nSteps; ++s) {
        KD_TRY(d0, barrier());
        for (int r = 0; r + 1 < world; ++r) {               // halos between rank r and r + 1
            kamino_dist *a = ranks[r], *b = ranks[r + 1];
            for (int f = 0; f < 3; ++f) {
                KD_TRY(d0, cudaMemcpyAsync(fieldThis(b, f) + (size_t)(b->lo - kHalo) * N, fieldThis(a, f) + (size_t)(a->hi - kHalo) * N,
                                           sizeof(float) * kHalo * N, cudaMemcpyDeviceToDevice, b->stream));
                KD_TRY(d0, cudaMemcpyAsync(fieldThis(a, f) + (size_t)a->hi * N, fieldThis(b, f) + (size_t)b->lo * N,
                                           sizeof(float) * kHalo * N, cudaMemcpyDeviceToDevice, a->stream));
            }
        }
        KD_TRY(d0, barrier());
        for (int r = 0; r < world; ++r) {
            cudaError_t e = enqueueToSpectrum(ranks[r]);
            if (e != cudaSuccess) return fail(d0, (int)e, "advection / geometric / divergence + FFT launch");
        }
        KD_TRY(d0, barrier());
        const size_t block = (size_t)d0->rows * d0->kper;
        if (!peerStores) {
            for (int r = 0; r < world; ++r)
                for (int p = 0; p < world; ++p)             // r's block for p -> p's packed rows of r
                    KD_TRY(d0, cudaMemcpyAsync(ranks[p]->packed + r * block, ranks[r]->specSend + p * block, sizeof(float2) * block,
                                               cudaMemcpyDeviceToDevice, ranks[p]->stream));
            KD_TRY(d0, barrier());
        }
        for (int r = 0; r < world; ++r) {
            cudaError_t e = enqueueSolve(ranks[r]);
            if (e != cudaSuccess) return fail(d0, (int)e, "theta solve launch");
        }
        KD_TRY(d0, barrier());
        if (!peerStores) {
            for (int r = 0; r < world; ++r)
                for (int p = 0; p < world; ++p)             // r's solution rows of p (+1) -> p's specBack block r
                    KD_TRY(d0, cudaMemcpyAsync(ranks[p]->specBack + (size_t)r * (d0->rows + 1) * d0->kper, ranks[r]->packed + (size_t)p * block,
                                               sizeof(float2) * (size_t)backRows(d0, p) * d0->kper, cudaMemcpyDeviceToDevice, ranks[p]->stream));
            KD_TRY(d0, barrier());
        }
        for (int r = 0; r < world; ++r) {
            cudaError_t e = enqueueInverse(ranks[r]);
            if (e != cudaSuccess) return fail(d0, (int)e, "inverse FFT + gradient launch");
        }
    }
    KD_TRY(d0, barrier());
    return 0;
}

int kamino_dist_sync(kamino_dist* d)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    DeviceGuard guard(d->device);
    KD_TRY(d, cudaStreamSynchronize(d->stream));
    return d->world > 1 ? checkViolation(d) : 0;
}

int kamino_dist_transport(kamino_dist* d, int setPeerStores, int* usesPeerStores, const char** note)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    if (setPeerStores >= 0) {
        const bool want = setPeerStores != 0;
        if (want && d->comm && !d->peersLinked) return fail(d, KAMINO_ERR_STATE, "peer memory is not mapped: " + d->peerNote);
        if (want && d->world > kMaxPeers) return fail(d, KAMINO_ERR_STATE, "more ranks than peer slots");
        DeviceGuard guard(d->device);
        KD_TRY(d, cudaStreamSynchronize(d->stream));
        d->peerStores = want && d->world > 1;
    }
    if (usesPeerStores) *usesPeerStores = d->peerStores ? 1 : 0;
    if (note) *note = d->peerNote.c_str();
    return 0;
}

int kamino_dist_stream(kamino_dist* d, void** cudaStream)
{
    if (!d || !cudaStream) return fail(d, KAMINO_ERR_INVALID, "null argument");
    *cudaStream = (void*)d->stream;
    return 0;
}

int kamino_dist_comm_stats(kamino_dist* d, int enable, double* haloSeconds, double* transposeSeconds, long* steps,
                           size_t* haloBytesPerStep, size_t* transposeBytesPerStep)
{
    if (!d) return fail(nullptr, KAMINO_ERR_INVALID, "null band context");
    if (haloSeconds) *haloSeconds = d->haloSeconds;
    if (transposeSeconds) *transposeSeconds = d->transposeSeconds;
    if (steps) *steps = d->timedSteps;
    const int neighbours = (d->rank > 0) + (d->rank < d->world - 1);
    if (haloBytesPerStep) *haloBytesPerStep = (size_t)neighbours * 3 * kHalo * d->g.nPhi * sizeof(float);          // sent (= received)
    if (transposeBytesPerStep) {
        size_t bytes = 0;
        for (int p = 0; p < d->world; ++p)
            if (p != d->rank) bytes += sizeof(float2) * (size_t)d->kper * (d->rows + backRows(d, p));                // sent to other ranks
        *transposeBytesPerStep = bytes;
    }
    if (enable >= 0 && (enable != 0) != d->timing) {
        d->timing = enable != 0;
        d->haloSeconds = d->transposeSeconds = 0.0;
        d->timedSteps = 0;
    }
    return 0;
}

} // extern "C"
