// Shared device/host definitions for the kamino_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kb {

// pi and 2*pi as the reference spells them (include/KaminoHeader.cuh:26-27): doubles.
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
// The fp32 neighbours of those constants from above. For an fp32 x:
//   (double)x > kPi   <=>  x >= kPiF      (kPi lies strictly between two fp32 values)
//   0 <= x < kTwoPiF   =>  floorf((float)((double)x / kTwoPi)) == 0
// (see DESIGN.md "exactness notes").
constexpr float kPiF = 3.14159274101257324f;      // 0x40490FDB
constexpr float kTwoPiF = 6.28318548202514648f;   // 0x40C90FDB

enum SampleKind { kVPhi = 0, kVTheta = 1, kCentered = 2 };

// Per-context constants, passed to every kernel by value (lands in the constant bank).
struct GridParams {
    int nTheta;        // rows of u_phi / density / pressure; u_theta has nTheta - 1
    int nPhi;          // 2 * nTheta, power of two
    int log2NPhi;
    float radius;      // radiusGlobal
    float dt;          // timeStepGlobal
    float h;           // gridLenGlobal = (float)(pi / nTheta)
    float invH;        // (float)(1.0 / (double)h)   (kernel/KaminoCore.cu:42)
    float halfH;       // 0.5f * h (exact)
    float cofTheta;    // dt / radius                (kernel/KaminoCore.cu:204)
    size_t cells;      // nTheta * nPhi : per-simulation stride of every field buffer
    long numParticles; // per simulation
    // theta band this launch works on (rows rowBegin .. rowBegin + rowCount - 1 of the GLOBAL grid;
    // the whole grid unless a kamino_band_* entry point narrowed it). Arrays and indices stay
    // global: a band-decomposed run keeps full-size buffers and only computes its rows.
    int rowBegin, rowCount;
    // experiment (KAMINO_PDL_TAIL=1, default off -> 0): number of trailing blocks of a launch that
    // release the programmatic dependents at their entry; filled per launch by launchChained
    int pdlTail;
};

// Field buffers of the whole batch; simulation b starts at ptr + b * cells.
struct FieldSet {
    const float* velPhi;
    const float* velTheta;
    const float* density;
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The five kernels of a step form a chain in which every kernel consumes what the previous one
// wrote. Inside a captured step graph they are launched with the programmatic-stream-
// serialization attribute: the launch of kernel k+1 is set up while kernel k drains (its blocks
// become schedulable when every block of kernel k has exited or is exiting), and every kernel
// calls pdlWait() before its first dependent global access (read OR write), which blocks until
// kernel k has completed and its writes are visible, so the chain stays transitively ordered.
// Stream capture turns the attribute into programmatic graph edges.
// Measured (r01j A/B): triggering the dependents EARLY (griddepcontrol.launch_dependents at
// kernel entry) is a large loss -- the waiting blocks of kernel k+1 take SM slots from the
// later waves of kernel k -- so no kernel triggers explicitly. Outside graph capture (phase
// entry points, which may follow a memcpy) launches are plain. KAMINO_PDL=0/1 switches it.
#ifdef __CUDACC__
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Tail trigger (experiment, off unless GridParams::pdlTail > 0). Without an explicit trigger the next
// kernel of the chain becomes schedulable only when the LAST block of this one exits; triggering at
// kernel entry for every block was measured as a loss (r01j: the dependents' waiting blocks take SM
// slots from the later waves). Here only the blocks of the last wave trigger, at their entry: by
// then every block of this grid is resident or done, so the dependents' blocks can only take slots
// nobody else needs, run their prologue (table loads, twiddle staging) and park in
// griddepcontrol.wait, which still returns only after this grid has completed and flushed.
__device__ __forceinline__ void pdlTriggerTail(const GridParams& g)
{
    if (g.pdlTail > 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned linear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        if (linear + (unsigned)g.pdlTail >= total) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
}

bool pdlEnabled();
bool pdlTailEnabled();                  // KAMINO_PDL_TAIL=1
void pdlSetCapturing(bool capturing);   // set by the context around graph capture

// Every kernel of the step takes GridParams first; the launcher fills its pdlTail field.
template <typename... KArgs, typename... Args>
cudaError_t launchChained(void (*kernel)(GridParams, KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          GridParams g, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdlEnabled() ? 1 : 0;
    g.pdlTail = 0;
    if (cfg.numAttrs && pdlTailEnabled()) {
        // one wave of this kernel = SMs x resident blocks per SM (occupancy query, cached per kernel)
        static int blocksPerWave = 0;          // one instance per (kernel signature) instantiation and call site
        static const void* cachedFor = nullptr;
        if (cachedFor != (const void*)kernel) {
            int perSm = 0, device = 0, sms = 0;
            cudaGetDevice(&device);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, (int)(block.x * block.y * block.z), smem) != cudaSuccess)
                perSm = 1;
            blocksPerWave = (perSm > 0 ? perSm : 1) * sms;
            cachedFor = (const void*)kernel;
        }
        g.pdlTail = blocksPerWave;
    }
    return cudaLaunchKernelEx(&cfg, kernel, g, KArgs(args)...);
}
#endif

#define KB_CUDA_OK(expr)                                                     \
    do {                                                                     \
        cudaError_t kb_err__ = (expr);                                       \
        if (kb_err__ != cudaSuccess) return (int)kb_err__;                   \
    } while (0)

} // namespace kb
