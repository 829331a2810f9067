#!/usr/bin/env python
"""bench.py -- steps/s and cell-updates/s of the per-timestep solver path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reps R]
                    [--impl b200|reference|cli] [--workload c2|c1|c3|c4|c5]

Workload (BASELINE.json configs[1], "C2"): one 512 x 1024 (theta x phi) simulation with
density advection and 1,048,352 passive tracer particles (particleDensity = 2, the closest
the configKamino.txt grammar gets to 1M), radius 5, dt 0.005, the reference's FBM initial
velocity, the synthetic density of SURVEY.md 8d, rand()-seeded particles. A "step" is one
KaminoSolver::stepForward (advection + particles -> geometric -> projection).

N > 1 (launched under torchrun, one rank per GPU): every rank steps its own independent
simulation of the same shape (ensemble sharding, no data-path collective); `value` is the
steps of all ranks divided by the slowest rank's time ("scaling": "weak"). The line also
carries `ensemble_c4`: BASELINE config 4 as written, 64 independent 256 x 512 simulations
sharded 64/N per rank (strong scaling, simulation-steps/s), and `banded`: BASELINE config 5, one
8192 x 16384 simulation theta-band decomposed over the N ranks (C++ step loop with NCCL halo
exchange and NCCL transposes, kamino_dist_*; ms/step, bytes and GB/s of the transposes).

Timing: W warm-up steps, one untimed K-step call (the same graph chunking as the timed
call; every step graph is pre-instantiated at context creation), then R repetitions of the
K-step window, each bracketed by barrier + synchronize and a CUDA-event pair on the
launching stream. Per repetition the slowest rank counts; `value` is K x simulations / the
MEDIAN repetition (min / max are reported beside it). Clocks are sampled in-process through
NVML (no fork) during the repetitions.

The JSON line (rank 0) carries:
  value / ms_per_step   K steps launched back to back as CUDA graphs, state resident in HBM
  e2e                   the same K steps through the reference-facing frame loop
                        (Kamino::run: upload of the initial state from pinned host memory,
                        10 steps per frame, read-back of u_phi, u_theta, density and the
                        particle coordinates after every frame -- what the reference's
                        writers copy back, kernel/KaminoSolver.cu:301-303,375)
  roofline              the dominant kernel: algorithmic bytes / its CUDA-event duration,
                        against the measured HBM peak and (L2-resident sizes) a measured L2
                        copy bandwidth
  cpu_baseline          the CPU oracle port (oracle/kamino_oracle.c, OpenMP) on a bounded sample
  cold                  per-step time with a 512 MB L2 flush before every step

--impl reference runs the UNMODIFIED reference CUDA sources (oracle/_ref/kamino_ref, built
by oracle/ref_harness/Makefile from /root/reference) on the same workload on the GPU: the
reference has no CPU implementation of this path -- its implementation IS the CUDA build,
so that is what the reference arm times (its own cudaEvent phase timers for `value`, the
same frame loop with its own pageable read-backs for `e2e`). With --gpus N rank 0 starts N
reference processes, one per GPU (the reference is single-device; an ensemble user runs N
copies), and aggregates them the way our arm aggregates its ranks.

--impl cli times the compiled drop-in executable (kaminogpu_b200/kamino, the C++ host
classes over the C ABI) on a configKamino.txt of the same workload with output off.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nTheta, particleDensity, total simulations, description)
    "c1": (128, 200.0, 1, "C1 128x256 + 6,552,200 particles"),
    "c2": (512, 2.0, 1, "C2 512x1024 + 1,048,352 particles"),
    "c3": (2048, 0.0, 1, "C3 2048x4096, no particles"),
    "c4": (256, 0.0, 64, "C4 ensemble of 64 x 256x512, no particles"),
    "c5": (8192, 0.0, 1, "C5 8192x16384 on ONE GPU, no particles, dt 0.0025"),
}
C4_SIMS = 64
# BASELINE config 5. The reference cannot run this size; dt <= 0.0025 keeps the equator rows out of the cubic's overflow
# band; scripts/c5_regime.py: finite and bounded over 12 steps with the reference's own initial field.
C5 = {"nTheta": 8192, "dt": 0.0025, "radius": 5.0, "steps": 5}
RADIUS, DT, STEPS_PER_FRAME = 5.0, 0.005, 10
BYTES_PER_CELL = {"advect": 24, "geometric": 16, "divergence_fft": 12, "tridiagonal": 8, "inverse_fft_gradient": 20}
KERNELS = ["advect", "geometric", "divergence_fft", "tridiagonal", "inverse_fft_gradient"]
BYTES_PER_PARTICLE = 16


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons while the device is under load (B200_PROFILING.md recipe),
    read in-process through NVML every 20 ms: no process is forked inside the timed region.
    Falls back to polling nvidia-smi when NVML cannot be loaded."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._thread = index, [], threading.Event(), None
        self.source = "nvml"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self.source = "nvidia-smi"

    def _sample(self):
        if self._nvml is not None:
            n = self._nvml
            sm = float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
            self.samples.append((sm, self._max, {k for k, bit in self.REASONS.items() if mask & bit}))
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            s = [x.strip() for x in out.strip().split(",")]
            if len(s) >= 6:
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                self.samples.append((float(s[0]), float(s[1]),
                                     {k for k, flag in zip(names, s[2:6]) if flag.lower().startswith("active")}))

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._stop.clear()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        sm = [s[0] for s in self.samples]
        reasons = set()
        for s in self.samples:
            reasons |= s[2]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(s[1] for s in self.samples) if self.samples else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def cpu_baseline(nTheta, particleDensity, budget_s=12.0):
    """The oracle port timed on the host cores: whole steps of the same workload until ~budget_s."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as oa
    import numpy as np
    lib = oa.lib()
    p = oa.params(nTheta, RADIUS, DT)
    u, v = oa.init_velocity(nTheta, RADIUS)
    rho = oa.synthetic_density(nTheta)
    pc = oa.seed_particles(nTheta, particleDensity) if particleDensity > 0 else None
    n = 0 if pc is None else pc.size // 2
    pressure = np.zeros(nTheta * 2 * nTheta, np.float32)
    step = lambda: lib.ko_step(ctypes.byref(p), oa.fptr(u), oa.fptr(v), oa.fptr(rho), oa.fptr(pressure),
                               oa.fptr(pc), ctypes.c_long(n), 0)
    step()                                           # warm-up
    t0 = time.perf_counter()
    done = 0
    while True:
        step()
        done += 1
        el = time.perf_counter() - t0
        if el >= budget_s or done >= 1000:
            break
    cores = int(lib.ko_num_threads())
    return {"value": done / el, "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": "%d full steps of the same %dx%d workload (%d particles) in %.1f s, OpenMP over rows / "
                      "wavenumbers / particles" % (done, nTheta, 2 * nTheta, n, el)}


def run_reference(args, rank, world):
    """The reference's own CUDA build on the same workload. Rank 0 only; with --gpus N it starts one
    reference process per GPU (CUDA_VISIBLE_DEVICES=i: the reference hard-codes device 0,
    kernel/KaminoSolver.cu:20) and aggregates them as our arm aggregates its ranks: all steps / the
    slowest process."""
    if rank != 0:
        return 0
    nTheta, pdens, batch, desc = WORKLOADS[args.workload]
    exe = os.path.join(ROOT, "oracle", "_ref", "kamino_ref")
    n = max(args.gpus, 1)
    base = {"impl": "reference", "metric": "sim steps/s", "unit": "steps/s", "higher_is_better": True,
            "n_gpus": n, "steps": args.steps, "warmup": args.warmup}
    if not os.path.exists(exe):
        print(json.dumps(dict(base, unavailable="oracle/_ref/kamino_ref not built (run __graft_entry__.build() "
                                               "where /root/reference is mounted)")))
        return 0
    if batch != 1:
        print(json.dumps(dict(base, unavailable="the reference runs one simulation per process")))
        return 0
    pd = max(pdens, 1.0)      # the reference aborts on an empty particle set
    cmd = [exe, "bench", str(nTheta), str(pd), str(DT), str(RADIUS), str(args.steps), str(STEPS_PER_FRAME), str(args.warmup)]
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    devices = visible.split(",") if visible else [str(i) for i in range(n)]
    with ClockSampler(0) as clocks:
        procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                  env=dict(os.environ, CUDA_VISIBLE_DEVICES=devices[i % len(devices)])) for i in range(n)]
        outs = [p.communicate(timeout=3000) for p in procs]
    lines = []
    for out, err in outs:
        got = None
        for l in out.splitlines():
            if l.startswith('{"ref_bench"'):
                got = json.loads(l)
        if got is None:
            print(json.dumps(dict(base, unavailable="reference run failed: " + (err.strip().splitlines() or ["?"])[-1][:200])))
            return 0
        lines.append(got)
    line = lines[0]
    cells = nTheta * 2 * nTheta
    v = n * min(l["steps_per_s"] for l in lines)
    e2e = n * min(l["e2e_steps_per_s"] for l in lines)
    frame_bytes = line["d2h_bytes_per_frame"]
    res = dict(base)
    res.update({
        "value": v, "ms_per_step": 1e3 * n / v, "cell_updates_per_s": v * cells, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + (" per GPU" if n > 1 else "") + " (reference CUDA build, sm_100, its own cudaEvent phase timers)",
                   "nTheta": nTheta, "nPhi": 2 * nTheta, "particles": line["particles"], "dt": DT, "radius": RADIUS,
                   "parallelism": "%d independent reference processes, one per GPU" % n if n > 1 else "single simulation"},
        "e2e": {"value": e2e, "unit": "steps/s",
                "h2d_bytes_per_step": frame_bytes / args.steps, "d2h_bytes_per_step": frame_bytes / STEPS_PER_FRAME},
        "phases_s": {"advection": line["advection_s"], "geometric": line["geometric_s"], "projection": line["projection_s"]},
        "per_process_steps_per_s": [l["steps_per_s"] for l in lines],
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": 0, "kind": "reference",
                         "sample": "the reference has no CPU path: its own CUDA build, %d steps on the GPU" % args.steps},
        "gpu_launches": 18 * args.steps * n,
        "clocks": clocks.summary(),
    })
    print(json.dumps(res))
    return 0


def run_cli(args, rank, world):
    """The compiled drop-in (kaminogpu_b200/kamino: C++ host classes over the C ABI) on a configKamino.txt
    of the same workload, grid / particle output off ("null"): steps/s from its own `Time spent` line
    (KaminoTimer around the frame loop, kernel/KaminoCore.cu:883-909)."""
    if rank != 0:
        return 0
    nTheta, pdens, batch, desc = WORKLOADS[args.workload]
    exe = os.path.join(ROOT, "kaminogpu_b200", "kamino")
    base = {"impl": "cli", "metric": "sim steps/s", "unit": "steps/s", "higher_is_better": True, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup}
    if batch != 1 or not os.path.exists(exe):
        print(json.dumps(dict(base, unavailable="kamino executable missing or ensemble workload")))
        return 0
    frames = max(args.steps // STEPS_PER_FRAME, 1)
    frame_dt = DT * (STEPS_PER_FRAME - 0.5)         # the reference's loop then takes STEPS_PER_FRAME steps per frame
    with tempfile.TemporaryDirectory() as tmp:
        cfg = os.path.join(tmp, "configKamino.txt")
        with open(cfg, "w") as f:
            f.write("%g %d %g %g %.9g %d 0.0 1 1 1 1 null null null null null\n" % (RADIUS, nTheta, max(pdens, 1.0), DT, frame_dt, frames))
        out = subprocess.run([exe, cfg], capture_output=True, text=True, timeout=3000, cwd=tmp)
    ms = None
    for l in out.stdout.splitlines():
        if l.startswith("Time spent:"):
            ms = float(l.split(":")[1].strip().rstrip("ms"))
    if ms is None:
        print(json.dumps(dict(base, unavailable="kamino run failed: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200])))
        return 0
    sys.path.insert(0, ROOT)
    from kaminogpu_b200.solver import steps_per_frame
    steps = sum(steps_per_frame(DT, frame_dt, frames))
    v = steps / (ms * 1e-3)
    print(json.dumps(dict(base, value=v, ms_per_step=1e3 / v, steps=steps, frames=frames, dtype="f32", data="synthetic",
                          config={"workload": desc + " (kamino executable, configKamino.txt, output off)", "nTheta": nTheta,
                                  "nPhi": 2 * nTheta, "dt": DT, "radius": RADIUS})))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--reps", type=int, default=30, help="repetitions of the timed K-step window (median reported)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "cli"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the secondary C4 (64 x 256x512 sharded) measurement")
    ap.add_argument("--no-banded", action="store_true", help="skip the secondary C5 (8192x16384 theta-band) measurement")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.impl == "cli":
        return run_cli(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    from kaminogpu_b200 import capi
    from kaminogpu_b200.solver import KaminoSolver

    if os.environ.get("KAMINO_DEBUG_STEP_MASK"):
        raise SystemExit("bench.py: KAMINO_DEBUG_STEP_MASK is set (timing instrumentation that skips kernels); refusing to run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- kaminogpu_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(values):
        if world == 1:
            return [float(x) for x in values]
        t = torch.tensor(list(values), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    nTheta, pdens, batch, desc = WORKLOADS[args.workload]
    nPhi = 2 * nTheta
    cells = nTheta * nPhi
    K, W, R = args.steps, max(args.warmup, 3), max(args.reps, 1)
    lib = capi.load()
    stream = torch.cuda.Stream()

    def synthetic_density(nT):
        rho = np.empty((nT, 2 * nT), np.float32)
        jj, ii = np.meshgrid(np.arange(nT), np.arange(2 * nT), indexing="ij")
        h = np.float32(np.pi / nT)
        rho[:] = 0.5 + 0.5 * np.sin(4.0 * ii * float(h)) * np.sin((jj + 0.5) * float(h)) ** 2
        return rho

    def timed_windows(solver, reps, sampler=None):
        """reps x (barrier, event, K steps, event, barrier) -> per-repetition seconds, max over ranks."""
        solver.stepForward(DT, nSteps=K)            # untimed: the chunking of the timed call, graphs resident
        barrier()
        pairs = []
        ctxm = sampler if sampler is not None else _Null()
        with ctxm:
            for _ in range(reps):
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                solver.stepForward(DT, nSteps=K)
                b.record(stream)
                barrier()
                pairs.append((a, b))
        return max_over_ranks([a.elapsed_time(b) * 1e-3 for a, b in pairs])

    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    dt_run = C5["dt"] if args.workload == "c5" else DT
    s = KaminoSolver(nPhi, nTheta, RADIUS, dt_run, device=local_rank, batch=batch)
    s.set_stream(stream.cuda_stream)           # so that torch.cuda.Event brackets our launches
    rho0 = synthetic_density(nTheta)
    for sim in range(batch):
        q = s.quantity(capi.DENSITY, sim)
        q.cpuBuffer[:] = rho0
        q.copyToGPU()
    if pdens > 0:
        s.initParticlesfromPic("", pdens)
    nPart = s.particles.numOfParticles if s.particles is not None else 0

    # pinned host mirrors for the end-to-end frame loop
    def pinned(n):
        p = ctypes.c_void_p()
        capi.check(lib.kamino_host_alloc(ctypes.byref(p), max(4 * n, 4)))
        return p
    hU, hV, hR = pinned(cells * batch), pinned((cells - nPhi) * batch), pinned(cells * batch)
    hP = pinned(2 * nPart * batch)
    pinU, pinV, pinR = pinned(cells), pinned(cells), pinned(cells)
    pinP = pinned(2 * nPart)
    ctypes.memmove(pinU, s.velPhi.cpuBuffer.ctypes.data, 4 * cells)
    ctypes.memmove(pinV, s.velTheta.cpuBuffer.ctypes.data, 4 * (cells - nPhi))
    ctypes.memmove(pinR, rho0.ctypes.data, 4 * cells)
    if nPart:
        ctypes.memmove(pinP, s.particles.coordCPUBuffer.ctypes.data, 8 * nPart)

    def upload_state():
        for sim in range(batch):
            capi.check(lib.kamino_upload_field_async(s._ctx, capi.VEL_PHI, sim, pinU), s._ctx)
            capi.check(lib.kamino_upload_field_async(s._ctx, capi.VEL_THETA, sim, pinV), s._ctx)
            capi.check(lib.kamino_upload_field_async(s._ctx, capi.DENSITY, sim, pinR), s._ctx)
            if nPart:
                capi.check(lib.kamino_upload_particles_async(s._ctx, sim, pinP), s._ctx)
    state_bytes = batch * 4 * (cells + (cells - nPhi) + cells + 2 * nPart)

    clocks = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        # ---- device-resident: W warm-up steps, one untimed K-step call, then R timed K-step windows ----
        s.stepForward(DT, nSteps=W)
        reps = timed_windows(s, R, clocks)
        seconds = statistics.median(reps)
        if len(clocks.samples) < 10:
            # the repetitions lasted only a few milliseconds: keep sampling over an untimed repetition of the
            # same load for half a second so that the clocks line has samples under load
            with clocks:
                t_load = time.perf_counter()
                while time.perf_counter() - t_load < 0.5:
                    s.stepForward(DT, nSteps=K)
                    stream.synchronize()
        upload_state()                      # thousands of extra steps later: back to the initial state
        s.stepForward(DT, nSteps=W)
        stream.synchronize()

        # ---- per-kernel CUDA-event times over K steps (same state, launched individually) ----
        nlaunch = int(lib.kamino_launches_per_step(s._ctx))
        kernels = KERNELS + (["advect_particles"] if nlaunch == len(KERNELS) + 1 else [])
        kern = (ctypes.c_float * len(kernels))()
        nprof = min(max(K, 50), 200)
        capi.check(lib.kamino_profile_steps(s._ctx, nprof, kern), s._ctx)
        kernel_s = {k: kern[i] / nprof for i, k in enumerate(kernels)}

        # ---- cold: 512 MB L2 flush before every step, each step timed on its own ----------
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        ncold = min(max(K, 20), 50)
        cold = []
        for _ in range(ncold):
            flush.add_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            s.stepForward(DT, nSteps=1)
            b.record(stream)
            stream.synchronize()
            cold.append(a.elapsed_time(b) * 1e-3)
        del flush
        cold_s = statistics.median(cold)

        # ---- measured copy bandwidths that bound the numbers above --------------------------
        def copy_gbs(nbytes, reps_, src=None, dst=None):
            a_ = src if src is not None else torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            b_ = dst if dst is not None else torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            for _ in range(3):
                b_.copy_(a_, non_blocking=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps_):
                b_.copy_(a_, non_blocking=True)
            e1.record(stream)
            stream.synchronize()
            return reps_ * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
        l2_gbs = 2.0 * max(copy_gbs(16 << 20, 200) for _ in range(3))   # 16 MB -> 16 MB, both L2-resident: read + write bytes, best of 3
        host_buf = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
        dev_buf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        barrier()                                        # all ranks copy at the same time: the host side is shared
        d2h_gbs = copy_gbs(64 << 20, 5, src=dev_buf, dst=host_buf)
        barrier()
        h2d_gbs = copy_gbs(64 << 20, 5, src=host_buf, dst=dev_buf)
        del host_buf, dev_buf
        d2h_gbs = -max_over_ranks([-d2h_gbs])[0]         # slowest rank
        h2d_gbs = -max_over_ranks([-h2d_gbs])[0]

        # ---- end to end: upload + frames of STEPS_PER_FRAME steps + read-backs --------------
        frames = max(K // STEPS_PER_FRAME, 1)
        upload_state()
        capi.check(lib.kamino_run_frames(s._ctx, 2, STEPS_PER_FRAME, hU, hV, hR, hP if nPart else None), s._ctx)   # warm-up
        e2e_reps = []
        for _ in range(min(R, 10)):
            barrier()
            t0 = time.perf_counter()
            upload_state()
            capi.check(lib.kamino_run_frames(s._ctx, frames, STEPS_PER_FRAME, hU, hV, hR, hP if nPart else None), s._ctx)
            barrier()
            e2e_reps.append(time.perf_counter() - t0)
        e2e_reps = max_over_ranks(e2e_reps)
        e2e_seconds = statistics.median(e2e_reps)
        e2e_steps = frames * STEPS_PER_FRAME

    # finite check on what came back (a diverged run would be meaningless)
    back = np.ctypeslib.as_array(ctypes.cast(hU, ctypes.POINTER(ctypes.c_float)), shape=(cells * batch,))
    finite = bool(np.isfinite(back).all())
    cold_s = max_over_ranks([cold_s])[0]
    s.close()

    # ---- secondary: BASELINE config 4 as written, 64 x (256 x 512) sharded 64 / N per rank (strong scaling) ----
    ensemble = None
    if not args.no_ensemble and args.workload == "c2" and C4_SIMS % world == 0:
        nT4 = 256
        per_rank = C4_SIMS // world
        e = KaminoSolver(2 * nT4, nT4, RADIUS, DT, device=local_rank, batch=per_rank)
        e.set_stream(stream.cuda_stream)
        rho4 = synthetic_density(nT4)
        for sim in range(per_rank):
            q = e.quantity(capi.DENSITY, sim)
            q.cpuBuffer[:] = rho4
            q.copyToGPU()
        with torch.cuda.stream(stream):
            e.stepForward(DT, nSteps=W)
            ereps = timed_windows(e, min(R, 10))
        e.close()
        esec = statistics.median(ereps)
        ensemble = {"workload": "C4: %d independent 256x512 simulations, %d per GPU, no collective" % (C4_SIMS, per_rank),
                    "value": K * C4_SIMS / esec, "unit": "simulation-steps/s", "scaling": "strong",
                    "ms_per_ensemble_step": esec / K * 1e3, "sims_per_gpu": per_rank,
                    "cell_updates_per_s": K * C4_SIMS / esec * nT4 * 2 * nT4, "reps": len(ereps)}

    # ---- secondary: BASELINE config 5, ONE 8192 x 16384 simulation theta-band decomposed over the N ranks ----------
    # (kamino_dist_*: C++ step loop, NCCL halo exchange + NCCL transposes around the theta solve; N = 1 runs the same
    # kernels on one band = the whole grid, so that the N = 1, 2, 4, 8 lines of a scaling run are comparable)
    banded = None
    if not args.no_banded and args.workload == "c2":
        from kaminogpu_b200.dist import DistributedSolver
        nT5, steps5 = C5["nTheta"], C5["steps"]
        d = DistributedSolver(nT5, C5["radius"], C5["dt"], device=local_rank)
        # synthetic smooth initial field of the reference's amplitude (its FBM initialiser is a serial host loop that
        # costs about two core-minutes at this size; kamino_init_velocity_host_rows provides it per band when wanted)
        t_init = time.perf_counter()
        h5 = np.pi / nT5
        th_u = ((np.arange(d.lo, d.hi) + 0.5) * h5).astype(np.float32)[:, None]
        th_v = ((np.arange(d.lo, d.lo + d.rows_of(capi.VEL_THETA)) + 1.0) * h5).astype(np.float32)[:, None]
        ph_u = ((np.arange(2 * nT5) - 0.5) * h5).astype(np.float32)[None, :]
        ph_v = (np.arange(2 * nT5) * h5).astype(np.float32)[None, :]
        d.upload(capi.VEL_PHI, 0.1 * np.sin(th_u) * np.cos(4 * ph_u) + 0.05 * np.sin(3 * th_u) * np.sin(7 * ph_u))
        d.upload(capi.VEL_THETA, 0.1 * np.sin(2 * th_v) * np.sin(3 * ph_v) + 0.03 * np.sin(5 * th_v) * np.cos(11 * ph_v))
        d.upload(capi.DENSITY, 0.5 + 0.5 * np.sin(4 * ph_v) * np.sin(th_u) ** 2)
        t_init = time.perf_counter() - t_init
        ext = torch.cuda.ExternalStream(d.cuda_stream)
        peer5, note5 = d.transport()
        d.step(3)
        d.sync()
        breps = []
        for _ in range(3):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ext)
            d.step(steps5)
            b.record(ext)
            d.sync()
            barrier()
            breps.append(a.elapsed_time(b) * 1e-3)
        breps = max_over_ranks(breps)
        d.comm_stats(enable=1)
        d.step(steps5)
        d.sync()
        cs = d.comm_stats(enable=0)
        back = d.download(capi.VEL_PHI)
        fin5 = max_over_ranks([0.0 if np.isfinite(back).all() else 1.0])[0] == 0.0
        tsec = max_over_ranks([cs["transpose_s"] / max(cs["steps"], 1), cs["halo_s"] / max(cs["steps"], 1)])
        mem = max_over_ranks([float(d.device_bytes)])[0]
        d.close()
        bsec = statistics.median(breps) / steps5
        banded = {"workload": "C5: one %dx%d simulation, theta-band decomposed over %d GPU%s" % (nT5, 2 * nT5, world, "s" if world > 1 else ""),
                  "ms_per_step": bsec * 1e3, "steps_per_s": 1.0 / bsec, "cell_updates_per_s": nT5 * 2 * nT5 / bsec, "scaling": "strong",
                  "steps": steps5, "reps": len(breps), "finite": fin5, "dt": C5["dt"], "radius": C5["radius"],
                  "initial_field": "analytic (two low wavenumber modes per component, max |u| 0.15)",
                  "device_gb_per_rank": mem / 1e9, "init_s": t_init,
                  "collectives": "none (one band)" if world == 1 else
                                 ("NCCL send/recv of 24-row halos; the two transposes of the half spectrum are peer-memory stores of the FFT and "
                                  "theta-solve kernels over NVLink (%s) + one NCCL barrier each" % note5 if peer5 else
                                  "NCCL send/recv: 24-row halos + two transposes of the half spectrum per step (%s)" % note5),
                  "peer_memory_transposes": peer5,
                  "transpose_bytes_sent_per_rank_per_step": cs["transpose_bytes_per_step"],
                  "halo_bytes_sent_per_rank_per_step": cs["halo_bytes_per_step"],
                  "transpose_ms_per_step": tsec[0] * 1e3, "halo_ms_per_step": tsec[1] * 1e3,
                  "transpose_note": "time between the kernels around each transpose: the barrier when the kernels store into peer memory, the NCCL send/recv otherwise",
                  "transpose_gbs_per_rank": (cs["transpose_bytes_per_step"] / tsec[0] / 1e9) if world > 1 and tsec[0] > 0 and not peer5 else None}

    if rank == 0:
        sims = batch * world
        value = K * sims / seconds
        frame_bytes = batch * 4 * (cells + (cells - nPhi) + cells + 2 * nPart)
        peak, peak_src = measured_peaks()
        dom = max(kernel_s, key=kernel_s.get)
        alg = {k: BYTES_PER_CELL[k] * cells * batch for k in KERNELS}
        if "advect_particles" in kernel_s:
            alg["advect_particles"] = BYTES_PER_PARTICLE * nPart * batch     # own kernel on a parallel graph branch
        else:
            alg["advect"] += BYTES_PER_PARTICLE * nPart * batch
        achieved = alg[dom] / kernel_s[dom] / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(dom)
        l2_resident = state_bytes * 2 < 100e6
        res = {
            "metric": "sim steps/s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": seconds / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "reps": {"n": len(reps), "statistic": "median of per-repetition max over ranks",
                     "ms_per_step_min": min(reps) / K * 1e3, "ms_per_step_max": max(reps) / K * 1e3},
            "cell_updates_per_s": value * cells * batch,
            "config": {"workload": desc + (" per GPU" if world > 1 else ""), "nTheta": nTheta, "nPhi": nPhi,
                       "particles": nPart, "batch_per_gpu": batch, "dt": dt_run, "radius": RADIUS,
                       "parallelism": "ensemble x%d (independent simulations, no collective)" % world if world > 1 else "single simulation",
                       "l2": "state (%.0f MB) is smaller than the 126 MB L2: consecutive steps of one simulation run "
                             "L2-resident by nature, so `value` is NOT flushed; `cold` flushes L2 (512 MB write) before every step"
                             % (state_bytes * 2 / 1e6) if l2_resident else
                             "state (%.0f MB) is larger than the 126 MB L2: every step streams from HBM" % (state_bytes * 2 / 1e6)},
            "e2e": {"value": e2e_steps * sims / e2e_seconds, "unit": "steps/s",
                    "h2d_bytes_per_step": state_bytes / e2e_steps, "d2h_bytes_per_step": frame_bytes / STEPS_PER_FRAME,
                    "steps_per_frame": STEPS_PER_FRAME, "frames": frames, "reps": len(e2e_reps),
                    "pcie_d2h_gbs": d2h_gbs, "pcie_h2d_gbs": h2d_gbs,
                    "pcie_note": "64 MB pinned copies, all %d ranks at the same time, slowest rank" % world,
                    "d2h_floor_ms_per_step": frame_bytes / STEPS_PER_FRAME / (d2h_gbs * 1e9) * 1e3},
            "cold": {"value": sims / cold_s, "unit": "steps/s", "ms_per_step": cold_s * 1e3, "steps": ncold},
            "gpu_launches": nlaunch * K,
            "kernel_us": {k: kernel_s[k] * 1e6 for k in kernel_s},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes": alg[dom],
                         "l2_copy_gbs": l2_gbs, "frac_l2": achieved / l2_gbs,
                         "note": ("algorithmic bytes / CUDA-event time of the kernel launched alone in stream order; "
                                  "working set L2-resident at this size: frac_l2 is against the measured L2 copy bandwidth "
                                  "(16 MB device-to-device copy, read + write bytes)") if l2_resident else
                                 "algorithmic bytes / CUDA-event time of the kernel"},
            "roofline_all": {k: {"achieved": alg[k] / kernel_s[k] / 1e9, "frac": alg[k] / kernel_s[k] / 1e9 / peak,
                                 "frac_l2": alg[k] / kernel_s[k] / 1e9 / l2_gbs} for k in kernel_s},
            "clocks": clocks.summary(),
            "finite": finite,
        }
        if cold_s * K < seconds * 0.98:
            res["warning"] = "cold step faster than the warm median: timing defect?"
        if ensemble is not None:
            res["ensemble_c4"] = ensemble
        if banded is not None:
            res["banded"] = banded
        if not args.no_cpu_baseline and world == 1:
            res["cpu_baseline"] = cpu_baseline(nTheta, pdens)
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
