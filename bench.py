#!/usr/bin/env python
"""bench.py -- steps/s and cell-updates/s of the per-timestep solver path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c1|c3|c4]

Workload (BASELINE.json configs[1], "C2"): one 512 x 1024 (theta x phi) simulation with
density advection and 1,048,352 passive tracer particles (particleDensity = 2, the closest
the configKamino.txt grammar gets to 1M), radius 5, dt 0.005, the reference's FBM initial
velocity, the synthetic density of SURVEY.md 8d, rand()-seeded particles. A "step" is one
KaminoSolver::stepForward (advection + particles -> geometric -> projection).

N > 1 (launched under torchrun, one rank per GPU): every rank steps its own independent
simulation of the same shape (ensemble sharding, no data-path collective); `value` is the
steps of all ranks divided by the slowest rank's time ("scaling": "weak").

The JSON line (rank 0) carries:
  value / ms_per_step   K steps launched back to back as CUDA graphs, state resident in HBM,
                        CUDA events on the launching stream, max over ranks
  e2e                   the same K steps through the reference-facing frame loop
                        (Kamino::run: upload of the initial state from pinned host memory,
                        10 steps per frame, read-back of u_phi, u_theta, density and the
                        particle coordinates after every frame -- what the reference's
                        writers copy back, kernel/KaminoSolver.cu:301-303,375)
  roofline              the dominant kernel: algorithmic bytes / its CUDA-event duration
  cpu_baseline          the CPU oracle port (oracle/kamino_oracle.c, OpenMP) on a bounded sample
  cold                  per-step time with a 512 MB L2 flush before every step

--impl reference runs the UNMODIFIED reference CUDA sources (oracle/_ref/kamino_ref, built
by oracle/ref_harness/Makefile from /root/reference) on the same workload on the GPU: the
reference has no CPU implementation of this path -- its implementation IS the CUDA build,
so that is what the reference arm times (its own cudaEvent phase timers for `value`, the
same frame loop with its own pageable read-backs for `e2e`).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nTheta, particleDensity, batch, description)
    "c1": (128, 200.0, 1, "C1 128x256 + 6,552,200 particles"),
    "c2": (512, 2.0, 1, "C2 512x1024 + 1,048,352 particles"),
    "c3": (2048, 0.0, 1, "C3 2048x4096, no particles"),
    "c4": (256, 0.0, 64, "C4 ensemble of 64 x 256x512, no particles"),
}
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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._thread = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            if len(s) < 6:
                continue
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except ValueError:
                continue
            for name, flag in zip(names, s[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
    """The reference's own CUDA build on the same workload (rank 0 only)."""
    if rank != 0:
        return 0
    nTheta, pdens, batch, desc = WORKLOADS[args.workload]
    exe = os.path.join(ROOT, "oracle", "_ref", "kamino_ref")
    base = {"impl": "reference", "metric": "sim steps/s", "unit": "steps/s", "higher_is_better": True,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup}
    if not os.path.exists(exe):
        print(json.dumps(dict(base, unavailable="oracle/_ref/kamino_ref not built (run __graft_entry__.build() "
                                               "where /root/reference is mounted)")))
        return 0
    if batch != 1:
        print(json.dumps(dict(base, unavailable="the reference runs one simulation per process")))
        return 0
    pd = max(pdens, 1.0)      # the reference aborts on an empty particle set
    with ClockSampler(0) as clocks:
        out = subprocess.run([exe, "bench", str(nTheta), str(pd), str(DT), str(RADIUS), str(args.steps),
                              str(STEPS_PER_FRAME), str(args.warmup)], capture_output=True, text=True, timeout=3000)
    line = None
    for l in out.stdout.splitlines():
        if l.startswith('{"ref_bench"'):
            line = json.loads(l)
    if line is None:
        print(json.dumps(dict(base, unavailable="reference run failed: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200])))
        return 0
    cells = nTheta * 2 * nTheta
    v = line["steps_per_s"]
    frame_bytes = line["d2h_bytes_per_frame"]
    res = dict(base)
    res.update({
        "value": v, "ms_per_step": 1e3 / v, "cell_updates_per_s": v * cells, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + " (reference CUDA build, sm_100, its own cudaEvent phase timers)",
                   "nTheta": nTheta, "nPhi": 2 * nTheta, "particles": line["particles"], "dt": DT, "radius": RADIUS},
        "e2e": {"value": line["e2e_steps_per_s"], "unit": "steps/s",
                "h2d_bytes_per_step": frame_bytes / args.steps, "d2h_bytes_per_step": frame_bytes / STEPS_PER_FRAME},
        "phases_s": {"advection": line["advection_s"], "geometric": line["geometric_s"], "projection": line["projection_s"]},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": 0, "kind": "reference",
                         "sample": "the reference has no CPU path: its own CUDA build, %d steps on the GPU" % args.steps},
        "gpu_launches": 18 * args.steps,
        "clocks": clocks.summary(),
    })
    print(json.dumps(res))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference(args, rank, world)

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

    nTheta, pdens, batch, desc = WORKLOADS[args.workload]
    nPhi = 2 * nTheta
    cells = nTheta * nPhi
    K, W = args.steps, max(args.warmup, 3)
    lib = capi.load()

    s = KaminoSolver(nPhi, nTheta, RADIUS, DT, device=local_rank, batch=batch)
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)           # so that torch.cuda.Event brackets our launches
    rho0 = np.empty((nTheta, nPhi), np.float32)
    jj, ii = np.meshgrid(np.arange(nTheta), np.arange(nPhi), indexing="ij")
    h = np.float32(np.pi / nTheta)
    rho0[:] = 0.5 + 0.5 * np.sin(4.0 * ii * float(h)) * np.sin((jj + 0.5) * float(h)) ** 2
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

    with torch.cuda.stream(stream):
        # ---- device-resident: W warm-up steps, then exactly K timed steps -----------------
        s.stepForward(DT, nSteps=W)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            barrier()
            ev0.record(stream)
            s.stepForward(DT, nSteps=K)
            ev1.record(stream)
            barrier()
        seconds = ev0.elapsed_time(ev1) * 1e-3
        # the timed region is a few tens of milliseconds (one nvidia-smi sample): keep sampling the clocks
        # over an untimed repetition of the same load for about a second and report both together
        with ClockSampler(local_rank) as clocks_load:
            t_load = time.perf_counter()
            while time.perf_counter() - t_load < 1.0:
                s.stepForward(DT, nSteps=K)
                stream.synchronize()
        clocks.samples.extend(clocks_load.samples)
        upload_state()                      # thousands of extra steps later: back to the initial state
        s.stepForward(DT, nSteps=W)
        stream.synchronize()

        # ---- per-kernel CUDA-event times over K steps (same state, launched individually) ----
        nlaunch = int(lib.kamino_launches_per_step(s._ctx))
        kernels = KERNELS + (["advect_particles"] if nlaunch == len(KERNELS) + 1 else [])
        kern = (ctypes.c_float * len(kernels))()
        nprof = min(K, 200)
        capi.check(lib.kamino_profile_steps(s._ctx, nprof, kern), s._ctx)
        kernel_s = {k: kern[i] / nprof for i, k in enumerate(kernels)}

        # ---- cold: 512 MB L2 flush before every step, each step timed on its own ----------
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        ncold = min(K, 50)
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

        # ---- end to end: upload + frames of STEPS_PER_FRAME steps + read-backs --------------
        frames = max(K // STEPS_PER_FRAME, 1)
        upload_state()
        capi.check(lib.kamino_run_frames(s._ctx, 2, STEPS_PER_FRAME, hU, hV, hR, hP if nPart else None), s._ctx)   # warm-up
        barrier()
        t0 = time.perf_counter()
        upload_state()
        capi.check(lib.kamino_run_frames(s._ctx, frames, STEPS_PER_FRAME, hU, hV, hR, hP if nPart else None), s._ctx)
        barrier()
        e2e_seconds = time.perf_counter() - t0
        e2e_steps = frames * STEPS_PER_FRAME

    # finite check on what came back (a diverged run would be meaningless)
    back = np.ctypeslib.as_array(ctypes.cast(hU, ctypes.POINTER(ctypes.c_float)), shape=(cells * batch,))
    finite = bool(np.isfinite(back).all())

    if world > 1:
        t = torch.tensor([seconds, e2e_seconds, cold_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        seconds, e2e_seconds, cold_s = (float(x) for x in t.tolist())

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
        res = {
            "metric": "sim steps/s", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": seconds / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "cell_updates_per_s": value * cells,
            "config": {"workload": desc + (" per GPU" if world > 1 else ""), "nTheta": nTheta, "nPhi": nPhi,
                       "particles": nPart, "batch_per_gpu": batch, "dt": DT, "radius": RADIUS,
                       "parallelism": "ensemble x%d (independent simulations, no collective)" % world if world > 1 else "single simulation",
                       "l2": "state (%.0f MB) is smaller than the 126 MB L2: consecutive steps of one simulation run "
                             "L2-resident by nature, so `value` is NOT flushed; `cold` flushes L2 (512 MB write) before every step"
                             % (state_bytes * 2 / 1e6)},
            "e2e": {"value": e2e_steps * sims / e2e_seconds, "unit": "steps/s",
                    "h2d_bytes_per_step": state_bytes / e2e_steps, "d2h_bytes_per_step": frame_bytes / STEPS_PER_FRAME,
                    "steps_per_frame": STEPS_PER_FRAME, "frames": frames},
            "cold": {"value": sims / cold_s, "unit": "steps/s", "ms_per_step": cold_s * 1e3, "steps": ncold},
            "gpu_launches": nlaunch * K,
            "kernel_us": {k: kernel_s[k] * 1e6 for k in kernel_s},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes": alg[dom],
                         "note": "algorithmic bytes / CUDA-event time of the kernel launched alone in stream order; "
                                 "working set L2-resident at this size" if state_bytes * 2 < 100e6 else
                                 "algorithmic bytes / CUDA-event time of the kernel"},
            "roofline_all": {k: {"achieved": alg[k] / kernel_s[k] / 1e9, "frac": alg[k] / kernel_s[k] / 1e9 / peak}
                             for k in kernel_s},
            "clocks": clocks.summary(),
            "finite": finite,
        }
        if not args.no_cpu_baseline and world == 1:
            res["cpu_baseline"] = cpu_baseline(nTheta, pdens)
        print(json.dumps(res))
    s.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
