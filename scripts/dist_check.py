#!/usr/bin/env python
"""Band-decomposed run over NCCL (kamino_dist_*, csrc/dist.cu) against the single-GPU step, on N GPUs.

    torchrun --nproc-per-node N scripts/dist_check.py <nTheta> <steps> [dt] [velocity scale] [timed steps]

Every rank runs the band-decomposed simulation AND (sizes permitting) the whole simulation on its own GPU,
then compares its band bit for bit. Prints ms/step of both and the NCCL accounting of rank 0.
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kaminogpu_b200 import capi                    # noqa: E402
from kaminogpu_b200.dist import DistributedSolver  # noqa: E402
from kaminogpu_b200.solver import KaminoSolver     # noqa: E402


def analytic_fields(nT, lo, hi):
    """Cheap smooth fields of the reference's amplitude, rows [lo, hi): u_phi, u_theta (row j = node (j+1)h), density."""
    N = 2 * nT
    h = np.pi / nT
    th_u = ((np.arange(lo, hi) + 0.5) * h)[:, None]
    ph_u = ((np.arange(N) - 0.5) * h)[None, :]
    u = (0.1 * np.sin(th_u) * np.cos(4 * ph_u) + 0.05 * np.sin(3 * th_u) * np.sin(7 * ph_u)).astype(np.float32)
    th_v = ((np.arange(lo, hi) + 1.0) * h)[:, None]
    ph_v = (np.arange(N) * h)[None, :]
    v = (0.1 * np.sin(2 * th_v) * np.sin(3 * ph_v) + 0.03 * np.sin(5 * th_v) * np.cos(11 * ph_v)).astype(np.float32)
    rho = (0.5 + 0.5 * np.sin(4 * ph_v) * np.sin(th_u) ** 2).astype(np.float32)
    return u, v, rho


def main():
    nT, steps = int(sys.argv[1]), int(sys.argv[2])
    dt = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005
    scale = np.float32(sys.argv[4]) if len(sys.argv) > 4 else np.float32(1.0)
    timed = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = 2 * nT
    d = DistributedSolver(nT, 5.0, dt, device=local)
    t0 = time.perf_counter()
    fbm = nT <= 2048          # the reference's FBM field costs ~2 core-minutes at 8192 x 16384: analytic field there
    if fbm:
        u, v = d.init_velocity()
        if scale != 1.0:
            d.upload(capi.VEL_PHI, u * scale)
            d.upload(capi.VEL_THETA, v[:d.rows_of(capi.VEL_THETA)] * scale)
    else:
        u, v, _ = analytic_fields(nT, d.lo, d.hi)
        d.upload(capi.VEL_PHI, u * scale)
        d.upload(capi.VEL_THETA, v[:d.rows_of(capi.VEL_THETA)] * scale)
    t_init = time.perf_counter() - t0
    h = np.float32(np.pi / nT)
    rho = analytic_fields(nT, d.lo, d.hi)[2]
    d.upload(capi.DENSITY, rho)
    d.step(steps)
    d.sync()
    got = {f: d.download(f) for f in (capi.VEL_PHI, capi.VEL_THETA, capi.DENSITY)}
    finite = all(np.isfinite(a).all() for a in got.values())
    # timing: `timed` more steps, max over ranks; then the same with the other transport of the transposes
    def timed_run():
        d.step(2); d.sync()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        d.step(timed)
        d.sync()
        dist.barrier()
        return (time.perf_counter() - t0) / timed * 1e3
    peer, note = d.transport()
    ms_dist = timed_run()
    ms_other = None
    if world > 1 and (peer or "IPC mappings" in note):
        d.transport(not peer)
        ms_other = timed_run()
        d.transport(peer)
    d.comm_stats(enable=1)
    d.step(timed)
    d.sync()
    stats = d.comm_stats(enable=0)
    line = "rank %d/%d nTheta %d rows [%d,%d) %.2f GB on device, init %.1f s, finite %s, %.3f ms/step banded (%s; %s)" % (
        rank, world, nT, d.lo, d.hi, d.device_bytes / 1e9, t_init, finite, ms_dist, "peer stores" if peer else "NCCL transposes", note)
    if ms_other is not None:
        line += ", %.3f ms/step with %s" % (ms_other, "NCCL transposes" if peer else "peer stores")
    if stats["steps"]:
        n = stats["steps"]
        line += " | halo %.3f ms (%.1f GB/s)" % (stats["halo_s"] / n * 1e3, stats["halo_bytes_per_step"] / max(stats["halo_s"] / n, 1e-9) / 1e9)
        if peer:
            line += " barriers of the peer-store transposes %.3f ms (%.1f MB stored into peers per step)" % (
                stats["transpose_s"] / n * 1e3, stats["transpose_bytes_per_step"] / 1e6)
        else:
            line += " transposes %.3f ms (%.1f GB/s sent)" % (
                stats["transpose_s"] / n * 1e3, stats["transpose_bytes_per_step"] / max(stats["transpose_s"] / n, 1e-9) / 1e9)
    # single-GPU run of the whole grid on this rank's GPU, compared on my band
    same = None
    try:
        with KaminoSolver(N, nT, 5.0, dt, device=local) as s:
            uu, vv, rr = analytic_fields(nT, 0, nT)
            if not fbm:
                s.velPhi.cpuBuffer[:] = uu; s.velTheta.cpuBuffer[:] = vv[:nT - 1]
            if scale != 1.0 or not fbm:
                s.velPhi.cpuBuffer[:] *= scale; s.velPhi.copyToGPU()
                s.velTheta.cpuBuffer[:] *= scale; s.velTheta.copyToGPU()
            s.density.cpuBuffer[:] = rr
            s.density.copyToGPU()
            del uu, vv, rr
            s.stepForward(nSteps=steps)
            s.sync()
            ref = {capi.VEL_PHI: s.velPhi.copyBackToCPU(), capi.VEL_THETA: s.velTheta.copyBackToCPU(), capi.DENSITY: s.density.copyBackToCPU()}
            same = all(np.array_equal(got[f].view(np.uint32), ref[f][d.lo:d.lo + got[f].shape[0]].view(np.uint32)) for f in got)
            maxu = float(np.abs(ref[capi.VEL_PHI]).max())
            t0 = time.perf_counter()
            s.stepForward(nSteps=timed)
            s.sync()
            ms_single = (time.perf_counter() - t0) / timed * 1e3
        line += " | single GPU %.3f ms/step, max|u_phi| %.3g, band bit-identical: %s" % (ms_single, maxu, same)
    except capi.KaminoError as e:
        line += " | single-GPU comparison unavailable: %s" % e
    print(line, flush=True)
    d.close()
    dist.destroy_process_group()
    if same is False or not finite:
        sys.exit(1)


if __name__ == "__main__":
    main()
