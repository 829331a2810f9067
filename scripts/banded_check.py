#!/usr/bin/env python
"""Multi-GPU check of the theta-band mode (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/banded_check.py [nTheta] [steps] [alltoall|spike]

Every rank steps its band of ONE simulation (NCCL halo exchange + all-to-all); rank 0 also runs the
ordinary single-GPU solver and the gathered bands must be bit-identical to it. Prints timings.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kaminogpu_b200 import banded                      # noqa: E402
from kaminogpu_b200.solver import KaminoSolver         # noqa: E402


def main():
    nT = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    solve = sys.argv[3] if len(sys.argv) > 3 else "alltoall"        # or "spike": reduced-interface theta solve
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = 2 * nT
    jj, ii = np.meshgrid(np.arange(nT), np.arange(N), indexing="ij")
    h = np.float32(np.pi / nT)
    rho0 = (0.5 + 0.5 * np.sin(4.0 * ii * float(h)) * np.sin((jj + 0.5) * float(h)) ** 2).astype(np.float32)
    dt = 0.005 if nT <= 2048 else 0.0025
    s = banded.DistributedBandedSolver(nT, 5.0, dt, device=local, solve=solve)
    s.r.solver.density.cpuBuffer[:] = rho0
    s.r.solver.density.copyToGPU()
    dist.barrier(); torch.cuda.synchronize()
    s.step(2)                                           # warm-up (NCCL channels)
    s.sync(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.step(steps)
    s.sync(); torch.cuda.synchronize(); dist.barrier()
    el = time.perf_counter() - t0
    u, v, rho = s.r.download_band()
    parts = [None] * world
    dist.all_gather_object(parts, (u, v, rho))
    ok = True
    if rank == 0:
        U, V, RHO = (np.concatenate([p[k] for p in parts], axis=0) for k in range(3))
        with KaminoSolver(N, nT, 5.0, dt, device=local) as ref:
            ref.density.cpuBuffer[:] = rho0
            ref.density.copyToGPU()
            ref.stepForward(dt, nSteps=2 + steps)
            ref.sync()
            ru, rv, rr = ref.velPhi.copyBackToCPU().copy(), ref.velTheta.copyBackToCPU().copy(), ref.density.copyBackToCPU().copy()
            t1 = time.perf_counter()
            ref.stepForward(dt, nSteps=steps)
            ref.sync()
            single = (time.perf_counter() - t1) / steps
        for name, a, b in (("velPhi", U, ru), ("velTheta", V, rv), ("density", RHO, rr)):
            # bit patterns, so that runs which have gone non-finite still compare (at 8192 x 16384 the
            # reference scheme's 1 / (h sin(theta)) pressure gradient amplifies fp32 noise ~1e7 x in the
            # polar rows: the CPU oracle shows max|u_phi| 0.09 -> 26 -> 52 -> 104 over the first steps)
            a32, b32 = np.ascontiguousarray(a, np.float32).view(np.uint32), np.ascontiguousarray(b, np.float32).view(np.uint32)
            same = np.array_equal(a32, b32)
            if solve != "spike":
                ok &= same
            finite = float(np.isfinite(b).mean())
            print("banded x%d vs single GPU, %d x %d, %d steps: %-8s %s (%.4f of the single-GPU values finite)" % (
                  world, nT, N, 2 + steps, name,
                  "bit-identical" if same else "DIFFERS in %d words" % int((a32 != b32).sum()), finite))
            if solve == "spike":                # a different operation order: compared at tolerance level
                fin = np.isfinite(a) & np.isfinite(b)
                rel = float(np.linalg.norm((a - b)[fin]) / max(np.linalg.norm(b[fin]), 1e-30))
                print("    spike mode: relative L2 difference %.2e" % rel)
                ok &= rel <= 1e-3
        print("banded x%d: %.3f ms/step   single GPU: %.3f ms/step" % (world, el / steps * 1e3, single * 1e3))
    s.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
