#!/usr/bin/env python
"""compute-sanitizer probe: one projection (or one band-decomposed step) at a given size with a cheap analytic field.
    compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_probe.py <nTheta> single|dist<P> [phase]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from dist_debug import fields                      # noqa: E402
from kaminogpu_b200 import capi, dist              # noqa: E402
from kaminogpu_b200.solver import KaminoSolver     # noqa: E402

nT, mode = int(sys.argv[1]), sys.argv[2]
phase = sys.argv[3] if len(sys.argv) > 3 else "projection"
u, v, rho = fields(nT)
if mode == "single":
    with KaminoSolver(2 * nT, nT, 5.0, 0.0025, initVelocity=False) as s:
        s.velPhi.cpuBuffer[:] = u; s.velPhi.copyToGPU()
        s.velTheta.cpuBuffer[:] = v; s.velTheta.copyToGPU()
        s.density.cpuBuffer[:] = rho; s.density.copyToGPU()
        {"projection": s.projection, "advection": s.advection, "geometric": s.geometric, "step": s.stepForward}[phase]()
        s.sync()
        print("single", phase, "done")
else:
    world = int(mode[4:])
    grp = dist.LocalGroup(nT, 5.0, 0.0025, world)
    grp.upload_global(capi.VEL_PHI, u); grp.upload_global(capi.VEL_THETA, v); grp.upload_global(capi.DENSITY, rho)
    grp.step(1)
    grp.sync()
    grp.close()
    print("dist", world, "done")
