#!/usr/bin/env python
"""Debug aid: virtual-rank band decomposition (kamino_dist_group_step) against kamino_step at a given size, with a cheap
analytic initial field; reports per field where the first differences are, and whether kamino_step itself repeats.
    python scripts/dist_debug.py <nTheta> <world> [steps] [dt]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kaminogpu_b200 import capi, dist              # noqa: E402
from kaminogpu_b200.solver import KaminoSolver     # noqa: E402


def fields(nT):
    N = 2 * nT
    h = np.pi / nT
    th_u = ((np.arange(nT) + 0.5) * h)[:, None]
    ph_u = ((np.arange(N) - 0.5) * h)[None, :]
    u = (0.1 * np.sin(th_u) * np.cos(4 * ph_u) + 0.05 * np.sin(3 * th_u) * np.sin(7 * ph_u)).astype(np.float32)
    th_v = ((np.arange(nT - 1) + 1.0) * h)[:, None]
    ph_v = (np.arange(N) * h)[None, :]
    v = (0.1 * np.sin(2 * th_v) * np.sin(3 * ph_v) + 0.03 * np.sin(5 * th_v) * np.cos(11 * ph_v)).astype(np.float32)
    rho = (0.5 + 0.5 * np.sin(4 * ph_v) * np.sin(th_u) ** 2).astype(np.float32)
    return u, v, rho


def single(nT, dt, u, v, rho, steps):
    with KaminoSolver(2 * nT, nT, 5.0, dt, initVelocity=False) as s:
        s.velPhi.cpuBuffer[:] = u; s.velPhi.copyToGPU()
        s.velTheta.cpuBuffer[:] = v; s.velTheta.copyToGPU()
        s.density.cpuBuffer[:] = rho; s.density.copyToGPU()
        s.stepForward(nSteps=steps)
        s.sync()
        return {capi.VEL_PHI: s.velPhi.copyBackToCPU().copy(), capi.VEL_THETA: s.velTheta.copyBackToCPU().copy(),
                capi.DENSITY: s.density.copyBackToCPU().copy(), capi.PRESSURE: s.pressure.copyBackToCPU().copy()}


def main():
    nT, world = int(sys.argv[1]), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    dt = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0025
    u, v, rho = fields(nT)
    a = single(nT, dt, u, v, rho, steps)
    b = single(nT, dt, u, v, rho, steps)
    names = ((capi.VEL_PHI, "u_phi"), (capi.VEL_THETA, "u_theta"), (capi.DENSITY, "density"), (capi.PRESSURE, "pressure"))
    for f, name in names:
        same = np.array_equal(a[f].view(np.uint32), b[f].view(np.uint32))
        print("single GPU repeatable %-8s %s" % (name, same), flush=True)
    grp = dist.LocalGroup(nT, 5.0, dt, world)
    try:
        grp.upload_global(capi.VEL_PHI, u); grp.upload_global(capi.VEL_THETA, v); grp.upload_global(capi.DENSITY, rho)
        grp.step(steps)
        grp.sync()
        for f, name in names:
            g = grp.gather(f)
            ref = a[f][:g.shape[0]]
            diff = g.view(np.uint32) != ref.view(np.uint32)
            rows = np.nonzero(diff.any(axis=1))[0]
            print("dist x%d vs single %-8s differing words %d of %d; rows %s ... %s; max abs diff %.3g (max |ref| %.3g)" % (
                world, name, int(diff.sum()), diff.size, rows[:6].tolist(), rows[-6:].tolist(),
                float(np.abs(g.astype(np.float64) - ref).max()), float(np.abs(ref).max())), flush=True)
            if rows.size:
                cols = np.nonzero(diff[rows[0]])[0]
                print("      first differing row %d: %d columns, first %s" % (rows[0], cols.size, cols[:8].tolist()))
    finally:
        grp.close()


if __name__ == "__main__":
    main()
