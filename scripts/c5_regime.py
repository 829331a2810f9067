#!/usr/bin/env python
"""Find a regime in which the 8192 x 16384 run (BASELINE config 5) stays finite: single GPU, a few steps per candidate.

    python scripts/c5_regime.py [nTheta] [steps]

The reference cannot run this size (kernel/KaminoCore.cu:779-784). With its semantics in fp32 the pressure
gradient's 1 / (h sin(theta)) reaches 1.4e7 in the polar rows, which turns round-off of p into O(1) u_phi noise
(r01: max|u_phi| 0.09 -> 26 -> 52 -> 104 at dt = 0.0025), so the candidates vary dt, the radius and the amplitude
of the initial velocity. Prints max|u_phi|, max|u_theta| after every step.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kaminogpu_b200.solver import KaminoSolver     # noqa: E402


def main():
    nT = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    N = 2 * nT
    u0 = v0 = None
    # (radius, dt, velocity scale)
    for radius, dt, scale in ((5.0, 0.0025, 1.0), (5.0, 0.0025, 1e-2), (5.0, 0.0025, 1e-4), (50.0, 0.0025, 1.0),
                              (500.0, 0.0025, 1.0), (5.0, 0.00025, 1e-2)):
        t0 = time.perf_counter()
        with KaminoSolver(N, nT, radius, dt, initVelocity=(u0 is None)) as s:
            if u0 is None:
                u0, v0 = s.velPhi.cpuBuffer.copy(), s.velTheta.cpuBuffer.copy()
                # the reference's initial field scales with 1 / radius (kernel/KaminoInitializer.cu:9-55): undo it below
                base_radius = radius
            k = np.float32(scale * base_radius / radius)
            s.velPhi.cpuBuffer[:] = u0 * k; s.velPhi.copyToGPU()
            s.velTheta.cpuBuffer[:] = v0 * k; s.velTheta.copyToGPU()
            jj = (np.arange(nT, dtype=np.float32) + 0.5) * np.float32(np.pi / nT)
            s.density.cpuBuffer[:] = (0.5 + 0.5 * np.sin(jj) ** 2)[:, None]
            s.density.copyToGPU()
            print("radius %g dt %g scale %g: start max|u_phi| %.3g max|u_theta| %.3g (setup %.1f s)" % (
                radius, dt, scale, np.abs(s.velPhi.cpuBuffer).max(), np.abs(s.velTheta.cpuBuffer).max(), time.perf_counter() - t0), flush=True)
            for step in range(1, steps + 1):
                s.stepForward(nSteps=1)
                u = s.velPhi.copyBackToCPU()
                v = s.velTheta.copyBackToCPU()
                fin = bool(np.isfinite(u).all() and np.isfinite(v).all())
                print("   step %2d max|u_phi| %.4g max|u_theta| %.4g polar-row max|u_phi| %.4g equator max|u_phi| %.4g finite %s" % (
                    step, np.nanmax(np.abs(u)), np.nanmax(np.abs(v)), np.nanmax(np.abs(u[:4])), np.nanmax(np.abs(u[nT // 2 - 2:nT // 2 + 2])), fin), flush=True)
                if not fin:
                    break


if __name__ == "__main__":
    main()
