#!/usr/bin/env python
"""Attribute the in-graph step time to kernels: time K graph-launched steps with subsets of the five
kernels captured (KAMINO_DEBUG_STEP_MASK, timing instrumentation only -- results of masked runs are
meaningless). Usage: python scripts/step_mask_timing.py [workload] [steps]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, ctypes
sys.path.insert(0, %r)
import torch
from kaminogpu_b200 import capi
from kaminogpu_b200.solver import KaminoSolver
nT, pd, K = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
s = KaminoSolver(2 * nT, nT, 5.0, 0.005, device=0, batch=1)
st = torch.cuda.Stream(); s.set_stream(st.cuda_stream)
if pd > 0: s.initParticlesfromPic("", pd)
with torch.cuda.stream(st):
    s.stepForward(0.005, nSteps=50); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); s.stepForward(0.005, nSteps=K); b.record(st); torch.cuda.synchronize()
print("US_PER_STEP", a.elapsed_time(b) * 1e3 / K)
''' % ROOT
W = {"c2": (512, 2.0), "c3": (2048, 0.0), "c1": (128, 200.0)}
def main():
    w = sys.argv[1] if len(sys.argv) > 1 else "c2"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    nT, pd = W[w]
    # bit 5 = the particle kernel (its own kernel on a parallel graph branch when the context holds particles)
    names = {63: "all", 31: "all but particles", 1: "advect (cells)", 32: "particles", 33: "advect + particles",
             2: "geometric", 4: "divergence_fft", 8: "tridiagonal", 16: "inverse_fft_gradient",
             28: "projection", 30: "geometric + projection", 62: "all but advect (cells)"}
    for mask, name in names.items():
        env = dict(os.environ, KAMINO_DEBUG_STEP_MASK=str(mask))
        out = subprocess.run([sys.executable, "-c", CHILD, str(nT), str(pd), str(K)], env=env, capture_output=True, text=True)
        us = [l.split()[1] for l in out.stdout.splitlines() if l.startswith("US_PER_STEP")]
        print("%s mask %2d %-22s %s us/step" % (w, mask, name, us[0] if us else "FAILED " + out.stderr[-300:]), flush=True)
if __name__ == "__main__":
    main()
