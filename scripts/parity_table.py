#!/usr/bin/env python
"""Per-size parity table against a LIVE run of the reference's own CUDA build (oracle/_ref/kamino_ref).

    python scripts/parity_table.py [out.md] [sizes...]        sizes among c1 c4 c2 c3 (default: all)

For every size: the reference is run once (phase dumps of step 1, then free-running with a dump every
10 steps), our build starts every phase from the reference's own state, and the table records per field
  * identical 32-bit words and relative L2 for advection / geometric / particles,
  * projection: ours vs reference, ours vs an fp64 evaluation of the reference's operator, the reference
    vs that fp64 evaluation, and ours with the theta solve in the reference's cyclic-reduction order
    (kamino_debug_project_cr) vs the reference,
  * free-run divergence after 1 / 10 / 100 steps.
Needs a GPU and oracle/_ref/kamino_ref. Test infrastructure (reads oracle/), not product code.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as oa                                   # noqa: E402
from kaminogpu_b200.solver import KaminoSolver            # noqa: E402
from test_parity_gpu import fp64_projection, words_equal, state, set_velocity   # noqa: E402

SIZES = {"c1": (128, 200, 100), "c4": (256, 1, 100), "c2": (512, 2, 100), "c3": (2048, 1, 10)}
EXE = os.path.join(ROOT, "oracle", "_ref", "kamino_ref")


def run_reference(nT, pdens, steps, out):
    r = subprocess.run([EXE, "dump", str(nT), str(pdens), "0.005", "5.0", str(steps), out, "-", "1"],
                       capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        raise SystemExit("kamino_ref failed: " + r.stderr[-500:])


def main():
    out_md = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_table.md"
    sizes = sys.argv[2:] or list(SIZES)
    rows_phase, rows_proj, rows_free = [], [], []
    for name in sizes:
        nT, pdens, steps = SIZES[name]
        N = 2 * nT
        with tempfile.TemporaryDirectory() as tmp:
            run_reference(nT, pdens, steps, tmp)
            ld = lambda tag, f: np.fromfile(os.path.join(tmp, "%s.%s.f32" % (tag, f)), dtype=np.float32)
            with KaminoSolver(N, nT, 5.0, 0.005) as s:
                assert np.array_equal(s.velPhi.cpuBuffer.ravel(), ld("init", "velPhi")), "initial velocity differs"
                s.density.cpuBuffer[:] = ld("init", "density").reshape(nT, N)
                s.density.copyToGPU()
                s.initParticlesfromPic("", pdens)
                assert np.array_equal(s.particles.coordCPUBuffer, ld("init", "particles")), "particle seeding differs"
                s.advection()
                st = state(s)
                for f in ("velPhi", "velTheta", "density", "particles"):
                    rows_phase.append((name, nT, "advection", f, words_equal(st[f], ld("s1_adv", f)), oa.rel_l2(st[f], ld("s1_adv", f))))
                s.geometric()
                st = state(s)
                for f in ("velPhi", "velTheta"):
                    rows_phase.append((name, nT, "geometric", f, words_equal(st[f], ld("s1_geo", f)), oa.rel_l2(st[f], ld("s1_geo", f))))
                # projection from the reference's own post-geometric state, both solve orders
                ug, vg = ld("s1_geo", "velPhi"), ld("s1_geo", "velTheta")
                u64, v64, p64 = fp64_projection(nT, ug, vg)
                exact = {"velPhi": u64, "velTheta": v64, "pressure": p64}
                got = {}
                for mode in ("lu", "cr"):
                    set_velocity(s, ug, vg)
                    (s.projection if mode == "lu" else s.projection_cr_order)()
                    stp = state(s)
                    got[mode] = {"velPhi": stp["velPhi"], "velTheta": stp["velTheta"], "pressure": s.pressure.copyBackToCPU().ravel().copy()}
                for f in ("velPhi", "velTheta", "pressure"):
                    ref = ld("s1_proj", f)
                    rows_proj.append((name, nT, f, oa.rel_l2(got["lu"][f], ref), oa.rel_l2(got["lu"][f], exact[f]),
                                      oa.rel_l2(ref, exact[f]), oa.rel_l2(got["cr"][f], ref), oa.rel_l2(got["cr"][f], exact[f])))
            # free run from the initial state
            with KaminoSolver(N, nT, 5.0, 0.005) as s:
                s.density.cpuBuffer[:] = ld("init", "density").reshape(nT, N)
                s.density.copyToGPU()
                s.initParticlesfromPic("", pdens)
                done = 0
                for k in (1, 10, 100):
                    if k > steps:
                        break
                    s.stepForward(nSteps=k - done)
                    done = k
                    st = state(s)
                    rows_free.append((name, nT, k) + tuple(oa.rel_l2(st[f], ld("s%d_proj" % k, f)) for f in ("velPhi", "velTheta", "density", "particles")))
        print("done", name, flush=True)
    with open(out_md, "w") as f:
        f.write("### Phase parity from the reference's own states (live reference CUDA build, same box)\n\n")
        f.write("| size | nTheta | phase | field | identical words | rel L2 |\n|---|---|---|---|---|---|\n")
        for r in rows_phase:
            f.write("| %s | %d | %s | %s | %.6f | %.2e |\n" % r)
        f.write("\n### Projection (one call from the reference's post-geometric state)\n\n")
        f.write("| size | nTheta | field | ours vs ref | ours vs fp64 | ref vs fp64 | ours (CR order) vs ref | ours (CR order) vs fp64 |\n|---|---|---|---|---|---|---|---|\n")
        for r in rows_proj:
            f.write("| %s | %d | %s | %.2e | %.2e | %.2e | %.2e | %.2e |\n" % r)
        f.write("\n### Free-run divergence from the reference (relative L2 after k steps)\n\n")
        f.write("| size | nTheta | steps | u_phi | u_theta | density | particles |\n|---|---|---|---|---|---|---|\n")
        for r in rows_free:
            f.write("| %s | %d | %d | %.2e | %.2e | %.2e | %.2e |\n" % r)
    print(open(out_md).read())


if __name__ == "__main__":
    main()
