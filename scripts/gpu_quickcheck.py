"""Quick GPU parity probe against the committed reference goldens (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_api as oa
from kaminogpu_b200.solver import KaminoSolver

def report(tag, got, ref):
    got = np.asarray(got).ravel(); ref = np.asarray(ref).ravel()
    nexact = int((got.view(np.uint32) == ref.view(np.uint32)).sum())
    print("  %-22s relL2 %.3e  maxabs %.3e  exact %d/%d" % (tag, oa.rel_l2(got, ref), np.abs(got - ref).max(), nexact, got.size))

for case in sys.argv[1:] or ["t16", "t32", "t64", "t128"]:
    g = oa.golden(case)
    nT = int(g["meta.nTheta"]); N = 2 * nT
    print("== case", case, "nTheta", nT)
    s = KaminoSolver(N, nT, float(g["meta.radius"]), float(g["meta.dt"]))
    report("init velPhi", s.velPhi.cpuBuffer, g["init.velPhi"])
    report("init velTheta", s.velTheta.cpuBuffer, g["init.velTheta"])
    s.density.cpuBuffer[:] = g["init.density"].reshape(nT, N); s.density.copyToGPU()
    s.initParticlesfromPic("", 0, coords=g["init.particles"])
    def dump(tag, pressure=False):
        report(tag + " velPhi", s.velPhi.copyBackToCPU(), g[tag + ".velPhi"])
        report(tag + " velTheta", s.velTheta.copyBackToCPU(), g[tag + ".velTheta"])
        report(tag + " density", s.density.copyBackToCPU(), g[tag + ".density"])
        report(tag + " particles", s.particles.copyBack2CPU(), g[tag + ".particles"])
        if pressure:
            report(tag + " pressure", s.pressure.copyBackToCPU(), g[tag + ".pressure"])
    s.advection(); dump("s1_adv")
    # isolate the next phases: start them from the reference's own state
    s.velPhi.cpuBuffer[:] = g["s1_adv.velPhi"].reshape(nT, N); s.velPhi.copyToGPU()
    s.velTheta.cpuBuffer[:] = g["s1_adv.velTheta"].reshape(nT - 1, N); s.velTheta.copyToGPU()
    s.geometric(); dump("s1_geo")
    s.velPhi.cpuBuffer[:] = g["s1_geo.velPhi"].reshape(nT, N); s.velPhi.copyToGPU()
    s.velTheta.cpuBuffer[:] = g["s1_geo.velTheta"].reshape(nT - 1, N); s.velTheta.copyToGPU()
    s.projection(); dump("s1_proj", True)
    print("  phase times", s.phase_times())
    s.close()
