"""GPU suite: the CUDA path (through the C ABI / the KaminoSolver mirror) against
  (1) the committed dumps of the reference's own CUDA build (tests/golden/ref_*.npz),
  (2) the CPU oracle on the same seeded inputs,
  (3) size-independent properties at the BASELINE.json sizes.

Tolerances (fp32 relative L2 per field, north star: <= 1e-5 per field per step):
  advection, geometric, particles: 1e-5 asserted; these phases reproduce the reference's
      arithmetic operation for operation and are expected bit-identical on the goldens
      (asserted as >= 99.9% identical words so that a single double-rounding tie cannot
      fail the suite; the exact fraction is printed).
  projection: u_theta and pressure 1e-5 against the reference dump. u_phi: the reference's own
      result is 5e-6 (nTheta=16) .. 8e-4 (nTheta=128) away from an fp64 evaluation of its own
      operator: it adds the n = 0 "identity solve" mode (the zonal-mean divergence, O(100)) into
      every pressure value inside cuFFT and subtracts it again (kernel/KaminoCore.cu:692-700),
      which leaves ~|U_0| * 2^-24 of absolute noise in p, and the phi gradient divides that by
      h*sin(theta) (3e-4 in the polar rows at nTheta=128). That noise depends on cuFFT's internal
      rounding and cannot be reproduced by any other FFT. Asserted instead: (a) our u_phi, u_theta
      and p are within 1e-5 of the fp64 evaluation (measured ~1e-6, i.e. ~500x closer than the
      reference is), and (b) |ours - reference| <= 1.5 x |reference - fp64| (the whole distance
      is the reference's own noise). See DESIGN.md "projection parity".
"""
import ctypes

import numpy as np
import pytest

import oracle_api as oa

pytestmark = pytest.mark.gpu

CASES = ["t16", "t32", "t64", "t128"]


@pytest.fixture(scope="module")
def K(built):
    from kaminogpu_b200 import capi, solver
    capi.load()
    return solver


def words_equal(a, b):
    a = np.ascontiguousarray(a, np.float32).ravel()
    b = np.ascontiguousarray(b, np.float32).ravel()
    return float((a.view(np.uint32) == b.view(np.uint32)).mean())


def make_solver(K, g, particles=True, **kw):
    nT = int(g["meta.nTheta"])
    N = 2 * nT
    s = K.KaminoSolver(N, nT, float(g["meta.radius"]), float(g["meta.dt"]), **kw)
    s.density.cpuBuffer[:] = g["init.density"].reshape(nT, N)
    s.density.copyToGPU()
    if particles:
        s.initParticlesfromPic("", 0, coords=g["init.particles"])
    return s


def set_velocity(s, u, v):
    s.velPhi.cpuBuffer[:] = np.asarray(u).reshape(s.nTheta, s.nPhi)
    s.velPhi.copyToGPU()
    s.velTheta.cpuBuffer[:] = np.asarray(v).reshape(s.nTheta - 1, s.nPhi)
    s.velTheta.copyToGPU()


def state(s):
    out = {"velPhi": s.velPhi.copyBackToCPU().ravel().copy(), "velTheta": s.velTheta.copyBackToCPU().ravel().copy(),
           "density": s.density.copyBackToCPU().ravel().copy()}
    if s.particles is not None:
        out["particles"] = s.particles.copyBack2CPU().copy()
    return out


def fp64_projection(nT, u, v, radius=5.0, dt=0.005):
    """fp64 evaluation of the reference's projection operator (rfft -> per-wavenumber
    tridiagonal solve with the fp32 coefficient tables -> irfft, n = 0 dropped)."""
    N = 2 * nT
    p = oa.params(nT, radius, dt)
    div = np.zeros(nT * N, np.float32)
    oa.lib().ko_divergence(ctypes.byref(p), oa.fptr(np.ascontiguousarray(u)), oa.fptr(np.ascontiguousarray(v)), oa.fptr(div))
    F = np.fft.rfft(div.reshape(nT, N).astype(np.float64), axis=1) / N
    U = np.zeros_like(F)
    from scipy.linalg import solve_banded
    for n in range(1, N // 2 + 1):
        a = np.zeros(nT, np.float32); b = np.zeros(nT, np.float32); c = np.zeros(nT, np.float32)
        oa.lib().ko_abc_row(ctypes.byref(p), n, oa.fptr(a), oa.fptr(b), oa.fptr(c))
        ab = np.zeros((3, nT))
        ab[0, 1:] = c[:-1]; ab[1] = b; ab[2, :-1] = a[1:]
        U[:, n] = solve_banded((1, 1), ab, F[:, n])
    pr = np.fft.irfft(U * N, n=N, axis=1)
    h = float(p.gridLen)
    theta = (np.arange(nT) + 0.5) * h
    uo = u.reshape(nT, N).astype(np.float64) - (pr - np.roll(pr, 1, axis=1)) / (h * np.sin(theta))[:, None]
    vo = v.reshape(nT - 1, N).astype(np.float64) - (pr[1:] - pr[:-1]) / h
    return uo.ravel(), vo.ravel(), pr.ravel()


# ---- (1) phase-by-phase against the reference's own CUDA build ------------------------------

@pytest.mark.parametrize("case", CASES)
def test_initial_velocity_is_the_references(K, case):
    g = oa.golden(case)
    with make_solver(K, g, particles=False) as s:
        st = state(s)
    assert np.array_equal(st["velPhi"], g["init.velPhi"])
    assert np.array_equal(st["velTheta"], g["init.velTheta"])


@pytest.mark.parametrize("case", CASES)
def test_advection_vs_reference_dump(K, case):
    g = oa.golden(case)
    with make_solver(K, g) as s:
        s.advection()
        st = state(s)
    for name in ("velPhi", "velTheta", "density", "particles"):
        ref = g["s1_adv." + name]
        e, w = oa.rel_l2(st[name], ref), words_equal(st[name], ref)
        print("advection %s %-9s relL2 %.2e identical words %.5f" % (case, name, e, w))
        assert e <= 1e-5
        assert w >= 0.999


@pytest.mark.parametrize("case", CASES)
def test_geometric_vs_reference_dump(K, case):
    g = oa.golden(case)
    with make_solver(K, g, particles=False) as s:
        set_velocity(s, g["s1_adv.velPhi"], g["s1_adv.velTheta"])
        s.geometric()
        st = state(s)
    for name in ("velPhi", "velTheta"):
        ref = g["s1_geo." + name]
        e, w = oa.rel_l2(st[name], ref), words_equal(st[name], ref)
        print("geometric %s %-9s relL2 %.2e identical words %.5f" % (case, name, e, w))
        assert e <= 1e-5
        assert w >= 0.999


@pytest.mark.parametrize("case", CASES)
def test_projection_vs_reference_dump(K, case):
    g = oa.golden(case)
    nT = int(g["meta.nTheta"])
    N = 2 * nT
    with make_solver(K, g, particles=False) as s:
        set_velocity(s, g["s1_geo.velPhi"], g["s1_geo.velTheta"])
        s.projection()
        st = state(s)
        pressure = s.pressure.copyBackToCPU().ravel().copy()
    u64, v64, p64 = fp64_projection(nT, g["s1_geo.velPhi"], g["s1_geo.velTheta"])
    ours = {"velPhi": st["velPhi"], "velTheta": st["velTheta"], "pressure": pressure}
    exact = {"velPhi": u64, "velTheta": v64, "pressure": p64}
    for name in ("velPhi", "velTheta", "pressure"):
        ref = g["s1_proj." + name]
        e_ref = oa.rel_l2(ours[name], ref)
        e_ours64 = oa.rel_l2(ours[name], exact[name])
        e_ref64 = oa.rel_l2(ref, exact[name])
        print("projection %s %-9s vs reference %.2e | vs fp64: ours %.2e reference %.2e"
              % (case, name, e_ref, e_ours64, e_ref64))
        # (a) closer to the exact operator than the north-star tolerance
        assert e_ours64 <= 1e-5
        # (b) the distance to the reference is the reference's own fp32 noise, not ours
        if name == "velPhi":
            assert e_ref <= 1.5 * e_ref64 + 1e-6
        else:
            assert e_ref <= 1e-5


@pytest.mark.parametrize("case", CASES)
def test_projection_in_cyclic_reduction_order_vs_reference_dump(K, case):
    """kamino_debug_project_cr: same FFTs, the theta solve in the reference's elimination order
    (kernel/tdm.cu:43-90). Pressure and u_theta then sit within the north-star 1e-5 of the reference's dump."""
    g = oa.golden(case)
    with make_solver(K, g, particles=False) as s:
        set_velocity(s, g["s1_geo.velPhi"], g["s1_geo.velTheta"])
        s.projection_cr_order()
        st = state(s)
        pressure = s.pressure.copyBackToCPU().ravel().copy()
    for name, got in (("velTheta", st["velTheta"]), ("pressure", pressure)):
        e = oa.rel_l2(got, g["s1_proj." + name])
        print("projection (CR order) %s %-9s vs reference %.2e" % (case, name, e))
        assert e <= 1e-5


@pytest.mark.parametrize("case", ["t16", "t32"])
def test_second_step_phases_vs_reference_dump(K, case):
    """Each phase of step 2 started from the reference's state (covers buffer-role swaps)."""
    g = oa.golden(case)
    nT = int(g["meta.nTheta"])
    N = 2 * nT
    with make_solver(K, g) as s:
        set_velocity(s, g["s1_proj.velPhi"], g["s1_proj.velTheta"])
        s.density.cpuBuffer[:] = g["s1_proj.density"].reshape(nT, N)
        s.density.copyToGPU()
        s.particles.coordCPUBuffer[:] = g["s1_proj.particles"]
        s.particles.copy2GPU()
        s.advection()
        st = state(s)
        for name in ("velPhi", "velTheta", "density", "particles"):
            assert oa.rel_l2(st[name], g["s2_adv." + name]) <= 1e-5, name
            assert words_equal(st[name], g["s2_adv." + name]) >= 0.999, name
        set_velocity(s, g["s2_adv.velPhi"], g["s2_adv.velTheta"])
        s.geometric()
        st = state(s)
        for name in ("velPhi", "velTheta"):
            assert oa.rel_l2(st[name], g["s2_geo." + name]) <= 1e-5, name


def test_free_run_divergence_100_steps(K):
    """100-step free run at 128 x 256 against the reference's own 100-step dump: reported;
    bounded loosely (chaotic amplification of fp32 round-off differences in the projection)."""
    g = oa.golden("t128")
    with make_solver(K, g) as s:
        s.stepForward(nSteps=1)
        st1 = state(s)
        s.stepForward(nSteps=9)
        st10 = state(s)
        s.stepForward(nSteps=90)
        st100 = state(s)
    for tag, st in (("s1_proj", st1), ("s10_proj", st10), ("s100_proj", st100)):
        for name in ("velPhi", "velTheta", "density", "particles"):
            print("free run %-9s %-9s relL2 vs reference %.2e" % (tag, name, oa.rel_l2(st[name], g[tag + "." + name])))
    assert oa.rel_l2(st1["velTheta"], g["s1_proj.velTheta"]) <= 1e-5
    assert oa.rel_l2(st1["density"], g["s1_proj.density"]) <= 1e-5
    assert oa.rel_l2(st1["particles"], g["s1_proj.particles"]) <= 1e-5
    assert oa.rel_l2(st100["density"], g["s100_proj.density"]) <= 5e-2
    assert oa.rel_l2(st100["velTheta"], g["s100_proj.velTheta"]) <= 2e-1
    assert np.isfinite(st100["velPhi"]).all()


# ---- (2) against the CPU oracle on the same inputs ---------------------------------------------

@pytest.mark.parametrize("kind", [oa.VPHI, oa.VTHETA, oa.CENTERED])
def test_sampler_indices_and_masks_bit_exact(K, kind):
    """Cell indices, weights, validated coordinates and the flipped / pole-branch predicates of
    the three samplers on adversarial coordinates: exactly the oracle's (IEEE-only) values."""
    nT = 64
    p = oa.params(nT)
    h = float(p.gridLen)
    rng = np.random.default_rng(11 + kind)
    phi = rng.uniform(-2 * np.pi, 4 * np.pi, 20000).astype(np.float32)
    theta = rng.uniform(-np.pi, 2 * np.pi, 20000).astype(np.float32)
    edge_t = np.array([0.0, -0.0, h / 2, h, -h / 2, np.pi, np.pi - h / 2, np.pi + h / 2, np.pi - h, np.pi + h,
                       2 * np.pi, np.float32(np.pi), np.nextafter(np.float32(np.pi), np.float32(4)),
                       np.nextafter(np.float32(np.pi), np.float32(0)), 1e-30, -1e-30, (nT - 1) * h, (nT - 1.5) * h,
                       (nT - 0.5) * h, 1.5 * h], dtype=np.float32)
    edge_p = np.array([0.0, -1e-7, 2 * np.pi, np.float32(2 * np.pi), np.nextafter(np.float32(2 * np.pi), np.float32(0)),
                       np.pi, -h / 2, h / 2, 2 * np.pi - h / 2, 4 * np.pi, -2 * np.pi, 127.5 * h], dtype=np.float32)
    ep, et = np.meshgrid(edge_p, edge_t)
    phi = np.concatenate([phi, ep.ravel()])
    theta = np.concatenate([theta, et.ravel()])
    with K.KaminoSolver(2 * nT, nT, 5.0, 0.005, initVelocity=False) as s:
        got = s.locate(kind, phi, theta)
    for k in range(phi.size):
        loc = oa.locate(p, kind, float(phi[k]), float(theta[k]))
        exp = (loc.phiIndex, loc.thetaIndex, np.float32(loc.alphaPhi), np.float32(loc.alphaTheta),
               np.float32(loc.phi), np.float32(loc.theta), loc.flipped | (loc.poleBranch << 1))
        have = (got["phiIndex"][k], got["thetaIndex"][k], got["alphaPhi"][k], got["alphaTheta"][k],
                got["phi"][k], got["theta"][k], got["flags"][k])
        assert all(np.asarray(a).tobytes() == np.asarray(b, dtype=np.asarray(a).dtype).tobytes() for a, b in zip(have, exp)), \
            (k, float(phi[k]), float(theta[k]), have, exp)


def test_one_step_at_c2_size_vs_oracle(K):
    """512 x 1024 with density and 1,048,352 particles (BASELINE config 2): advection against the
    oracle at full size."""
    nT = 512
    p = oa.params(nT)
    u, v = oa.init_velocity(nT)
    rho = oa.synthetic_density(nT)
    pc = oa.seed_particles(nT, 2.0)
    assert pc.size // 2 == 1048352
    uo, vo, ro, po, _ = oa.step(p, u, v, rho, pc, phase=1)
    with K.KaminoSolver(2 * nT, nT, 5.0, 0.005) as s:
        s.density.cpuBuffer[:] = rho.reshape(nT, 2 * nT)
        s.density.copyToGPU()
        s.initParticlesfromPic("", 2)
        assert np.array_equal(s.particles.coordCPUBuffer, pc)
        s.advection()
        st = state(s)
    for name, ref in (("velPhi", uo), ("velTheta", vo), ("density", ro), ("particles", po)):
        e = oa.rel_l2(st[name], ref)
        print("C2 advection %-9s relL2 vs oracle %.2e" % (name, e))
        assert e <= 1e-5


# ---- (3) structure and properties ------------------------------------------------------------------

def test_graph_steps_equal_phase_calls(K):
    g = oa.golden("t32")
    with make_solver(K, g) as a, make_solver(K, g) as b:
        for _ in range(3):
            a.advection(); a.geometric(); a.projection()
        b.stepForward(nSteps=3)
        b.sync()
        sa, sb = state(a), state(b)
    for name in sa:
        assert np.array_equal(sa[name], sb[name]), name


def test_step_chunking_is_consistent(K):
    """kamino_step(13) (graphs of 10 + 2 + 1 steps) == 13 x kamino_step(1)."""
    g = oa.golden("t16")
    with make_solver(K, g) as a, make_solver(K, g) as b:
        a.stepForward(nSteps=13)
        for _ in range(13):
            b.stepForward(nSteps=1)
        sa, sb = state(a), state(b)
    for name in sa:
        assert np.array_equal(sa[name], sb[name]), name


def test_ensemble_members_match_single_runs(K):
    """batch = 3 simulations with different densities/particles in one context == 3 single runs."""
    g = oa.golden("t32")
    nT, N = 32, 64
    rng = np.random.default_rng(3)
    rhos = [g["init.density"].reshape(nT, N) * np.float32(1 + k) for k in range(3)]
    parts = [np.stack([rng.uniform(0, 2 * np.pi, 1000), rng.uniform(0, np.pi, 1000)], axis=1).astype(np.float32).ravel()
             for _ in range(3)]
    vels = [(g["init.velPhi"] * np.float32(1 - 0.25 * k), g["init.velTheta"] * np.float32(1 + 0.5 * k)) for k in range(3)]
    singles = []
    for k in range(3):
        with K.KaminoSolver(N, nT, 5.0, 0.005) as s:
            set_velocity(s, *vels[k])
            s.density.cpuBuffer[:] = rhos[k]; s.density.copyToGPU()
            s.initParticlesfromPic("", 0, coords=parts[k])
            s.stepForward(nSteps=4)
            singles.append(state(s))
    from kaminogpu_b200 import capi
    with K.KaminoSolver(N, nT, 5.0, 0.005, batch=3) as s:
        s.initParticlesfromPic("", 0, coords=parts[0])
        for k in range(3):
            for field, val in ((capi.VEL_PHI, vels[k][0]), (capi.VEL_THETA, vels[k][1]), (capi.DENSITY, rhos[k])):
                q = s.quantity(field, k)
                q.cpuBuffer[:] = np.asarray(val).reshape(q.cpuBuffer.shape)
                q.copyToGPU()
            capi.check(s._lib.kamino_upload_particles(s._ctx, k, parts[k].ctypes.data), s._ctx)
        s.stepForward(nSteps=4)
        for k in range(3):
            for field, name in ((capi.VEL_PHI, "velPhi"), (capi.VEL_THETA, "velTheta"), (capi.DENSITY, "density")):
                got = s.quantity(field, k).copyBackToCPU().ravel()
                assert np.array_equal(got, singles[k][name]), (k, name)
            pc = np.zeros(2000, np.float32)
            capi.check(s._lib.kamino_download_particles(s._ctx, k, pc.ctypes.data), s._ctx)
            assert np.array_equal(pc, singles[k]["particles"]), k


def test_particle_tail_and_range(K):
    """Particle counts that are not a multiple of the block size are handled (the reference has
    no tail guard, kernel/KaminoCore.cu:323), coordinates stay in [0, 2 pi] x [0, pi]."""
    g = oa.golden("t32")
    rng = np.random.default_rng(5)
    for n in (1, 255, 257, 1000):
        pc = np.stack([rng.uniform(0, 2 * np.pi, n), rng.uniform(0, np.pi, n)], axis=1).astype(np.float32).ravel()
        with make_solver(K, g, particles=False) as s:
            s.initParticlesfromPic("", 0, coords=pc)
            s.stepForward(nSteps=5)
            out = s.particles.copyBack2CPU().reshape(-1, 2)
        assert out.shape[0] == n and np.isfinite(out).all()
        assert (out[:, 0] >= 0).all() and (out[:, 0] <= np.float32(2 * np.pi)).all()
        assert (out[:, 1] >= 0).all() and (out[:, 1] <= np.float32(np.pi)).all()
        p = oa.params(32)
        _, _, _, po, _ = oa.step(p, g["init.velPhi"], g["init.velTheta"], None, pc, phase=1)
        with make_solver(K, g, particles=False) as s:
            s.initParticlesfromPic("", 0, coords=pc)
            s.advection()
            assert oa.rel_l2(s.particles.copyBack2CPU(), po) <= 1e-6


def test_zero_particles_and_no_density_change(K):
    g = oa.golden("t16")
    with make_solver(K, g, particles=False) as s:
        s.stepForward(nSteps=2)
        st = state(s)
    assert oa.rel_l2(st["velTheta"], g["s2_proj.velTheta"]) <= 1e-5
    assert oa.rel_l2(st["density"], g["s2_proj.density"]) <= 1e-5


def test_geometric_is_phi_shift_equivariant(K):
    """The geometric update depends on theta only: rolling the input along phi rolls the output
    (bit-exact)."""
    g = oa.golden("t64")
    nT, N = 64, 128
    u = g["s1_adv.velPhi"].reshape(nT, N)
    v = g["s1_adv.velTheta"].reshape(nT - 1, N)
    outs = []
    for shift in (0, 37):
        with make_solver(K, g, particles=False) as s:
            set_velocity(s, np.roll(u, shift, axis=1), np.roll(v, shift, axis=1))
            s.geometric()
            st = state(s)
        outs.append((st["velPhi"].reshape(nT, N), st["velTheta"].reshape(nT - 1, N)))
    assert np.array_equal(np.roll(outs[0][0], 37, axis=1), outs[1][0])
    assert np.array_equal(np.roll(outs[0][1], 37, axis=1), outs[1][1])


@pytest.mark.parametrize("nT", [512, 2048])
def test_projection_properties_at_full_size(K, nT):
    """BASELINE sizes (512 x 1024, 2048 x 4096): the projection is linear (P(2u) = 2 P(u) exactly
    in binary floating point), removes most of the divergence, leaves a zonal (phi-independent)
    flow untouched, and the pressure has zero zonal mean (n = 0 is never projected)."""
    N = 2 * nT
    p = oa.params(nT)
    u, v = oa.init_velocity(nT)
    with K.KaminoSolver(N, nT, 5.0, 0.005, initVelocity=False) as s:
        def project(uu, vv):
            set_velocity(s, uu, vv)
            s.projection()
            st = state(s)
            return st["velPhi"], st["velTheta"], s.pressure.copyBackToCPU().copy()
        u1, v1, p1 = project(u, v)
        u2, v2, p2 = project(u * np.float32(2), v * np.float32(2))
        assert np.array_equal(u2, u1 * np.float32(2)) and np.array_equal(v2, v1 * np.float32(2))
        assert np.abs(p1.astype(np.float64).mean(axis=1)).max() <= 1e-5 * np.abs(p1).max()
        d0 = np.zeros(nT * N, np.float32); d1 = np.zeros(nT * N, np.float32)
        oa.lib().ko_divergence(ctypes.byref(p), oa.fptr(u), oa.fptr(v), oa.fptr(d0))
        oa.lib().ko_divergence(ctypes.byref(p), oa.fptr(u1), oa.fptr(v1), oa.fptr(d1))
        k0 = d0.reshape(nT, N)[2:-2].astype(np.float64); k1 = d1.reshape(nT, N)[2:-2].astype(np.float64)
        k0 -= k0.mean(axis=1, keepdims=True); k1 -= k1.mean(axis=1, keepdims=True)
        print("nTheta %d: non-zonal divergence norm %.3e -> %.3e" % (nT, np.linalg.norm(k0), np.linalg.norm(k1)))
        assert np.linalg.norm(k1) <= 0.2 * np.linalg.norm(k0)
        # zonal flow: u_phi depends on theta only, u_theta = 0 -> divergence has only n = 0 -> unchanged
        uz = np.repeat(np.sin((np.arange(nT) + 0.5) * float(p.gridLen)).astype(np.float32), N)
        vz = np.zeros((nT - 1) * N, np.float32)
        u3, v3, p3 = project(uz, vz)
        assert np.array_equal(u3, uz) and np.array_equal(v3, vz)


@pytest.mark.parametrize("nT", [512, 2048])
def test_advection_properties_at_full_size(K, nT):
    """Uniform density stays uniform to 1 ulp; zero velocity is a fixed point of the advection."""
    N = 2 * nT
    with K.KaminoSolver(N, nT, 5.0, 0.005) as s:
        s.density.cpuBuffer[:] = np.float32(0.75)
        s.density.copyToGPU()
        s.advection()
        rho = s.density.copyBackToCPU()
        assert np.abs(rho - np.float32(0.75)).max() <= 6e-8
    # zero velocity: advection alone is the identity on density (the geometric phase is NOT a
    # fixed point at u = 0 in the reference either: its Cardano branch returns the difference
    # of two O(1/|G|) terms, kernel/KaminoCore.cu:443-452)
    with K.KaminoSolver(N, nT, 5.0, 0.005, initVelocity=False) as s:
        rho0 = oa.synthetic_density(nT).reshape(nT, N)
        s.density.cpuBuffer[:] = rho0
        s.density.copyToGPU()
        s.advection()
        st = state(s)
        assert not st["velPhi"].any() and not st["velTheta"].any()
        assert oa.rel_l2(st["density"], rho0) <= 1e-6


def test_run_frames_matches_stepping(K):
    from kaminogpu_b200 import capi
    g = oa.golden("t32")
    nT, N = 32, 64
    with make_solver(K, g) as a, make_solver(K, g) as b:
        a.stepForward(nSteps=6)
        sa = state(a)
        n = b.particles.numOfParticles
        hU = np.zeros(nT * N, np.float32); hV = np.zeros((nT - 1) * N, np.float32)
        hR = np.zeros(nT * N, np.float32); hP = np.zeros(2 * n, np.float32)
        capi.check(b._lib.kamino_run_frames(b._ctx, 2, 3, hU.ctypes.data, hV.ctypes.data, hR.ctypes.data,
                                            hP.ctypes.data), b._ctx)
    assert np.array_equal(hU, sa["velPhi"]) and np.array_equal(hV, sa["velTheta"])
    assert np.array_equal(hR, sa["density"]) and np.array_equal(hP, sa["particles"])


def test_error_reporting(K):
    from kaminogpu_b200 import capi
    with K.KaminoSolver(64, 32, 5.0, 0.005, initVelocity=False) as s:
        buf = np.zeros(64 * 32, np.float32)
        assert s._lib.kamino_upload_field(s._ctx, 9, 0, buf.ctypes.data) == capi_err("INVALID")
        assert s._lib.kamino_upload_field(s._ctx, 0, 5, buf.ctypes.data) == capi_err("INVALID")
        assert b"range" in s._lib.kamino_last_error(s._ctx)
        assert s._lib.kamino_step(s._ctx, -1) == capi_err("INVALID")


def capi_err(name):
    return {"INVALID": 10001, "NO_DEVICE": 10002, "STATE": 10003}[name]


# BASELINE.json configs 1-4 (C5 cannot be launched by the reference, kernel/KaminoCore.cu:779-784)
LIVE_CASES = {"c1": (128, 200), "c4": (256, 1), "c2": (512, 2), "c3": (2048, 1)}


@pytest.mark.timeout(900)
@pytest.mark.parametrize("case", list(LIVE_CASES))
def test_live_reference_build_at_baseline_sizes(K, tmp_path, case):
    """Run the reference's own CUDA build (oracle/_ref/kamino_ref, compiled from /root/reference by
    oracle/ref_harness) HERE at every BASELINE size it can launch -- C1 128 x 256 with its real 6,552,200
    particles, one member of the C4 ensemble (256 x 512), C2 512 x 1024 with 1,048,352 particles, C3
    2048 x 4096 -- and compare phase by phase from its own states.

    Measured (r02c, profiles/r02_parity_table.md): advection, particles and geometric are 100.0000 %
    bit-identical at all four sizes. Projection: with the theta solve done in the reference's
    cyclic-reduction order (kamino_debug_project_cr) the pressure is 2.2e-6 .. 7.2e-6 from the reference's,
    i.e. the distance of the default (LU) pressure -- 3.7e-6 .. 1.3e-4 -- is the reference's CR rounding.
    u_theta and u_phi differentiate the pressure (1/h, 1/(h sin(theta))), which amplifies the rounding noise
    cuFFT leaves in p when it adds and re-subtracts the n = 0 mode (kernel/KaminoCore.cu:692-700): the
    reference's own u_theta / u_phi are 3.4e-6 .. 1.5e-4 / 7.7e-4 .. 1.1e-1 from an fp64 evaluation of
    its operator, ours are 1.2e-7 .. 1.3e-6 / 1.2e-6 .. 1.8e-5, and |ours - reference| equals
    |reference - fp64| to three digits. The bars below state exactly that."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "kamino_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/kamino_ref not built (needs /root/reference at build time)")
    nT, pdens = LIVE_CASES[case]
    N = 2 * nT
    out = subprocess.run([exe, "dump", str(nT), str(pdens), "0.005", "5.0", "1", str(tmp_path), "-", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    ld = lambda tag, f: np.fromfile(str(tmp_path / ("%s.%s.f32" % (tag, f))), dtype=np.float32)
    with K.KaminoSolver(N, nT, 5.0, 0.005) as s:
        assert np.array_equal(s.velPhi.cpuBuffer.ravel(), ld("init", "velPhi"))
        s.density.cpuBuffer[:] = ld("init", "density").reshape(nT, N)
        s.density.copyToGPU()
        s.initParticlesfromPic("", pdens)
        assert np.array_equal(s.particles.coordCPUBuffer, ld("init", "particles"))
        s.advection()
        st = state(s)
        for name in ("velPhi", "velTheta", "density", "particles"):
            w = words_equal(st[name], ld("s1_adv", name))
            print("live reference %s advection %-9s identical words %.6f" % (case, name, w))
            assert oa.rel_l2(st[name], ld("s1_adv", name)) <= 1e-5 and w >= 0.9999
        s.geometric()
        st = state(s)
        for name in ("velPhi", "velTheta"):
            w = words_equal(st[name], ld("s1_geo", name))
            print("live reference %s geometric %-9s identical words %.6f" % (case, name, w))
            assert oa.rel_l2(st[name], ld("s1_geo", name)) <= 1e-5 and w >= 0.9999
        # projection from the reference's own post-geometric state, default (LU) and CR-order theta solve
        ug, vg = ld("s1_geo", "velPhi"), ld("s1_geo", "velTheta")
        u64, v64, p64 = fp64_projection(nT, ug, vg)
        exact = {"velPhi": u64, "velTheta": v64, "pressure": p64}
        got = {}
        for mode in ("lu", "cr"):
            set_velocity(s, ug, vg)
            (s.projection if mode == "lu" else s.projection_cr_order)()
            stp = state(s)
            got[mode] = {"velPhi": stp["velPhi"], "velTheta": stp["velTheta"], "pressure": s.pressure.copyBackToCPU().ravel().copy()}
        for name in ("velPhi", "velTheta", "pressure"):
            ref = ld("s1_proj", name)
            e_ref, e64, r64 = oa.rel_l2(got["lu"][name], ref), oa.rel_l2(got["lu"][name], exact[name]), oa.rel_l2(ref, exact[name])
            e_cr = oa.rel_l2(got["cr"][name], ref)
            print("live reference %s projection %-9s vs reference %.2e | vs fp64: ours %.2e reference %.2e | CR order vs reference %.2e"
                  % (case, name, e_ref, e64, r64, e_cr))
            # ours is the more accurate evaluation of the reference's operator ...
            assert e64 <= 1e-5 or e64 <= 0.2 * r64
            # ... and what separates the two builds is the reference's own distance from it
            assert e_ref <= 1.1 * r64 + 1e-6
        # the pressure gap is the reference's cyclic-reduction rounding: gone when we solve in its order
        assert oa.rel_l2(got["cr"]["pressure"], ld("s1_proj", "pressure")) <= 1e-5


def test_cli_runs_config_file(K, tmp_path):
    """The compiled drop-in: `kamino configKamino.txt` with the reference's grammar, output on,
    produces plain (uncompressed, as Partio::write does for a .bgeo name) classic-bgeo frames and the reference's progress lines."""
    import gzip
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = tmp_path / "configKamino.txt"
    (tmp_path / "out").mkdir()
    cfg.write_text("5.0 32 4.0 0.005 0.041666668 2 0.0 1 1 1 1 %s/out/f %s/out/p null null null\n" % (tmp_path, tmp_path))
    out = subprocess.run([os.path.join(root, "kaminogpu_b200", "kamino"), str(cfg)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "Frame 2 is ready" in out.stdout and "frames per second" in out.stdout
    for stem, npts in (("f", 32 * 64), ("p", 8192)):
        for frame in (0, 1, 2):
            raw = open(str(tmp_path / "out" / ("%s%d.bgeo" % (stem, frame))), "rb").read()
            assert raw[:5] == b"BgeoV"
            version, points = struct.unpack(">ii", raw[5:13])
            assert version == 5 and points == npts


@pytest.mark.timeout(600)
def test_reference_executable_and_drop_in_on_the_shipped_scenario(K, tmp_path):
    """BASELINE config 1 through the executables: the reference's own `main` (oracle/_ref/kamino_ref_exe: all nine
    reference TUs incl. kernel/main.cu, Partio / OpenCV replaced by the link shim, i.e. output off) and the drop-in
    `kamino` on the same configKamino.txt -- the Kamino defaults of include/KaminoGPU.cuh:40-44: 128 x 256, particle
    density 200 (6,552,200 particles), dt 0.005, 10 frames of 1/24 s (100 steps). Both must run to completion and
    announce the same frames; the drop-in's step count per frame is the reference's loop rule
    (kernel/KaminoCore.cu:886-895). State-level parity at this size is test_live_reference_build_at_baseline_sizes[c1]."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_exe = os.path.join(root, "oracle", "_ref", "kamino_ref_exe")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/kamino_ref_exe not built (needs /root/reference at build time)")
    (tmp_path / "ref").mkdir()
    (tmp_path / "ours").mkdir()
    tokens = "5.0 128 200.0 0.005 0.041666668 10 0.0 1 1 1 1 %s %s null null null\n"
    (tmp_path / "ref" / "configKamino.txt").write_text(tokens % ("f", "p"))            # the shim's Partio::write is a no-op
    (tmp_path / "ours" / "configKamino.txt").write_text(tokens % ("null", "null"))     # output off
    ref = subprocess.run([ref_exe, "configKamino.txt"], cwd=str(tmp_path / "ref"), capture_output=True, text=True, timeout=400)
    assert ref.returncode == 0, ref.stderr[-500:]
    ours = subprocess.run([os.path.join(root, "kaminogpu_b200", "kamino"), "configKamino.txt"], cwd=str(tmp_path / "ours"),
                          capture_output=True, text=True, timeout=400)
    assert ours.returncode == 0, ours.stderr[-500:]
    frames = lambda out: [l.strip() for l in out.splitlines() if l.startswith("Frame ")]
    assert frames(ref.stdout) == frames(ours.stdout) == ["Frame %d is ready" % i for i in range(1, 11)]
    for out in (ref.stdout, ours.stdout):
        assert "Initializing velocity..." in out and "frames per second" in out
    assert sum(K.steps_per_frame(0.005, 0.041666668, 10)) == 100


def test_cli_image_driven_initialisation(K, tmp_path):
    """densityImage / colorImage tokens of configKamino.txt (kernel/main.cu:40-45): the density and the
    particle colours come from the mirrored, resized image (kernel/KaminoSolver.cu:243-277,
    kernel/KaminoParticles.cu:64-72). Frame 0 is written before any step, so its density attribute must
    equal the committed cv2-derived golden bit for bit."""
    import gzip
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    image = os.path.join(root, "tests", "golden", "images", "smooth.png")
    cfg = tmp_path / "configKamino.txt"
    (tmp_path / "out").mkdir()
    cfg.write_text("5.0 32 1.0 0.005 0.041666668 1 0.0 1 1 1 1 %s/out/f %s/out/p %s null %s\n" % (tmp_path, tmp_path, image, image))
    out = subprocess.run([os.path.join(root, "kaminogpu_b200", "kamino"), str(cfg)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "No density image provided" not in out.stderr and "No particle color image provided" not in out.stderr
    want = np.load(os.path.join(root, "tests", "golden", "image_init.npz"))["smooth.png.32x64.density"]

    def read_bgeo(path):
        """-> {attribute name: (points x count) float32}, points (positions are implicit, 4 floats)."""
        import struct
        raw = open(path, "rb").read()
        assert raw[:5] == b"BgeoV"          # plain bytes: Partio compresses only *.gz names
        version, points, _, _, _, nattr, _, _, _ = struct.unpack(">9i", raw[5:41])
        pos, attrs = 41, []
        for _ in range(nattr):
            (ln,) = struct.unpack(">H", raw[pos:pos + 2]); pos += 2
            name = raw[pos:pos + ln].decode(); pos += ln
            count, _type = struct.unpack(">HI", raw[pos:pos + 6]); pos += 6 + 4 * count
            attrs.append((name, count))
        per = 4 + sum(c for _, c in attrs)
        data = np.frombuffer(raw[pos:pos + 4 * per * points], dtype=">f4").reshape(points, per).astype(np.float32)
        out, col = {}, 4
        for name, count in attrs:
            out[name] = data[:, col:col + count]; col += count
        return out

    grid = read_bgeo(str(tmp_path / "out" / "f0.bgeo"))
    # the writer walks the grid theta-major or phi-major; either way the multiset of values is the golden's
    got = np.sort(grid["density"].ravel())
    assert np.array_equal(got.view(np.uint32), np.sort(want.ravel()).view(np.uint32))
    parts = read_bgeo(str(tmp_path / "out" / "p0.bgeo"))
    assert parts["color"].shape == (2048, 3) and parts["color"].max() > 0.0      # particles are not black


@pytest.mark.timeout(600)
def test_cli_restart_from_checkpoint_continues_bit_for_bit(K, tmp_path):
    """KAMINO_CHECKPOINT / KAMINO_RESTART (raw state checkpoint): frames 1-2 with a
    checkpoint, then a second process resuming for frames 3-4, must write the bytes an uninterrupted
    4-frame run writes."""
    import gzip
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "kaminogpu_b200", "kamino")

    def run(tag, frames, env_extra):
        d = tmp_path / tag
        d.mkdir()
        cfg = d / "configKamino.txt"
        cfg.write_text("5.0 32 4.0 0.005 0.041666668 %d 0.0 1 1 1 1 %s/f %s/p null null null\n" % (frames, d, d))
        out = subprocess.run([exe, str(cfg)], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env_extra))
        assert out.returncode == 0, out.stderr
        return d, out.stdout

    whole, _ = run("whole", 4, {})
    ck = str(tmp_path / "state.ck")
    first, _ = run("first", 2, {"KAMINO_CHECKPOINT": ck})
    assert os.path.exists(ck)
    # the resumed run uses the same output directory layout; only frames 3 and 4 are written by it
    second, stdout = run("second", 4, {"KAMINO_RESTART": ck})
    assert "Resuming after frame 2" in stdout
    for stem in ("f", "p"):
        for frame in (1, 2):
            assert open(str(first / ("%s%d.bgeo" % (stem, frame))), "rb").read() == open(str(whole / ("%s%d.bgeo" % (stem, frame))), "rb").read()
        for frame in (3, 4):
            assert open(str(second / ("%s%d.bgeo" % (stem, frame))), "rb").read() == open(str(whole / ("%s%d.bgeo" % (stem, frame))), "rb").read()
        assert not os.path.exists(str(second / ("%s0.bgeo" % stem)))


# ---- theta-band decomposition, C++ path (csrc/dist.cu, kamino_dist_*): virtual ranks on one GPU ----------

def _single_gpu_steps(K, nT, steps, scale=1.0):
    N = 2 * nT
    rho0 = oa.synthetic_density(nT).reshape(nT, N)
    with K.KaminoSolver(N, nT, 5.0, 0.005) as s:
        u0, v0 = s.velPhi.cpuBuffer.copy() * np.float32(scale), s.velTheta.cpuBuffer.copy() * np.float32(scale)
        set_velocity(s, u0, v0)
        s.density.cpuBuffer[:] = rho0
        s.density.copyToGPU()
        s.stepForward(nSteps=steps)
        st = state(s)
        pr = s.pressure.copyBackToCPU().copy()
    return u0, v0, rho0, st, pr


@pytest.mark.parametrize("peer_stores", [False, True], ids=["staged", "peer-stores"])
@pytest.mark.parametrize("nT,world", [(128, 4), (256, 2), (512, 8), (128, 1)])
def test_dist_virtual_ranks_bit_identical_to_single_gpu(K, nT, world, peer_stores):
    """kamino_dist_*: band-sized buffers, the FFTs writing / reading the transposes' wire layouts, the theta
    solve on the rank's wavenumber band with band-only LU tables. P virtual ranks on one device
    (kamino_dist_group_step) against kamino_step: u_phi, u_theta, density and pressure bit-identical after 3 steps.
    Both transports of the transposes: staged buffers moved by device copies (where the NCCL ranks send / receive), and
    peer stores (the FFT kernel and the theta solve write straight into the sibling ranks' buffers, as the NCCL ranks do
    through CUDA IPC mappings)."""
    from kaminogpu_b200 import capi, dist
    u0, v0, rho0, st, pr = _single_gpu_steps(K, nT, 3)
    grp = dist.LocalGroup(nT, 5.0, 0.005, world, peer_stores=peer_stores)
    try:
        grp.upload_global(capi.VEL_PHI, u0)
        grp.upload_global(capi.VEL_THETA, v0)
        grp.upload_global(capi.DENSITY, rho0)
        grp.step(2)
        grp.step(1)
        grp.sync()
        got = {"velPhi": grp.gather(capi.VEL_PHI), "velTheta": grp.gather(capi.VEL_THETA), "density": grp.gather(capi.DENSITY),
               "pressure": grp.gather(capi.PRESSURE)}
        held = [r.device_bytes for r in grp.ranks]
    finally:
        grp.close()
    want = dict(st, pressure=pr)
    for name in ("velPhi", "velTheta", "density", "pressure"):
        w = words_equal(got[name], want[name])
        print("dist x%d nTheta %d %s %-9s identical words %.6f" % (world, nT, "peer-stores" if peer_stores else "staged", name, w))
        assert np.array_equal(got[name].ravel().view(np.uint32), np.asarray(want[name], np.float32).ravel().view(np.uint32)), name
    if world > 1:
        # band-sized memory: a rank holds (rows + 2 x 24 halo) rows of 7 field buffers, 3 spectrum buffers, 1/P of the tables
        assert max(held) < 0.8 * held[0] * world or world == 2


def test_dist_init_velocity_rows_matches_full_field(K):
    from kaminogpu_b200 import capi, dist
    nT, world = 128, 4
    u, v = oa.init_velocity(nT)
    grp = dist.LocalGroup(nT, 5.0, 0.005, world)
    try:
        grp.init_velocity()
        assert np.array_equal(grp.gather(capi.VEL_PHI).ravel(), u)
        assert np.array_equal(grp.gather(capi.VEL_THETA).ravel(), v)
    finally:
        grp.close()


def test_dist_reports_a_backtrace_that_leaves_the_halo(K):
    """A theta-CFL the 24-row halo cannot cover must not pass silently (r01 ADVICE): the sampler clamps the
    gather into the resident rows, raises the device flag and kamino_dist_sync returns KAMINO_ERR_STATE."""
    from kaminogpu_b200 import capi, dist
    nT, world = 128, 4
    u, v = oa.init_velocity(nT)
    grp = dist.LocalGroup(nT, 5.0, 0.005, world)
    try:
        grp.upload_global(capi.VEL_PHI, u)
        grp.upload_global(capi.VEL_THETA, v * np.float32(60.0))         # theta-CFL ~ 40 rows
        grp.upload_global(capi.DENSITY, oa.synthetic_density(nT))
        grp.step(1)
        with pytest.raises(capi.KaminoError) as err:
            grp.sync()
        assert err.value.code == capi_err("STATE") and "halo" in str(err.value)
        # an ordinary run does not trip it
        grp.upload_global(capi.VEL_PHI, u)
        grp.upload_global(capi.VEL_THETA, v)
        grp.step(2)
        grp.sync()
    finally:
        grp.close()


# ---- GPU-side initialisers ----------------------------------------------------------------

@pytest.mark.parametrize("nT", [128, 512])
def test_device_velocity_initialiser_against_the_host_one(K, nT):
    """kamino_init_velocity_device restates the reference's FBM initialiser as device code, operation for operation;
    only CUDA's double sin() inside the lattice hash can differ from glibc's (where the two results straddle an fp32
    rounding boundary). Measured identical fraction is printed; bars: >= 99.99 % identical words, relative L2 <= 1e-6."""
    u, v = oa.init_velocity(nT)
    with K.KaminoSolver(2 * nT, nT, 5.0, 0.005, initVelocity=False) as s:
        s.initialize_velocity_on_device()
        st = state(s)
    for name, want in (("velPhi", u), ("velTheta", v)):
        w, e = words_equal(st[name], want), oa.rel_l2(st[name], want)
        print("device initialiser nTheta %d %-8s identical words %.6f relL2 %.2e" % (nT, name, w, e))
        assert w >= 0.9999 and e <= 1e-6


def test_device_velocity_initialiser_of_a_band_equals_the_whole_grid(K):
    from kaminogpu_b200 import capi, dist
    nT, world = 128, 4
    with K.KaminoSolver(2 * nT, nT, 5.0, 0.005, initVelocity=False) as s:
        s.initialize_velocity_on_device()
        st = state(s)
    grp = dist.LocalGroup(nT, 5.0, 0.005, world)
    try:
        for r in grp.ranks:
            r.init_velocity_on_device()
        assert np.array_equal(grp.gather(capi.VEL_PHI).ravel(), st["velPhi"])
        assert np.array_equal(grp.gather(capi.VEL_THETA).ravel(), st["velTheta"])
    finally:
        grp.close()


def test_device_particle_seeding_is_the_reference_lattice_with_counter_based_jitter(K):
    nT, dens = 64, 4.0
    spacing = np.float32(np.pi / nT / 2.0)
    def seeded(seed):
        with K.KaminoSolver(2 * nT, nT, 5.0, 0.005, initVelocity=False) as s:
            s.seed_particles_on_device(dens, seed)
            return s.particles.copyBack2CPU().reshape(-1, 2).copy()
    a, b, c = seeded(1234), seeded(1234), seeded(99)
    numTheta = int(np.float32(2.0) * nT)
    assert a.shape == (2 * numTheta * numTheta, 2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)            # reproducible per seed
    i, j = np.divmod(np.arange(a.shape[0]), numTheta)                   # index i * numTheta + j (KaminoParticles.cu:59)
    assert np.all(np.abs(a[:, 0] - i * spacing) <= spacing / 2 * 1.0001) and np.all(np.abs(a[:, 1] - j * spacing) <= spacing / 2 * 1.0001)
    assert a.min() >= 0.0                                               # clamped at 0 (:49-54)
    jitter = (a[:, 0] - i * spacing)[i > 0] / (spacing / 2)
    assert abs(jitter.mean()) < 0.02 and 0.55 < jitter.std() < 0.60     # uniform on [-1, 1]: std 1 / sqrt(3)
    with K.KaminoSolver(2 * nT, nT, 5.0, 0.005, initVelocity=False) as s:      # wrong particle count is refused
        from kaminogpu_b200 import capi
        capi.check(s._lib.kamino_alloc_particles(s._ctx, 10), s._ctx)
        assert s._lib.kamino_seed_particles_device(s._ctx, ctypes.c_float(dens), ctypes.c_ulonglong(1)) == capi_err("STATE")


# ---- theta-band decomposition, Python prototype + SPIKE research path (banded.py): virtual ranks on one GPU ----

@pytest.mark.parametrize("nT,world", [(128, 4), (256, 2)])
def test_banded_step_is_bit_identical_to_single_gpu(K, nT, world):
    """P virtual ranks (kaminogpu_b200.banded.LocalGroup: the per-rank phases of the multi-GPU
    driver with device copies in place of NCCL) against the ordinary single-context step:
    u_phi, u_theta and density bit-identical after 3 steps."""
    from kaminogpu_b200 import banded
    N = 2 * nT
    rho0 = oa.synthetic_density(nT).reshape(nT, N)
    with K.KaminoSolver(N, nT, 5.0, 0.005) as s:
        s.density.cpuBuffer[:] = rho0
        s.density.copyToGPU()
        s.stepForward(0.005, nSteps=3)
        s.sync()
        ref = state(s)
    grp = banded.LocalGroup(nT, 5.0, 0.005, world)
    try:
        for r in grp.ranks:
            r.solver.density.cpuBuffer[:] = rho0
            r.solver.density.copyToGPU()
        grp.step(3)
        u, v, rho = grp.gather()
    finally:
        grp.close()
    assert np.array_equal(u.ravel(), ref["velPhi"])
    assert np.array_equal(v.ravel(), ref["velTheta"])
    assert np.array_equal(rho.ravel(), ref["density"])


@pytest.mark.timeout(180)
@pytest.mark.parametrize("nT,world", [(128, 4), (256, 2)])
def test_banded_spike_mode_tracks_single_gpu(K, nT, world):
    """Reduced-interface (SPIKE) theta solve of the band-decomposed run (banded.SpikeInterface; band-local
    LU solves + a 2P x 2P interface system per wavenumber instead of the two all-to-all transposes): same
    systems, different operation order, so the comparison with the single-context step is at tolerance
    level. Bars after 2 steps: density and u_theta 1e-5 relative L2; u_phi 2e-4 (the 1/(h sin(theta))
    amplification of pressure round-off in the polar rows, as in the projection tests)."""
    from kaminogpu_b200 import banded
    N = 2 * nT
    rho0 = oa.synthetic_density(nT).reshape(nT, N)
    with K.KaminoSolver(N, nT, 5.0, 0.005) as s:
        s.density.cpuBuffer[:] = rho0
        s.density.copyToGPU()
        s.stepForward(0.005, nSteps=2)
        s.sync()
        ref = state(s)
    grp = banded.LocalGroup(nT, 5.0, 0.005, world, solve="spike")
    try:
        for r in grp.ranks:
            r.solver.density.cpuBuffer[:] = rho0
            r.solver.density.copyToGPU()
        grp.step(2)
        u, v, rho = grp.gather()
    finally:
        grp.close()
    for name, got, tol in (("velPhi", u, 2e-4), ("velTheta", v, 1e-5), ("density", rho, 1e-5)):
        err = oa.rel_l2(got.ravel(), ref[name])
        print("banded spike x%d, nTheta %d: %-8s relL2 vs single GPU %.2e" % (world, nT, name, err))
        assert err <= tol, name
