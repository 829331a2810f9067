"""Band decomposition: host-side logic and the communication layer under gloo (world size 2, CPU).

The GPU kernels are not involved: a CPU stand-in of a rank (BandLayout over CPU tensors) goes
through the same halo exchange / all-to-all / spectrum-row code paths as the NCCL run
(kaminogpu_b200/banded.py: exchange, all_to_all) and is checked against what a single global
array says every rank should hold afterwards.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_band_plan_ranges():
    from kaminogpu_b200.banded import BandPlan, HALO, ADVECT_EXTRA, GEO_EXTRA
    p = BandPlan(8192, 8)
    assert p.rows == 1024 and p.kper == 1024
    assert p.band(0) == (0, 1024) and p.band(7) == (7168, 8192)
    assert p.clipped(0, ADVECT_EXTRA) == (0, 1040) and p.clipped(7, GEO_EXTRA) == (7160, 8192)
    assert p.clipped(3, ADVECT_EXTRA) == (3056, 4112)
    # what the advection recomputes must be covered by the exchanged halo (backtraces reach < 4 rows)
    assert ADVECT_EXTRA + 4 <= HALO and GEO_EXTRA + 1 <= ADVECT_EXTRA and HALO <= 32
    with pytest.raises(ValueError):
        BandPlan(128, 8)        # 16-row bands are thinner than the halo
    with pytest.raises(ValueError):
        BandPlan(100, 3)
    assert BandPlan(64, 1).rows == 64   # a single band needs no halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nT, failures):
    import torch
    import torch.distributed as dist
    from kaminogpu_b200 import banded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        N = 2 * nT
        plan = banded.BandPlan(nT, world)

        class CpuRank(banded.BandLayout):
            pass

        r = CpuRank()
        r.plan, r.rank, r.world = plan, rank, world
        r.lo, r.hi = plan.band(rank)
        rows, kper = plan.rows, plan.kper
        jj = torch.arange(nT, dtype=torch.float32)[:, None]
        ii = torch.arange(N, dtype=torch.float32)[None, :]
        truth = [1000.0 * f + jj * 3.0 + ii * 0.001 for f in range(3)]          # global fields
        mine = []
        for t in truth:
            m = torch.full_like(t, -1.0)
            m[r.lo:r.hi] = t[r.lo:r.hi]                                         # I only hold my band
            mine.append(m)
        r.fields = lambda: mine
        banded.exchange(r.halo_sends(), r.halo_recvs(), dist)
        for t, m in zip(truth, mine):
            a, b = max(r.lo - banded.HALO, 0), min(r.hi + banded.HALO, nT)
            assert torch.equal(m[a:b], t[a:b]), "halo rows wrong on rank %d" % rank
            if a > 0:
                assert (m[:a] == -1).all()
            if b < nT:
                assert (m[b:] == -1).all()

        # spectrum transposes: S[theta][k][re/im] = theta + 1e-4 k (+0.5 for im)
        kk = torch.arange(plan.half, dtype=torch.float32)[None, :, None]
        S = (torch.arange(nT, dtype=torch.float32)[:, None, None] + 1e-4 * kk + torch.tensor([0.0, 0.5])[None, None, :])
        r.spectrum = torch.full_like(S, -1.0)
        r.spectrum[r.lo:r.hi] = S[r.lo:r.hi]
        r.send = torch.empty((world, rows, kper, 2))
        r.recv = torch.empty((world, rows, kper, 2))
        banded.all_to_all(r.recv, r.pack_forward(), dist)
        packed = r.recv.view(nT, kper, 2)
        assert torch.equal(packed, S[:, rank * kper:(rank + 1) * kper]), "forward transpose wrong on rank %d" % rank
        packed.mul_(2.0)                                                        # the "solve"
        banded.all_to_all(r.send, r.recv, dist)
        r.unpack_backward(r.send)
        assert torch.equal(r.spectrum[r.lo:r.hi], 2.0 * S[r.lo:r.hi]), "backward transpose wrong on rank %d" % rank
        sends = [(rank - 1, r.spectrum_row(r.lo))] if rank > 0 else []
        recvs = [(rank + 1, r.spectrum_row(r.hi))] if rank < world - 1 else []
        banded.exchange(sends, recvs, dist)
        if rank < world - 1:
            assert torch.equal(r.spectrum[r.hi], 2.0 * S[r.hi])
    except Exception as e:           # noqa: BLE001 - reported to the parent
        failures.put("rank %d: %r" % (rank, e))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_halo_exchange_and_transposes_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rank, 2, port, 64, failures)) for rank in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0, "worker exited with %r" % p.exitcode
    msgs = []
    while not failures.empty():
        msgs.append(failures.get())
    assert not msgs, msgs


# ---- reduced-interface (SPIKE) theta solve: the algebra and its communication under gloo --------------------

def _theta_system(nT, n):
    """a, b, c of wavenumber n as the reference builds them (kernel/KaminoSolver.cu:128-153), in fp64,
    Neumann-folded at the poles; a[0] and c[-1] are 0 after the fold."""
    h = np.pi / nT
    theta = (np.arange(nT) + 0.5) * h
    cot = np.cos(theta) / np.sin(theta) / (2.0 * h)
    a = 1.0 / h ** 2 - cot
    c = 1.0 / h ** 2 + cot
    b = -2.0 / h ** 2 - n * n / np.sin(theta) ** 2
    b[0] += a[0]; a[0] = 0.0
    b[-1] += c[-1]; c[-1] = 0.0
    return a, b, c


def _thomas(a, b, c, d):
    """d: [rows][...]; solves along axis 0."""
    n = len(b)
    cp = np.zeros(n)
    dp = np.zeros_like(d)
    cp[0] = c[0] / b[0]
    dp[0] = d[0] / b[0]
    for i in range(1, n):
        m = b[i] - a[i] * cp[i - 1]
        cp[i] = c[i] / m
        dp[i] = (d[i] - a[i] * dp[i - 1]) / m
    x = np.zeros_like(d)
    x[-1] = dp[-1]
    for i in range(n - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


def _spike_worker(rank, world, port, nT, failures):
    import torch
    import torch.distributed as dist
    from kaminogpu_b200 import banded
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        half = nT                                         # nPhi / 2 wavenumber slots; slot 0 = Nyquist
        plan = banded.BandPlan(nT, world)
        systems = [_theta_system(nT, half if k == 0 else k) for k in range(half)]
        rng = np.random.default_rng(5)
        rhs = rng.standard_normal((nT, half, 2))
        truth = np.empty_like(rhs)
        for k, (a, b, c) in enumerate(systems):
            truth[:, k] = _thomas(a, b, c, rhs[:, k])

        class CpuRank(banded.SpikeInterface):
            pass

        r = CpuRank()
        r.torch, r.rank, r.world = torch, rank, world
        r.lo, r.hi = plan.band(rank)
        r.spectrum = torch.full((nT, half, 2), float("nan"), dtype=torch.float64)   # other bands are never read
        r.spectrum[r.lo:r.hi] = torch.from_numpy(rhs[r.lo:r.hi])
        if r.hi < nT:
            r.spectrum[r.hi] = 0.0

        def local_solve():
            band = r.spectrum[r.lo:r.hi].numpy()
            for k, (a, b, c) in enumerate(systems):
                al, cl = a[r.lo:r.hi].copy(), c[r.lo:r.hi].copy()
                al[0] = 0.0
                cl[-1] = 0.0
                band[:, k] = _thomas(al, b[r.lo:r.hi], cl, band[:, k])
        r.local_solve = local_solve
        # the couplings are the same for every wavenumber (they do not contain n)
        r.coupling = lambda: (systems[1][0][r.lo] if r.lo > 0 else 0.0, systems[1][2][r.hi - 1] if r.hi < nT else 0.0)
        saved = r.spectrum[r.lo:r.hi].clone()
        r.spike_setup(dist)
        assert torch.equal(r.spectrum[r.lo:r.hi], saved), "setup must leave the band untouched"
        for _ in range(2):                                 # twice: the per-step path has no hidden state
            r.spectrum[r.lo:r.hi] = torch.from_numpy(rhs[r.lo:r.hi])
            r.spike_solve(dist)
            got = r.spectrum[r.lo:r.hi].numpy()
            err = np.linalg.norm(got - truth[r.lo:r.hi]) / np.linalg.norm(truth[r.lo:r.hi])
            assert err < 1e-10, "rank %d: band differs from the global solve by %.2e" % (rank, err)
            if r.hi < nT:
                row = r.spectrum[r.hi].numpy()
                assert np.linalg.norm(row - truth[r.hi]) <= 1e-10 * np.linalg.norm(truth[r.hi]), "next row wrong"
    except Exception as e:           # noqa: BLE001 - reported to the parent
        failures.put("rank %d: %r" % (rank, e))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_spike_solve_matches_global_solve(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    failures = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spike_worker, args=(rank, world, port, 128, failures)) for rank in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(280)
        assert p.exitcode == 0, "worker exited with %r" % p.exitcode
    msgs = []
    while not failures.empty():
        msgs.append(failures.get())
    assert not msgs, msgs
