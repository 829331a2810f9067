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
