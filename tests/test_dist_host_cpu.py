"""Band-decomposed C++ path (kamino_dist_*, csrc/dist.cu): what can be checked without a GPU.

  * the partition arithmetic and the argument validation of the C ABI (no device: NO_DEVICE, never a CPU path);
  * the per-band host initialiser against the full field (bit-identical);
  * the one torch.distributed use on that path -- handing rank 0's 128-byte NCCL unique id to the other ranks --
    under gloo at world size 2.
The step loop itself (NCCL send / recv around the kernels) is covered on the GPU: virtual ranks in
tests/test_parity_gpu.py, real ranks by scripts/dist_check.py.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle_api as oa


def test_band_partition():
    from kaminogpu_b200.dist import band_of
    assert band_of(8192, 8, 0) == (0, 1024) and band_of(8192, 8, 7) == (7168, 8192)
    assert band_of(512, 2, 1) == (256, 512) and band_of(128, 1, 0) == (0, 128)
    for bad in ((128, 8, 0), (8192, 3, 0), (100, 2, 0)):
        with pytest.raises(ValueError):
            band_of(*bad)


def test_dist_create_rejects_bad_arguments_and_has_no_cpu_path(built):
    from kaminogpu_b200 import capi
    lib = capi.load()
    h = ctypes.c_void_p()
    create = lambda nT, rank, world: lib.kamino_dist_create(ctypes.byref(h), 0, nT, ctypes.c_float(5.0), ctypes.c_float(0.005), rank, world, None)
    assert create(100, 0, 1) == 10001 and b"power of two" in lib.kamino_dist_last_error(None)
    assert create(512, 2, 2) == 10001            # rank out of range
    assert create(512, 0, 3) == 10001            # world not a power of two
    assert create(128, 0, 8) == 10001            # 16-row bands are thinner than the halo
    import torch
    if not torch.cuda.is_available():
        assert create(512, 0, 2) == 10002        # KAMINO_ERR_NO_DEVICE: no CPU fallback
        assert h.value is None
    assert lib.kamino_dist_step(None, 1) == 10001 and lib.kamino_dist_sync(None) == 10001


@pytest.mark.parametrize("nT,world", [(64, 2), (128, 4)])
def test_band_initialiser_equals_the_full_field(built, nT, world):
    from kaminogpu_b200 import capi
    from kaminogpu_b200.dist import band_of
    lib = capi.load()
    u, v = oa.init_velocity(nT)
    u, v = u.reshape(nT, 2 * nT), v.reshape(nT - 1, 2 * nT)
    for rank in range(world):
        lo, hi = band_of(nT, world, rank)
        ub = np.full((hi - lo, 2 * nT), np.nan, np.float32)
        vb = np.full((hi - lo, 2 * nT), np.nan, np.float32)
        assert lib.kamino_init_velocity_host_rows(nT, ctypes.c_float(5.0), lo, hi - lo, ub.ctypes.data, vb.ctypes.data) == 0
        assert np.array_equal(ub, u[lo:hi])
        nv = min(hi, nT - 1) - lo
        assert np.array_equal(vb[:nv], v[lo:lo + nv])
    assert lib.kamino_init_velocity_host_rows(nT, ctypes.c_float(5.0), nT - 8, 16, ub.ctypes.data, vb.ctypes.data) == 10001


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from kaminogpu_b200.dist import broadcast_unique_id
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        calls = []

        def make_id():
            calls.append(rank)
            return bytes((7 * k + 3) % 256 for k in range(128))
        got = broadcast_unique_id(make_id)
        out.put((rank, got, calls))
    finally:
        dist.destroy_process_group()


def test_unique_id_reaches_every_rank_under_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port, world = _free_port(), 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = bytes((7 * k + 3) % 256 for k in range(128))
    for rank, got, calls in results:
        assert got == want
        assert calls == ([0] if rank == 0 else [])        # only rank 0 asks NCCL for an id
