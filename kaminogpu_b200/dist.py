"""Theta-band decomposition of one simulation over P GPUs: thin ctypes mirror of the kamino_dist_* C ABI
(include/kamino_b200.h, csrc/dist.cu). The step loop, the NCCL halo exchange and the NCCL transposes around
the theta solve all run in C++; this module only creates the ranks and moves host arrays. torch.distributed
is used for ONE thing: broadcasting the 128-byte NCCL unique id at start-up.

    DistributedSolver   one rank of a run launched with one process per GPU (torchrun)
    LocalGroup          P virtual ranks on one device (device copies instead of NCCL): tests
"""
import ctypes

import numpy as np

from . import capi


def _check(code, handle=None):
    if code != 0:
        msg = capi.load().kamino_dist_last_error(handle)
        raise capi.KaminoError(code, msg.decode() if msg else "")


class _Rank:
    """One kamino_dist handle plus host-side convenience."""

    def __init__(self, nTheta, radius, dt, rank, world, device, unique_id):
        self.lib = capi.load()
        self.nTheta, self.nPhi, self.rank, self.world = nTheta, 2 * nTheta, rank, world
        self.radius = float(radius)
        h = ctypes.c_void_p()
        _check(self.lib.kamino_dist_create(ctypes.byref(h), device, nTheta, ctypes.c_float(radius), ctypes.c_float(dt),
                                           rank, world, unique_id))
        self.handle = h
        lo, hi, k0, k1, nbytes = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        _check(self.lib.kamino_dist_shape(h, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(k0), ctypes.byref(k1),
                                          ctypes.byref(nbytes)), h)
        self.lo, self.hi, self.device_bytes = lo.value, hi.value, nbytes.value

    def close(self):
        if getattr(self, "handle", None):
            self.lib.kamino_dist_destroy(self.handle)
            self.handle = None

    def rows_of(self, field):
        end = min(self.hi, self.nTheta - 1) if field == capi.VEL_THETA else self.hi
        return end - self.lo

    def upload(self, field, rows):
        """rows: this rank's rows of the field (rows_of(field) x nPhi)."""
        a = np.ascontiguousarray(rows, dtype=np.float32)
        assert a.size == self.rows_of(field) * self.nPhi, "upload: wrong number of rows for this band"
        _check(self.lib.kamino_dist_upload(self.handle, field, a.ctypes.data_as(ctypes.c_void_p)), self.handle)

    def upload_global(self, field, whole):
        """whole: the global field; the rank keeps its own rows."""
        whole = np.asarray(whole, dtype=np.float32).reshape(-1, self.nPhi)
        self.upload(field, whole[self.lo:self.lo + self.rows_of(field)])

    def download(self, field):
        out = np.empty((self.rows_of(field), self.nPhi), dtype=np.float32)
        _check(self.lib.kamino_dist_download(self.handle, field, out.ctypes.data_as(ctypes.c_void_p)), self.handle)
        return out

    def init_velocity(self):
        """The reference's FBM initial velocity (kernel/KaminoInitializer.cu:3-85), this band's rows only."""
        u = np.empty((self.hi - self.lo, self.nPhi), np.float32)
        v = np.empty((self.hi - self.lo, self.nPhi), np.float32)
        capi.check(self.lib.kamino_init_velocity_host_rows(self.nTheta, ctypes.c_float(self.radius), self.lo, self.hi - self.lo,
                                                           u.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p)))
        self.upload(capi.VEL_PHI, u)
        self.upload(capi.VEL_THETA, v[:self.rows_of(capi.VEL_THETA)])
        return u, v

    def transport(self, peer_stores=None):
        """-> (uses peer-memory transposes, note). peer_stores = True / False switches (collective over the ranks)."""
        uses, note = ctypes.c_int(), ctypes.c_char_p()
        _check(self.lib.kamino_dist_transport(self.handle, -1 if peer_stores is None else int(bool(peer_stores)),
                                              ctypes.byref(uses), ctypes.byref(note)), self.handle)
        return bool(uses.value), (note.value or b"").decode()

    def init_velocity_on_device(self):
        _check(self.lib.kamino_dist_init_velocity_device(self.handle), self.handle)

    def sync(self):
        _check(self.lib.kamino_dist_sync(self.handle), self.handle)

    def comm_stats(self, enable=-1):
        hs, ts, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_long()
        hb, tb = ctypes.c_size_t(), ctypes.c_size_t()
        _check(self.lib.kamino_dist_comm_stats(self.handle, enable, ctypes.byref(hs), ctypes.byref(ts), ctypes.byref(n),
                                               ctypes.byref(hb), ctypes.byref(tb)), self.handle)
        return {"halo_s": hs.value, "transpose_s": ts.value, "steps": n.value,
                "halo_bytes_per_step": hb.value, "transpose_bytes_per_step": tb.value}


def band_of(nTheta, world, rank):
    """(first row, one past the last row) of a rank: the partition kamino_dist_create uses (csrc/dist.cu)."""
    if world < 1 or world & (world - 1) or nTheta % world:
        raise ValueError("world must be a power of two that divides nTheta")
    rows = nTheta // world
    if world > 1 and (rows < 32 or rows % 8):
        raise ValueError("bands need at least 32 rows and a multiple of 8 wavenumbers per rank")
    return rank * rows, (rank + 1) * rows


def broadcast_unique_id(make_id, device=None):
    """Rank 0 calls make_id() -> 128 bytes (kamino_dist_unique_id); every rank of the torch.distributed group returns them.
    The one use of torch.distributed on the band-decomposed path."""
    import torch
    import torch.distributed as dist
    payload = bytes(make_id()) if dist.get_rank() == 0 else bytes(128)
    assert len(payload) == 128
    dev = torch.device("cuda", device) if (dist.get_backend() == "nccl" and device is not None) else torch.device("cpu")
    t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


class DistributedSolver(_Rank):
    """One rank under torch.distributed (any backend that can broadcast 128 bytes). Collective constructor."""

    def __init__(self, nTheta, radius, dt, device=0):
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        lib = capi.load()
        ident = None
        if world > 1:
            def make_id():
                buf = (ctypes.c_ubyte * 128)()
                _check(lib.kamino_dist_unique_id(buf))
                return bytes(buf)
            ident = (ctypes.c_ubyte * 128)(*broadcast_unique_id(make_id, device))
        super().__init__(nTheta, radius, dt, rank, world, device, ident)
        assert (self.lo, self.hi) == band_of(nTheta, world, rank)

    def step(self, nSteps=1):
        _check(self.lib.kamino_dist_step(self.handle, nSteps), self.handle)

    @property
    def cuda_stream(self):
        p = ctypes.c_void_p()
        _check(self.lib.kamino_dist_stream(self.handle, ctypes.byref(p)), self.handle)
        return p.value


class LocalGroup:
    """P virtual ranks in one process on one device (kamino_dist_group_step)."""

    def __init__(self, nTheta, radius, dt, world, device=0, peer_stores=False):
        """peer_stores: the kernels store straight into the sibling ranks' buffers (the peer-memory transposes of the
        NCCL ranks) instead of staging + device copies."""
        self.ranks = [_Rank(nTheta, radius, dt, r, world, device, None) for r in range(world)]
        self.world, self.nTheta, self.nPhi = world, nTheta, 2 * nTheta
        if peer_stores and world > 1:
            for r in self.ranks:
                assert r.transport(True)[0]
        self._handles = (ctypes.c_void_p * world)(*[r.handle for r in self.ranks])

    def close(self):
        for r in self.ranks:
            r.close()

    def upload_global(self, field, whole):
        for r in self.ranks:
            r.upload_global(field, whole)

    def init_velocity(self):
        for r in self.ranks:
            r.init_velocity()

    def step(self, nSteps=1):
        _check(capi.load().kamino_dist_group_step(self._handles, self.world, nSteps), self.ranks[0].handle)

    def sync(self):
        """Synchronise every rank (each clears its own halo-violation flag), then raise the first error."""
        first = None
        for r in self.ranks:
            try:
                r.sync()
            except capi.KaminoError as e:
                first = first or e
        if first is not None:
            raise first

    def gather(self, field):
        return np.concatenate([r.download(field) for r in self.ranks], axis=0)
