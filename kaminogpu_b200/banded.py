"""Theta-band decomposition of ONE simulation over P GPUs (BASELINE config 5).

The reference is single-GPU (device 0 hard-coded, kernel/KaminoSolver.cu:20) and cannot even
launch its theta solve above nTheta = 2048 (kernel/KaminoCore.cu:779-782), so this mode has no
reference counterpart; its contract is bit-identity with this package's own single-GPU step,
which the tests check.

One process per GPU (torch.distributed, NCCL over NVLink). Rank r owns the theta rows
[r nT/P, (r+1) nT/P) of u_phi, u_theta, density and pressure. Every rank keeps FULL-SIZE
buffers (global row indices everywhere, 46 B/cell: 6.2 GB at 8192 x 16384) and asks each kernel of
the step for its rows only (kamino_band_* entry points of the C ABI). Per step:

  1. halo exchange   HALO = 24 rows of u_phi, u_theta, density from each neighbour (P2P send/recv)
  2. advection       on [lo-16, hi+16): the 16 extra rows are recomputed locally instead of being
                     exchanged a second time (backtraces reach < 4 rows: theta-CFL < 1, RK midpoint,
                     bilinear stencil)
  3. geometric       on [lo-8, hi+8)   (its staggered re-averaging reaches one row)
  4. divergence+FFT  on [lo, hi)       -> spectrum rows [lo, hi), all wavenumbers
  5. all-to-all      [band rows][all k] -> [all rows][k band]           (NCCL all_to_all_single)
  6. theta solve     for the rank's wavenumber band, in place on the packed receive buffer
  7. all-to-all back, plus one spectrum row from the next rank (the theta gradient of the last
     band row needs p of row hi)
  8. inverse FFT + gradient on [lo, hi)

Tracer particles are not band-decomposed (they would migrate between ranks); banded contexts
carry none.

`LocalGroup` runs the same per-rank phases for P virtual ranks inside one process, on one GPU,
replacing the NCCL calls by device copies: that is what the single-GPU parity test drives.
"""
import ctypes

import numpy as np

from . import capi
from .solver import KaminoSolver

# NOTE: this module is the Python RESEARCH PROTOTYPE of the band decomposition (kept as the vehicle of the reduced-interface
# "SPIKE" theta solve). The product path is C++: csrc/dist.cu behind kamino_dist_* (kaminogpu_b200/dist.py mirrors it), with
# band-sized buffers, peer-memory / NCCL transposes and a device-side check that no backtrace leaves the halo. Here every
# rank keeps full-size buffers and the halo width is only checked on the host (check_halo below).
HALO = 24          # rows exchanged with each neighbour
ADVECT_EXTRA = 16  # rows beyond the band that the advection recomputes
GEO_EXTRA = 8      # rows beyond the band that the geometric phase recomputes


class BandPlan:
    """Row / wavenumber ranges of every rank."""

    def __init__(self, nTheta, world):
        if nTheta % world:
            raise ValueError("nTheta must be divisible by the number of bands")
        self.nTheta, self.world = nTheta, world
        self.rows = nTheta // world
        self.half = nTheta                    # nPhi / 2 wavenumber slots
        self.kper = self.half // world
        if world > 1 and (self.rows < 32 or self.rows % 8 or self.kper % 8):
            raise ValueError("bands need >= 32 rows, a multiple of 8 rows and of 8 wavenumbers per rank")

    def band(self, r):
        return r * self.rows, (r + 1) * self.rows

    def clipped(self, r, extra):
        lo, hi = self.band(r)
        return max(lo - extra, 0), min(hi + extra, self.nTheta)


class _CudaArray:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class BandLayout:
    """Which rows / spectrum blocks a rank sends and receives: pure tensor slicing, shared by the
    GPU rank (BandRank) and the CPU stand-in of the gloo tests. Needs: plan, rank, world, lo, hi,
    fields() -> [u_phi, u_theta, density] full-size [nTheta][nPhi] tensors, spectrum
    [nTheta][nPhi/2][2], send / recv [world][rows][kper][2]."""

    # ---- halos: (tensor, row range) pairs ------------------------------------------------------
    def halo_sends(self):
        """[(to_rank, tensor)] : my top HALO rows go up, my bottom HALO rows go down."""
        ops = []
        for f in self.fields():
            if self.rank > 0:
                ops.append((self.rank - 1, f[self.lo:self.lo + HALO]))
            if self.rank < self.world - 1:
                ops.append((self.rank + 1, f[self.hi - HALO:self.hi]))
        return ops

    def halo_recvs(self):
        ops = []
        for f in self.fields():
            if self.rank > 0:
                ops.append((self.rank - 1, f[self.lo - HALO:self.lo]))
            if self.rank < self.world - 1:
                ops.append((self.rank + 1, f[self.hi:self.hi + HALO]))
        return ops

    def pack_forward(self):
        """[band rows][all k] -> send[dest][band rows][k of dest]."""
        rows, kper, P = self.plan.rows, self.plan.kper, self.world
        self.send.copy_(self.spectrum[self.lo:self.hi].view(rows, P, kper, 2).permute(1, 0, 2, 3))
        return self.send

    def unpack_backward(self, packed):
        """packed[src][band rows][k of src] -> spectrum[band rows][all k]."""
        rows, kper, P = self.plan.rows, self.plan.kper, self.world
        self.spectrum[self.lo:self.hi].view(rows, P, kper, 2).copy_(packed.permute(1, 0, 2, 3))

    def spectrum_row(self, j):
        return self.spectrum[j]


# ---- reduced-interface ("SPIKE") theta solve: no transposes, one small all-gather per step --------------
#
# Every wavenumber's theta system A x = d is tridiagonal over ALL rows, which is why the default mode
# transposes the half spectrum with two all-to-alls (2 x nTheta x nPhi/2 x 8 bytes per step: 2 x 537 MB at
# 8192 x 16384). Splitting the rows into the P bands,
#     A_p x_p + a_lo x_(lo-1) e_first + c_(hi-1) x_hi e_last = d_p ,
# so with g_p = A_p^-1 d_p (one band-local solve per step) and the right-hand-side independent spikes
# v_p = A_p^-1 (a_lo e_first), w_p = A_p^-1 (c_(hi-1) e_last) (computed once),
#     x_p = g_p - v_p x_(lo-1) - w_p x_hi .
# Taking the first and last row of that identity for every band gives a 2P x 2P system per wavenumber
# in the interface values (t_p, b_p) = (x_lo, x_(hi-1)); its matrix does not depend on the right-hand
# side, so its inverse is computed once (fp64) and a step needs: the local solve, an all-gather of two
# rows of the band's half spectrum per rank (2 x nPhi/2 x 8 bytes: 131 KB at 8192 x 16384), one small
# matrix-vector product per wavenumber and one fused correction of the band. x_hi = t_(p+1) also
# provides the extra spectrum row the theta gradient of the last band row needs. The result equals the
# global solve up to round-off (tests/test_banded_comm_cpu.py: 1e-11 against a direct fp64 solve), but
# it is a different operation order than the single-GPU kernel, so this mode is compared at tolerance
# level, not bit for bit; the band-local solves run the kernel of the full solve at nTheta / P rows.
class SpikeInterface:
    """The communication / algebra part, over torch tensors (GPU ranks and the CPU stand-in alike).

    Needs from the host object: rank, world, lo, hi, spectrum [nTheta][half][2], local_solve() (in place
    on spectrum[lo:hi]), coupling() -> (a_lo, c_hi_minus_1) (0.0 where the band touches a pole)."""

    # The two collectives are kept out of the phase methods so that the same code serves the NCCL run,
    # the gloo stand-in and the virtual ranks of LocalGroup (which gathers by plain copies).
    def spike_setup_local(self):
        """Spikes of my band from two unit right-hand sides; returns my four interface rows [4][half]."""
        torch = self.torch
        band = self.spectrum[self.lo:self.hi]
        m = band.shape[0]
        a_lo, c_hi = self.coupling()
        saved = band.clone()
        spikes = []
        for row, coeff in ((0, a_lo), (m - 1, c_hi)):
            band.zero_()
            if coeff != 0.0:
                band[row, :, 0] = coeff
                self.local_solve()
            spikes.append(band[:, :, 0].clone())
        band.copy_(saved)
        self.spike_v, self.spike_w = spikes                       # [m][half]
        return torch.stack([self.spike_v[0], self.spike_v[m - 1], self.spike_w[0], self.spike_w[m - 1]]).contiguous()

    def spike_setup_finish(self, everyone):
        """everyone[p] = rank p's interface rows (v_first, v_last, w_first, w_last)."""
        torch = self.torch
        band = self.spectrum[self.lo:self.hi]
        half, P = band.shape[1], self.world
        device = band.device
        M = torch.zeros((half, 2 * P, 2 * P), dtype=torch.float64, device=device)
        eye = torch.arange(2 * P, device=device)
        M[:, eye, eye] = 1.0
        for p in range(P):
            vt, vb, wt, wb = (everyone[p][k].to(torch.float64) for k in range(4))
            if p > 0:
                M[:, 2 * p, 2 * p - 1] = vt
                M[:, 2 * p + 1, 2 * p - 1] = vb
            if p < P - 1:
                M[:, 2 * p, 2 * p + 2] = wt
                M[:, 2 * p + 1, 2 * p + 2] = wb
        self.reduced_inverse = torch.linalg.inv(M)                 # [half][2P][2P], fp64
        self._interface = torch.empty((2, half, 2), dtype=band.dtype, device=device)

    def spike_local(self):
        """Band-local solve in place (band <- g); returns my interface rows [2][half][2] (first, last)."""
        band = self.spectrum[self.lo:self.hi]
        self.local_solve()
        self._interface[0].copy_(band[0])
        self._interface[1].copy_(band[band.shape[0] - 1])
        return self._interface

    def spike_finish(self, gathered):
        """gathered[p] = rank p's interface rows. Solves the reduced system and corrects the band;
        spectrum[hi] (if the band is not the last one) receives the solution's next row."""
        torch = self.torch
        band = self.spectrum[self.lo:self.hi]
        P, p = self.world, self.rank
        # right-hand side of the reduced system: [half][2P][re/im], unknown order t_0, b_0, t_1, b_1, ...
        rhs = torch.stack([g[k] for g in gathered for k in (0, 1)], dim=1).to(torch.float64)
        z = torch.matmul(self.reduced_inverse, rhs)                # [half][2P][2]
        zero = torch.zeros_like(z[:, 0])
        above = z[:, 2 * p - 1] if p > 0 else zero                 # x_(lo-1) = b_(p-1)
        below = z[:, 2 * p + 2] if p < P - 1 else zero             # x_hi     = t_(p+1)
        above, below = above.to(band.dtype), below.to(band.dtype)
        band.sub_(self.spike_v[:, :, None] * above[None] + self.spike_w[:, :, None] * below[None])
        if p < P - 1:
            self.spectrum[self.hi].copy_(below)

    # ---- with a torch.distributed-style process group ---------------------------------------------------
    def spike_setup(self, dist):
        mine = self.spike_setup_local()
        everyone = [self.torch.empty_like(mine) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(everyone, mine)
        else:
            everyone[0] = mine
        self.spike_setup_finish(everyone)
        self._gathered = [self.torch.empty_like(self._interface) for _ in range(self.world)]

    def spike_solve(self, dist):
        """spectrum[lo:hi] holds the right-hand sides on entry and the solution on exit."""
        mine = self.spike_local()
        if self.world > 1:
            dist.all_gather(self._gathered, mine)
        else:
            self._gathered[0].copy_(mine)
        self.spike_finish(self._gathered)


class BandRank(BandLayout, SpikeInterface):
    """The state and compute phases of one rank (no communication in here)."""

    def __init__(self, nTheta, radius, dt, rank, world, device=0):
        import torch
        self.torch = torch
        self.plan = BandPlan(nTheta, world)
        self.rank, self.world, self.device = rank, world, device
        self.nTheta, self.nPhi = nTheta, 2 * nTheta
        self.solver = KaminoSolver(self.nPhi, nTheta, radius, dt, device=device)
        self.lib, self.ctx = self.solver._lib, self.solver._ctx
        self.lo, self.hi = self.plan.band(rank)
        half = self.plan.half
        p = ctypes.c_void_p()
        capi.check(self.lib.kamino_spectrum_device_ptr(self.ctx, 0, ctypes.byref(p)), self.ctx)
        self.spectrum = self._wrap(p.value, (nTheta, half, 2))
        rows, kper = self.plan.rows, self.plan.kper
        dev = torch.device("cuda", device)
        self.send = torch.empty((world, rows, kper, 2), dtype=torch.float32, device=dev)
        self.recv = torch.empty((world, rows, kper, 2), dtype=torch.float32, device=dev)

    def close(self):
        self.solver.close()

    def _wrap(self, ptr, shape):
        return self.torch.as_tensor(_CudaArray(ptr, shape), device=self.torch.device("cuda", self.device))

    def set_stream(self, cuda_stream_ptr):
        self.solver.set_stream(cuda_stream_ptr)
        self._kernels_on_torch_stream = True        # kernels and torch ops are ordered on one stream from here on

    def fields(self):
        """Full-size torch views of the this-step u_phi, u_theta, density (pointers move with the swaps)."""
        out = []
        for field in (capi.VEL_PHI, capi.VEL_THETA, capi.DENSITY):
            p, pitch = ctypes.c_void_p(), ctypes.c_size_t()
            capi.check(self.lib.kamino_field_device_ptr(self.ctx, field, 0, 0, ctypes.byref(p), ctypes.byref(pitch)),
                       self.ctx)
            out.append(self._wrap(p.value, (self.nTheta, self.nPhi)))
        return out

    # ---- compute phases ----------------------------------------------------------------------------
    def advect_to_spectrum(self):
        a0, a1 = self.plan.clipped(self.rank, ADVECT_EXTRA)
        g0, g1 = self.plan.clipped(self.rank, GEO_EXTRA)
        capi.check(self.lib.kamino_band_advect(self.ctx, a0, a1 - a0), self.ctx)
        capi.check(self.lib.kamino_band_geometric(self.ctx, g0, g1 - g0), self.ctx)
        capi.check(self.lib.kamino_band_divergence_fft(self.ctx, self.lo, self.hi - self.lo), self.ctx)

    def solve(self, packed):
        """packed = [all rows][my k band] (the receive buffer of the forward all-to-all), in place."""
        kper = self.plan.kper
        capi.check(self.lib.kamino_band_tridiagonal(self.ctx, ctypes.c_void_p(packed.data_ptr()), kper,
                                                    self.rank * kper, kper), self.ctx)

    def inverse(self):
        capi.check(self.lib.kamino_band_inverse_fft_gradient(self.ctx, self.lo, self.hi - self.lo), self.ctx)

    # ---- reduced-interface (SPIKE) mode: see SpikeInterface ---------------------------------------------
    def prepare_local_solver(self):
        capi.check(self.lib.kamino_band_solver_prepare(self.ctx, self.lo, self.hi - self.lo), self.ctx)

    def local_solve(self):
        # the spike setup interleaves torch writes to the band with this kernel: unless both run on one
        # stream (DistributedBandedSolver), order them through the host (LocalGroup's virtual ranks)
        shared = getattr(self, "_kernels_on_torch_stream", False)
        if not shared:
            self.torch.cuda.synchronize()
        capi.check(self.lib.kamino_band_local_solve(self.ctx), self.ctx)
        if not shared:
            self.solver.sync()

    def coupling(self):
        """(a of my first row, c of my last row): the couplings to the neighbouring bands, 0 at the poles."""
        a, c = ctypes.c_float(0.0), ctypes.c_float(0.0)
        if self.lo > 0:
            capi.check(self.lib.kamino_tridiagonal_coefficients(self.ctx, self.lo, ctypes.byref(a), None), self.ctx)
        if self.hi < self.nTheta:
            capi.check(self.lib.kamino_tridiagonal_coefficients(self.ctx, self.hi - 1, None, ctypes.byref(c)), self.ctx)
        return float(a.value), float(c.value)

    # ---- host access (tests, start-up) ----------------------------------------------------------------
    def download_band(self):
        """numpy copies of my rows of u_phi, u_theta (clipped to nTheta-1 rows), density."""
        self.solver.sync()
        u, v, rho = [f.cpu().numpy() for f in self.fields()]
        return u[self.lo:self.hi].copy(), v[self.lo:min(self.hi, self.nTheta - 1)].copy(), rho[self.lo:self.hi].copy()


# ---- communication: free functions over torch tensors, so that they also run on CPU with gloo ----------

def exchange(sends, recvs, dist):
    """Post all receives and sends of one neighbour exchange ((peer, tensor) pairs) and wait."""
    ops = [dist.P2POp(dist.irecv, t, peer) for peer, t in recvs] + [dist.P2POp(dist.isend, t, peer) for peer, t in sends]
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def all_to_all(recv, send, dist):
    """recv[src] <- send[dest] of rank src; both [world][...] contiguous with equal blocks."""
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv.view(-1), send.view(-1))
        return
    # gloo (CPU tests): pairwise exchange
    rank, world = dist.get_rank(), dist.get_world_size()
    recv[rank].copy_(send[rank])
    ops = []
    for peer in range(world):
        if peer != rank:
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer))
            ops.append(dist.P2POp(dist.isend, send[peer], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class DistributedBandedSolver:
    """One rank of a band-decomposed simulation under torch.distributed."""

    def __init__(self, nTheta, radius, dt, device=0, solve="alltoall"):
        """solve = "alltoall" (default: transposes around the full theta solve, bit-identical to the
        single-GPU step) or "spike" (reduced-interface solve, no transposes; see SpikeInterface)."""
        import torch.distributed as dist
        if solve not in ("alltoall", "spike"):
            raise ValueError("solve must be 'alltoall' or 'spike'")
        self.solve = solve
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.r = BandRank(nTheta, radius, dt, self.rank, self.world, device)
        # kernels, torch copies and the NCCL calls are all ordered on ONE non-default torch stream
        # (kamino_set_stream treats the NULL / legacy default stream as "use the context's own")
        torch = self.r.torch
        self.stream = torch.cuda.Stream(device=torch.device("cuda", device))
        self.r.set_stream(self.stream.cuda_stream)
        self._graph, self._eager_done = None, 0      # None: not captured yet; False: capture unavailable
        if solve == "spike" and self.world > 1:
            with torch.cuda.stream(self.stream):
                self.r.prepare_local_solver()
                self.r.spike_setup(dist)

    def close(self):
        if self._graph:                       # release the captured NCCL work before the communicator goes away
            self.stream.synchronize()
            self._graph = False
            self.r.torch.cuda.synchronize()
        self.r.close()

    # Two steps (after which every buffer role is back where it started) are captured into one CUDA
    # graph -- kernels, the torch pack / unpack copies and the NCCL calls alike -- and replayed: the
    # eager loop spends ~350 us per step on the host (ten launches and four NCCL calls issued from
    # Python), which hides the GPU time of every grid below 2048 x 4096 (r01m). The first steps run
    # eagerly (NCCL creates its point-to-point channels lazily), a throw-away capture of one
    # all-to-all checks that this NCCL build can be captured at all, and any failure keeps the
    # eager loop. Measured on 2 B200s (r01o): 0.336 -> 0.248 ms/step at 512 x 1024 and bit-identical, but
    # 0.427 -> 0.731 ms/step at 2048 x 4096, and the processes did not exit cleanly while the graph
    # still held the NCCL work, so it is OPT-IN: KAMINO_BANDED_GRAPH=1.
    def step(self, nSteps=1):
        import os
        torch = self.r.torch
        with torch.cuda.stream(self.stream):
            if self.world > 1 and os.environ.get("KAMINO_BANDED_GRAPH", "0") == "1" and self._graph is not False:
                eager = min(nSteps, max(0, 2 - self._eager_done))
                self._step(eager)
                self._eager_done += eager
                nSteps -= eager
                if nSteps >= 2 and self._graph is None:
                    self._graph = self._capture_two_steps()
                while nSteps >= 2 and self._graph:
                    self._graph.replay()
                    nSteps -= 2
            self._step(nSteps)

    def _capture_two_steps(self):
        torch, dist = self.r.torch, self.dist
        self.stream.synchronize()
        try:
            probe = torch.cuda.CUDAGraph()
            a, b = torch.zeros_like(self.r.send), torch.zeros_like(self.r.recv)
            with torch.cuda.graph(probe, stream=self.stream, capture_error_mode="thread_local"):
                all_to_all(b, a, dist)
            probe.replay()
            self.stream.synchronize()
        except Exception:                      # this torch / NCCL pair cannot capture collectives
            return False
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=self.stream, capture_error_mode="thread_local"):
            self._step(2)
        return graph

    def sync(self):
        self.stream.synchronize()

    def check_halo(self):
        """Raise if the theta displacement of a backtrace can leave the exchanged halo (r01 ADVICE: it used to read stale
        rows silently): max |u_theta| dt / (R h) rows per stage, two stages, plus the bilinear stencil."""
        r = self.r
        v = r.fields()[1]
        rows = float(v[max(r.lo - 1, 0):r.hi].abs().max()) * r.solver.frameDuration / (r.solver.radius * float(r.solver.gridLen))
        if 2.0 * rows + 2.0 > HALO - ADVECT_EXTRA:
            raise RuntimeError("theta-CFL %.2f rows per stage: a backtrace can leave the %d-row halo of the band decomposition"
                               % (rows, HALO))

    def _step(self, nSteps):
        r, dist = self.r, self.dist
        for k in range(nSteps):
            if self.world > 1 and k % 8 == 0:
                self.check_halo()
            if self.world > 1:
                exchange(r.halo_sends(), r.halo_recvs(), dist)
            r.advect_to_spectrum()
            if self.world > 1 and self.solve == "spike":
                r.spike_solve(dist)
                r.inverse()
                continue
            if self.world > 1:
                all_to_all(r.recv, r.pack_forward(), dist)
                packed = r.recv
            else:
                packed = r.pack_forward()
            r.solve(packed.view(r.nTheta, r.plan.kper, 2))
            if self.world > 1:
                all_to_all(r.send, packed, dist)
                r.unpack_backward(r.send)
                # the theta gradient of my last row needs the pressure spectrum of row hi
                sends = [(self.rank - 1, r.spectrum_row(r.lo))] if self.rank > 0 else []
                recvs = [(self.rank + 1, r.spectrum_row(r.hi))] if self.rank < self.world - 1 else []
                exchange(sends, recvs, dist)
            else:
                r.unpack_backward(packed)
            r.inverse()


class LocalGroup:
    """P virtual ranks in one process on one GPU: the same phases, device copies instead of NCCL."""

    def __init__(self, nTheta, radius, dt, world, device=0, solve="alltoall"):
        self.ranks = [BandRank(nTheta, radius, dt, r, world, device) for r in range(world)]
        self.world = world
        self.solve = solve
        if solve == "spike" and world > 1:
            for r in self.ranks:
                r.prepare_local_solver()
            self.sync()
            rows = [r.spike_setup_local() for r in self.ranks]
            self.sync()
            for r in self.ranks:
                r.spike_setup_finish(rows)
            self.sync()

    def close(self):
        for r in self.ranks:
            r.close()

    def sync(self):
        for r in self.ranks:
            r.solver.sync()
        self.ranks[0].torch.cuda.synchronize()

    def step(self, nSteps=1):
        R = self.ranks
        for _ in range(nSteps):
            self.sync()
            sends = {r.rank: r.halo_sends() for r in R}
            for r in R:
                # my k-th receive from `peer` pairs with peer's k-th send to me (same field order)
                mine = r.halo_recvs()
                for peer in {p for p, _ in mine}:
                    src = [t for to, t in sends[peer] if to == r.rank]
                    dst = [t for frm, t in mine if frm == peer]
                    for s, d in zip(src, dst):
                        d.copy_(s)
            self.sync()
            for r in R:
                r.advect_to_spectrum()
            self.sync()
            if self.solve == "spike" and self.world > 1:
                rows = [r.spike_local().clone() for r in R]
                self.sync()
                for r in R:
                    r.spike_finish(rows)
                self.sync()
                for r in R:
                    r.inverse()
                continue
            packs = [r.pack_forward() for r in R]
            self.sync()
            for r in R:
                for q in R:
                    r.recv[q.rank].copy_(packs[q.rank][r.rank])
            self.sync()
            for r in R:
                r.solve(r.recv.view(r.nTheta, r.plan.kper, 2))
            self.sync()
            for r in R:
                for q in R:
                    r.send[q.rank].copy_(q.recv[r.rank])
            self.sync()
            for r in R:
                r.unpack_backward(r.send)
            self.sync()
            for r in R:
                if r.rank < self.world - 1:
                    r.spectrum_row(r.hi).copy_(R[r.rank + 1].spectrum_row(r.hi))
            self.sync()
            for r in R:
                r.inverse()
        self.sync()

    def gather(self):
        """The global u_phi, u_theta, density assembled from the bands (numpy)."""
        parts = [r.download_band() for r in self.ranks]
        return tuple(np.concatenate([p[k] for p in parts], axis=0) for k in range(3))
