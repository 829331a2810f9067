"""CPU evidence for the projection-parity argument: the reference's theta solve -- cyclic reduction
in fp32, restated in oracle/kamino_oracle.c from kernel/tdm.cu:3-96 -- is itself an order of magnitude further from an
fp64 solve of the same fp32 coefficients than a plain LU (Thomas) solve in fp32 is once the grid has 512 rows (the two
are level at 128 rows), and the gap grows with nTheta (1.3e-4 at 2048 rows on the GPU, profiles/r02_parity_table.md). That is
why the product (an LU solve) cannot sit within 1e-5 of the reference's pressure at 512 rows and above, and why it does once
it solves in the reference's order (kamino_debug_project_cr, GPU suite)."""
import ctypes

import numpy as np
import pytest
from scipy.linalg import solve_banded

import oracle_api as oa


def _systems(nT):
    """fp32 coefficients of every wavenumber (kernel/KaminoSolver.cu:117-163) and a smooth right-hand side: the half
    spectrum of the divergence of the reference's initial field."""
    N = 2 * nT
    p = oa.params(nT)
    u, v = oa.init_velocity(nT)
    div = np.zeros(nT * N, np.float32)
    oa.lib().ko_divergence(ctypes.byref(p), oa.fptr(u), oa.fptr(v), oa.fptr(div))
    F = (np.fft.rfft(div.reshape(nT, N).astype(np.float64), axis=1) / N)[:, 1:]          # n = 1 .. N/2
    a = np.zeros((N // 2, nT), np.float32); b = np.zeros_like(a); c = np.zeros_like(a)
    for n in range(1, N // 2 + 1):
        oa.lib().ko_abc_row(ctypes.byref(p), n, oa.fptr(a[n - 1]), oa.fptr(b[n - 1]), oa.fptr(c[n - 1]))
    return a, b, c, np.ascontiguousarray(F.real.T.astype(np.float32))                    # [wavenumber][row]


def _fp64(a, b, c, d):
    out = np.empty(d.shape, np.float64)
    nT = d.shape[1]
    for k in range(d.shape[0]):
        ab = np.zeros((3, nT))
        ab[0, 1:] = c[k, :-1]; ab[1] = b[k]; ab[2, :-1] = a[k, 1:]
        out[k] = solve_banded((1, 1), ab, d[k].astype(np.float64))
    return out


def _thomas_fp32(a, b, c, d):
    """Plain LU in fp32, every wavenumber at once (vector over the first axis)."""
    nT = d.shape[1]
    bp = b.copy(); y = d.copy()
    for i in range(1, nT):
        w = (a[:, i] / bp[:, i - 1]).astype(np.float32)
        bp[:, i] = (b[:, i] - w * c[:, i - 1]).astype(np.float32)
        y[:, i] = (y[:, i] - w * y[:, i - 1]).astype(np.float32)
    x = np.empty_like(y)
    x[:, -1] = y[:, -1] / bp[:, -1]
    for i in range(nT - 2, -1, -1):
        x[:, i] = ((y[:, i] - c[:, i] * x[:, i + 1]) / bp[:, i]).astype(np.float32)
    return x


def _cyclic_reduction(a, b, c, d):
    x = np.zeros_like(d)
    nT = d.shape[1]
    for k in range(d.shape[0]):
        aa, bb, cc, dd = a[k].copy(), b[k].copy(), c[k].copy(), d[k].copy()
        oa.lib().ko_cyclic_reduction(nT, oa.fptr(aa), oa.fptr(bb), oa.fptr(cc), oa.fptr(dd), oa.fptr(x[k]))
    return x


@pytest.mark.parametrize("nT", [128, 512])
def test_cyclic_reduction_is_the_less_accurate_solve(nT):
    a, b, c, d = _systems(nT)
    exact = _fp64(a, b, c, d)
    e_cr = oa.rel_l2(_cyclic_reduction(a, b, c, d), exact)
    e_lu = oa.rel_l2(_thomas_fp32(a, b, c, d), exact)
    print("nTheta %d: theta solve vs fp64 -- cyclic reduction (reference order) %.2e, LU %.2e" % (nT, e_cr, e_lu))
    assert e_lu <= 5e-6                              # an fp32 LU stays inside the north-star tolerance at both sizes
    if nT == 128:
        assert e_cr <= 1e-5                          # measured 3.4e-6: at the shipped size the reference's solve is fine too
    else:
        assert e_cr > 1e-5 and e_cr >= 3.0 * e_lu    # measured 2.2e-5 vs 2.7e-6 at 512 rows (the GPU builds differ by 1.7e-5)
