"""Fused CG building blocks on the CPU model of tests/cuda_emul: spmv_dot_kernel (apply! + dot in
one pass) and axpby_norm2_kernel (broadcast update + norm in one pass) against the CPU oracle's separate operations."""
import numpy as np
import pytest

import emul
from helpers import O, oracle_problem, synthetic


def _problem(kind, N, **kw):
    c, m, A = oracle_problem(kind, N, 1, **kw)
    rp, ci, va = A.csr(1)
    return rp - 1, ci - 1, va


@pytest.mark.parametrize("kind,N,cap", [("7pt", 9, 2048), ("27pt", 5, 2048), ("powerlaw", 3000, 2048), ("powerlaw", 3000, 256)])
@pytest.mark.parametrize("tma", [1, 0])
def test_spmv_dot_matches_apply_then_dot(kind, N, cap, tma):
    rp, ci, va = _problem(kind, N, **({"max_len": 1200} if kind == "powerlaw" else {}))
    n = len(rp) - 1
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    w = synthetic.vector_block(1, n, 1, seed=3)
    for alpha, beta, wv in ((1.0, 0.0, x), (3.0, 0.5, w)):  # W = X is the p'Ap of CG
        y_ref = O.local_apply_raw(y0.copy(order="F"), rp + 1, ci + 1, va, x, O.NO_TRANS, alpha, beta)
        y_plain, _ = emul.local_apply(rp, ci, va, x, y0, alpha, beta, cap=cap, tma=tma)
        y_got, dots = emul.spmv_dot(rp, ci, va, x, wv, y0, alpha, beta, cap=cap, tma=tma, repeats=3)
        assert np.array_equal(y_got, y_plain.ravel())  # the fused kernel computes y exactly like the plain one
        d_ref = O.dot_raw(wv, y_ref)[0]
        scale = O.dot_raw(np.abs(wv), np.abs(y_ref))[0]
        assert dots[0] == dots[1] == dots[2]  # deterministic, and the ticket re-arms
        assert abs(dots[0] - d_ref) <= 1e-12 * scale


def test_spmv_dot_empty_rows_and_tiny_matrix():
    rp = np.array([0, 0, 2, 2, 3], dtype=np.int32)
    ci = np.array([0, 3, 1], dtype=np.int32)
    va = np.array([2.0, -1.0, 4.0])
    x = np.array([1.0, 2.0, 3.0, 4.0])
    w = np.array([10.0, 20.0, 30.0, 40.0])
    y, dots = emul.spmv_dot(rp, ci, va, x, w, np.ones(4), 1.0, 0.0)
    assert y.tolist() == [0.0, -2.0, 0.0, 8.0] and dots[0] == 20.0 * -2.0 + 40.0 * 8.0


@pytest.mark.parametrize("n,k,blocks", [(1, 1, 1), (1000, 1, 1), (5000, 3, 4), (257, 2, 7)])
def test_axpby_norm2(n, k, blocks):
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    a, b = -0.37, 1.0
    y, sums = emul.axpby_norm2(y0, a, x, b, blocks=blocks, repeats=2)
    y1 = a * x + b * y0        # numpy evaluates (a*x) + (b*y) unfused, like the reference's broadcast
    y2 = a * x + b * y1
    assert np.array_equal(y, y2)  # the update itself is bit-exact
    for rep, yr in enumerate((y1, y2)):
        ref = O.norm_raw(yr, 2.0)
        assert (np.abs(sums[rep] - ref) <= 1e-12 * ref).all()


@pytest.mark.parametrize("n,blocks", [(1, 1), (1000, 2), (4099, 5)])
def test_update_kernels_with_device_resident_coefficients(n, blocks):
    """jp_solver_kernels.cuh: the coefficient scale * s[num] / s[den] is formed on the device; the update is the
    unfused (a*x) + (b*y) of the host-scalar kernel, bit for bit"""
    x = synthetic.vector_block(1, n, 1, seed=1).ravel()
    y0 = synthetic.vector_block(1, n, 1, seed=2).ravel()
    s = np.array([3.7, -0.9, 123.0, 0.25])
    # CG's  x += (rr / pAp) p :  a = s[0] / s[1], b = 1
    y, s1 = emul.axpby_dev(y0, x, s, (1.0, 0, 1), (1.0, -1, -1), blocks=blocks)
    assert np.array_equal(y, (s[0] / s[1]) * x + 1.0 * y0) and np.array_equal(s1, s)
    # CG's  r -= (rr / pAp) Ap  with ||r||^2 into slot 2
    y, s2 = emul.axpby_dev(y0, x, s, (-1.0, 0, 1), (1.0, -1, -1), with_norm=True, out_slot=2, blocks=blocks, repeats=1)
    r = (-1.0 * s[0] / s[1]) * x + 1.0 * y0
    assert np.array_equal(y, r)
    ref = O.norm_raw(r.reshape(-1, 1), 2.0)[0]
    assert abs(s2[2] - ref) <= 1e-12 * ref and np.array_equal(s2[[0, 1, 3]], s[[0, 1, 3]])
    # CG's  p = r + (rr_new / rr) p :  a = 1, b = s[2] / s[0]; twice on the same workspace
    y, s3 = emul.axpby_dev(y0, x, s, (1.0, -1, -1), (1.0, 2, 0), blocks=blocks, repeats=2)
    b = s[2] / s[0]
    assert np.array_equal(y, 1.0 * x + b * (1.0 * x + b * y0))
