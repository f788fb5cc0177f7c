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
