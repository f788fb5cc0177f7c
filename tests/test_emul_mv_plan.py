"""The MultiVector kernels (csrc/jp_mv_kernels.cuh), the transfer kernels (csrc/jp_plan_kernels.cuh) and the transposed
localApply kernel executed on the CPU model of tests/cuda_emul, against the CPU oracle / the reference's definitions:
dot, norm (src/MultiVector.jl:84-146), fill!, scale!, broadcast (:52-77, 497-502), copyAndPermute / packAndPrepare /
unpackAndCombine (:308-371) and localApply TRANS (src/CSRMatrix.jl:1147-1160)."""
import numpy as np
import pytest

import emul
from helpers import O, oracle_problem, synthetic


@pytest.mark.parametrize("n,k,blocks", [(1, 1, 1), (1000, 3, 2), (5000, 1, 4), (4099, 2, 3)])
def test_dot_and_norm_sums(n, k, blocks):
    x = synthetic.vector_block(1, n, k, seed=1)
    y = synthetic.vector_block(1, n, k, seed=2)
    d = emul.mv_reduce(0, x, y, blocks=blocks)
    assert np.array_equal(d[0], d[1])  # deterministic, the ticket re-arms
    assert (np.abs(d[0] - O.dot_raw(x, y)) <= 1e-12 * O.dot_raw(np.abs(x), np.abs(y))).all()
    s2 = emul.mv_reduce(1, x, blocks=blocks)
    assert np.array_equal(s2[0], s2[1]) and (np.abs(s2[0] - O.norm_raw(x, 2.0)) <= 1e-12 * O.norm_raw(x, 2.0)).all()
    s3 = emul.mv_reduce(2, np.abs(x), p=3.0, blocks=blocks)
    assert (np.abs(s3[0] - O.norm_raw(np.abs(x), 3.0)) <= 1e-12 * O.norm_raw(np.abs(x), 3.0)).all()


def test_reference_golden_dot_norm_values():
    """test/DenseMultiVectorTests.jl:76-77, 103-111: dot(ones, ones) == n, norm(ones, 2) == sqrt(n), norm(twos, 3)"""
    n = 777
    ones = np.ones((n, 3))
    assert emul.mv_reduce(0, ones, ones)[0].tolist() == [float(n)] * 3
    assert np.sqrt(emul.mv_reduce(1, ones)[0]).tolist() == [np.sqrt(n)] * 3
    assert np.allclose(emul.mv_reduce(2, 2 * ones, p=3.0)[0] ** (1 / 3), (8.0 * n) ** (1 / 3), rtol=1e-15)


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 3), (2049, 2)])
def test_elementwise_kernels_are_bit_exact(n, k):
    x = synthetic.vector_block(1, n, k, seed=1)
    y = synthetic.vector_block(1, n, k, seed=2)
    assert np.array_equal(emul.mv_elementwise(0, y, a=-3.25), np.full((n, k), -3.25))
    assert np.array_equal(emul.mv_elementwise(1, y, a=1.7), y * 1.7)
    ak = np.arange(1, k + 1) * 0.3
    assert np.array_equal(emul.mv_elementwise(2, y, alpha_k=ak), y * ak)
    assert np.array_equal(emul.mv_elementwise(3, y, x=x, a=-0.37, b=1.25), -0.37 * x + 1.25 * y)  # (a*x) + (b*y), unfused


def test_transfer_kernels_follow_the_reference_definitions():
    rng = np.random.default_rng(4)
    n_src, n_tgt, k, m = 50, 40, 3, 17
    src = rng.standard_normal((n_src, k))
    tgt = rng.standard_normal((n_tgt, k))
    lids = rng.integers(0, n_src, size=m)
    # packAndPrepare (src/MultiVector.jl:339-347): LID-major, vector-minor
    _, buf = emul.plan_op(0, src, np.zeros((m, k)), lids)
    assert np.array_equal(buf, src[lids, :])
    # unpackAndCombine (:363-369), distinct target LIDs
    tl = rng.permutation(n_tgt)[:m]
    got, _ = emul.plan_op(1, tgt, buf, tl)
    exp = tgt.copy()
    exp[tl, :] = buf
    assert np.array_equal(got, exp)
    # scatter-add with repeated LIDs (a row exported to several ranks): every contribution arrives
    tl2 = rng.integers(0, 5, size=m)
    got, _ = emul.plan_op(2, tgt, buf, tl2)
    exp = tgt.copy()
    np.add.at(exp, tl2, buf)
    assert np.allclose(got, exp, rtol=1e-14, atol=1e-14)
    # copyAndPermute's permute half (:320-323) and its ADD twin
    to, frm = rng.permutation(n_tgt)[:m], rng.integers(0, n_src, size=m)
    got, _ = emul.plan_op(3, tgt, src, to, frm)
    exp = tgt.copy()
    exp[to, :] = src[frm, :]
    assert np.array_equal(got, exp)
    got, _ = emul.plan_op(4, tgt, src, to, frm)
    exp = tgt.copy()
    exp[to, :] += src[frm, :]
    assert np.array_equal(got, exp)
    got, _ = emul.plan_op(5, tgt, src, n_ids=23)
    exp = tgt.copy()
    exp[:23, :] += src[:23, :]
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("kind,N,k", [("7pt", 6, 1), ("powerlaw", 900, 3)])
def test_transposed_local_apply(kind, N, k):
    c, m, A = oracle_problem(kind, N, 1, max_len=300)
    rp, ci, va = A.csr(1)
    n = len(rp) - 1
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    alpha, beta = 3.0, 0.5
    ref = O.local_apply_raw(y0.copy(order="F"), rp, ci, va, x, O.TRANS, alpha, beta)
    # the transposed apply is the row-block kernel on A^T (csrc/jp_transpose.cu builds it by a stable sort by column)
    rpt, cit, vat = emul.transpose_csr(rp - 1, ci - 1, va, n)
    got, _ = emul.local_apply(rpt, cit, vat, x, y0, alpha, beta)
    bound = 1e-12 * (alpha * O.local_apply_raw(np.zeros_like(y0), rp, ci, np.abs(va), np.abs(x), O.TRANS, 1.0, 0.0) + beta * np.abs(y0))
    assert (np.abs(got - ref) <= bound + 1e-300).all()


def test_reference_golden_transposed_apply():
    """test/CSRMatrixTests.jl:152-158: A = [[2,3],[5,7]], apply!(Y, A, X, TRANS, 3, 0.5), Y = diag(1,2), X = 2s"""
    rp, ci, va = np.array([0, 2, 4]), np.array([0, 1, 0, 1]), np.array([2.0, 3.0, 5.0, 7.0])
    rpt, cit, vat = emul.transpose_csr(rp, ci, va, 2)
    y, _ = emul.local_apply(rpt, cit, vat, np.full((2, 2), 2.0), np.array([[1.0, 0.0], [0.0, 2.0]]), 3.0, 0.5)
    assert np.array_equal(y, np.array([[42.5, 42.0], [60.0, 61.0]]))
