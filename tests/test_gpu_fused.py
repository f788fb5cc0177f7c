"""Fused CG building blocks through the C ABI on the GPU: jp_apply_dot and jp_mv_axpby_norm2
against the separate apply! / dot / broadcast / norm calls and the CPU oracle, and the CG driver built from them."""
import numpy as np
import pytest

from helpers import O, jp, synthetic

pytestmark = pytest.mark.gpu


def _serial_problem(kind, N):
    comm = jp.SerialComm()
    n = N if kind == "powerlaw" else N ** synthetic.STENCILS[kind][0]
    m = jp.BlockMap(n, comm)
    return m, n, synthetic.build_matrix(jp, m, kind, N)


@pytest.mark.parametrize("kind,N", [("7pt", 40), ("27pt", 20), ("powerlaw", 200000)])
def test_apply_dot_matches_apply_then_dot(kind, N):
    m, n, A = _serial_problem(kind, N)
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    w = synthetic.vector_block(1, n, 1, seed=3)
    X, W = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, w)
    for alpha, beta, Wv, wv in ((1.0, 0.0, None, x), (3.0, 0.5, W, w)):
        Y1, Y2 = jp.DenseMultiVector(m, y0), jp.DenseMultiVector(m, y0)
        A.apply_(Y1, X, jp.NO_TRANS, alpha, beta)
        d1 = (X if Wv is None else Wv).dot(Y1)[0, 0]
        for _ in range(3):  # the kernel re-arms its ticket: repeated calls give the same answer
            Y2.data = y0
            d2 = A.apply_dot_(Y2, X, Wv, jp.NO_TRANS, alpha, beta)[0, 0]
            assert np.array_equal(Y1.data, Y2.data)
            y_ref = O.local_apply_raw(y0.copy(order="F"), A.rowOffsets, A.localIndices1D, A.values, x, O.NO_TRANS, alpha, beta)
            scale = O.dot_raw(np.abs(wv), np.abs(y_ref))[0]
            assert abs(d2 - O.dot_raw(wv, y_ref)[0]) <= 1e-12 * scale
            assert abs(d2 - d1) <= 1e-12 * scale


def test_apply_dot_multivector_and_transposed_fall_back_to_two_kernels():
    m, n, A = _serial_problem("7pt", 12)
    x = synthetic.vector_block(1, n, 3, seed=1)
    X = jp.DenseMultiVector(m, x)
    for mode in (jp.NO_TRANS, jp.TRANS):
        Y1, Y2 = jp.DenseMultiVector(m, 3), jp.DenseMultiVector(m, 3)
        A.apply_(Y1, X, mode, 2.0, 0.0)
        d1 = X.dot(Y1)
        d2 = A.apply_dot_(Y2, X, None, mode, 2.0, 0.0)
        if mode == jp.NO_TRANS:
            assert np.array_equal(Y1.data, Y2.data) and np.array_equal(d1, d2)
        else:  # the transposed kernel scatters with atomics: the summation order differs from run to run
            ref = np.abs(Y1.data).max()
            assert np.abs(Y1.data - Y2.data).max() <= 1e-12 * ref and (np.abs(d1 - d2) <= 1e-12 * np.abs(x).T @ np.abs(Y1.data) @ np.ones(3)).all()


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (300001, 2)])
def test_axpby_norm2_matches_axpby_then_norm(n, k):
    m = jp.BlockMap(n, jp.SerialComm())
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    X, Y1, Y2 = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0), jp.DenseMultiVector(m, y0)
    for _ in range(2):
        Y1.axpby_(-0.37, X, 1.0)
        n1 = Y1.norm(2)
        n2 = Y2.axpby_norm2_(-0.37, X, 1.0)
        assert np.array_equal(Y1.data, Y2.data)  # the update is bit-exact
        ref = np.sqrt(O.norm_raw(Y1.data, 2.0))
        assert (np.abs(n2[0] - ref) <= 1e-12 * ref).all() and (np.abs(n2 - n1) <= 1e-12 * ref).all()


def test_cg_fused_and_unfused_solve_the_laplacian():
    m, n, A = _serial_problem("7pt", 16)
    xs = synthetic.vector_block(1, n, 1, seed=7)
    B = A.apply(jp.DenseMultiVector(m, xs))
    results = []
    for fused in (True, False):
        X = jp.DenseMultiVector(m, 1)
        its, rel = jp.cg(A, B, X, tol=1e-12, maxiter=500, fused=fused)
        assert rel <= 1e-12 and its < 200
        err = np.abs(X.data - xs).max()
        assert err <= 1e-9, err
        results.append((its, X.data))
    assert abs(results[0][0] - results[1][0]) <= 1


def test_cg_with_device_resident_scalars():
    """jp_*_dev: CG without host round trips inside the iteration against the host-scalar CG and the known solution"""
    m, n, A = _serial_problem("7pt", 16)
    xs = synthetic.vector_block(1, n, 1, seed=7)
    B = A.apply(jp.DenseMultiVector(m, xs))
    X1, X2 = jp.DenseMultiVector(m, 1), jp.DenseMultiVector(m, 1)
    its1, rel1 = jp.cg(A, B, X1, tol=1e-12, maxiter=500, fused=True)
    its2, rel2 = jp.cg_device(A, B, X2, tol=1e-12, maxiter=500, check_every=5)
    assert rel2 <= 1e-12 and its1 <= its2 <= its1 + 6
    assert np.abs(X2.data - xs).max() <= 1e-9 and np.abs(X2.data - X1.data).max() <= 1e-9
    # the building blocks one by one
    s = jp.Scalars(4)
    s.upload([3.0, -2.0, 0.0, 0.5])
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    Xv, Yv = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0)
    lib = jp._lib.lib()
    jp._lib.check(lib.jp_mv_axpby_norm2sq_dev(jp.SerialComm().handle, Yv.h, Xv.h, s.h, -1.0, 0, 1, 1.0, -1, -1, 2))
    r = (-1.0 * 3.0 / -2.0) * x + 1.0 * y0
    assert np.array_equal(Yv.data, r)
    v = s.download()
    ref = O.norm_raw(r, 2.0)[0]
    assert abs(v[2] - ref) <= 1e-12 * ref and v[0] == 3.0 and v[3] == 0.5
    jp._lib.check(lib.jp_mv_dot_dev(jp.SerialComm().handle, Xv.h, Yv.h, s.h, 3))
    d = s.download()[3]
    assert abs(d - O.dot_raw(x, r)[0]) <= 1e-12 * O.dot_raw(np.abs(x), np.abs(r))[0]
