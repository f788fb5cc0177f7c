"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the host mirror and the
C ABI, against the CPU oracle on the same seeded inputs and against the reference's golden vectors.

Tolerances: integer/index results bit-exact.  Float64 SpMV/SpMM: |y_gpu - y_ref| <= 1e-12 * (|alpha| sum|a_ij||x_j|
+ |beta||y0_i|) component-wise (the north star's 1e-12 relative, measured against the magnitude of the terms
because stencil rows cancel); where the kernel keeps the reference's summation order the result is required to
be bit-identical.  dot/norm: 1e-12 relative.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import O, global_rows, jp, oracle_problem
from jpetra_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu

TOL = 1e-12


def serial_matrix(kind, N, **kw):
    c = jp.SerialComm()
    m = jp.BlockMap(global_rows(kind, N), c)
    return c, m, synthetic.build_matrix(jp, m, kind, N, **kw)


def abs_bound(A, x, y0, alpha, beta):
    """|alpha| * |A| |x| + |beta| |y0| per component, through the oracle's localApply on absolute values"""
    b = O.local_apply_raw(np.abs(y0).copy(), A.rowOffsets, A.localIndices1D, np.abs(A.values), np.abs(x), O.NO_TRANS, abs(alpha), abs(beta))
    return b


def check_apply(A, x, y0, alpha, beta, exact):
    m = A.getRowMap()
    X = jp.DenseMultiVector(m, x)
    Y = jp.DenseMultiVector(m, y0)
    assert A.apply_(Y, X, jp.NO_TRANS, alpha, beta) is Y
    got = Y.data
    ref = O.local_apply_raw(np.array(y0, order="F", copy=True), A.rowOffsets, A.localIndices1D, A.values, x, O.NO_TRANS, alpha, beta)
    assert np.array_equal(X.data, x)  # X is not mutated
    bound = abs_bound(A, x, y0, alpha, beta)
    err = np.abs(got - ref)
    assert (err <= TOL * bound + 1e-300).all(), float((err / (bound + 1e-300)).max())
    if exact:
        assert np.array_equal(got, ref)
    return got


# ---- reference golden vectors through the GPU path -----------------------------------------------------------
def test_golden_csrmatrix_serial():
    """test/CSRMatrixTests.jl:120-174"""
    c = jp.SerialComm()
    m = jp.BlockMap(2, 2, c)
    A = jp.CSRMatrix(m, 2, jp.STATIC_PROFILE)
    A.insertGlobalValues(1, [1, 2], [2, 3])
    A.insertGlobalValues(2, [1, 2], [5, 7])
    A.fillComplete()
    assert A.localIndices1D.tolist() == [1, 2, 1, 2]
    Y = jp.DenseMultiVector(m, np.array([[1.0, 0], [0, 2]]))
    X = jp.DenseMultiVector(m, np.full((2, 2), 2.0))
    assert A.apply_(Y, X, jp.TRANS, 3.0, 0.5) is Y
    assert np.array_equal(X.data, np.full((2, 2), 2.0))
    assert np.array_equal(Y.data, np.array([[42.5, 42], [60, 61]]))
    Y = A * X
    assert isinstance(Y, jp.DenseMultiVector) and Y.getMap() is X.getMap()
    assert np.array_equal(Y.data, np.array([[10.0, 10], [24, 24]]))
    Y = jp.DenseMultiVector(m, np.array([[1.0, 0], [0, 2]]))
    assert A.mul_(Y, X) is Y
    assert np.array_equal(Y.data, np.array([[10.0, 10], [24, 24]]))
    Y = jp.DenseMultiVector(m, np.array([[1.0, 0], [0, 2]]))
    A.apply_(Y, X, jp.NO_TRANS, 3.0, 0.5)
    assert np.array_equal(Y.data, np.array([[30.5, 30], [72, 73]]))  # the commented-out expectation at :143-149
    # alpha == 0 shortcut (src/RowMatrix.jl:271-278)
    Y = jp.DenseMultiVector(m, np.array([[1.0, 0], [0, 2]]))
    A.apply_(Y, X, jp.NO_TRANS, 0.0, 0.5)
    assert np.array_equal(Y.data, np.array([[0.5, 0], [0, 1.0]]))
    A.apply_(Y, X, jp.NO_TRANS, 0.0, 0.0)
    assert np.array_equal(Y.data, np.zeros((2, 2)))
    with pytest.raises(jp.InvalidArgumentError):
        A.apply_(X, X)


def test_golden_dense_multivector():
    """test/DenseMultiVectorTests.jl:56-144 on SerialComm"""
    n = 8
    c = jp.SerialComm()
    m = jp.BlockMap(n, n, c)
    v = jp.DenseMultiVector(m, np.ones((n, 3)))
    assert v.scale_(5.0) is v and np.array_equal(v.data, 5 * np.ones((n, 3)))
    v = jp.DenseMultiVector(m, np.ones((n, 3)))
    v.scale_([2.0, 3.0, 4.0])
    assert np.array_equal(v.data, np.ones((n, 3)) * [2.0, 3.0, 4.0])
    assert v.getVectorCopy(2).tolist() == [3.0] * n
    v = jp.DenseMultiVector(m, np.ones((n, 3)))
    assert np.array_equal(v.dot(v), np.full((1, 3), float(n)))
    v.fill_(8)
    assert np.array_equal(v.data, 8 * np.ones((n, 3)))
    v2 = v.copy()
    assert v2 == v and v2.h != v.h
    z = jp.DenseMultiVector(m, 3)
    assert np.array_equal(z.data, np.zeros((n, 3)))
    assert z.copyto_(v) is z and np.array_equal(z.data, v.data)
    v = jp.DenseMultiVector(m, 10.0 * np.ones((n, 3)))
    v.commReduce()
    assert np.array_equal(v.data, 10.0 * np.ones((n, 3)))
    v = jp.DenseMultiVector(m, np.ones((n, 3)))
    assert np.array_equal(v.norm(2), np.full((1, 3), np.sqrt(n)))
    assert np.allclose(v.norm(3), np.full((1, 3), n ** (1 / 3)), rtol=1e-15, atol=0)
    v = jp.DenseMultiVector(m, 2 * np.ones((n, 3)))
    assert np.array_equal(v.norm(2), np.full((1, 3), np.sqrt(4 * n)))
    assert np.allclose(v.norm(3), np.full((1, 3), (8 * n) ** (1 / 3)), rtol=1e-15, atol=0)
    data = np.arange(1, 3 * n + 1, dtype=np.float64).reshape((n, 3), order="F")
    for plan, f in ((jp.Import(m, m), jp.doImport), (jp.Export(m, m), jp.doExport), (jp.Import(m, m), jp.doExport), (jp.Export(m, m), jp.doImport)):
        src = jp.DenseMultiVector(m, data)
        tgt = jp.DenseMultiVector(m, 3)
        f(src, tgt, plan, jp.REPLACE)
        assert np.array_equal(tgt.data, data)
    with pytest.raises(jp.InvalidArgumentError):  # src/MultiVector.jl:88-93
        v.dot(jp.DenseMultiVector(m, 2))
    with pytest.raises(jp.InvalidArgumentError):
        jp.DenseMultiVector(m, np.ones((n + 1, 2)))


# ---- SpMV / SpMM against the oracle ---------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,N", [("5pt", 1000), ("7pt", 64), ("27pt", 40), ("7pt", 3), ("5pt", 1)])
def test_spmv_matches_oracle(kind, N):
    """BASELINE config 1 (2-D 5-point 1000x1000, SerialComm, single vector) and smaller stencils: thread-per-row
    blocks keep the reference's summation order, so the result must be bit-identical."""
    c, m, A = serial_matrix(kind, N)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    check_apply(A, x, y0, 3.0, 0.5, exact=True)
    check_apply(A, x, y0, 1.0, 0.0, exact=True)
    # beta == 0 must overwrite NaN/Inf in Y (src/Operator.jl:42)
    ynan = np.full((n, 1), np.nan)
    got = check_apply(A, x, np.ones((n, 1)), 1.0, 0.0, exact=True)
    Y = jp.DenseMultiVector(m, ynan)
    A.apply_(Y, jp.DenseMultiVector(m, x), jp.NO_TRANS, 1.0, 0.0)
    assert np.array_equal(Y.data, got)


@pytest.mark.parametrize("kind,N,k", [("27pt", 24, 8), ("27pt", 24, 32), ("7pt", 20, 2), ("5pt", 50, 3), ("27pt", 12, 11)])
def test_spmm_matches_oracle(kind, N, k):
    c, m, A = serial_matrix(kind, N)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    check_apply(A, x, y0, 3.0, 0.5, exact=False)
    check_apply(A, x, y0, 1.0, 0.0, exact=False)


@pytest.mark.parametrize("k", [1, 4])
def test_powerlaw_matches_oracle(k):
    """row-length-adaptive paths: warp-per-row blocks and rows longer than the shared-memory stage"""
    n = 60000
    c, m, A = serial_matrix("powerlaw", n, max_len=9000)
    info = A.device_info()
    assert info["max_row_len"] > 2048  # at least one row exceeds the stage and takes the streamed path
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    check_apply(A, x, y0, 3.0, 0.5, exact=False)


def test_empty_rows_and_empty_matrix():
    c = jp.SerialComm()
    m = jp.BlockMap(6, c)
    A = jp.CSRMatrix(m, [2, 0, 1, 0, 0, 3], jp.STATIC_PROFILE)
    A.insertGlobalValues(1, [1, 6], [1.0, 2.0])
    A.insertGlobalValues(3, [3], [4.0])
    A.insertGlobalValues(6, [2, 4, 5], [1.0, 1.0, 1.0])
    A.fillComplete(m, m)
    x = np.arange(1.0, 7.0).reshape(6, 1)
    y0 = np.ones((6, 1))
    check_apply(A, x, y0, 2.0, 3.0, exact=True)
    m0 = jp.BlockMap(0, c)
    A0 = jp.CSRMatrix(m0, 0, jp.STATIC_PROFILE)
    A0.fillComplete(m0, m0)
    Y = jp.DenseMultiVector(m0, 1)
    A0.apply_(Y, jp.DenseMultiVector(m0, 1))
    assert Y.data.shape == (0, 1)


def test_trans_matches_oracle():
    c, m, A = serial_matrix("powerlaw", 5000, max_len=300)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, 3, seed=1)
    y0 = synthetic.vector_block(1, n, 3, seed=2)
    X, Y = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0)
    A.apply_(Y, X, jp.TRANS, 3.0, 0.5)
    ref = O.local_apply_raw(np.array(y0, order="F", copy=True), A.rowOffsets, A.localIndices1D, A.values, x, O.TRANS, 3.0, 0.5)
    bound = O.local_apply_raw(np.abs(y0), A.rowOffsets, A.localIndices1D, np.abs(A.values), np.abs(x), O.TRANS, 3.0, 0.5)
    assert (np.abs(Y.data - ref) <= TOL * bound).all()


# ---- device copy of the CSR is exactly what fillComplete produced -------------------------------------------------------
def test_device_csr_is_bit_exact():
    c, m, A = serial_matrix("powerlaw", 20000, max_len=500)
    rp = np.empty(len(A.rowOffsets), dtype=np.int32)
    ci = np.empty(len(A.localIndices1D), dtype=np.int32)
    va = np.empty(len(A.values), dtype=np.float64)
    _lib.check(_lib.lib().jp_csr_download(A.h, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(va)))
    assert np.array_equal(rp + 1, A.rowOffsets) and np.array_equal(ci + 1, A.localIndices1D) and np.array_equal(va, A.values)
    _, _, OA = oracle_problem("powerlaw", 20000, 1, max_len=500)
    orp, oci, ova = OA.csr(1)
    assert np.array_equal(rp + 1, orp) and np.array_equal(ci + 1, oci) and np.array_equal(va, ova)


# ---- dot / norm / elementwise ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(1, 1), (1000, 3), (1 << 20, 1), ((1 << 20) + 7, 8)])
def test_dot_norm_match_oracle(n, k):
    c = jp.SerialComm()
    m = jp.BlockMap(n, c)
    a = synthetic.vector_block(1, n, k, seed=1)
    b = synthetic.vector_block(1, n, k, seed=2)
    A_, B_ = jp.DenseMultiVector(m, a), jp.DenseMultiVector(m, b)
    d = A_.dot(B_)
    ref = O.dot_raw(a, b)
    scale = O.dot_raw(np.abs(a), np.abs(b))
    assert d.shape == (1, k) and (np.abs(d[0] - ref) <= TOL * scale).all()
    nr = A_.norm(2)
    refn = np.sqrt(O.norm_raw(a, 2.0))
    assert (np.abs(nr[0] - refn) <= TOL * refn).all()
    n3 = jp.DenseMultiVector(m, np.abs(a)).norm(3)
    ref3 = O.norm_raw(np.abs(a), 3.0) ** (1 / 3)
    assert (np.abs(n3[0] - ref3) <= 1e-11 * ref3).all()
    # exact sums as a second anchor (the oracle's own sequential rounding error grows like sqrt(n))
    import math
    for j in range(k):
        exact = math.fsum((a[:, j] * b[:, j]).tolist()) if n <= 1 << 16 else None
        if exact is not None:
            assert abs(d[0, j] - exact) <= TOL * scale[j]


def test_elementwise_bit_exact():
    n, k = 100003, 3
    c = jp.SerialComm()
    m = jp.BlockMap(n, c)
    x = synthetic.vector_block(1, n, k, seed=1)
    y = synthetic.vector_block(1, n, k, seed=2)
    Y, X = jp.DenseMultiVector(m, y), jp.DenseMultiVector(m, x)
    Y.axpby_(0.3, X, 1.7)
    oc = O.OComm(1)
    om = O.OMap.uniform(oc, n)
    oy, ox = O.OMV.from_arrays(om, [y]), O.OMV.from_arrays(om, [x])
    oy.axpby(0.3, ox, 1.7)
    assert np.array_equal(Y.data, oy.get(1))
    Y.scale_(1.0 / 3.0)
    oy.scale(1.0 / 3.0)
    assert np.array_equal(Y.data, oy.get(1))
    Y.scale_([0.1, 0.2, 0.7])
    oy.scale([0.1, 0.2, 0.7])
    assert np.array_equal(Y.data, oy.get(1))
    Y.fill_(-2.5)
    assert np.array_equal(Y.data, np.full((n, k), -2.5))


# ---- halo machinery on one GPU: a periodic problem whose wrap-around columns are "remote" columns served by a self
#      message (pack kernel -> exchange -> interior/boundary split -> halo kernel) ----------------------------------------
def _periodic_plan_problem(n, k):
    """1-D periodic 3-point stencil on one rank.  Column map = [1..n | n+1 (= row n's right neighbour, GID 1),
    n+2 (= row 1's left neighbour, GID n)]; the two halo entries are exported by this rank to itself."""
    lib = _lib.lib()
    c = jp.SerialComm()
    rowptr = np.arange(0, 3 * n + 1, 3, dtype=np.int32) + 1
    col = np.empty((n, 3), dtype=np.int32)
    idx = np.arange(1, n + 1)
    col[:, 0], col[:, 1], col[:, 2] = idx - 1, idx, idx + 1
    col[0, 0] = n + 2   # left neighbour of row 1 is GID n -> halo slot 2
    col[n - 1, 2] = n + 1  # right neighbour of row n is GID 1 -> halo slot 1
    col.sort(axis=1)
    vals = np.empty((n, 3))
    rng = np.random.default_rng(3)
    vals[:] = rng.standard_normal((n, 3))
    hA = C.c_int64()
    _lib.check(lib.jp_csr_create(n, n + 2, n, 3 * n, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(vals), 1, C.byref(hA)))
    remote = np.array([n + 1, n + 2], dtype=np.int32)
    export = np.array([1, n], dtype=np.int32)
    one = np.array([1], dtype=np.int32)
    two = np.array([2], dtype=np.int64)
    hP = C.c_int64()
    _lib.check(lib.jp_plan_create(c.handle, n, n + 2, n, None, None, 0, _lib.ptr(remote), 2, _lib.ptr(export), 2, _lib.ptr(one), _lib.ptr(two), 1,
                                  _lib.ptr(one), _lib.ptr(two), 1, 1, C.byref(hP)))
    return c, hA.value, hP.value, rowptr, col, vals


@pytest.mark.parametrize("n,k", [(5000, 1), (777, 4)])
def test_fused_apply_with_self_halo(n, k):
    lib = _lib.lib()
    c, hA, hP, rowptr, col, vals = _periodic_plan_problem(n, k)
    m = jp.BlockMap(n, c)
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    X, Y = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0)
    info = np.zeros(8, dtype=np.int64)
    _lib.check(lib.jp_csr_info(hA, _lib.ptr(info)))
    assert info[4] >= 1 and info[5] >= 1  # both interior and boundary row blocks exist
    for _ in range(3):  # repeated applies reuse the plan buffers
        Yw = jp.DenseMultiVector(m, y0)
        _lib.check(lib.jp_apply(c.handle, Yw.h, hA, X.h, hP, jp.NO_TRANS, 3.0, 0.5))
        got = Yw.data
    xcol = np.vstack([x, x[0:1, :], x[n - 1:n, :]])  # what doImport would have built
    ref = O.local_apply_raw(np.array(y0, order="F", copy=True), rowptr, col.ravel(), vals.ravel(), xcol, O.NO_TRANS, 3.0, 0.5)
    if k == 1:
        assert np.array_equal(got, ref)
    bound = O.local_apply_raw(np.abs(y0), rowptr, col.ravel(), np.abs(vals).ravel(), np.abs(xcol), O.NO_TRANS, 3.0, 0.5)
    assert (np.abs(got - ref) <= TOL * bound).all()
    # the generic doTransfer path builds the same column-map multivector
    mc = jp.BlockMap(n + 2, c)
    XC = jp.DenseMultiVector(mc, k)
    _lib.check(lib.jp_transfer(hP, X.h, XC.h, jp.INSERT, 0))
    assert np.array_equal(XC.data, xcol)
    # waiting without a post is an InvalidStateError (src/MPIDistributor.jl:574-576)
    with pytest.raises(jp.InvalidStateError):
        _lib.check(lib.jp_transfer_wait(hP))
    # host-buffer entry point gives the same numbers
    yh = np.array(y0, order="F", copy=True)
    _lib.check(lib.jp_apply_host(c.handle, hA, hP, _lib.ptr(np.asfortranarray(x)), _lib.ptr(yh), k, jp.NO_TRANS, 3.0, 0.5))
    assert np.array_equal(yh, got)
    _lib.check(lib.jp_plan_destroy(hP))
    _lib.check(lib.jp_csr_destroy(hA))


def test_general_column_map_path():
    """a column map with unused local columns (permutes): jp_apply falls back to import + single launch"""
    c = jp.SerialComm()
    m = jp.BlockMap(6, c)
    A = jp.CSRMatrix(m, [3, 1, 2, 1, 3, 2], jp.STATIC_PROFILE)
    for g, (ci, va) in {1: ([3, 1, 2], [1.0, 2.0, 4.0]), 2: ([2], [5.0]), 3: ([6, 5], [1.0, 2.0]), 4: ([1], [9.0]), 5: ([1, 2, 3], [1.0, 1.5, 1.25]),
                        6: ([6, 1], [7.0, 8.0])}.items():
        A.insertGlobalValues(g, ci, va)
    A.fillComplete(m, m)
    assert A.importer is not None and len(A.importer.permuteToLIDs()) == 2  # GID 4 is never referenced
    x = np.arange(1.0, 13.0).reshape((6, 2), order="F")
    X, Y = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, 2)
    A.apply_(Y, X)
    dense = np.zeros((6, 6))
    for g, (ci, va) in {1: ([3, 1, 2], [1.0, 2.0, 4.0]), 2: ([2], [5.0]), 3: ([6, 5], [1.0, 2.0]), 4: ([1], [9.0]), 5: ([1, 2, 3], [1.0, 1.5, 1.25]),
                        6: ([6, 1], [7.0, 8.0])}.items():
        for cc, vv in zip(ci, va):
            dense[g - 1, cc - 1] += vv
    assert np.array_equal(Y.data, dense @ x)


def test_host_buffer_apply_matches_device_apply():
    c, m, A = serial_matrix("7pt", 48)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    ref = O.local_apply_raw(np.array(y0, order="F", copy=True), A.rowOffsets, A.localIndices1D, A.values, x, O.NO_TRANS, 3.0, 0.5)
    yh = A.apply_host(x, np.array(y0, order="F", copy=True), jp.NO_TRANS, 3.0, 0.5)
    assert np.array_equal(yh, ref)
    assert np.array_equal(A.apply_host(x), O.local_apply_raw(np.zeros((n, 1)), A.rowOffsets, A.localIndices1D, A.values, x))


@pytest.mark.parametrize("kind,N,k", [("7pt", 96, 1), ("5pt", 900, 2), ("powerlaw", 700000, 1)])
def test_pipelined_host_buffer_apply(kind, N, k):
    """large enough (>= 4 MB of X) for jp_apply_host to take its pipelined path: chunked H2D of X, row-block chunks
    launched as soon as the columns they read are on the device, chunked D2H of Y; must equal the device path"""
    c, m, A = serial_matrix(kind, N, max_len=500)
    n = m.numMyElements()
    assert n * k * 8 >= 4 << 20
    x = synthetic.vector_block(1, n, k, seed=1)
    X, Y = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, k)
    A.apply_(Y, X)
    want = Y.data
    for _ in range(2):
        yh = np.full((n, k), np.nan, order="F")
        A.apply_host(x, yh)
        assert np.array_equal(yh, want)
    # beta != 0 takes the unpipelined path and must still agree
    y0 = synthetic.vector_block(1, n, k, seed=2)
    Y2 = jp.DenseMultiVector(m, y0)
    A.apply_(Y2, X, jp.NO_TRANS, 3.0, 0.5)
    assert np.array_equal(A.apply_host(x, np.array(y0, order="F", copy=True), jp.NO_TRANS, 3.0, 0.5), Y2.data)


def test_invalid_handles_and_arguments():
    lib = _lib.lib()
    with pytest.raises(jp.InvalidArgumentError):
        _lib.check(lib.jp_mv_fill(12345, 1.0))
    c = jp.SerialComm()
    m = jp.BlockMap(4, c)
    v = jp.DenseMultiVector(m, 1)
    with pytest.raises(jp.InvalidArgumentError):
        _lib.check(lib.jp_csr_info(v.h, None))  # wrong kind of handle
    rp = np.array([1, 2, 1], dtype=np.int32)
    h = C.c_int64()
    with pytest.raises(jp.InvalidArgumentError):  # decreasing row pointers
        _lib.check(lib.jp_csr_create(2, 2, 2, 0, _lib.ptr(rp), None, None, 1, C.byref(h)))
    n = C.c_int64()
    _lib.check(lib.jp_launch_count(C.byref(n)))
    assert n.value > 0


def test_transposed_apply_with_importer_serial():
    """TRANS through an importer (a column map with permutes, one rank): the reverse-mode ADD path of
    apply_trans_distributed, against a dense A^T X and against the oracle"""
    c = jp.SerialComm()
    m = jp.BlockMap(6, c)
    rows = {1: ([3, 1, 2], [1.0, 2.0, 4.0]), 2: ([2], [5.0]), 3: ([6, 5], [1.0, 2.0]), 4: ([1], [9.0]), 5: ([1, 2, 3], [1.0, 1.5, 1.25]),
            6: ([6, 1], [7.0, 8.0])}
    A = jp.CSRMatrix(m, [3, 1, 2, 1, 3, 2], jp.STATIC_PROFILE)
    oc = O.OComm(1)
    om = O.OMap.uniform(oc, 6)
    OA = O.OMatrix(om, [[3, 1, 2, 1, 3, 2]])
    dense = np.zeros((6, 6))
    for g, (ci, va) in rows.items():
        A.insertGlobalValues(g, ci, va)
        OA.insert(1, g, ci, va)
        for cc, vv in zip(ci, va):
            dense[g - 1, cc - 1] += vv
    A.fillComplete(m, m)
    OA.fill_complete(om, om)
    assert A.importer is not None
    x = np.arange(1.0, 13.0).reshape((6, 2), order="F")
    y0 = np.linspace(-1, 1, 12).reshape((6, 2), order="F")
    X, Y = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0)
    A.apply_(Y, X, jp.TRANS, 3.0, 0.5)
    OX, OY = O.OMV.from_arrays(om, [x]), O.OMV.from_arrays(om, [y0])
    OA.apply(OY, OX, O.TRANS, 3.0, 0.5)
    ref = 0.5 * y0 + 3.0 * (dense.T @ x)
    assert np.allclose(Y.data, ref, rtol=1e-13, atol=1e-13) and np.allclose(OY.get(1), ref, rtol=1e-13, atol=1e-13)


def test_spmm_default_path_selection(monkeypatch):
    """diagonal-structured matrices (stencils) take the diagonal kernel, random ones and odd leading dimensions the gather
    kernel; with the diagonal copy switched off a banded matrix takes the run-staged kernel"""
    c, m, A = serial_matrix("27pt", 16)
    n = m.numMyElements()
    check_apply(A, synthetic.vector_block(1, n, 5, seed=1), synthetic.vector_block(1, n, 5, seed=2), 3.0, 0.5, exact=False)
    info = A.spmm_info()
    assert info["dia_state"] == 1 and info["dia_diagonals"] == 27 and info["dia_rows"] > n // 2 and info["runs_state"] == -1, info
    c, m, A = serial_matrix("27pt", 15)  # n = 3375 is odd: X's columns are not all 16-byte aligned -> gather kernel
    n = m.numMyElements()
    check_apply(A, synthetic.vector_block(1, n, 5, seed=1), synthetic.vector_block(1, n, 5, seed=2), 3.0, 0.5, exact=False)
    c, m, A = serial_matrix("powerlaw", 20000)
    n = m.numMyElements()
    check_apply(A, synthetic.vector_block(1, n, 5, seed=1), synthetic.vector_block(1, n, 5, seed=2), 3.0, 0.5, exact=False)
    info = A.spmm_info()
    assert info["runs_state"] == -1 and info["dia_state"] == -1, info
    monkeypatch.setenv("JP_SPMM_DIA", "0")
    c, m, A = serial_matrix("27pt", 16)
    n = m.numMyElements()
    check_apply(A, synthetic.vector_block(1, n, 5, seed=1), synthetic.vector_block(1, n, 5, seed=2), 3.0, 0.5, exact=False)
    info = A.spmm_info()
    assert info["dia_state"] == -1 and info["runs_state"] == 1, info


@pytest.mark.parametrize("kind,N,k", [("27pt", 20, 8), ("27pt", 20, 11), ("27pt", 12, 32), ("7pt", 24, 6), ("5pt", 90, 3)])
def test_spmm_diagonal_and_run_staged_kernels(monkeypatch, kind, N, k):
    """the two SpMM kernels that stage X tiles with TMA bulk copies -- spmm_dia_kernel (diagonal-major values, the default
    for stencils) and spmm_runs_kernel (CSR with aligned column runs, the fallback for banded matrices) -- against the
    oracle, and against each other (on the rows the diagonal kernel serves both add a row's entries in ascending column order
    with FMAs -- bit-identical, tests/test_emul_spmv.py; the first and last rows go through the gather kernel's lanes-per-row
    tree, so the comparison here is to rounding)"""
    c, m, A = serial_matrix(kind, N)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, k, seed=1)
    y0 = synthetic.vector_block(1, n, k, seed=2)
    got_dia = check_apply(A, x, y0, 3.0, 0.5, exact=False)
    check_apply(A, x, y0, 1.0, 0.0, exact=False)
    assert A.spmm_info()["dia_state"] == 1, A.spmm_info()
    monkeypatch.setenv("JP_SPMM_DIA", "0")
    c, m, A2 = serial_matrix(kind, N)
    got_runs = check_apply(A2, x, y0, 3.0, 0.5, exact=False)
    check_apply(A2, x, y0, 1.0, 0.0, exact=False)
    info = A2.spmm_info()
    assert info["runs_state"] == 1 and info["runs_blocks"] > 0, info  # the run-staged kernel really ran
    assert np.abs(got_dia - got_runs).max() <= 1e-13 * np.abs(got_runs).max()


def test_full_size_7pt_256_properties_and_oracle():
    """BASELINE configs[1] at full size on one GPU: size-independent properties, then the whole vector against the
    oracle (bit-exact: every block of a stencil is thread-per-row in the reference's summation order)."""
    N = 256
    c, m, A = serial_matrix("7pt", N)
    n = m.numMyElements()
    assert n == N ** 3 and len(A.values) == 7 * N ** 3 - 6 * N ** 2
    ones = jp.DenseMultiVector(m, np.ones((n, 1)))
    Y = jp.DenseMultiVector(m, 1)
    A.apply_(Y, ones)
    rs = Y.data[:, 0].reshape((N, N, N))  # [k, j, i]
    # row sums of the Dirichlet-truncated Laplacian = number of missing neighbours
    idx = np.arange(N)
    edge = ((idx == 0) | (idx == N - 1)).astype(np.float64)
    expect = edge[:, None, None] + edge[None, :, None] + edge[None, None, :]
    assert np.array_equal(rs, expect)
    x = synthetic.vector_block(1, n, 1, seed=1)
    z = synthetic.vector_block(1, n, 1, seed=2)
    X, Z = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, z)
    AX, AZ = jp.DenseMultiVector(m, 1), jp.DenseMultiVector(m, 1)
    A.apply_(AX, X)
    A.apply_(AZ, Z)
    # symmetry: z.(A x) == x.(A z); linearity: A(2x - 3z) == 2 Ax - 3 Az
    s1, s2 = Z.dot(AX)[0, 0], X.dot(AZ)[0, 0]
    scale = 12.0 * np.abs(x).T @ np.abs(z)
    assert abs(s1 - s2) <= 1e-12 * float(scale[0, 0]) + 1e-6 * 0 + 1e-9
    W = jp.DenseMultiVector(m, 2.0 * x - 3.0 * z)
    AW = jp.DenseMultiVector(m, 1)
    A.apply_(AW, W)
    comb = 2.0 * AX.data - 3.0 * AZ.data
    assert np.abs(AW.data - comb).max() <= 1e-12 * 12 * 5
    # the full vector against the oracle's localApply, alpha=3 beta=0.5
    y0 = z.copy(order="F")
    Yb = jp.DenseMultiVector(m, y0)
    A.apply_(Yb, X, jp.NO_TRANS, 3.0, 0.5)
    ref = O.local_apply_raw(y0, A.rowOffsets, A.localIndices1D, A.values, x, O.NO_TRANS, 3.0, 0.5)
    assert np.array_equal(Yb.data, ref)
    # norm / dot at n = 16.7M: the oracle's own sequential rounding error is ~sqrt(n)*eps, so anchor on per-vector
    # partial agreement at 1e-12 of sum|x||y| and on the exact (math.fsum) value
    import math
    d = X.dot(Z)[0, 0]
    exact = math.fsum((x[:, 0] * z[:, 0]).tolist())
    assert abs(d - exact) <= 1e-12 * float((np.abs(x).T @ np.abs(z))[0, 0])
    assert abs(d - O.dot_raw(x, z)[0]) <= 1e-11 * float((np.abs(x).T @ np.abs(z))[0, 0])


@pytest.mark.parametrize("key,kind,size", [("7pt_64_serial", "7pt", 64), ("5pt_300_serial", "5pt", 300), ("27pt_20_serial", "27pt", 20)])
def test_against_committed_golden_digests(key, kind, size):
    """the CUDA path against tests/golden/fixtures.json (no oracle call): SHA-256 of the Float64 result of
    apply!(Y, A, X, NO_TRANS, 3, 0.5) and of A*X must equal the committed digests; dot / norm within 1e-12"""
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures.json")))[key]
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    c, m, A = serial_matrix(kind, size)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    X, Y = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0)
    A.apply_(Y, X, jp.NO_TRANS, 3.0, 0.5)
    assert dig(Y.data) == g["apply_3_05_sha256"]
    AX = A * X
    assert dig(AX.data) == g["ax_sha256"]
    assert abs(X.dot(AX)[0, 0] - g["dot_x_ax"]) <= 1e-12 * float((np.abs(x).T @ np.abs(AX.data))[0, 0])
    assert abs(AX.norm(2)[0, 0] - g["norm2_ax"]) <= 1e-12 * g["norm2_ax"]


@pytest.mark.parametrize("slab_mb", ["1", "0"])
def test_spmv_column_slabs_on_a_large_irregular_matrix(monkeypatch, slab_mb):
    """a power-law matrix whose x (3.2 MB) is larger than two column slabs of 1 MB takes spmv_cb_kernel (nonzeros in
    slab-major order, segmented warp sums, carries; no atomics); JP_CB_SLAB_MB=0 keeps the row-block kernel.  NO_TRANS and
    TRANS (the cached transposed matrix is swept the same way), k = 1; k > 1 stays with the gather SpMM kernel"""
    monkeypatch.setenv("JP_CB_SLAB_MB", slab_mb)
    c, m, A = serial_matrix("powerlaw", 400_000, max_len=3000)
    n = m.numMyElements()
    x = synthetic.vector_block(1, n, 1, seed=1)
    y0 = synthetic.vector_block(1, n, 1, seed=2)
    got = check_apply(A, x, y0, 3.0, 0.5, exact=False)
    check_apply(A, x, y0, 1.0, 0.0, exact=False)
    info = A.spmm_info()
    assert (info["colblock_state"], info["colblock_slabs"]) == ((1, 4) if slab_mb == "1" else (-1, 0)), info
    assert np.array_equal(check_apply(A, x, y0, 3.0, 0.5, exact=False), got)  # deterministic
    X, Yt = jp.DenseMultiVector(m, x), jp.DenseMultiVector(m, y0)
    A.apply_(Yt, X, jp.TRANS, 3.0, 0.5)
    ref = O.local_apply_raw(np.array(y0, order="F", copy=True), A.rowOffsets, A.localIndices1D, A.values, x, O.TRANS, 3.0, 0.5)
    bound = O.local_apply_raw(np.abs(y0).copy(order="F"), A.rowOffsets, A.localIndices1D, np.abs(A.values), np.abs(x), O.TRANS, 3.0, 0.5)
    assert (np.abs(Yt.data - ref) <= TOL * bound + 1e-300).all()
    x3 = synthetic.vector_block(1, n, 3, seed=1)
    check_apply(A, x3, synthetic.vector_block(1, n, 3, seed=2), 3.0, 0.5, exact=False)
