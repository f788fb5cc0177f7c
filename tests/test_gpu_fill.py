"""Device fillComplete through the C ABI (jp_csr_fill_complete) on the GPU: the cases of tests/fill_cases.py against the
CPU oracle (bit-exact column map / local indices / merged values for 1-4 simulated ranks), and the host mirror's
fillComplete(device=True) against its host path, including apply! on the result."""
import ctypes as C

import numpy as np
import pytest

import fill_cases
from helpers import jp, synthetic

pytestmark = pytest.mark.gpu


def gpu_fill(entry_ptr, gids, vals, gmin, gmax, dom_min, dom_n):
    lib = jp._lib.lib()
    ptr = np.ascontiguousarray(entry_ptr, dtype=np.int64)
    g = np.ascontiguousarray(gids, dtype=np.int64)
    v = np.ascontiguousarray(vals, dtype=np.float64)
    n_rows = len(ptr) - 1
    info = np.zeros(4, dtype=np.int64)
    h = C.c_int64()
    jp._lib.check(lib.jp_csr_fill_complete(n_rows, jp._lib.ptr(ptr), jp._lib.ptr(g), jp._lib.ptr(v), gmin, gmax, dom_min, dom_n, C.byref(h),
                                           jp._lib.ptr(info)))
    n_cols, n_local_used, n_same, nnz = (int(x) for x in info)
    out = dict(n_cols=n_cols, n_local_used=n_local_used, n_same=n_same, nnz=nnz, rowmax=None,
               rowptr=np.empty(n_rows + 1, dtype=np.int32), colind=np.empty(nnz, dtype=np.int32), vals=np.empty(nnz, dtype=np.float64),
               colmap=np.empty(n_cols, dtype=np.int64))
    jp._lib.check(lib.jp_csr_download(h.value, jp._lib.ptr(out["rowptr"]), jp._lib.ptr(out["colind"]), jp._lib.ptr(out["vals"])))
    jp._lib.check(lib.jp_csr_col_map(h.value, jp._lib.ptr(out["colmap"])))
    dinfo = np.zeros(8, dtype=np.int64)
    jp._lib.check(lib.jp_csr_info(h.value, jp._lib.ptr(dinfo)))
    assert (dinfo[0], dinfo[1], dinfo[2], dinfo[3]) == (n_rows, n_cols, n_same, nnz)
    jp._lib.check(lib.jp_csr_destroy(h.value))
    return out


@pytest.mark.parametrize("kind,N,P", [("5pt", 40, 1), ("7pt", 12, 3), ("27pt", 9, 2), ("powerlaw", 3000, 1), ("powerlaw", 3000, 4)])
def test_synthetic_problems_match_oracle(kind, N, P):
    fill_cases.synthetic_case(gpu_fill, kind, N, P)


def test_duplicates_long_rows_and_unused_local_columns():
    fill_cases.duplicates_long_rows_case(gpu_fill)


def test_domain_map_differs_from_row_map_and_empty_rank():
    fill_cases.domain_differs_case(gpu_fill)


def test_empty_matrix_and_bad_gid():
    fill_cases.empty_and_bad_gid_case(gpu_fill, jp.InvalidArgumentError)


@pytest.mark.parametrize("kind,N", [("7pt", 24), ("27pt", 12), ("powerlaw", 20000)])
def test_mirror_device_fill_equals_host_fill(kind, N):
    comm = jp.SerialComm()
    n = N if kind == "powerlaw" else N ** synthetic.STENCILS[kind][0]
    m = jp.BlockMap(n, comm)
    Ah = synthetic.build_matrix(jp, m, kind, N)
    Ad = synthetic.build_matrix(jp, m, kind, N, fill=False)
    Ad.fillComplete(m, m, device=True)
    assert Ad.h and Ad._csr_host is None  # the device path ran; nothing was packed on the host
    assert Ad.getColMap().sameAs(Ah.getColMap()) and Ad.getImporter() is None
    assert Ad.getLocalNumEntries() == Ah.getLocalNumEntries()
    assert np.array_equal(Ad.rowOffsets, Ah.rowOffsets) and np.array_equal(Ad.localIndices1D, Ah.localIndices1D)
    assert np.array_equal(Ad.values, Ah.values)
    assert Ad.device_info() == Ah.device_info()  # same row blocks
    x = synthetic.vector_block(1, n, 2, seed=1)
    X = jp.DenseMultiVector(m, x)
    Yh, Yd = Ah.apply(X), Ad.apply(X)
    assert np.array_equal(Yh.data, Yd.data)


def test_mirror_device_fill_rejects_remote_columns_on_one_process():
    comm = jp.SerialComm()
    m = jp.BlockMap(4, comm)
    A = jp.CSRMatrix(m, 2, jp.STATIC_PROFILE)
    for g in range(1, 5):
        A.insertGlobalValues(g, [g, g + 3], [1.0, 2.0])  # column 7 is owned by nobody
    with pytest.raises(RuntimeError):
        A.fillComplete(device=True)
