"""ctypes loader of tests/cuda_emul/_build/libjp_emul.so: the product's *_kernels.cuh and *_pipeline.h compiled for the
CPU execution model of tests/cuda_emul/cuda_emul.h.  TEST INFRASTRUCTURE ONLY (imported by tests/test_emul_*.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emul")
_lib = None


class EmulError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)
        L = C.CDLL(os.path.join(_DIR, "_build", "libjp_emul.so"))
        L.emul_last_error.restype = C.c_char_p
        L.emul_launch_count.restype = C.c_uint64
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise EmulError(rc, lib().emul_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fill_complete(entry_ptr, gids, vals, gmin, gmax, dom_min, dom_n):
    """device fillComplete (csrc/jp_fill_pipeline.h) on the CPU model -> dict with 0-based rowptr/colind, vals,
    colmap GIDs, rowmax and the counts"""
    ptr = np.ascontiguousarray(entry_ptr, dtype=np.int64)
    g = np.ascontiguousarray(gids, dtype=np.int64)
    v = np.ascontiguousarray(vals, dtype=np.float64)
    n_rows = len(ptr) - 1
    info = np.zeros(4, dtype=np.int64)
    h = C.c_void_p()
    _check(lib().emul_fill_complete(C.c_int32(n_rows), _p(ptr), _p(g), _p(v), C.c_int64(gmin), C.c_int64(gmax), C.c_int64(dom_min),
                                    C.c_int32(dom_n), _p(info), C.byref(h)))
    n_cols, n_local_used, n_same, nnz = (int(x) for x in info)
    out = dict(n_cols=n_cols, n_local_used=n_local_used, n_same=n_same, nnz=nnz,
               rowptr=np.empty(n_rows + 1, dtype=np.int32), colind=np.empty(nnz, dtype=np.int32),
               vals=np.empty(nnz, dtype=np.float64), colmap=np.empty(n_cols, dtype=np.int64), rowmax=np.empty(n_rows, dtype=np.int32))
    lib().emul_fill_get(h, _p(out["rowptr"]), _p(out["colind"]), _p(out["vals"]), _p(out["colmap"]), _p(out["rowmax"]))
    lib().emul_fill_free(h)
    return out


def local_apply(rowptr0, colind0, vals, x, y, alpha=1.0, beta=0.0, n_local_cols=None, halo_split=False, variant=0, cap=2048,
                max_rows=256, T=1, tma=1, allow_direct=1):
    """y = beta*y + alpha*A*x through csrc/jp_csr_kernels.cuh on the CPU model.  x: (n_cols, k) column-map multivector.
    halo_split: rows >= n_local_cols of x are handed over as the LID-major halo buffer (the HALO=true kernels).
    Returns (y_new, stats) or (None, stats) when the tile tables are unusable for this matrix."""
    rp = np.ascontiguousarray(rowptr0, dtype=np.int32)
    ci = np.ascontiguousarray(colind0, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float64)
    x = np.asfortranarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1, order="F")
    n_cols, k = x.shape
    n_rows = len(rp) - 1
    yy = np.asfortranarray(np.array(y, dtype=np.float64).reshape(n_rows, k), dtype=np.float64).copy(order="F")
    nloc = n_cols if n_local_cols is None else int(n_local_cols)
    if halo_split:
        xl = np.asfortranarray(x[:nloc, :]).copy(order="F")
        halo = np.ascontiguousarray(x[nloc:, :])  # (n_remote, k) row-major = LID-major, vector-minor
        x_rows, hp = nloc, _p(halo)
    else:
        xl, x_rows, hp = x, n_cols, None
    stats = np.zeros(4, dtype=np.int64)
    rc = lib().emul_local_apply(C.c_int32(n_rows), C.c_int32(n_cols), C.c_int32(nloc), _p(rp), _p(ci), _p(va), _p(xl), C.c_int32(x_rows), hp,
                                _p(yy), C.c_int32(k), C.c_double(alpha), C.c_double(beta), C.c_int32(variant), C.c_int32(cap),
                                C.c_int32(max_rows), C.c_int32(T), C.c_int32(tma), C.c_int32(allow_direct), _p(stats))
    st = dict(zip(("blocks", "direct", "subwarp", "longrow"), (int(s) for s in stats)))
    return (None if rc != 0 else yy), st


def spmv_dot(rowptr0, colind0, vals, x, w, y, alpha=1.0, beta=0.0, cap=2048, max_rows=256, tma=1, repeats=2):
    """spmv_dot_kernel on the CPU model: returns (y_new, [dot per repeat])"""
    rp = np.ascontiguousarray(rowptr0, dtype=np.int32)
    ci = np.ascontiguousarray(colind0, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float64)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).ravel())
    yy = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel()).copy()
    out = np.zeros(repeats, dtype=np.float64)
    rc = lib().emul_spmv_dot(C.c_int32(len(rp) - 1), C.c_int32(len(x)), _p(rp), _p(ci), _p(va), _p(x), _p(w), _p(yy), C.c_double(alpha),
                             C.c_double(beta), C.c_int32(cap), C.c_int32(max_rows), C.c_int32(tma), C.c_int32(repeats), _p(out))
    assert rc == 0, f"emul_spmv_dot rc={rc}"
    return yy, out


def axpby_norm2(y, a, x, b, blocks=3, repeats=1):
    """axpby_norm2_kernel on the CPU model, applied `repeats` times: returns (y_final, sums[repeats, k])"""
    yy = np.asfortranarray(np.array(y, dtype=np.float64))
    if yy.ndim == 1:
        yy = yy.reshape(-1, 1, order="F")
    xx = np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(yy.shape, order="F"))
    yy = yy.copy(order="F")
    n, k = yy.shape
    out = np.zeros((repeats, k), dtype=np.float64)
    rc = lib().emul_axpby_norm2(_p(yy), _p(xx), C.c_int64(n), C.c_int32(k), C.c_double(a), C.c_double(b), C.c_int32(blocks),
                                C.c_int32(repeats), _p(out))
    assert rc == 0
    return yy, out


def _f(a, shape=None):
    a = np.asfortranarray(np.array(a, dtype=np.float64))
    if a.ndim == 1:
        a = a.reshape(-1, 1, order="F")
    return a.copy(order="F")


def mv_reduce(op, x, y=None, p=2.0, blocks=3, repeats=2):
    """reduce_kernel<op> (0 dot, 1 sum of squares, 2 sum of p-th powers) -> sums[repeats, k]"""
    x = _f(x)
    n, k = x.shape
    yy = _f(y) if y is not None else None
    out = np.zeros((repeats, k), dtype=np.float64)
    rc = lib().emul_mv_reduce(C.c_int32(op), _p(x), _p(yy) if yy is not None else None, C.c_int64(n), C.c_int32(k), C.c_double(p),
                              C.c_int32(blocks), C.c_int32(repeats), _p(out))
    assert rc == 0
    return out


def mv_elementwise(op, y, x=None, alpha_k=None, a=0.0, b=0.0, blocks=3):
    """0 fill(y, a), 1 scale(y, a), 2 scale_vec(y, alpha_k), 3 axpby(y, a, x, b) -> new y"""
    yy = _f(y)
    n, k = yy.shape
    xx = _f(x) if x is not None else None
    ak = np.ascontiguousarray(alpha_k, dtype=np.float64) if alpha_k is not None else None
    rc = lib().emul_mv_elementwise(C.c_int32(op), _p(yy), _p(xx) if xx is not None else None, _p(ak) if ak is not None else None,
                                   C.c_int64(n), C.c_int32(k), C.c_double(a), C.c_double(b), C.c_int32(blocks))
    assert rc == 0
    return yy


def plan_op(op, mv, other, lids_a=None, lids_b=None, n_ids=None, blocks=2):
    """transfer kernels with 0-based LIDs (see emul_plan_op); returns (mv, other) after the launch.  For ops 0-2 `other`
    is the (n_ids, k) LID-major buffer; for ops 3-5 the second multivector."""
    m = _f(mv)
    n_mv, k = m.shape
    la = np.ascontiguousarray(lids_a, dtype=np.int32) if lids_a is not None else None
    lb = np.ascontiguousarray(lids_b, dtype=np.int32) if lids_b is not None else None
    if n_ids is None:
        n_ids = len(la)
    if op <= 2:
        o = np.ascontiguousarray(np.array(other, dtype=np.float64).reshape(n_ids, k))  # row-major = LID-major, vector-minor
        n_other = n_ids
    else:
        o = _f(other)
        n_other = o.shape[0]
    rc = lib().emul_plan_op(C.c_int32(op), _p(m), C.c_int64(n_mv), _p(o), C.c_int64(n_other), _p(la) if la is not None else None,
                            _p(lb) if lb is not None else None, C.c_int32(n_ids), C.c_int32(k), C.c_int32(blocks))
    assert rc == 0
    return m, o


def spmm_dia(rowptr0, colind0, vals, x, y, alpha=1.0, beta=0.0, parity=0):
    """spmm_dia_kernel on the CPU model with the library's plan (jp_dia_plan.h) and the device builders (dia_mark_kernel,
    dia_fill_kernel).  Returns (y_new, row_lo, row_hi): rows [row_lo, row_hi) are computed by the diagonal kernel, the
    others are left as they were (the library serves them with the gather kernel); (None, 0, 0) when not eligible.
    parity: the parity of the first row served (the library takes it from the CSR row block it starts at)."""
    rp = np.ascontiguousarray(rowptr0, dtype=np.int32)
    ci = np.ascontiguousarray(colind0, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float64)
    xx, yy = _f(x), _f(y).copy(order="F")
    rng = np.zeros(2, dtype=np.int32)
    rc = lib().emul_spmm_dia(C.c_int32(len(rp) - 1), C.c_int32(xx.shape[0]), _p(rp), _p(ci), _p(va), _p(xx), _p(yy), C.c_int32(xx.shape[1]),
                             C.c_double(alpha), C.c_double(beta), C.c_int32(parity), _p(rng))
    if rc != 0:
        return None, 0, 0
    return yy, int(rng[0]), int(rng[1])


def spmv_cb(rowptr0, colind0, vals, x, y, alpha=1.0, beta=0.0, slab_cols=1024):
    """column-slab SpMV (spmv_cb_kernel + spmv_cb_carry_kernel and the device builders) on the CPU model; k = 1"""
    rp = np.ascontiguousarray(rowptr0, dtype=np.int32)
    ci = np.ascontiguousarray(colind0, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float64)
    xx = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    yy = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel()).copy()
    rc = lib().emul_spmv_cb(C.c_int32(len(rp) - 1), C.c_int32(len(xx)), _p(rp), _p(ci), _p(va), _p(xx), _p(yy), C.c_double(alpha), C.c_double(beta),
                            C.c_int32(slab_cols))
    assert rc == 0
    return yy.reshape(-1, 1)


def transpose_csr(rowptr0, colind0, vals, n_cols):
    """A^T as CSR the way csrc/jp_transpose.cu builds it: a STABLE sort of the nonzero positions by column index, so the
    entries of a row of A^T are in ascending row order of A (0-based arrays in, 0-based arrays out)"""
    rp = np.asarray(rowptr0, dtype=np.int64)
    ci = np.asarray(colind0, dtype=np.int64)
    row_of = np.repeat(np.arange(len(rp) - 1, dtype=np.int64), np.diff(rp))
    order = np.argsort(ci, kind="stable")
    rpt = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(np.bincount(ci, minlength=n_cols), out=rpt[1:])
    return rpt.astype(np.int32), row_of[order].astype(np.int32), np.asarray(vals, dtype=np.float64)[order]


def set_schedule(mode: str = "forward", seed: int = 0):
    """order in which a block's runnable threads get their turn: forward | reverse | random (seeded)"""
    lib().emul_set_schedule(C.c_int({"forward": 0, "reverse": 1, "random": 2}[mode]), C.c_uint64(seed))


def selftest_neighbour(nthreads: int, with_barrier: bool):
    out = np.full(nthreads, -1, dtype=np.int32)
    lib().emul_selftest_neighbour(_p(out), C.c_int(nthreads), C.c_int(int(with_barrier)))
    return out


def axpby_dev(y, x, s, a, b, with_norm=False, out_slot=0, blocks=3, repeats=1):
    """axpby_dev_kernel / axpby_norm2_dev_kernel: a, b = (scale, num_slot, den_slot) with -1 for "no slot".
    Returns (y_new, s_new)."""
    yy = np.ascontiguousarray(np.array(y, dtype=np.float64).ravel()).copy()
    xx = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    ss = np.ascontiguousarray(np.array(s, dtype=np.float64).ravel()).copy()
    rc = lib().emul_axpby_dev(_p(yy), _p(xx), C.c_int64(len(yy)), _p(ss), C.c_int32(len(ss)), C.c_double(a[0]), C.c_int32(a[1]), C.c_int32(a[2]),
                              C.c_double(b[0]), C.c_int32(b[1]), C.c_int32(b[2]), C.c_int32(int(with_norm)), C.c_int32(out_slot),
                              C.c_int32(blocks), C.c_int32(repeats))
    assert rc == 0
    return yy, ss
