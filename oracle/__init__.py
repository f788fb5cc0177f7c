"""ctypes loader for the CPU oracle (oracle/jp_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs -- never from the product
package.  All ids are 1-based like the reference; "ranks" are simulated
(``OComm(P)``), every per-rank argument is a list indexed by pid-1.

Parity pinning: see the header of jp_oracle.cpp and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libjp_oracle.so")

INVALID_ARGUMENT = 1
INVALID_STATE = 2

# reference enum values (src/Operator.jl:10, src/DistObject.jl:21)
NO_TRANS, TRANS, CONJ_TRANS = 1, 2, 3
ADD, INSERT, REPLACE, ABSMAX, ZERO = 1, 2, 3, 4, 5


class OracleInvalidArgument(Exception):
    """InvalidArgumentError of the reference (src/Utils.jl:37-43)."""


class OracleInvalidState(Exception):
    """InvalidStateError of the reference (src/Utils.jl:45-51)."""


class OracleError(Exception):
    pass


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jp_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_comm_create", "orc_dist_create", "orc_plan_distributor", "orc_matrix_col_map",
                     "orc_matrix_importer", "orc_matrix_exporter"):
            getattr(L, name).restype = C.c_void_p
        for name in ("orc_dist_field_len", "orc_dist_result_bytes", "orc_plan_field_len", "orc_mv_local_length",
                     "orc_matrix_local_nnz", "orc_matrix_local_rows", "orc_map_gid"):
            getattr(L, name).restype = C.c_int64
        L.orc_map_lid.restype = C.c_int32
        _lib = L
    return _lib


def _check(rc):
    if rc == 0:
        return
    msg = lib().orc_last_error().decode()
    if rc == INVALID_ARGUMENT:
        raise OracleInvalidArgument(msg)
    if rc == INVALID_STATE:
        raise OracleInvalidState(msg)
    raise OracleError(msg)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _flat(per_rank, dtype):
    offs = np.zeros(len(per_rank) + 1, dtype=np.int64)
    for i, a in enumerate(per_rank):
        offs[i + 1] = offs[i] + len(a)
    flat = np.concatenate([np.asarray(a, dtype=dtype).ravel() for a in per_rank]) if len(per_rank) else np.zeros(0, dtype)
    return np.ascontiguousarray(flat, dtype=dtype), offs


class OComm:
    """SimComm of P logical ranks (the analogue of `mpiexec -n P`)."""

    def __init__(self, P: int):
        self.P = P
        self.h = C.c_void_p(lib().orc_comm_create(P))

    def _reduce(self, op, per_rank):
        a = np.ascontiguousarray(per_rank)
        n = a.shape[1]
        if np.issubdtype(a.dtype, np.floating):
            a = a.astype(np.float64)
            out = np.empty_like(a)
            _check(lib().orc_reduce_f64(self.h, op, _p(a, C.c_double), C.c_int64(n), _p(out, C.c_double)))
        else:
            a = a.astype(np.int64)
            out = np.empty_like(a)
            _check(lib().orc_reduce_i64(self.h, op, _p(a, C.c_int64), C.c_int64(n), _p(out, C.c_int64)))
        return out

    def sumAll(self, per_rank):
        return self._reduce(0, per_rank)

    def maxAll(self, per_rank):
        return self._reduce(1, per_rank)

    def minAll(self, per_rank):
        return self._reduce(2, per_rank)

    def scanSum(self, per_rank):
        return self._reduce(3, per_rank)

    def gatherAll(self, per_rank):
        flat, offs = _flat(per_rank, np.int64)
        out = np.empty(int(offs[-1]), dtype=np.int64)
        _check(lib().orc_gather_all_i64(self.h, _p(flat, C.c_int64), _p(offs, C.c_int64), _p(out, C.c_int64)))
        return out

    def broadcastAll(self, per_rank, root):
        a = _i64(per_rank)
        out = np.empty_like(a)
        _check(lib().orc_broadcast_all_i64(self.h, _p(a, C.c_int64), C.c_int64(a.shape[1]), root, _p(out, C.c_int64)))
        return out


def compute_offsets_const(n, num_ents):
    out = np.empty(n, dtype=np.int64)
    _check(lib().orc_compute_offsets_const(_p(out, C.c_int64), C.c_int64(n), C.c_int64(num_ents)))
    return out


def compute_offsets(n, num_ents):
    out = np.empty(n, dtype=np.int64)
    e = _i64(num_ents)
    tot = C.c_int64(0)
    _check(lib().orc_compute_offsets(_p(out, C.c_int64), C.c_int64(n), _p(e, C.c_int64), C.c_int64(len(e)), C.byref(tot)))
    return tot.value, out


class OMap:
    INFO = ("numMyElements", "minMyGID", "maxMyGID", "minAllGID", "maxAllGID", "minLID", "maxLID", "linearMap",
            "distributedGlobal", "numGlobalElements")

    def __init__(self, comm: OComm, handle, owned=True):
        self.comm = comm
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.owned = owned

    @classmethod
    def uniform(cls, comm, N):
        h = C.c_void_p()
        _check(lib().orc_map_uniform(comm.h, C.c_int64(N), C.byref(h)))
        return cls(comm, h)

    @classmethod
    def linear(cls, comm, N, num_my):
        nm = np.ascontiguousarray(num_my, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().orc_map_linear(comm.h, C.c_int64(N), _p(nm, C.c_int32), C.byref(h)))
        return cls(comm, h)

    @classmethod
    def from_gids(cls, comm, N, gids_per_rank):
        flat, offs = _flat(gids_per_rank, np.int64)
        h = C.c_void_p()
        _check(lib().orc_map_gids(comm.h, C.c_int64(N), _p(flat, C.c_int64), _p(offs, C.c_int64), C.byref(h)))
        return cls(comm, h)

    def info(self, pid):
        out = np.empty(10, dtype=np.int64)
        lib().orc_map_info(self.h, pid, _p(out, C.c_int64))
        return dict(zip(self.INFO, (int(v) for v in out)))

    def lid(self, pid, gid):
        return int(lib().orc_map_lid(self.h, pid, C.c_int64(gid)))

    def gid(self, pid, lid):
        return int(lib().orc_map_gid(self.h, pid, C.c_int64(lid)))

    def my_gids(self, pid):
        out = np.empty(self.info(pid)["numMyElements"], dtype=np.int64)
        lib().orc_map_my_gids(self.h, pid, _p(out, C.c_int64))
        return out

    def remote_id_list(self, pid, gids):
        g = _i64(gids)
        pids = np.empty(len(g), dtype=np.int32)
        lids = np.empty(len(g), dtype=np.int32)
        _check(lib().orc_map_remote_id_list(self.h, pid, _p(g, C.c_int64), C.c_int64(len(g)), _p(pids, C.c_int32), _p(lids, C.c_int32)))
        return pids, lids

    def same_as(self, other):
        return bool(lib().orc_map_same_as(self.h, other.h))


class ODist:
    FIELDS = {"procs_to": 0, "lengths_to": 1, "starts_to": 2, "indices_to": 3, "procs_from": 4, "lengths_from": 5,
              "starts_from": 6, "scalars": 7, "exportGIDs": 8, "exportPIDs": 9}

    def __init__(self, comm: OComm, handle=None):
        self.comm = comm
        self.h = C.c_void_p(lib().orc_dist_create(comm.h)) if handle is None else C.c_void_p(handle)

    def create_from_sends(self, export_pids_per_rank):
        flat, offs = _flat(export_pids_per_rank, np.int32)
        out = np.empty(self.comm.P, dtype=np.int64)
        _check(lib().orc_dist_create_from_sends(self.h, _p(flat, C.c_int32), _p(offs, C.c_int64), _p(out, C.c_int64)))
        return out

    def create_from_recvs(self, remote_gids_per_rank, remote_pids_per_rank):
        g, offs = _flat(remote_gids_per_rank, np.int64)
        p, offs2 = _flat(remote_pids_per_rank, np.int32)
        if not np.array_equal(offs, offs2):
            raise OracleInvalidArgument("remote lists must be the same length")
        _check(lib().orc_dist_create_from_recvs(self.h, _p(g, C.c_int64), _p(p, C.c_int32), _p(offs, C.c_int64)))
        return ([self.field(pid, "exportGIDs") for pid in range(1, self.comm.P + 1)],
                [self.field(pid, "exportPIDs") for pid in range(1, self.comm.P + 1)])

    def field(self, pid, name, reverse=False):
        f = self.FIELDS[name]
        n = lib().orc_dist_field_len(self.h, pid, f, int(reverse))
        if n < 0:
            _check(3)
        out = np.empty(n, dtype=np.int64)
        _check(lib().orc_dist_field_get(self.h, pid, f, int(reverse), _p(out, C.c_int64)))
        return out

    def scalars(self, pid, reverse=False):
        s = self.field(pid, "scalars", reverse)
        return dict(numSends=int(s[0]), numRecvs=int(s[1]), selfMsg=int(s[2]), totalRecvLength=int(s[3]), numExports=int(s[4]))

    def resolve_posts(self, objs_per_rank, dtype=np.int64, reverse=False):
        arrs = [np.ascontiguousarray(a, dtype=dtype) for a in objs_per_rank]
        elem = arrs[0].dtype.itemsize * (int(np.prod(arrs[0].shape[1:])) if arrs[0].ndim > 1 else 1)
        offs = np.zeros(len(arrs) + 1, dtype=np.int64)
        for i, a in enumerate(arrs):
            offs[i + 1] = offs[i] + a.shape[0]
        raw = np.concatenate([a.view(np.uint8).ravel() for a in arrs]) if arrs else np.zeros(0, np.uint8)
        raw = np.ascontiguousarray(raw)
        self._dtype, self._tail = arrs[0].dtype, arrs[0].shape[1:]
        _check(lib().orc_dist_resolve_posts(self.h, _p(raw, C.c_uint8), _p(offs, C.c_int64), elem, int(reverse)))

    def resolve_waits(self, reverse=False):
        _check(lib().orc_dist_resolve_waits(self.h, int(reverse)))
        out = []
        for pid in range(1, self.comm.P + 1):
            nb = lib().orc_dist_result_bytes(self.h, pid)
            buf = np.empty(nb, dtype=np.uint8)
            lib().orc_dist_result_get(self.h, pid, _p(buf, C.c_uint8))
            out.append(buf.view(self._dtype).reshape((-1,) + tuple(self._tail)))
        return out

    def resolve(self, objs_per_rank, dtype=np.int64, reverse=False):
        self.resolve_posts(objs_per_rank, dtype, reverse)
        return self.resolve_waits(reverse)


class OPlan:
    """Import or Export plan."""
    FIELDS = {"numSameIDs": 0, "permuteToLIDs": 1, "permuteFromLIDs": 2, "remoteLIDs": 3, "exportLIDs": 4,
              "exportPIDs": 5, "isLocallyComplete": 6}

    def __init__(self, comm, handle, owned=True):
        self.comm = comm
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.owned = owned

    @classmethod
    def make_import(cls, source: OMap, target: OMap):
        h = C.c_void_p()
        _check(lib().orc_import_create(source.h, target.h, C.byref(h)))
        return cls(source.comm, h)

    @classmethod
    def make_export(cls, source: OMap, target: OMap):
        h = C.c_void_p()
        _check(lib().orc_export_create(source.h, target.h, C.byref(h)))
        return cls(source.comm, h)

    def field(self, pid, name):
        f = self.FIELDS[name]
        n = lib().orc_plan_field_len(self.h, pid, f)
        out = np.empty(n, dtype=np.int64)
        lib().orc_plan_field_get(self.h, pid, f, _p(out, C.c_int64))
        return int(out[0]) if name in ("numSameIDs", "isLocallyComplete") else out

    @property
    def distributor(self):
        return ODist(self.comm, lib().orc_plan_distributor(self.h))


class OMV:
    def __init__(self, map_: OMap, k: int, handle=None):
        self.map, self.k = map_, k
        if handle is None:
            h = C.c_void_p()
            _check(lib().orc_mv_create(map_.h, k, C.byref(h)))
            handle = h
        self.h = handle

    @classmethod
    def from_arrays(cls, map_: OMap, arrays_per_rank):
        k = np.asarray(arrays_per_rank[0]).reshape(len(arrays_per_rank[0]), -1).shape[1] if len(arrays_per_rank[0]) else 1
        mv = cls(map_, k)
        for pid, a in enumerate(arrays_per_rank, start=1):
            mv.set(pid, a)
        return mv

    def local_length(self, pid):
        return int(lib().orc_mv_local_length(self.h, pid))

    def set(self, pid, arr):
        n = self.local_length(pid)
        a = np.asfortranarray(np.asarray(arr, dtype=np.float64).reshape(n, self.k))
        lib().orc_mv_set(self.h, pid, a.ctypes.data_as(C.POINTER(C.c_double)))

    def get(self, pid):
        n = self.local_length(pid)
        out = np.empty((n, self.k), dtype=np.float64, order="F")
        lib().orc_mv_get(self.h, pid, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def fill(self, v):
        lib().orc_mv_fill(self.h, C.c_double(v))

    def scale(self, alpha):
        if np.isscalar(alpha):
            lib().orc_mv_scale(self.h, C.c_double(alpha))
        else:
            a = np.ascontiguousarray(alpha, dtype=np.float64)
            lib().orc_mv_scale_vec(self.h, _p(a, C.c_double))

    def axpby(self, a, x: "OMV", b):
        lib().orc_mv_axpby(self.h, C.c_double(a), x.h, C.c_double(b))

    def dot(self, other, partials=False):
        out = np.empty(self.k, dtype=np.float64)
        part = np.empty((self.map.comm.P, self.k), dtype=np.float64)
        _check(lib().orc_mv_dot(self.h, other.h, _p(out, C.c_double), _p(part, C.c_double)))
        return (out, part) if partials else out

    def norm(self, n=2, partials=False):
        out = np.empty(self.k, dtype=np.float64)
        part = np.empty((self.map.comm.P, self.k), dtype=np.float64)
        _check(lib().orc_mv_norm(self.h, C.c_double(n), _p(out, C.c_double), _p(part, C.c_double)))
        return (out, part) if partials else out

    def comm_reduce(self):
        _check(lib().orc_mv_comm_reduce(self.h))


def do_transfer(src: OMV, tgt: OMV, plan: OPlan, reversed_=False, cm=INSERT):
    _check(lib().orc_do_transfer(src.h, tgt.h, plan.h, int(reversed_), cm))


class OMatrix:
    def __init__(self, row_map: OMap, nnz_per_row_per_rank):
        self.row_map = row_map
        self.comm = row_map.comm
        flat, offs = _flat(nnz_per_row_per_rank, np.int32)
        h = C.c_void_p()
        _check(lib().orc_matrix_create(row_map.h, _p(flat, C.c_int32), _p(offs, C.c_int64), C.byref(h)))
        self.h = h

    def insert(self, pid, grow, cols, vals):
        c = _i64(cols)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        _check(lib().orc_matrix_insert(self.h, pid, C.c_int64(grow), _p(c, C.c_int64), _p(v, C.c_double), C.c_int64(len(c))))

    def insert_rows(self, pid, rows, rowptr0, cols, vals):
        r, rp, c = _i64(rows), _i64(rowptr0), _i64(cols)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        _check(lib().orc_matrix_insert_rows(self.h, pid, C.c_int64(len(r)), _p(r, C.c_int64), _p(rp, C.c_int64), _p(c, C.c_int64), _p(v, C.c_double)))

    def fill_complete(self, domain_map: OMap = None, range_map: OMap = None):
        self.domain_map = domain_map or self.row_map
        self.range_map = range_map or self.row_map
        _check(lib().orc_matrix_fill_complete(self.h, self.domain_map.h, self.range_map.h))

    @property
    def col_map(self):
        return OMap(self.comm, lib().orc_matrix_col_map(self.h), owned=False)

    @property
    def importer(self):
        h = lib().orc_matrix_importer(self.h)
        return OPlan(self.comm, h, owned=False) if h else None

    @property
    def exporter(self):
        h = lib().orc_matrix_exporter(self.h)
        return OPlan(self.comm, h, owned=False) if h else None

    def csr(self, pid):
        n = int(lib().orc_matrix_local_rows(self.h, pid))
        nnz = int(lib().orc_matrix_local_nnz(self.h, pid))
        rp = np.empty(n + 1, dtype=np.int32)
        ci = np.empty(nnz, dtype=np.int32)
        va = np.empty(nnz, dtype=np.float64)
        lib().orc_matrix_get_csr(self.h, pid, _p(rp, C.c_int32), _p(ci, C.c_int32), _p(va, C.c_double))
        return rp, ci, va

    def apply(self, Y: OMV, X: OMV, mode=NO_TRANS, alpha=1.0, beta=0.0):
        _check(lib().orc_apply(Y.h, self.h, X.h, mode, C.c_double(alpha), C.c_double(beta)))
        return Y

    def time_apply(self, Y: OMV, X: OMV, alpha=1.0, beta=0.0, iters=3, warmup=1):
        s = C.c_double(0)
        _check(lib().orc_time_apply(Y.h, self.h, X.h, C.c_double(alpha), C.c_double(beta), iters, warmup, C.byref(s)))
        return s.value


def local_apply_raw(y, rowptr1, colind1, vals, x, mode=NO_TRANS, alpha=1.0, beta=0.0):
    """localApply (src/CSRMatrix.jl:1123-1162) on raw column-major arrays of one rank; y is updated in place."""
    y = np.asfortranarray(y, dtype=np.float64)
    x = np.asfortranarray(x, dtype=np.float64)
    if y.ndim == 1:
        y = y.reshape(-1, 1, order="F")
    if x.ndim == 1:
        x = x.reshape(-1, 1, order="F")
    rp = np.ascontiguousarray(rowptr1, dtype=np.int32)
    ci = np.ascontiguousarray(colind1, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float64)
    lib().orc_local_apply_raw(y.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(y.shape[0]), C.c_int32(len(rp) - 1),
                              _p(rp, C.c_int32), _p(ci, C.c_int32), _p(va, C.c_double),
                              x.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(x.shape[0]), y.shape[1], mode,
                              C.c_double(alpha), C.c_double(beta))
    return y


def dot_raw(a, b):
    a = np.asfortranarray(a, dtype=np.float64).reshape(len(a), -1, order="F")
    b = np.asfortranarray(b, dtype=np.float64).reshape(len(b), -1, order="F")
    out = np.empty(a.shape[1], dtype=np.float64)
    lib().orc_dot_raw(a.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(a.shape[0]), a.shape[1], _p(out, C.c_double))
    return out


def norm_raw(a, p=2.0):
    """local sum of x^p (before the allreduce and the root), src/MultiVector.jl:122-142"""
    a = np.asfortranarray(a, dtype=np.float64).reshape(len(a), -1, order="F")
    out = np.empty(a.shape[1], dtype=np.float64)
    lib().orc_norm_raw(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(a.shape[0]), a.shape[1], C.c_double(p), _p(out, C.c_double))
    return out


def time_local_apply_raw(rowptr1, colind1, vals, x, alpha=1.0, beta=0.0, iters=3, warmup=1):
    """seconds per localApply on ONE core (the SerialComm configuration), timing only the reference loop"""
    x = np.asfortranarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1, order="F")
    rp = np.ascontiguousarray(rowptr1, dtype=np.int32)
    ci = np.ascontiguousarray(colind1, dtype=np.int32)
    va = np.ascontiguousarray(vals, dtype=np.float64)
    y = np.zeros((len(rp) - 1, x.shape[1]), dtype=np.float64, order="F")
    f = lib().orc_time_local_apply_raw
    f.restype = C.c_double
    return float(f(y.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(y.shape[0]), C.c_int32(len(rp) - 1), _p(rp, C.c_int32),
                   _p(ci, C.c_int32), _p(va, C.c_double), x.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(x.shape[0]), x.shape[1],
                   C.c_double(alpha), C.c_double(beta), iters, warmup))
