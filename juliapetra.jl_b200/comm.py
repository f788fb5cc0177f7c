"""Communication layer: the abstract `Comm` of the reference (src/Comm.jl:12-57) and its backends.

* `SerialComm`  -- src/SerialComm.jl:76-133: every collective is the identity.
* `NCCLComm`    -- the replacement of `MPIComm` (src/MPIComm.jl:1-91): one rank per GPU, collectives through
                   the C ABI (NCCL over NVLink).  PIDs are 1-based (src/MPIComm.jl:81-83).
* `TorchDistComm` -- the same interface over an initialised `torch.distributed` process group; used for host-side
                   setup collectives in world_size>1 CPU tests (gloo) where no GPU exists.  It has no device side.

Collectives take and return 1-D numpy arrays (float64 or int64); the scalar convenience forms of
src/Comm.jl:72-119 are the `*Scalar` helpers on the base class.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .errors import InvalidArgumentError


def _canon(vals):
    a = np.atleast_1d(np.asarray(vals))
    if np.issubdtype(a.dtype, np.floating):
        return np.ascontiguousarray(a, dtype=np.float64), _lib.F64
    if np.issubdtype(a.dtype, np.bool_):
        return np.ascontiguousarray(a, dtype=np.int64), _lib.I64
    if np.issubdtype(a.dtype, np.integer):
        return np.ascontiguousarray(a, dtype=np.int64), _lib.I64
    raise InvalidArgumentError(f"collectives support float64 / int64 arrays, got {a.dtype}")


class Comm:
    """Abstract base, src/Comm.jl:56.  Required methods: barrier, broadcastAll, gatherAll, sumAll, maxAll, minAll,
    scanSum, myPid, numProc, createDistributor (src/Comm.jl:15-53)."""

    def barrier(self):
        raise NotImplementedError

    def broadcastAll(self, myvals, root: int):
        raise NotImplementedError

    def gatherAll(self, myVals):
        raise NotImplementedError

    def sumAll(self, partialsums):
        raise NotImplementedError

    def maxAll(self, partialmaxes):
        raise NotImplementedError

    def minAll(self, partialmins):
        raise NotImplementedError

    def scanSum(self, myvals):
        raise NotImplementedError

    def myPid(self) -> int:
        raise NotImplementedError

    def numProc(self) -> int:
        raise NotImplementedError

    def createDistributor(self):
        from .distributor import Distributor
        return Distributor(self)

    # setup-time record exchange used by Distributor.resolve on host arrays (the MPI point-to-point traffic of
    # src/MPIDistributor.jl:487-561, expressed as one all-to-all-v)
    def alltoallv(self, send: np.ndarray, sendcounts, recvcounts) -> np.ndarray:
        raise NotImplementedError

    # device side: handle of the jp comm object (fails loudly without a GPU)
    @property
    def handle(self) -> int:
        raise NotImplementedError

    # scalar convenience wrappers, src/Comm.jl:72-119
    def sumAllScalar(self, v):
        return self.sumAll([v])[0].item()

    def maxAllScalar(self, v):
        return self.maxAll([v])[0].item()

    def minAllScalar(self, v):
        return self.minAll([v])[0].item()

    def scanSumScalar(self, v):
        return self.scanSum([v])[0].item()

    def __repr__(self):  # src/Comm.jl:59-62
        return f"{type(self).__name__} with PID {self.myPid()} and {self.numProc()} processes"


class _DeviceCommMixin:
    _h = None

    def _create_handle(self) -> int:
        raise NotImplementedError

    @property
    def handle(self) -> int:
        if self._h is None:
            self._h = self._create_handle()
        return self._h


class SerialComm(_DeviceCommMixin, Comm):
    """src/SerialComm.jl:76-133"""

    def _create_handle(self):
        h = C.c_int64()
        _lib.check(_lib.lib().jp_comm_create_serial(C.byref(h)))
        return h.value

    def barrier(self):
        pass

    def broadcastAll(self, myvals, root):
        if root != 1:
            raise InvalidArgumentError("SerialComm can only accept PID of 1")
        return _canon(myvals)[0]

    def gatherAll(self, myVals):
        return _canon(myVals)[0]

    def sumAll(self, v):
        return _canon(v)[0]

    def maxAll(self, v):
        return _canon(v)[0]

    def minAll(self, v):
        return _canon(v)[0]

    def scanSum(self, v):
        return _canon(v)[0]

    def myPid(self):
        return 1

    def numProc(self):
        return 1

    def alltoallv(self, send, sendcounts, recvcounts):
        return np.array(send, copy=True)

    def createDistributor(self):  # src/SerialComm.jl:129-131
        from .distributor import SerialDistributor
        return SerialDistributor(self)

    def __eq__(self, other):
        return isinstance(other, SerialComm)

    def __hash__(self):
        return hash("SerialComm")


class NCCLComm(_DeviceCommMixin, Comm):
    """One rank per GPU; every collective goes through libjpetra_b200.so (NCCL).  `unique_id` is the 128-byte id
    made by rank 0 (`NCCLComm.uniqueId()`) and distributed by the host bootstrap (torch.distributed, a file, ...)."""

    def __init__(self, unique_id: bytes, nranks: int, rank0: int):
        if len(unique_id) != 128:
            raise InvalidArgumentError("unique_id must be 128 bytes")
        self._uid, self._n, self._r = bytes(unique_id), int(nranks), int(rank0)
        _ = self.handle  # communicator creation is itself collective: do it eagerly, in order

    @staticmethod
    def uniqueId() -> bytes:
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.lib().jp_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def fromTorchDistributed(cls):
        """Bootstrap from an initialised torch.distributed group (plumbing only): rank 0's id is broadcast."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        _lib.init()
        box = [cls.uniqueId() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], world, rank)

    def _create_handle(self):
        h = C.c_int64()
        buf = (C.c_uint8 * 128).from_buffer_copy(self._uid)
        _lib.check(_lib.lib().jp_comm_create_nccl(buf, self._n, self._r, C.byref(h)))
        return h.value

    def myPid(self):
        return self._r + 1

    def numProc(self):
        return self._n

    def barrier(self):
        _lib.check(_lib.lib().jp_comm_barrier(self.handle))

    def _allreduce(self, op, vals):
        a, dt = _canon(vals)
        out = np.empty_like(a)
        _lib.check(_lib.lib().jp_comm_allreduce(self.handle, op, dt, _lib.ptr(a), _lib.ptr(out), a.size))
        return out

    def sumAll(self, v):
        return self._allreduce(_lib.SUM, v)

    def maxAll(self, v):
        return self._allreduce(_lib.MAX, v)

    def minAll(self, v):
        return self._allreduce(_lib.MIN, v)

    def scanSum(self, v):
        a, dt = _canon(v)
        out = np.empty_like(a)
        _lib.check(_lib.lib().jp_comm_scan_sum(self.handle, dt, _lib.ptr(a), _lib.ptr(out), a.size))
        return out

    def gatherAll(self, myVals):
        a, dt = _canon(myVals)
        counts = np.zeros(self._n, dtype=np.int64)
        one = np.array([a.size], dtype=np.int64)
        ones = np.ones(self._n, dtype=np.int64)
        _lib.check(_lib.lib().jp_comm_gather_all(self.handle, _lib.I64, _lib.ptr(one), 1, _lib.ptr(ones), _lib.ptr(counts)))
        out = np.empty(int(counts.sum()), dtype=a.dtype)
        _lib.check(_lib.lib().jp_comm_gather_all(self.handle, dt, _lib.ptr(a), a.size, _lib.ptr(counts), _lib.ptr(out)))
        return out

    def broadcastAll(self, myvals, root):
        a, dt = _canon(myvals)
        out = a.copy()
        _lib.check(_lib.lib().jp_comm_broadcast(self.handle, dt, _lib.ptr(out), out.size, int(root)))
        return out

    def alltoallv(self, send, sendcounts, recvcounts):
        send = np.ascontiguousarray(send)
        elem = send.dtype.itemsize * (int(np.prod(send.shape[1:])) if send.ndim > 1 else 1)
        sc, rc = _lib.as_i64(sendcounts), _lib.as_i64(recvcounts)
        recv = np.empty((int(rc.sum()),) + send.shape[1:], dtype=send.dtype)
        _lib.check(_lib.lib().jp_comm_alltoallv(self.handle, elem, _lib.ptr(send), _lib.ptr(sc), _lib.ptr(recv), _lib.ptr(rc)))
        return recv


class TorchDistComm(Comm):
    """Host-only Comm over torch.distributed (gloo in the CPU tests).  Same semantics as MPIComm
    (src/MPIComm.jl:42-91); no device handle."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self._group = dist, group
        self._r, self._n = dist.get_rank(group), dist.get_world_size(group)

    @property
    def handle(self):
        raise InvalidArgumentError("TorchDistComm has no device side; use NCCLComm on a GPU box")

    def myPid(self):
        return self._r + 1

    def numProc(self):
        return self._n

    def barrier(self):
        self._dist.barrier(self._group)

    def _reduce(self, op, vals):
        import torch
        a, _ = _canon(vals)
        t = torch.from_numpy(a.copy())
        self._dist.all_reduce(t, op=op, group=self._group)
        return t.numpy()

    def sumAll(self, v):
        return self._reduce(self._dist.ReduceOp.SUM, v)

    def maxAll(self, v):
        return self._reduce(self._dist.ReduceOp.MAX, v)

    def minAll(self, v):
        return self._reduce(self._dist.ReduceOp.MIN, v)

    def _gather_objects(self, obj):
        out = [None] * self._n
        self._dist.all_gather_object(out, obj, group=self._group)
        return out

    def scanSum(self, v):
        a, _ = _canon(v)
        parts = self._gather_objects(a)
        acc = parts[0].copy()
        for p in range(1, self._r + 1):
            acc = acc + parts[p]
        return acc

    def gatherAll(self, myVals):
        a, _ = _canon(myVals)
        return np.concatenate(self._gather_objects(a))

    def broadcastAll(self, myvals, root):
        if not 1 <= root <= self._n:
            raise InvalidArgumentError("broadcastAll: root PID out of range")
        a, _ = _canon(myvals)
        box = [a if self._r == root - 1 else None]
        self._dist.broadcast_object_list(box, src=root - 1, group=self._group)
        return np.asarray(box[0])

    def alltoallv(self, send, sendcounts, recvcounts):
        send = np.ascontiguousarray(send)
        offs = np.concatenate([[0], np.cumsum(sendcounts)]).astype(np.int64)
        mine = [send[offs[p]:offs[p + 1]] for p in range(self._n)]
        everyone = self._gather_objects(mine)
        parts = [everyone[q][self._r] for q in range(self._n)]
        for q in range(self._n):
            assert len(parts[q]) == int(recvcounts[q]), "alltoallv: recvcounts do not match what the peers sent"
        return np.concatenate(parts) if parts else send[:0]
