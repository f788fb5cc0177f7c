"""DenseMultiVector with device-resident storage (src/DenseMultiVector.jl:6-60), the MultiVector operations of
src/MultiVector.jl:52-161 and the DistObject transfer entry points doImport / doExport
(src/DistObject.jl:86-164).  Every operation runs in libjpetra_b200.so on the GPU; `.data` is a host copy.

Naming: Julia's `f!` is `f_` here (`fill_`, `scale_`, `apply_`, `mul_`).
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _lib
from .blockmap import BlockMap
from .errors import InvalidArgumentError
from .importexport import Export, Import


class CombineMode(enum.IntEnum):  # src/DistObject.jl:21
    ADD = 1
    INSERT = 2
    REPLACE = 3
    ABSMAX = 4
    ZERO = 5


ADD, INSERT, REPLACE, ABSMAX, ZERO = CombineMode.ADD, CombineMode.INSERT, CombineMode.REPLACE, CombineMode.ABSMAX, CombineMode.ZERO


class DenseMultiVector:
    """DenseMultiVector(map, numVecs[, zeroOut])   src/DenseMultiVector.jl:20-28
       DenseMultiVector(map, data::Matrix)         src/DenseMultiVector.jl:35-41 (the array is copied to the device)"""

    def __init__(self, map_: BlockMap, data_or_numVecs, zeroOut: bool = True):
        self.map = map_
        n = map_.numMyElements()
        if np.ndim(data_or_numVecs) == 0:
            k = int(data_or_numVecs)
            host = None
        else:
            host = np.asarray(data_or_numVecs, dtype=np.float64)
            if host.ndim == 1:
                host = host.reshape(-1, 1)
            if host.shape[0] != n:
                raise InvalidArgumentError("Length of vectors does not match local length indicated by map")
            k = host.shape[1]
        self.numVectors_ = k
        h = C.c_int64()
        _lib.check(_lib.lib().jp_mv_create(n, k, C.byref(h)))  # zero-filled
        self.h = h.value
        if host is not None:
            self.data = host

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.load().jp_mv_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- required MultiVector interface (src/MultiVector.jl:15-28) ----
    def getMap(self):
        return self.map

    def numVectors(self):
        return self.numVectors_

    def localLength(self):  # src/MultiVector.jl:45
        return self.map.numMyElements()

    def globalLength(self):
        return self.map.numGlobalElements()

    def getComm(self):
        return self.map.getComm()

    def getLocalArray(self) -> np.ndarray:
        """host copy of the local array, column-major localLength x numVectors"""
        out = np.empty((self.localLength(), self.numVectors_), dtype=np.float64, order="F")
        _lib.check(_lib.lib().jp_mv_download(self.h, _lib.ptr(out)))
        return out

    @property
    def data(self) -> np.ndarray:
        return self.getLocalArray()

    @data.setter
    def data(self, host):
        a = np.asfortranarray(np.asarray(host, dtype=np.float64).reshape(self.localLength(), self.numVectors_))
        _lib.check(_lib.lib().jp_mv_upload(self.h, _lib.ptr(a)))

    def similar(self):
        return DenseMultiVector(self.map, self.numVectors_)

    def copy(self):  # src/DenseMultiVector.jl:50-52
        out = DenseMultiVector(self.map, self.numVectors_)
        _lib.check(_lib.lib().jp_mv_copy(out.h, self.h))
        return out

    def copyto_(self, src: "DenseMultiVector"):  # copyto!(dest, src), src/DenseMultiVector.jl:54-60
        _lib.check(_lib.lib().jp_mv_copy(self.h, src.h))
        self.map = src.map
        return self

    def getVectorCopy(self, i: int) -> np.ndarray:  # 1-based column
        return self.getLocalArray()[:, i - 1].copy()

    getVectorView = getVectorCopy

    # ---- src/MultiVector.jl:52-77 ----
    def fill_(self, value):
        _lib.check(_lib.lib().jp_mv_fill(self.h, float(value)))
        return self

    def scale_(self, alpha):
        if np.ndim(alpha) == 0:
            _lib.check(_lib.lib().jp_mv_scale(self.h, float(alpha)))
        else:
            a = _lib.as_f64(alpha)
            if len(a) != self.numVectors_:
                raise InvalidArgumentError("scale!: need one factor per vector")
            _lib.check(_lib.lib().jp_mv_scale_vec(self.h, _lib.ptr(a)))
        return self

    def axpby_(self, a, x: "DenseMultiVector", b):
        """self = a*x + b*self, what `@. y = a*x + b*y` lowers to (broadcast copyto!, src/MultiVector.jl:497-502)"""
        _lib.check(_lib.lib().jp_mv_axpby(self.h, float(a), x.h, float(b)))
        return self

    def axpby_norm2_(self, a, x: "DenseMultiVector", b) -> np.ndarray:
        """self = a*x + b*self and norm(self, 2) of the result in one pass over the data (jp_mv_axpby_norm2): the
        residual update of a CG iteration fused with its norm.  Returns what `norm(self, 2)` would."""
        out = np.empty((1, self.numVectors_), dtype=np.float64)
        _lib.check(_lib.lib().jp_mv_axpby_norm2(self.getComm().handle, self.h, float(a), x.h, float(b), _lib.ptr(out)))
        return out

    # ---- src/MultiVector.jl:84-161 ----
    def dot(self, other: "DenseMultiVector") -> np.ndarray:
        out = np.empty((1, self.numVectors_), dtype=np.float64)
        _lib.check(_lib.lib().jp_mv_dot(self.getComm().handle, self.h, other.h, _lib.ptr(out)))
        return out

    def norm(self, n=2) -> np.ndarray:
        out = np.empty((1, self.numVectors_), dtype=np.float64)
        if n == 2:
            _lib.check(_lib.lib().jp_mv_norm2(self.getComm().handle, self.h, _lib.ptr(out)))
        else:
            _lib.check(_lib.lib().jp_mv_normp(self.getComm().handle, self.h, float(n), _lib.ptr(out)))
        return out

    def commReduce(self):
        if self.map.distributedGlobal():
            raise InvalidArgumentError("Cannot reduce distributed MultiVector")
        _lib.check(_lib.lib().jp_mv_comm_reduce(self.getComm().handle, self.h))
        return self

    def __eq__(self, other):  # src/MultiVector.jl:463-469
        return (isinstance(other, DenseMultiVector) and self.map.sameAs(other.map)
                and np.array_equal(self.getLocalArray(), other.getLocalArray()))

    __hash__ = None


def dot(x: DenseMultiVector, y: DenseMultiVector):
    return x.dot(y)


def norm(x: DenseMultiVector, n=2):
    return x.norm(n)


def _transfer(source, target, plan, cm, reverse):
    _lib.check(_lib.lib().jp_transfer(plan.device_plan(), source.h, target.h, int(cm), int(reverse)))


def doImport(source: DenseMultiVector, target: DenseMultiVector, plan, cm=INSERT):
    """doImport(source, target, importer, cm): forward with an Import (src/DistObject.jl:86-91), reverse with an
    Export (:122-129); `None` is the trivial plan (:93-99)."""
    if plan is None:
        return
    if isinstance(plan, Import):
        _transfer(source, target, plan, cm, False)
    elif isinstance(plan, Export):
        _transfer(source, target, plan, cm, True)
    else:
        raise InvalidArgumentError("doImport needs an Import or Export plan")


def doExport(source: DenseMultiVector, target: DenseMultiVector, plan, cm=INSERT):
    """doExport(source, target, exporter, cm): forward with an Export (src/DistObject.jl:104-113), reverse with an
    Import (:147-157)."""
    if plan is None:
        return
    if isinstance(plan, Export):
        _transfer(source, target, plan, cm, False)
    elif isinstance(plan, Import):
        _transfer(source, target, plan, cm, True)
    else:
        raise InvalidArgumentError("doExport needs an Import or Export plan")
