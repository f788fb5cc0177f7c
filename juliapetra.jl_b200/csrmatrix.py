"""CSRMatrix -- host mirror of the reference's operator (src/CSRMatrix.jl, src/RowMatrix.jl, src/Operator.jl).

Host side (numpy, one-time): STATIC_PROFILE storage with global column indices, `insertGlobalValues`,
`fillComplete` = column-map build (src/CSRGraphInternalMethods.jl:852-982), GID->LID remap (:636-712), per-row
stable sort + duplicate merge (src/CSRMatrix.jl:549-606), Import/Export construction (:351-366) and packing to
CSR (src/CSRMatrix.jl:403-484).  The column map, local indices and Import lists are bit-exact parity targets.
Device side (every `apply_`): the packed CSR lives in HBM (jp_csr_create) and `apply_` is one C-ABI call
(jp_apply) that overlaps the halo exchange with the interior rows.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
import os

import numpy as np

from . import _lib
from .blockmap import BlockMap
from .errors import BackendError, InvalidArgumentError, InvalidStateError
from .importexport import Export, Import
from .multivector import DenseMultiVector
from .utils import computeOffsets


# ---- the four C-ABI calls of the device fillComplete, as module-level functions so that the CPU tests can run the
#      multi-rank glue around them against the CPU execution model of the kernels (tests/dist_cases.py) ----
def _device_fill_complete(n_rows, ptr0, gcols, vals, gmin, gmax, dom_min, n_dom):
    """jp_csr_fill_complete -> (csr handle, [n_cols, n_local_used, n_same, nnz]); InvalidArgumentError for a bad GID"""
    h, info = C.c_int64(), np.zeros(4, dtype=np.int64)
    _lib.check(_lib.lib().jp_csr_fill_complete(n_rows, _lib.ptr(ptr0), _lib.ptr(gcols), _lib.ptr(vals), gmin, gmax, dom_min, n_dom,
                                               C.byref(h), _lib.ptr(info)))
    return h.value, [int(v) for v in info]


def _device_col_map(h, n_cols):
    gids = np.empty(n_cols, dtype=np.int64)
    _lib.check(_lib.lib().jp_csr_col_map(h, _lib.ptr(gids)))
    return gids


def _device_csr_download(h, n_rows, nnz):
    rp = np.empty(n_rows + 1, dtype=np.int32)
    ci = np.empty(nnz, dtype=np.int32)
    va = np.empty(nnz, dtype=np.float64)
    _lib.check(_lib.lib().jp_csr_download(h, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(va)))
    return rp, ci, va


def _device_csr_destroy(h):
    _lib.lib().jp_csr_destroy(h)


class TransposeMode(enum.IntEnum):  # src/Operator.jl:10
    NO_TRANS = 1
    TRANS = 2
    CONJ_TRANS = 3


NO_TRANS, TRANS, CONJ_TRANS = TransposeMode.NO_TRANS, TransposeMode.TRANS, TransposeMode.CONJ_TRANS
STATIC_PROFILE = "STATIC_PROFILE"  # src/CSRGraphConstructors.jl:10-21 (the only profile on this path)


def isTransposed(mode) -> bool:  # src/Operator.jl:17
    return int(mode) != NO_TRANS


class CSRMatrix:
    """CSRMatrix{Data}(rowMap, numEntPerRow, STATIC_PROFILE) -- src/CSRMatrix.jl:82-135; numEntPerRow is a scalar
    maximum or one count per local row."""

    def __init__(self, rowMap: BlockMap, numEntPerRow, pftype=STATIC_PROFILE):
        if pftype != STATIC_PROFILE:
            raise InvalidArgumentError("only STATIC_PROFILE storage is on the hot path")
        self.rowMap = rowMap
        self.comm = rowMap.getComm()
        n = rowMap.numMyElements()
        if np.ndim(numEntPerRow) == 0:
            counts = np.full(n, int(numEntPerRow), dtype=np.int64)
        else:
            counts = np.asarray(numEntPerRow, dtype=np.int64).ravel()
            if len(counts) != n:
                raise InvalidArgumentError("numEntPerRow must have one entry per local row")
        # allocateIndices -> computeOffsets (src/CSRGraphInternalMethods.jl:302-348, src/Utils.jl:12-26)
        self._rowOffsets = np.empty(n + 1, dtype=np.int64)
        total = computeOffsets(self._rowOffsets, counts) if n else 0
        if n == 0:
            self._rowOffsets[:] = 1
        self._globalIndices1D = np.zeros(total, dtype=np.int64)
        self._values = np.zeros(total, dtype=np.float64)
        self._numRowEntries = np.zeros(n, dtype=np.int64)
        self.domainMap = self.rangeMap = self.colMap = None
        self.importer = self.exporter = None
        self._fillComplete = False
        self.h = None  # device CSR
        # packed local CSR after fillComplete (1-based, like LocalCSRGraph/LocalCSRMatrix); after a device
        # fillComplete the host copy is downloaded on first use
        self._csr_host = None
        self._nnz = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.load().jp_csr_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- fill ----
    def insertGlobalValues(self, globalRow: int, indices, values):
        """src/CSRMatrix.jl:617-665"""
        self.insertGlobalValuesBatch(np.array([globalRow], dtype=np.int64), np.array([0, len(indices)], dtype=np.int64), indices, values)

    def insertGlobalValuesBatch(self, globalRows, rowptr0, indices, values):
        """The same insert for many distinct owned rows at once: row globalRows[i] receives
        indices/values[rowptr0[i]:rowptr0[i+1]] (appended in order, like repeated insertGlobalValues calls)."""
        if self._fillComplete:
            raise InvalidStateError("Cannot insert into a fill complete matrix (resumeFill is out of scope)")
        rows = np.asarray(globalRows, dtype=np.int64).ravel()
        rp = np.asarray(rowptr0, dtype=np.int64).ravel()
        idx = np.asarray(indices, dtype=np.int64).ravel()
        val = np.asarray(values, dtype=np.float64).ravel()
        if len(rp) != len(rows) + 1 or rp[-1] != len(idx) or len(idx) != len(val):
            raise InvalidArgumentError("insertGlobalValues: inconsistent row pointer / index / value lengths")
        lr = self.rowMap.lid(rows)
        if (lr == 0).any():
            # insertNonownedGlobalValues + globalAssemble are non-functional in the reference (src/CSRMatrix.jl:292-397)
            raise InvalidStateError("insertGlobalValues on a non-owned row: non-local assembly is out of scope")
        if len(np.unique(lr)) != len(lr):
            raise InvalidArgumentError("insertGlobalValuesBatch: every row may appear once per call")
        cnt = np.diff(rp)
        cur = self._numRowEntries[lr - 1]
        alloc = self._rowOffsets[lr] - self._rowOffsets[lr - 1]
        if (cur + cnt > alloc).any():
            raise InvalidArgumentError("number of new indices exceed statically allocated graph structure")
        base = self._rowOffsets[lr - 1] - 1 + cur  # 0-based slot of the first new entry of each row
        dest = np.repeat(base - rp[:-1], cnt) + np.arange(len(idx), dtype=np.int64)
        self._globalIndices1D[dest] = idx
        self._values[dest] = val
        self._numRowEntries[lr - 1] = cur + cnt

    # ---- state ----
    def isFillComplete(self):
        return self._fillComplete

    def isFillActive(self):
        return not self._fillComplete

    def getRowMap(self):
        return self.rowMap

    def getColMap(self):
        return self.colMap

    def getDomainMap(self):
        return self.domainMap

    def getRangeMap(self):
        return self.rangeMap

    def getImporter(self):  # src/CSRGraphExternalMethods.jl:11
        return self.importer

    def getExporter(self):  # src/CSRGraphExternalMethods.jl:12
        return self.exporter

    def getLocalNumRows(self):
        return self.rowMap.numMyElements()

    def getLocalNumEntries(self):
        return int(self._numRowEntries.sum()) if not self._fillComplete else self._nnz

    def _host_csr(self):
        if self._csr_host is None and self._fillComplete and self.h:
            rp, ci, va = _device_csr_download(self.h, self.getLocalNumRows(), self._nnz)
            self._csr_host = (rp + 1, ci + 1, va)
        return self._csr_host

    @property
    def rowOffsets(self):
        c = self._host_csr()
        return None if c is None else c[0]

    @property
    def localIndices1D(self):
        c = self._host_csr()
        return None if c is None else c[1]

    @property
    def values(self):
        c = self._host_csr()
        return None if c is None else c[2]

    def getLocalRowCopy(self, localRow: int):
        s, e = self.rowOffsets[localRow - 1] - 1, self.rowOffsets[localRow] - 1
        return self.localIndices1D[s:e].copy(), self.values[s:e].copy()

    # ---- fillComplete ----
    def fillComplete(self, domainMap: BlockMap | None = None, rangeMap: BlockMap | None = None, device: bool | None = None, **plist):
        """fillComplete(matrix[, domainMap, rangeMap]) -- src/CSRMatrix.jl:706-747.
        device=False: column map / remap / sort+merge / pack in numpy on the host, then the upload (jp_csr_create).
        device=True (or JP_FILL=device): the entries are uploaded and the same steps run as CUDA kernels
        (jp_csr_fill_complete); needs a contiguous domain map, otherwise the host path is taken.  Both give
        bit-identical column maps, local indices and values."""
        if device is None:
            device = os.environ.get("JP_FILL", "host") == "device"
        dom = domainMap if domainMap is not None else self.rowMap
        if device and self.colMap is None and getattr(dom, "contiguousByPid_", False) and self._fillCompleteDevice(domainMap, rangeMap):
            return self
        self._fillCompleteHost(domainMap, rangeMap)
        self._upload()
        return self

    def _fillCompleteDevice(self, domainMap=None, rangeMap=None) -> bool:
        """Device fillComplete; False when the serial colMap == domMap special case of the reference
        (src/CSRGraphInternalMethods.jl:918-925) applies to a non-square matrix, which the host path handles."""
        if self._fillComplete:
            raise InvalidStateError("Matrix cannot be fill complete when fillComplete(...) is called")
        dom = domainMap if domainMap is not None else self.rowMap
        comm = self.comm
        ptr0, gcols, vals = self._packed_global()
        ptr0, gcols, vals = _lib.as_i64(ptr0), _lib.as_i64(gcols), _lib.as_f64(vals)
        n_rows, n_dom = self.getLocalNumRows(), dom.numMyElements()
        h, info = 0, [0, 0, 0, 0]
        error = False
        try:
            h, info = _device_fill_complete(n_rows, ptr0, gcols, vals, dom.minAllGID(), dom.maxAllGID(),
                                            dom.minMyGID() if n_dom else dom.minAllGID(), n_dom)
        except InvalidArgumentError:
            error = True  # a column GID nobody owns: remotePIDs == 0 in the reference (:936-939)
        n_cols, n_local_used, n_same, nnz = info
        serial = comm.numProc() == 1
        if not error and serial and n_cols > n_local_used:
            error = True  # remote GIDs on one process, :918-921
        if bool(comm.maxAllScalar(int(error))):
            if h:
                _device_csr_destroy(h)
            raise RuntimeError("makeColMap reports an error on at least one process")
        if serial and n_local_used == n_rows and n_local_used != n_dom:
            _device_csr_destroy(h)
            return False
        self.domainMap = dom
        self.rangeMap = rangeMap if rangeMap is not None else self.rowMap
        if serial and n_local_used == n_rows:
            self.colMap = dom  # :922-924
        else:
            self.colMap = BlockMap(dom.numGlobalElements(), _device_col_map(h, n_cols), comm)  # :981
        if not self.domainMap.sameAs(self.colMap):
            self.importer = Import(self.domainMap, self.colMap)
        if not self.rangeMap.sameAs(self.rowMap):
            self.exporter = Export(self.rowMap, self.rangeMap)
        expect_same = self.importer.numSameIDs() if self.importer is not None else n_cols
        if expect_same != n_same:
            _device_csr_destroy(h)
            raise BackendError(f"device fillComplete: numSameIDs {n_same} disagrees with the Import ({expect_same})")
        self.h = h
        self._nnz = nnz
        self._globalIndices1D = self._values = None
        self._fillComplete = True
        if self.importer is not None:
            self.importer.device_plan()
        return True

    def _packed_global(self):
        """entries actually inserted, row by row in insertion order: (ptr0, gcols, vals)"""
        n = self.getLocalNumRows()
        cnt = self._numRowEntries
        ptr0 = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(cnt, out=ptr0[1:])
        if n and np.array_equal(self._rowOffsets[1:] - self._rowOffsets[:-1], cnt):
            return ptr0, self._globalIndices1D, self._values  # every row full: storage is already packed
        src = np.repeat(self._rowOffsets[:-1] - 1 - ptr0[:-1], cnt) + np.arange(int(ptr0[-1]), dtype=np.int64)
        return ptr0, self._globalIndices1D[src], self._values[src]

    def _makeColMap(self, ptr0, gcols):
        """__makeColMap, src/CSRGraphInternalMethods.jl:852-982 (+ the maxAll of the error flag,
        src/CSRGraphExternalMethods.jl:619-636).  Remote GIDs are ordered by owner PID and, inside one owner, by
        ascending GID (the reference iterates an untyped Set there; SURVEY.md 8c deviation 1)."""
        domMap, comm = self.domainMap, self.comm
        localNumRows = self.getLocalNumRows()
        error = False
        lid_dom = domMap.lid(gcols) if len(gcols) else np.zeros(0, dtype=np.int64)
        numDomainElts = domMap.numMyElements()
        gidIsLocal = np.zeros(max(localNumRows, numDomainElts), dtype=bool)
        gidIsLocal[lid_dom[lid_dom != 0] - 1] = True
        numLocalColGIDs = int(gidIsLocal.sum())
        remoteColGIDs = np.unique(gcols[lid_dom == 0])
        colMap = None
        if comm.numProc() == 1:  # serial short circuit, :918-925
            if len(remoteColGIDs) != 0:
                error = True
            if numLocalColGIDs == localNumRows:
                colMap = domMap
        if colMap is None and not error:
            remotePIDs = domMap.remoteIDList(remoteColGIDs)[0] if len(remoteColGIDs) else np.zeros(0, dtype=np.int64)
            if (remotePIDs == 0).any():
                error = True
            order = np.argsort(remotePIDs, kind="stable")  # :940-942
            remoteColGIDs = remoteColGIDs[order]
            if numLocalColGIDs == numDomainElts:
                localColGIDs = domMap.myGlobalElements()
            else:
                localColGIDs = domMap.myGlobalElements()[gidIsLocal[:numDomainElts]]
            myColumns = np.concatenate([localColGIDs, remoteColGIDs]).astype(np.int64)
        if bool(comm.maxAllScalar(int(error))):
            raise RuntimeError("makeColMap reports an error on at least one process")
        if colMap is None:
            colMap = BlockMap(domMap.numGlobalElements(), myColumns, comm)  # :981
        return colMap

    @staticmethod
    def _sortAndMerge(ptr0, lcols, vals):
        """sortAndMergeIndicesAndValues, src/CSRMatrix.jl:549-606: per row, stable sort by local index and sum
        duplicates left to right.  Rows already strictly increasing are left untouched."""
        nnz = len(lcols)
        n = len(ptr0) - 1
        if nnz == 0:
            return ptr0.copy(), lcols, vals
        bad = np.zeros(nnz, dtype=bool)
        bad[1:] = lcols[1:] <= lcols[:-1]
        starts = ptr0[:-1][np.diff(ptr0) > 0]
        bad[starts] = False  # the first entry of a row is never out of order
        if not bad.any():
            return ptr0.copy(), lcols, vals
        rowid = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr0))
        dirty_rows = np.unique(rowid[bad])
        sel = np.nonzero(np.isin(rowid, dirty_rows))[0]  # entries of rows that need work (ascending position)
        key = rowid[sel] * (int(lcols.max()) + 1) + lcols[sel]
        order = np.argsort(key, kind="stable")
        lcols = lcols.copy()
        vals = vals.copy()
        lcols[sel] = lcols[sel][order]
        vals[sel] = vals[sel][order]
        key = key[order]
        dup = np.zeros(len(sel), dtype=bool)
        dup[1:] = key[1:] == key[:-1]
        if not dup.any():
            return ptr0.copy(), lcols, vals
        # merge: every duplicate is added into the first entry of its run, left to right (:588-597)
        head = np.maximum.accumulate(np.where(dup, 0, np.arange(len(sel))))
        v = vals[sel]
        np.add.at(v, head[dup], v[dup])  # in-order, unbuffered: ((v0+v1)+v2)...
        vals[sel] = v
        keep = np.ones(nnz, dtype=bool)
        keep[sel[dup]] = False
        newcnt = np.bincount(rowid[keep], minlength=n).astype(np.int64)
        newptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(newcnt, out=newptr[1:])
        return newptr, lcols[keep], vals[keep]

    def _fillCompleteHost(self, domainMap=None, rangeMap=None):
        if self._fillComplete:
            raise InvalidStateError("Matrix cannot be fill complete when fillComplete(...) is called")
        self.domainMap = domainMap if domainMap is not None else self.rowMap
        self.rangeMap = rangeMap if rangeMap is not None else self.rowMap
        ptr0, gcols, vals = self._packed_global()
        if self.colMap is None:
            self.colMap = self._makeColMap(ptr0, gcols)
        # makeIndicesLocal, src/CSRGraphInternalMethods.jl:636-712
        lcols = self.colMap.lid(gcols) if len(gcols) else np.zeros(0, dtype=np.int64)
        numBad = int((lcols == 0).sum())
        if numBad != 0:
            raise InvalidArgumentError(f"When converting column indices from global to local, we encountered {numBad} "
                                       "indices that do not live in the column map on this process")
        ptr0, lcols, vals = self._sortAndMerge(ptr0, lcols, vals)
        # makeImportExport, src/CSRGraphInternalMethods.jl:351-366
        if not self.domainMap.sameAs(self.colMap):
            self.importer = Import(self.domainMap, self.colMap)
        if not self.rangeMap.sameAs(self.rowMap):
            self.exporter = Export(self.rowMap, self.rangeMap)
        # fillLocalGraphAndMatrix, src/CSRMatrix.jl:403-484: packed 1-based CSR with LID (int32) indices
        if len(lcols) and (len(lcols) >= 2**31 - 64 or self.colMap.numMyElements() >= 2**31 - 1):
            raise InvalidArgumentError("local matrix does not fit Int32 LIDs")
        self._csr_host = ((ptr0 + 1).astype(np.int32), lcols.astype(np.int32), np.ascontiguousarray(vals, dtype=np.float64))
        self._nnz = len(self._csr_host[2])
        self._globalIndices1D = self._values = None
        self._fillComplete = True

    def _upload(self):
        n_cols = self.colMap.numMyElements()
        n_local_cols = self.importer.numSameIDs() if self.importer is not None else n_cols
        h = C.c_int64()
        _lib.check(_lib.lib().jp_csr_create(self.getLocalNumRows(), n_cols, n_local_cols, len(self.values), _lib.ptr(self.rowOffsets),
                                            _lib.ptr(self.localIndices1D), _lib.ptr(self.values), 1, C.byref(h)))
        self.h = h.value
        if self.importer is not None:
            self.importer.device_plan()

    def device_info(self):
        info = np.zeros(8, dtype=np.int64)
        _lib.check(_lib.lib().jp_csr_info(self.h, _lib.ptr(info)))
        return dict(zip(("n_rows", "n_cols", "n_local_cols", "nnz", "interior_blocks", "boundary_blocks", "max_row_len",
                         "threads_per_row"), (int(v) for v in info)))

    def spmm_info(self):
        info = np.zeros(10, dtype=np.int64)
        _lib.check(_lib.lib().jp_csr_spmm_info(self.h, _lib.ptr(info)))
        return dict(zip(("tile_state", "runs_state", "runs_blocks", "runs_widest_tile", "dia_state", "dia_diagonals", "dia_rows", "dia_tile_width",
                         "colblock_state", "colblock_slabs"), (int(v) for v in info)))

    # ---- Operator (src/Operator.jl:35-88, src/RowMatrix.jl:261-357, 406-416) ----
    def apply_(self, Y: DenseMultiVector, X: DenseMultiVector, mode=NO_TRANS, alpha=1.0, beta=0.0):
        """apply!(Y, A, X, mode, alpha, beta): Y = beta*Y + alpha*op(A)*X"""
        if self.isFillActive():
            raise InvalidStateError("Cannot call apply(...) until fillComplete(...)")
        comm = self.comm
        alpha, beta = float(alpha), float(beta)
        YIsReplicated = not Y.getMap().distributedGlobal()
        if alpha != 0.0 and YIsReplicated and comm.myPid() != 1:
            beta = 0.0  # src/RowMatrix.jl:283-287
        plan = self.importer.device_plan() if self.importer is not None else 0
        if self.exporter is not None:
            # rangeMap != rowMap (src/RowMatrix.jl:298-308, 324-330; the reference's lines use undefined names, so the
            # semantics are Tpetra's): the product is formed on the row map and exported into Y with ADD
            _lib.check(_lib.lib().jp_apply_export(comm.handle, Y.h, self.h, X.h, plan, self.exporter.device_plan(), int(mode), alpha, beta))
        else:
            _lib.check(_lib.lib().jp_apply(comm.handle, Y.h, self.h, X.h, plan, int(mode), alpha, beta))
        if alpha != 0.0 and YIsReplicated and comm.numProc() > 1:
            Y.commReduce()  # src/RowMatrix.jl:353-355 (identity on SerialComm)
        return Y

    def apply_dot_(self, Y: DenseMultiVector, X: DenseMultiVector, W: DenseMultiVector | None = None, mode=NO_TRANS, alpha=1.0, beta=0.0):
        """apply!(Y, A, X, mode, alpha, beta) followed by dot(W, Y) (W defaults to X: the p'Ap of a CG iteration), as ONE
        C-ABI call (jp_apply_dot): no host round trip between the two, and for k = 1 without an importer the SpMV kernel
        accumulates the dot product itself.  Returns the 1 x numVectors array `dot` returns (src/MultiVector.jl:84-108)."""
        if self.isFillActive():
            raise InvalidStateError("Cannot call apply(...) until fillComplete(...)")
        W = X if W is None else W
        comm = self.comm
        if self.exporter is not None or (not Y.getMap().distributedGlobal() and comm.numProc() > 1):
            # the exporter's ADD, or the commReduce of a replicated Y (src/RowMatrix.jl:353-355), sits between apply! and
            # dot: not fusable
            self.apply_(Y, X, mode, alpha, beta)
            return W.dot(Y)
        out = np.empty((1, Y.numVectors()), dtype=np.float64)
        plan = self.importer.device_plan() if self.importer is not None else 0
        _lib.check(_lib.lib().jp_apply_dot(comm.handle, Y.h, self.h, X.h, plan, int(mode), float(alpha), float(beta), W.h, _lib.ptr(out)))
        return out

    def apply(self, X: DenseMultiVector, mode=NO_TRANS, alpha=1.0):  # src/Operator.jl:80-88
        Y = DenseMultiVector(self.domainMap if isTransposed(mode) else self.rangeMap, X.numVectors())
        return self.apply_(Y, X, mode, alpha, 0.0)

    def mul_(self, Y: DenseMultiVector, X: DenseMultiVector):  # LinearAlgebra.mul!, src/RowMatrix.jl:406-409
        return self.apply_(Y, X, NO_TRANS, 1.0, 0.0)

    def __mul__(self, X: DenseMultiVector):  # A*X, src/RowMatrix.jl:411-416
        return self.apply(X)

    def apply_host(self, X: np.ndarray, Y: np.ndarray | None = None, mode=NO_TRANS, alpha=1.0, beta=0.0) -> np.ndarray:
        """apply! on HOST arrays (column-major n x k), the call a host that keeps Array{Float64,2} data makes: the
        H2D copy of X, the kernels and the D2H copy of Y all happen inside this one C-ABI call (jp_apply_host)."""
        if self.isFillActive():
            raise InvalidStateError("Cannot call apply(...) until fillComplete(...)")
        if self.exporter is not None:
            raise InvalidStateError("apply_host: a matrix whose range map differs from its row map needs device multivectors (apply_)")
        X = np.asfortranarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1, order="F")
        k = X.shape[1]
        ny = self.colMap.numMyElements() if isTransposed(mode) else self.getLocalNumRows()
        if Y is None:
            Y = np.zeros((ny, k), dtype=np.float64, order="F")
        if not (Y.flags.f_contiguous and Y.dtype == np.float64 and Y.shape == (ny, k)):
            raise InvalidArgumentError("apply_host: Y must be a column-major float64 array of shape (rows, k)")
        plan = self.importer.device_plan() if self.importer is not None else 0
        _lib.check(_lib.lib().jp_apply_host(self.comm.handle, self.h, plan, _lib.ptr(X), _lib.ptr(Y), k, int(mode), float(alpha), float(beta)))
        return Y


def apply_(Y, A: CSRMatrix, X, mode=NO_TRANS, alpha=1.0, beta=0.0):
    """apply!(Y, A, X[, mode][, alpha][, beta]) with the default arguments of src/Operator.jl:66-72"""
    return A.apply_(Y, X, mode, alpha, beta)
