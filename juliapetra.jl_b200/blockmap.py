"""BlockMap -- the row/column partition of the reference (src/BlockMap.jl, src/BlockMapData.jl:40-64) and the
linear/replicated BasicDirectory behind `remoteIDList` (src/BasicDirectory.jl:21-33,122-186).

GIDs, LIDs and PIDs are 1-based integers exactly as in the reference; `lid`/`gid` also accept numpy arrays
(vectorised forms used by fillComplete on 10^8-entry index lists).
"""
from __future__ import annotations

import numpy as np

from .comm import Comm
from .errors import InvalidArgumentError, InvalidStateError


class BlockMap:
    """BlockMap(numGlobalElements, comm)                   uniform linear        src/BlockMap.jl:40-73
       BlockMap(numGlobalElements, numMyElements, comm)    user-defined linear   src/BlockMap.jl:84-126
       BlockMap(numGlobalElements, myGlobalElements, comm) arbitrary GID list    src/BlockMap.jl:205-286"""

    def __init__(self, numGlobalElements, *args):
        if len(args) == 1:
            comm, = args
            self._init_uniform(int(numGlobalElements), comm)
        elif len(args) == 2 and np.ndim(args[0]) == 0:
            self._init_linear(int(numGlobalElements), int(args[0]), args[1])
        elif len(args) == 2:
            self._init_gids(int(numGlobalElements), np.asarray(args[0], dtype=np.int64).ravel(), args[1])
        else:
            raise InvalidArgumentError("BlockMap(numGlobalElements, [numMyElements | myGlobalElements], comm)")

    # ---- the 19-field record of src/BlockMapData.jl:40-64 (fields that this path reads) ----
    def _blank(self, numGlobalElements, comm: Comm):
        self.comm = comm
        self.directory = None
        self._myGlobalElements = None  # materialised lazily for linear maps (src/BlockMap.jl:621-633)
        self.numGlobalElements_ = numGlobalElements
        self.numMyElements_ = 0
        self.minAllGID_ = self.maxAllGID_ = self.minMyGID_ = self.maxMyGID_ = 0
        self.minLID_ = self.maxLID_ = 0
        self.linearMap_ = False
        self.contiguousByPid_ = False  # GIDs ascend with the owner PID (uniform / user-linear constructors)
        self.distributedGlobal_ = False
        self.oneToOneIsDetermined = False
        self.oneToOne = False
        # lookup table of a non-linear map: contiguous prefix [first, first+_contig) then sorted (gid -> lid)
        self._contig = 0
        self._sorted_gids = None
        self._sorted_lids = None

    def _init_uniform(self, N, comm):
        if N < 0:
            raise InvalidArgumentError(f"NumGlobalElements = {N}.  Should be >= 0")
        self._blank(N, comm)
        P = comm.numProc()
        self.linearMap_ = True
        self.contiguousByPid_ = True
        myPIDVal = comm.myPid() - 1
        numMy = N // P
        remainder = N % P
        startIndex = myPIDVal * (numMy + 1)
        if myPIDVal < remainder:
            numMy += 1
        else:
            startIndex -= (myPIDVal - remainder)
        self.numMyElements_ = numMy
        self.minAllGID_ = 1
        self.maxAllGID_ = self.minAllGID_ + N - 1
        self.minMyGID_ = startIndex + 1
        self.maxMyGID_ = self.minMyGID_ + numMy - 1
        self.distributedGlobal_ = self._isDistributedGlobal(N, numMy)
        self._endOfConstructorOps()

    def _init_linear(self, N, numMy, comm):
        if N < -1:
            raise InvalidArgumentError(f"NumGlobalElements = {N}.  Should be >= -1")
        if numMy < 0:
            raise InvalidArgumentError(f"NumMyElements = {numMy}. Should be >= 0")
        self._blank(N, comm)
        self.numMyElements_ = numMy
        self.linearMap_ = True
        self.contiguousByPid_ = True
        self.distributedGlobal_ = self._isDistributedGlobal(N, numMy)
        if not self.distributedGlobal_ or comm.numProc() == 1:
            self.numGlobalElements_ = numMy
            self.minAllGID_ = 1
            self.maxAllGID_ = numMy
            self.minMyGID_ = 1
            self.maxMyGID_ = numMy
        else:
            self.numGlobalElements_ = int(comm.sumAllScalar(numMy))
            self.minAllGID_ = 1
            self.maxAllGID_ = self.numGlobalElements_
            self.maxMyGID_ = int(comm.scanSumScalar(numMy))
            startIndex = self.maxMyGID_ - numMy
            self.minMyGID_ = startIndex + 1
            self.maxMyGID_ = self.minMyGID_ + numMy - 1
        if N != -1 and N != self.numGlobalElements_:  # checkValidNGE, src/BlockMap.jl:424-431
            raise InvalidArgumentError(f"Invalid NumGlobalElements.  NumGlobalElements = {N}.  Should equal "
                                       f"{self.numGlobalElements_}, or be set to -1 to compute automatically")
        self._endOfConstructorOps()

    def _init_gids(self, N, gids, comm):
        self._blank(0, comm)
        numMy = len(gids)
        self.numMyElements_ = numMy
        linear = 1
        if numMy > 0:
            self._myGlobalElements = gids.copy()
            self.minMyGID_ = int(gids.min())
            self.maxMyGID_ = int(gids.max())
            if numMy > 1 and not np.all(gids[1:] == gids[:-1] + 1):
                linear = 0
        else:
            self._myGlobalElements = np.zeros(0, dtype=np.int64)
            self.minMyGID_ = 1
            self.maxMyGID_ = 0
        # linearity as the commented-out code and test/MPIBlockMapTests.jl:48 intend (src/BlockMap.jl:236-239)
        self.linearMap_ = bool(comm.minAllScalar(linear))
        if comm.numProc() == 1:
            self.numGlobalElements_ = numMy
            self.minAllGID_ = self.minMyGID_
            self.maxAllGID_ = self.maxMyGID_
        else:
            lo = -self.minMyGID_ if numMy > 0 else np.iinfo(np.int64).min  # -(Inf) for an empty rank
            r = comm.maxAll(np.array([lo, self.maxMyGID_], dtype=np.int64))
            self.minAllGID_ = int(-r[0])
            self.maxAllGID_ = int(r[1])
            if N != -1:
                self.numGlobalElements_ = N
            elif self.linearMap_:
                self.numGlobalElements_ = int(comm.sumAllScalar(numMy))
            else:
                self.numGlobalElements_ = int(len(np.unique(comm.gatherAll(gids))))
        self.distributedGlobal_ = self._isDistributedGlobal(self.numGlobalElements_, numMy)
        self._endOfConstructorOps()

    def _isDistributedGlobal(self, numGlobal, numMy):  # src/BlockMap.jl:366-375
        if self.comm.numProc() > 1:
            localReplicated = 1 if numGlobal == numMy else 0
            return not bool(self.comm.minAllScalar(localReplicated))
        return False

    def _endOfConstructorOps(self):  # src/BlockMap.jl:377-422
        self.minLID_ = 1
        self.maxLID_ = max(self.numMyElements_, 1)
        if self.linearMap_ or self.numMyElements_ == 0:
            return
        g = self._myGlobalElements
        # GlobalToLocalSetup: the reference ends up hashing every GID; here a contiguous prefix is resolved
        # arithmetically and the rest through a sorted table with last-insert-wins, which is the same function
        brk = np.nonzero(g[1:] != g[:-1] + 1)[0]
        self._contig = int(brk[0]) + 1 if len(brk) else len(g)
        rest = g[self._contig:]
        if len(rest) and ((rest >= g[0]) & (rest < g[0] + self._contig)).any():
            # a duplicate of a prefix GID later in the list overwrites it in the reference's hash (only the
            # first element is resolved arithmetically there, src/BlockMap.jl:399-413): keep that behaviour
            self._contig = 1
            rest = g[1:]
        if len(rest):
            order = np.argsort(rest, kind="stable")
            sg = rest[order]
            sl = (order + self._contig + 1).astype(np.int64)
            last = np.ones(len(sg), dtype=bool)
            last[:-1] = sg[1:] != sg[:-1]  # duplicates: the later insert overwrites (Dict semantics)
            self._sorted_gids, self._sorted_lids = sg[last], sl[last]
            # a GID of the tail that also lies in the prefix range is overwritten by the tail entry in the reference's hash;
            # the prefix test in lid() is applied first there as well (src/BlockMap.jl:553-556), so prefix wins

    # ---- accessors (src/BlockMap.jl:433-522, 578-665) ----
    def getComm(self):
        return self.comm

    def numGlobalElements(self):
        return self.numGlobalElements_

    def numMyElements(self):
        return self.numMyElements_

    def minAllGID(self):
        return self.minAllGID_

    def maxAllGID(self):
        return self.maxAllGID_

    def minMyGID(self):
        return self.minMyGID_

    def maxMyGID(self):
        return self.maxMyGID_

    def minLID(self):
        return self.minLID_

    def maxLID(self):
        return self.maxLID_

    def linearMap(self):
        return self.linearMap_

    def distributedGlobal(self):
        return self.distributedGlobal_

    def uniqueGIDs(self):
        return self.isOneToOne()

    def isOneToOne(self):  # src/BlockMap.jl:733-752
        if not self.oneToOneIsDetermined:
            if self.comm.numProc() < 2:
                self.oneToOne = True
            else:
                if self.directory is None:
                    self.directory = BasicDirectory(self)
                self.oneToOne = self.directory.gidsAllUniquelyOwned()
            self.oneToOneIsDetermined = True
        return self.oneToOne

    def myGlobalElements(self) -> np.ndarray:  # src/BlockMap.jl:621-633
        if self._myGlobalElements is None:
            self._myGlobalElements = self.minMyGID_ + np.arange(self.numMyElements_, dtype=np.int64)
        return self._myGlobalElements

    def myGlobalElementIDs(self) -> np.ndarray:  # src/BlockMap.jl:719-729
        return self.myGlobalElements().copy()

    def lid(self, gid):
        """lid(map, gid): local id of a GID or 0 (src/BlockMap.jl:547-560). Scalar or array."""
        if np.ndim(gid) == 0:
            return int(self.lid(np.array([gid], dtype=np.int64))[0])
        g = np.asarray(gid, dtype=np.int64)
        out = np.zeros(g.shape, dtype=np.int64)
        inrange = (g >= self.minMyGID_) & (g <= self.maxMyGID_)
        if self.linearMap_:
            out[inrange] = g[inrange] - self.minMyGID_ + 1
            return out
        if self.numMyElements_ == 0:
            return out
        first = int(self._myGlobalElements[0])
        pre = inrange & (g >= first) & (g < first + self._contig)
        out[pre] = g[pre] - first + 1
        if self._sorted_gids is not None:
            rest = inrange & ~pre
            if rest.any():
                q = g[rest]
                pos = np.searchsorted(self._sorted_gids, q)
                pos[pos == len(self._sorted_gids)] = 0
                hit = self._sorted_gids[pos] == q
                out[rest] = np.where(hit, self._sorted_lids[pos], 0)
        return out

    def gid(self, lid):
        """gid(map, lid): global id of an LID or 0 (src/BlockMap.jl:567-576). Scalar or array."""
        if np.ndim(lid) == 0:
            return int(self.gid(np.array([lid], dtype=np.int64))[0])
        l = np.asarray(lid, dtype=np.int64)
        out = np.zeros(l.shape, dtype=np.int64)
        if self.numMyElements_ == 0:
            return out
        ok = (l >= self.minLID_) & (l <= self.maxLID_)
        if self.linearMap_:
            out[ok] = l[ok] + self.minMyGID_ - 1
        else:
            out[ok] = self._myGlobalElements[l[ok] - 1]
        return out

    def myGID(self, gid) -> bool:  # src/BlockMap.jl:441
        return self.lid(gid) != 0

    def myLID(self, lid) -> bool:  # src/BlockMap.jl:449
        return self.gid(lid) != 0

    def remoteIDList(self, gidList):
        """(owning PIDs, LIDs on the owner) for a list of GIDs -- src/BlockMap.jl:526-539"""
        if self.directory is None:
            self.directory = BasicDirectory(self)
        return self.directory.getDirectoryEntries(self, np.asarray(gidList, dtype=np.int64).ravel())

    def sameAs(self, other: "BlockMap") -> bool:  # src/BlockMap.jl:668-702
        if self is other:
            return True
        if (self.minAllGID_ != other.minAllGID_ or self.maxAllGID_ != other.maxAllGID_
                or self.numGlobalElements_ != other.numGlobalElements_):
            return False
        mySameMap = 1
        if self.numMyElements_ != other.numMyElements_:
            mySameMap = 0
        if self.linearMap_ and other.linearMap_:
            if self.minMyGID_ != other.minMyGID_:
                mySameMap = 0
        elif mySameMap:
            if not np.array_equal(self.myGlobalElements(), other.myGlobalElements()):
                mySameMap = 0
        return bool(self.comm.minAllScalar(mySameMap))


class BasicDirectory:
    """src/BasicDirectory.jl:8-43: replicated and linear cases (the general hashed directory is non-functional in
    the reference: undefined names at :45-119 and :187-245)."""

    def __init__(self, map_: BlockMap):
        self.map = map_
        comm = map_.getComm()
        self.allMinGIDs = None
        if not map_.distributedGlobal():
            self.entryOnMultipleProcs = comm.numProc() != 1
        elif map_.linearMap():
            allMin = comm.gatherAll(np.array([map_.minMyGID()], dtype=np.int64))
            self.allMinGIDs = np.concatenate([allMin, [1 + map_.maxAllGID()]]).astype(np.int64)
            self.entryOnMultipleProcs = len(np.unique(self.allMinGIDs)) != len(self.allMinGIDs)
        else:
            raise InvalidStateError("BasicDirectory: the general (non-linear, distributed) directory is non-functional "
                                    "in the reference (src/BasicDirectory.jl:45-119)")

    def gidsAllUniquelyOwned(self) -> bool:
        return not self.entryOnMultipleProcs

    def getDirectoryEntries(self, map_: BlockMap, globalEntries: np.ndarray):
        """src/BasicDirectory.jl:122-186 -> (procs, localEntries)"""
        n = len(globalEntries)
        if not map_.distributedGlobal():
            lids = map_.lid(globalEntries) if n else np.zeros(0, dtype=np.int64)
            procs = np.where(lids != 0, map_.getComm().myPid(), 0).astype(np.int64)
            return procs, lids
        # linear: the reference scans allMinGIDs from proc 1 upward; a sorted search finds the same owner
        lst = self.allMinGIDs.copy()
        order = np.argsort(lst, kind="stable")
        lst = lst[order]
        if n and ((globalEntries < map_.minAllGID()).any() or (globalEntries > map_.maxAllGID()).any()):
            bad = globalEntries[(globalEntries < map_.minAllGID()) | (globalEntries > map_.maxAllGID())][0]
            raise InvalidArgumentError(f"GID={bad} out of valid range [{map_.minAllGID()}, {map_.maxAllGID()}]")
        numProc = map_.getComm().numProc()
        # first proc1 (1-based position in the sorted list) with lst[proc1] <= gid < lst[proc1+1]
        pos = np.searchsorted(lst, globalEntries, side="right")  # lst[pos-1] <= gid < lst[pos]
        pos = np.minimum(pos, numProc)
        procs = (order[pos - 1] + 1).astype(np.int64)
        lids = (globalEntries - lst[pos - 1] + 1).astype(np.int64)
        return procs, lids
