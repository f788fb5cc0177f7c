"""Distributor -- the communication plan behind Import/Export (src/Distributor.jl:10-86).

`Distributor` restates the plan construction of `MPIDistributor` (src/MPIDistributor.jl:69-299, 423-449): the plan
fields `procs_to / lengths_to / starts_to / indices_to / procs_from / lengths_from / starts_from / numSends /
numRecvs / selfMsg / totalRecvLength` are bit-exact parity targets.  What changes is how the plan is learnt and
executed: the reference's Reduce+Scatter, ANY_SOURCE receives and ready-sends (:205-251) become one gatherAll of a
P-vector of message lengths, and the run-time exchange of MultiVector data happens on the device
(jp_plan_create / jp_transfer, see multivector.py) with static message sizes.  `resolve` on host arrays (used at
setup time for GID lists) is an all-to-all-v through the Comm.

`SerialDistributor` is src/SerialComm.jl:8-73.
"""
from __future__ import annotations

import numpy as np

from .errors import InvalidArgumentError, InvalidStateError


class SerialDistributor:
    """src/SerialComm.jl:8-73: identity plan, PIDs must all be 1."""

    def __init__(self, comm=None):
        self.comm = comm
        self.post = None
        self.reversePost = None
        self._n = 0

    def getComm(self):
        return self.comm

    def createFromSends(self, exportPIDs) -> int:
        exportPIDs = np.asarray(exportPIDs, dtype=np.int64).ravel()
        if (exportPIDs != 1).any():
            raise InvalidArgumentError("SerialDistributor can only accept PID of 1")
        self._n = len(exportPIDs)
        return self._n

    def createFromRecvs(self, remoteGIDs, remotePIDs):
        remoteGIDs = np.asarray(remoteGIDs, dtype=np.int64).ravel()
        remotePIDs = np.asarray(remotePIDs, dtype=np.int64).ravel()
        if (remotePIDs != 1).any():
            raise InvalidArgumentError("SerialDistributor can only accept PID of 1")
        self._n = len(remoteGIDs)
        return remoteGIDs, remotePIDs

    def resolve(self, exportObjs):
        return exportObjs

    def resolveReverse(self, exportObjs):
        return exportObjs

    def resolvePosts(self, exportObjs):
        self.post = exportObjs

    def resolveWaits(self):
        if self.post is None:
            raise InvalidStateError("Must post before waiting")
        result, self.post = self.post, None
        return result

    def resolveReversePosts(self, exportObjs):
        self.reversePost = exportObjs

    def resolveReverseWaits(self):
        if self.reversePost is None:
            raise InvalidStateError("Must reverse post before reverse waiting")
        result, self.reversePost = self.reversePost, None
        return result

    # plan fields in the MPIDistributor vocabulary (one self message), for jp_plan_create
    def plan_fields(self):
        n = self._n
        one = np.array([1], dtype=np.int64) if n else np.zeros(0, dtype=np.int64)
        ln = np.array([n], dtype=np.int64) if n else np.zeros(0, dtype=np.int64)
        return dict(procs_to=one, lengths_to=ln, starts_to=one.copy(), indices_to=np.zeros(0, dtype=np.int64),
                    procs_from=one.copy(), lengths_from=ln.copy(), starts_from=one.copy())


class Distributor:
    """Plan record of src/MPIDistributor.jl:12-65 for a multi-rank Comm."""

    def __init__(self, comm):
        self.comm = comm
        z = lambda: np.zeros(0, dtype=np.int64)
        self.lengths_to, self.procs_to, self.indices_to, self.starts_to = z(), z(), z(), z()
        self.lengths_from, self.procs_from, self.starts_from = z(), z(), z()
        self.numRecvs = self.numSends = self.numExports = 0
        self.selfMsg = 0
        self.maxSendLength = 0
        self.totalRecvLength = 0
        self.planReverse = None
        self._posted = None
        self._has_plan = False

    def getComm(self):
        return self.comm

    # ---- src/MPIDistributor.jl:69-193 ----
    def _createSendStructure(self, pid: int, nProcs: int, exportPIDs: np.ndarray):
        numExports = len(exportPIDs)
        self.numExports = numExports
        active = exportPIDs >= 1
        if (exportPIDs > nProcs).any():
            raise InvalidArgumentError("export PID larger than the number of processes")
        starts = np.bincount(exportPIDs[active], minlength=nProcs + 1)[1:nProcs + 1].astype(np.int64)  # starts[p-1]
        nactive = int(active.sum())
        noSendBuff = not bool((exportPIDs[1:] < exportPIDs[:-1]).any())
        numDeadIndices = numExports - nactive
        self.selfMsg = int(starts[pid - 1] != 0)
        self.indices_to = np.zeros(0, dtype=np.int64)
        procs = np.nonzero(starts > 0)[0] + 1  # ascending PIDs with something to send
        self.numSends = len(procs)
        self.procs_to = procs.astype(np.int64)
        self.lengths_to = starts[procs - 1].astype(np.int64)
        if noSendBuff:
            # grouped by processor (ascending, dead indices first): starts_to walks the export list
            self.starts_to = (numDeadIndices + 1 + np.concatenate([[0], np.cumsum(self.lengths_to)[:-1]])).astype(np.int64) \
                if self.numSends else np.zeros(0, dtype=np.int64)
            self.maxSendLength = 0  # :128 assigns a local; the field stays 0
        else:
            # not grouped: indices_to lists, per destination in ascending PID order, the 1-based export positions
            pos = np.nonzero(active)[0]
            order = np.argsort(exportPIDs[pos], kind="stable")
            self.indices_to = (pos[order] + 1).astype(np.int64)
            self.starts_to = (1 + np.concatenate([[0], np.cumsum(self.lengths_to)[:-1]])).astype(np.int64) \
                if self.numSends else np.zeros(0, dtype=np.int64)
            other = self.lengths_to[self.procs_to != pid]
            self.maxSendLength = int(other.max()) if len(other) else 0
        self.numSends -= self.selfMsg

    # ---- src/MPIDistributor.jl:196-273 ----
    def _computeRecvs(self, myProc: int, nProcs: int):
        sendLen = np.zeros(nProcs, dtype=np.int64)
        sendLen[self.procs_to - 1] = self.lengths_to
        M = self.comm.gatherAll(sendLen).reshape(nProcs, nProcs)  # M[src, dst]
        col = M[:, myProc - 1]
        src = np.nonzero(col > 0)[0]
        self.procs_from = (src + 1).astype(np.int64)  # sortperm(procs_from), self included (:254-256)
        self.lengths_from = col[src].astype(np.int64)
        totalRecvs = len(src)
        self.starts_from = (1 + np.concatenate([[0], np.cumsum(self.lengths_from)[:-1]])).astype(np.int64) \
            if totalRecvs else np.zeros(0, dtype=np.int64)
        self.totalRecvLength = int(self.lengths_from.sum())
        self.numRecvs = totalRecvs - self.selfMsg

    # ---- src/MPIDistributor.jl:423-435 ----
    def createFromSends(self, exportPIDs) -> int:
        exportPIDs = np.asarray(exportPIDs, dtype=np.int64).ravel()
        pid, nProcs = self.comm.myPid(), self.comm.numProc()
        self._createSendStructure(pid, nProcs, exportPIDs)
        self._computeRecvs(pid, nProcs)
        self.planReverse = None
        self._has_plan = True
        return self.totalRecvLength

    # ---- src/MPIDistributor.jl:438-449 with computeSends :275-299 ----
    def createFromRecvs(self, remoteGIDs, remotePIDs):
        remoteGIDs = np.asarray(remoteGIDs, dtype=np.int64).ravel()
        remotePIDs = np.asarray(remotePIDs, dtype=np.int64).ravel()
        if len(remoteGIDs) != len(remotePIDs):
            raise InvalidArgumentError("remote lists must be the same length")
        tmpPlan = Distributor(self.comm)
        importObjs = np.empty((len(remoteGIDs), 2), dtype=np.int64)  # (GID, myPid) tuples
        importObjs[:, 0] = remoteGIDs
        importObjs[:, 1] = self.comm.myPid()
        numExports = tmpPlan.createFromSends(remotePIDs.copy())
        exportObjs = tmpPlan.resolve(importObjs)
        assert len(exportObjs) == numExports
        exportGIDs = exportObjs[:, 0].copy()
        exportPIDs = exportObjs[:, 1].copy()
        self.createFromSends(exportPIDs)
        return exportGIDs, exportPIDs

    # ---- src/MPIDistributor.jl:304-343 ----
    def _reverse(self) -> "Distributor":
        if self.planReverse is None:
            r = Distributor(self.comm)
            r.lengths_to, r.procs_to, r.starts_to = self.lengths_from, self.procs_from, self.starts_from
            r.indices_to = np.zeros(0, dtype=np.int64)
            r.lengths_from, r.procs_from, r.starts_from = self.lengths_to, self.procs_to, self.starts_to
            r.numSends, r.numRecvs, r.selfMsg = self.numRecvs, self.numSends, self.selfMsg
            other = self.lengths_from[self.procs_from != self.comm.myPid()]
            r.maxSendLength = int(other.max()) if len(other) else 0
            r.totalRecvLength = int(self.lengths_to.sum())
            r._has_plan = True
            self.planReverse = r
        return self.planReverse

    # ---- exchange of host records: src/MPIDistributor.jl:451-590 ----
    def _exchange(self, exportObjs: np.ndarray) -> np.ndarray:
        if not self._has_plan:
            raise InvalidStateError("Distributor has no plan (createFromSends / createFromRecvs not called)")
        exportObjs = np.ascontiguousarray(exportObjs)
        nProcs = self.comm.numProc()
        total = int(self.lengths_to.sum())
        if len(self.indices_to) == 0:  # already grouped by processor (:469-474): consecutive slices
            if len(exportObjs) < total:
                raise InvalidArgumentError("resolve: exportObjs shorter than the plan's sends")
            send = exportObjs[:total]
        else:  # gather through indices_to (:475-484)
            send = exportObjs[self.indices_to - 1]
        sendcounts = np.zeros(nProcs, dtype=np.int64)
        sendcounts[self.procs_to - 1] = self.lengths_to
        recvcounts = np.zeros(nProcs, dtype=np.int64)
        recvcounts[self.procs_from - 1] = self.lengths_from
        return self.comm.alltoallv(send, sendcounts, recvcounts)  # ascending source PID == procs_from order (:582-589)

    def resolvePosts(self, exportObjs):
        self._posted = self._exchange(np.asarray(exportObjs))

    def resolveWaits(self):
        if self._posted is None:
            raise InvalidStateError("Cannot resolve waits when no posts have been made")
        result, self._posted = self._posted, None
        return result

    def resolve(self, exportObjs):  # src/Distributor.jl:72-75
        self.resolvePosts(exportObjs)
        return self.resolveWaits()

    def resolveReversePosts(self, exportObjs):
        if len(self.indices_to) != 0:
            raise InvalidStateError("Cannot do reverse comm when data is not blocked by processor")
        self._reverse().resolvePosts(exportObjs)

    def resolveReverseWaits(self):
        if self.planReverse is None:
            raise InvalidStateError("Cannot resolve reverse waits if there is no reverse plan")
        return self.planReverse.resolveWaits()

    def resolveReverse(self, exportObjs):  # src/Distributor.jl:83-86
        self.resolveReversePosts(exportObjs)
        return self.resolveReverseWaits()

    def plan_fields(self):
        return dict(procs_to=self.procs_to, lengths_to=self.lengths_to, starts_to=self.starts_to, indices_to=self.indices_to,
                    procs_from=self.procs_from, lengths_from=self.lengths_from, starts_from=self.starts_from)
