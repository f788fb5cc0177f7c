"""Import / Export communication plans (src/ImportExportData.jl:1-21, src/Import.jl:82-244,
src/Export.jl:18-166).  The index lists `numSameIDs, permuteToLIDs, permuteFromLIDs, remoteLIDs, exportLIDs,
exportPIDs` are bit-exact parity targets (1-based, as in the reference); `device_plan()` uploads them once to the
device (jp_plan_create) where doTransfer runs.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib
from .blockmap import BlockMap
from .errors import InvalidArgumentError


class ImportExportData:
    """src/ImportExportData.jl:1-21"""

    def __init__(self, source: BlockMap, target: BlockMap):
        z = lambda: np.zeros(0, dtype=np.int64)
        self.source, self.target = source, target
        self.permuteToLIDs, self.permuteFromLIDs, self.remoteLIDs = z(), z(), z()
        self.exportLIDs, self.exportPIDs = z(), z()
        self.numSameIDs = 0
        self.distributor = source.getComm().createDistributor()
        self.isLocallyComplete = True
        self._device_plan = None


def _common_prefix(a: np.ndarray, b: np.ndarray) -> int:
    n = min(len(a), len(b))
    if n == 0:
        return 0
    neq = np.nonzero(a[:n] != b[:n])[0]
    return int(neq[0]) if len(neq) else n


class _Plan:
    """shared getters (src/Import.jl:246-317, src/Export.jl:168-238)"""
    data: ImportExportData

    def sourceMap(self):
        return self.data.source

    def targetMap(self):
        return self.data.target

    def distributor(self):
        return self.data.distributor

    def isLocallyComplete(self):
        return self.data.isLocallyComplete

    def numSameIDs(self):
        return self.data.numSameIDs

    def permuteToLIDs(self):
        return self.data.permuteToLIDs

    def permuteFromLIDs(self):
        return self.data.permuteFromLIDs

    def remoteLIDs(self):
        return self.data.remoteLIDs

    def exportLIDs(self):
        return self.data.exportLIDs

    def exportPIDs(self):
        return self.data.exportPIDs

    def device_plan(self) -> int:
        """jp plan handle: the lists above + the distributor plan on the device (created once, fails loudly
        without a GPU).  export LIDs are handed over in send order (grouped by procs_to)."""
        d = self.data
        if d._device_plan is None:
            f = d.distributor.plan_fields()
            export = d.exportLIDs
            if len(f["indices_to"]):
                export = export[f["indices_to"] - 1]
            comm = d.source.getComm()
            h = C.c_int64()
            pt, pf = _lib.as_i32(d.permuteToLIDs), _lib.as_i32(d.permuteFromLIDs)
            rl, el = _lib.as_i32(d.remoteLIDs), _lib.as_i32(export)
            pto, lto = _lib.as_i32(f["procs_to"]), _lib.as_i64(f["lengths_to"])
            pfr, lfr = _lib.as_i32(f["procs_from"]), _lib.as_i64(f["lengths_from"])
            _lib.check(_lib.lib().jp_plan_create(
                comm.handle, d.source.numMyElements(), d.target.numMyElements(), d.numSameIDs, _lib.ptr(pt), _lib.ptr(pf), len(pt),
                _lib.ptr(rl), len(rl), _lib.ptr(el), len(el), _lib.ptr(pto), _lib.ptr(lto), len(pto), _lib.ptr(pfr), _lib.ptr(lfr),
                len(pfr), 1, C.byref(h)))
            d._device_plan = h.value
        return d._device_plan


class Import(_Plan):
    """Import(source, target[, remotePIDs]) -- src/Import.jl:82-93"""

    def __init__(self, source: BlockMap, target: BlockMap, remotePIDs=None):
        self.data = self.importData = ImportExportData(source, target)
        remoteGIDs = self._setupSamePermuteRemote()
        if source.distributedGlobal():
            self._setupExport(remoteGIDs, remotePIDs)

    def _setupSamePermuteRemote(self):  # src/Import.jl:97-150 (getIDSources)
        data = self.data
        source, target = data.source, data.target
        sourceGIDs, targetGIDs = source.myGlobalElements(), target.myGlobalElements()
        numSame = _common_prefix(sourceGIDs, targetGIDs)
        data.numSameIDs = numSame
        tgtLIDs = np.arange(numSame + 1, len(targetGIDs) + 1, dtype=np.int64)
        rest = targetGIDs[numSame:]
        srcLIDs = source.lid(rest) if len(rest) else np.zeros(0, dtype=np.int64)
        local = srcLIDs != 0
        data.permuteToLIDs = tgtLIDs[local]
        data.permuteFromLIDs = srcLIDs[local]
        remoteGIDs = rest[~local].copy()
        data.remoteLIDs = tgtLIDs[~local]
        if len(data.remoteLIDs) != 0 and not source.distributedGlobal():
            data.isLocallyComplete = False
            warnings.warn("Target has remote LIDs but source is not distributed globally.  Importing a submap of the target map")
        return remoteGIDs

    def _setupExport(self, remoteGIDs, userRemotePIDs):  # src/Import.jl:152-244
        data = self.data
        source = data.source
        if userRemotePIDs is not None:
            remoteProcIDs = np.asarray(userRemotePIDs, dtype=np.int64).ravel().copy()
            if len(remoteProcIDs) != len(remoteGIDs):
                raise InvalidArgumentError("remotePIDs must either be null or match the size of remoteGIDs.")
        else:
            remoteProcIDs, owner_lids = source.remoteIDList(remoteGIDs)
            missing = owner_lids == 0
            if missing.any():
                data.isLocallyComplete = False
                warnings.warn("Source map was un-able to figure out which process owns one or more of the GIDs in the list of remote GIDs.")
                keep = remoteProcIDs != 0
                remoteProcIDs, remoteGIDs, data.remoteLIDs = remoteProcIDs[keep], remoteGIDs[keep], data.remoteLIDs[keep]
        order = np.argsort(remoteProcIDs, kind="stable")  # sortperm, :226
        remoteProcIDs = remoteProcIDs[order]
        remoteGIDs = remoteGIDs[order]
        data.remoteLIDs = data.remoteLIDs[order]  # the reference permutes a shadowing local here (:229)
        exportGIDs, exportPIDs = data.distributor.createFromRecvs(remoteGIDs, remoteProcIDs)  # :231
        data.exportPIDs = np.asarray(exportPIDs, dtype=np.int64)
        if len(exportGIDs) > 0:
            data.exportLIDs = source.lid(np.asarray(exportGIDs, dtype=np.int64))


class Export(_Plan):
    """Export(source, target) -- src/Export.jl:18-45"""

    def __init__(self, source: BlockMap, target: BlockMap, remotePIDs=None):
        self.data = self.exportData = ImportExportData(source, target)
        exportGIDs = self._setupSamePermuteExport()
        if source.distributedGlobal():
            self._setupRemote(exportGIDs)

    def _setupSamePermuteExport(self):  # src/Export.jl:50-139
        data = self.data
        source, target = data.source, data.target
        sourceGIDs, targetGIDs = source.myGlobalElements(), target.myGlobalElements()
        numSame = _common_prefix(sourceGIDs, targetGIDs)
        data.numSameIDs = numSame
        srcLIDs = np.arange(numSame + 1, len(sourceGIDs) + 1, dtype=np.int64)
        rest = sourceGIDs[numSame:]
        tgtLIDs = target.lid(rest) if len(rest) else np.zeros(0, dtype=np.int64)
        local = tgtLIDs != 0
        data.permuteToLIDs = tgtLIDs[local]
        data.permuteFromLIDs = srcLIDs[local]
        exportGIDs = rest[~local].copy()
        data.exportLIDs = srcLIDs[~local]
        if len(data.exportLIDs) != 0 and not source.distributedGlobal():
            data.isLocallyComplete = False
            warnings.warn("Source has export LIDs but source not distributed globally.  Exporting to a submap of the target map.")
        if source.distributedGlobal():
            exportPIDs, _ = target.remoteIDList(exportGIDs)
            if (exportPIDs == 0).any():
                warnings.warn("The source Map has GIDs not found in the target Map")
                data.isLocallyComplete = False
                keep = exportPIDs != 0
                exportGIDs, data.exportLIDs, exportPIDs = exportGIDs[keep], data.exportLIDs[keep], exportPIDs[keep]
            data.exportPIDs = exportPIDs
        return exportGIDs

    def _setupRemote(self, exportGIDs):  # src/Export.jl:141-166
        data = self.data
        order = np.argsort(data.exportPIDs, kind="stable")
        data.exportPIDs = data.exportPIDs[order]
        data.exportLIDs = data.exportLIDs[order]
        exportGIDs = exportGIDs[order]
        numRemoteIDs = data.distributor.createFromSends(data.exportPIDs)
        remoteGIDs = data.distributor.resolve(exportGIDs)
        lids = data.target.lid(np.asarray(remoteGIDs, dtype=np.int64)) if len(remoteGIDs) else np.zeros(0, dtype=np.int64)
        out = np.zeros(numRemoteIDs, dtype=np.int64)
        out[:len(lids)] = lids
        data.remoteLIDs = out
