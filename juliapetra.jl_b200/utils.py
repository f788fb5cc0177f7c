"""src/Utils.jl:5-26 -- row-count -> 1-based row pointers."""
import numpy as np

from .errors import InvalidArgumentError


def computeOffsets(rowPtrs: np.ndarray, numEnts):
    """computeOffsets(rowPtrs, numEnts): scalar form fills rowPtrs[i] = numEnts*(i-1)+1 (src/Utils.jl:5-11);
    array form writes the running 1-based offsets and returns the total count (src/Utils.jl:12-26)."""
    numOffsets = len(rowPtrs)
    if np.isscalar(numEnts):
        rowPtrs[:] = int(numEnts) * np.arange(numOffsets, dtype=np.int64) + 1
        return rowPtrs
    numEnts = np.asarray(numEnts)
    numCounts = len(numEnts)
    if numCounts >= numOffsets:
        raise InvalidArgumentError(f"length(numEnts) = {numCounts} >= length(rowPtrs) = {numOffsets}")
    csum = np.cumsum(numEnts, dtype=np.int64)
    rowPtrs[0] = 1
    rowPtrs[1:numCounts + 1] = csum + 1
    total = int(csum[-1]) if numCounts else 0
    rowPtrs[numCounts + 1:] = total + 1
    return total
