"""juliapetra.jl_b200 -- B200-native backend for JuliaPetra's distributed sparse matrix-vector hot path.

Host mirror of the reference's interface for the path (same names, argument meaning and error behaviour):
  Comm / SerialComm / NCCLComm, BlockMap, Import / Export, Distributor, DenseMultiVector, CSRMatrix,
  apply_ (apply!), dot, norm, doImport / doExport.
All arithmetic runs in libjpetra_b200.so (hand-written sm_100a CUDA + NCCL) through the C ABI declared in
include/jpetra_b200.h; nothing here falls back to the CPU.

The directory name contains a dot, so it is imported through the `jpetra_b200` shim at the repo root:
    import jpetra_b200 as jp
"""
from . import _lib
from .blockmap import BasicDirectory, BlockMap
from .cg import Scalars, cg, cg_device
from .comm import Comm, NCCLComm, SerialComm, TorchDistComm
from .csrmatrix import (CONJ_TRANS, NO_TRANS, STATIC_PROFILE, TRANS, CSRMatrix, TransposeMode, apply_, isTransposed)
from .distributor import Distributor, SerialDistributor
from .errors import BackendError, InvalidArgumentError, InvalidStateError
from .importexport import Export, Import, ImportExportData
from .multivector import (ABSMAX, ADD, INSERT, REPLACE, ZERO, CombineMode, DenseMultiVector, doExport, doImport, dot, norm)
from .utils import computeOffsets

__all__ = [
    "Comm", "SerialComm", "NCCLComm", "TorchDistComm", "BlockMap", "BasicDirectory", "Distributor", "SerialDistributor",
    "Import", "Export", "ImportExportData", "DenseMultiVector", "CSRMatrix", "TransposeMode", "CombineMode", "NO_TRANS", "TRANS",
    "CONJ_TRANS", "ADD", "INSERT", "REPLACE", "ABSMAX", "ZERO", "STATIC_PROFILE", "apply_", "dot", "norm", "doImport", "doExport",
    "computeOffsets", "isTransposed", "cg", "cg_device", "Scalars", "InvalidArgumentError", "InvalidStateError", "BackendError",
]
