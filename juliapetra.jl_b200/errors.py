"""Error convention of the reference (src/Utils.jl:37-51)."""


class InvalidArgumentError(Exception):
    """Thrown when an argument is invalid (src/Utils.jl:37-43)."""


class InvalidStateError(Exception):
    """Thrown when an object is in a state that does not allow the call (src/Utils.jl:45-51)."""


class BackendError(RuntimeError):
    """CUDA / NCCL failure reported by libjpetra_b200.so (status 3 / 4). There is no CPU fallback."""
