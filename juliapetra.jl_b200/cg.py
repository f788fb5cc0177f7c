"""Conjugate gradients on the Operator interface -- the solver loop BASELINE configs[4] times ("CG-style loop: apply! +
dot + norm2 + axpy").  The reference ships no solver; this is the canonical caller of its hot path
(apply!, dot, norm, broadcast updates: src/RowMatrix.jl:261-357, src/MultiVector.jl:84-146, 497-502).

fused=True uses the fused building blocks of SURVEY.md 8f item 3: `apply_dot_` (apply! + p'Ap in one call / one kernel)
and `axpby_norm2_` (r -= alpha*Ap fused with norm(r)), i.e. 4 launches and 2 host synchronisations per iteration
instead of 6 and 3.  Both variants perform the same floating-point operations on the vectors; only the reduction order
of the scalars differs (within the 1e-12 contract).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .multivector import DenseMultiVector


def cg(A, b: DenseMultiVector, x: DenseMultiVector, tol: float = 1e-10, maxiter: int = 1000, fused: bool = True):
    """Solve A x = b for a symmetric positive definite CSRMatrix (single right-hand side); x holds the initial guess
    and the result.  Returns (iterations, ||r|| / ||b||)."""
    if b.numVectors() != 1 or x.numVectors() != 1:
        raise ValueError("cg: single right-hand side")
    r = b.copy()
    Ap = DenseMultiVector(A.getRangeMap(), 1)
    A.apply_(Ap, x)                      # r = b - A x
    r.axpby_(-1.0, Ap, 1.0)
    p = r.copy()
    bnorm = float(b.norm(2)[0, 0]) or 1.0
    rnorm = float(r.norm(2)[0, 0])
    rr = rnorm * rnorm
    it = 0
    while it < maxiter and rnorm > tol * bnorm:
        if fused:
            pAp = float(A.apply_dot_(Ap, p)[0, 0])
        else:
            A.apply_(Ap, p)
            pAp = float(p.dot(Ap)[0, 0])
        alpha = rr / pAp
        x.axpby_(alpha, p, 1.0)          # @. x = x + alpha*p
        if fused:
            rnorm = float(r.axpby_norm2_(-alpha, Ap, 1.0)[0, 0])
        else:
            r.axpby_(-alpha, Ap, 1.0)    # @. r = r - alpha*Ap
            rnorm = float(r.norm(2)[0, 0])
        rr_new = rnorm * rnorm
        p.axpby_(1.0, r, rr_new / rr)    # @. p = r + beta*p
        rr = rr_new
        it += 1
    return it, rnorm / bnorm


class Scalars:
    """A small device-resident array of doubles (jp_scalars_*): reductions leave their results here and update kernels
    read their coefficients from it, so a solver iteration needs no host round trip."""

    def __init__(self, n: int):
        h = C.c_int64()
        _lib.check(_lib.lib().jp_scalars_create(int(n), C.byref(h)))
        self.h, self.n = h.value, int(n)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.load().jp_scalars_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def download(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        _lib.check(_lib.lib().jp_scalars_download(self.h, _lib.ptr(out)))
        return out

    def upload(self, values):
        a = _lib.as_f64(values)
        if len(a) != self.n:
            raise ValueError("Scalars.upload: wrong length")
        _lib.check(_lib.lib().jp_scalars_upload(self.h, _lib.ptr(a)))


def cg_device(A, b: DenseMultiVector, x: DenseMultiVector, tol: float = 1e-10, maxiter: int = 1000, check_every: int = 10):
    """The same conjugate-gradient iteration as `cg`, with every scalar kept on the device (jp_*_dev): p'Ap, ||r||^2 and
    the coefficients alpha = rr / p'Ap, beta = rr_new / rr never travel to the host, so an iteration is four launches of
    pure stream work; the host looks at the residual every `check_every` iterations.  Same algorithm and kernels as
    `cg(fused=True)` (same reductions, same unfused updates); the scalars differ in the last bits because the host
    version squares the square root `norm` returns.  It may run up to check_every - 1 iterations past convergence.
    Returns (iterations, ||r|| / ||b||).
    Matrices with an exporter or a replicated range map take `cg`."""
    if b.numVectors() != 1 or x.numVectors() != 1:
        raise ValueError("cg_device: single right-hand side")
    if A.exporter is not None:
        raise ValueError("cg_device: rangeMap != rowMap is served by cg()")
    lib = _lib.lib()
    comm = A.comm
    plan = A.importer.device_plan() if A.importer is not None else 0
    RR, PAP, RR_NEW, BB = 0, 1, 2, 3
    s = Scalars(4)
    r = b.copy()
    Ap = DenseMultiVector(A.getRangeMap(), 1)
    A.apply_(Ap, x)                       # r = b - A x
    r.axpby_(-1.0, Ap, 1.0)
    p = r.copy()
    _lib.check(lib.jp_mv_dot_dev(comm.handle, b.h, b.h, s.h, BB))
    _lib.check(lib.jp_mv_dot_dev(comm.handle, r.h, r.h, s.h, RR))
    v = s.download()
    bnorm = float(np.sqrt(v[BB])) or 1.0
    rnorm = float(np.sqrt(v[RR]))
    it = 0
    while it < maxiter and rnorm > tol * bnorm:
        for _ in range(min(check_every, maxiter - it)):
            _lib.check(lib.jp_apply_dot_dev(comm.handle, Ap.h, A.h, p.h, plan, 1.0, 0.0, p.h, s.h, PAP))      # Ap = A p; s[PAP] = p'Ap
            _lib.check(lib.jp_mv_axpby_dev(x.h, p.h, s.h, 1.0, RR, PAP, 1.0, -1, -1))                          # x += (rr/pAp) p
            _lib.check(lib.jp_mv_axpby_norm2sq_dev(comm.handle, r.h, Ap.h, s.h, -1.0, RR, PAP, 1.0, -1, -1, RR_NEW))  # r -= (rr/pAp) Ap
            _lib.check(lib.jp_mv_axpby_dev(p.h, r.h, s.h, 1.0, -1, -1, 1.0, RR_NEW, RR))                        # p = r + (rr_new/rr) p
            _lib.check(lib.jp_scalars_copy(s.h, RR, RR_NEW))
            it += 1
        rnorm = float(np.sqrt(s.download()[RR]))
    return it, rnorm / bnorm
