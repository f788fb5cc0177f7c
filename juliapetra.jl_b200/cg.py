"""Conjugate gradients on the Operator interface -- the solver loop BASELINE configs[4] times ("CG-style loop: apply! +
dot + norm2 + axpy").  The reference ships no solver; this is the canonical caller of its hot path
(apply!, dot, norm, broadcast updates: src/RowMatrix.jl:261-357, src/MultiVector.jl:84-146, 497-502).

fused=True uses the fused building blocks of SURVEY.md 8f item 3: `apply_dot_` (apply! + p'Ap in one call / one kernel)
and `axpby_norm2_` (r -= alpha*Ap fused with norm(r)), i.e. 4 launches and 2 host synchronisations per iteration
instead of 6 and 3.  Both variants perform the same floating-point operations on the vectors; only the reduction order
of the scalars differs (within the 1e-12 contract).
"""
from __future__ import annotations

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
