"""Builds the synthetic benchmark/test problems inside the CPU oracle, for P simulated ranks.
Test infrastructure / CPU-baseline only.  The row generator is the shared pure function of the global row id
(juliapetra.jl_b200/synthetic.py); everything after it (insert, fillComplete, Import, apply!) is the oracle's."""
import numpy as np

from . import OComm, OMap, OMatrix


def _synthetic():
    import jpetra_b200  # noqa: F401  (registers the package)
    from jpetra_b200 import synthetic
    return synthetic


def problem_rows(kind, N, first, last, **kw):
    s = _synthetic()
    if kind == "powerlaw":
        return s.powerlaw_rows(N, first, last, kw.get("mean", 16.0), kw.get("max_len", 10000))
    return s.stencil_rows(kind, N, first, last)


def global_rows(kind, N):
    return N if kind == "powerlaw" else N ** _synthetic().STENCILS[kind][0]


def oracle_problem(kind, N, P, chunk_rows=1 << 21, abs_vals=False, **kw):
    """(comm, rowMap, filled OMatrix) for P simulated ranks, rows split by BlockMap(n, comm); abs_vals=True builds |A|
    (same structure, absolute values: the operator of the component-wise error bounds)"""
    s = _synthetic()
    c = OComm(P)
    n = global_rows(kind, N)
    m = OMap.uniform(c, n)
    infos = [m.info(pid) for pid in range(1, P + 1)]
    counts = []
    for info in infos:
        if kind == "powerlaw":
            r = np.arange(info["minMyGID"] - 1, info["maxMyGID"], dtype=np.int64)
            counts.append(s.powerlaw_row_lengths(r, N, kw.get("mean", 16.0), kw.get("max_len", 10000)))
        else:
            counts.append(np.full(info["numMyElements"], len(s._offsets(kind)), dtype=np.int64))
    A = OMatrix(m, counts)
    for pid, info in enumerate(infos, start=1):
        lo, last = info["minMyGID"], info["maxMyGID"]
        while lo <= last and info["numMyElements"] > 0:
            hi = min(last, lo + chunk_rows - 1)
            rows, rp, cols, vals = problem_rows(kind, N, lo, hi, **kw)
            A.insert_rows(pid, rows, rp, cols, np.abs(vals) if abs_vals else vals)
            lo = hi + 1
    A.fill_complete(m, m)
    return c, m, A
