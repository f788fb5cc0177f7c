"""Shared helpers: build the same synthetic problem in the CPU oracle and in the host mirror."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import jpetra_b200 as jp  # noqa: E402
from jpetra_b200 import synthetic  # noqa: E402
import oracle as O  # noqa: E402


def problem_rows(kind, N, first, last, **kw):
    if kind == "powerlaw":
        return synthetic.powerlaw_rows(N, first, last, kw.get("mean", 16.0), kw.get("max_len", 10000))
    return synthetic.stencil_rows(kind, N, first, last)


def global_rows(kind, N):
    return N if kind == "powerlaw" else N ** synthetic.STENCILS[kind][0]


def oracle_problem(kind, N, P, **kw):
    """(comm, rowMap, filled OMatrix) in the oracle for P simulated ranks"""
    c = O.OComm(P)
    n = global_rows(kind, N)
    m = O.OMap.uniform(c, n)
    per_rank = []
    for pid in range(1, P + 1):
        info = m.info(pid)
        rows, rp, cols, vals = problem_rows(kind, N, info["minMyGID"], info["maxMyGID"], **kw) if info["numMyElements"] else (
            np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0))
        per_rank.append((rows, rp, cols, vals))
    A = O.OMatrix(m, [np.diff(r[1]) for r in per_rank])
    for pid, (rows, rp, cols, vals) in enumerate(per_rank, start=1):
        if len(rows):
            A.insert_rows(pid, rows, rp, cols, vals)
    A.fill_complete(m, m)
    return c, m, A


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def assert_plan_equal(mirror_plan, oracle_plan, pid):
    for f in ("permuteToLIDs", "permuteFromLIDs", "remoteLIDs", "exportLIDs", "exportPIDs"):
        got = np.asarray(getattr(mirror_plan, f)())
        exp = oracle_plan.field(pid, f)
        assert np.array_equal(got, exp), (f, pid, got[:10], exp[:10])
    assert mirror_plan.numSameIDs() == oracle_plan.field(pid, "numSameIDs")
    d, od = mirror_plan.distributor().plan_fields(), oracle_plan.distributor
    for f in ("procs_to", "lengths_to", "starts_to", "indices_to", "procs_from", "lengths_from", "starts_from"):
        assert np.array_equal(np.asarray(d[f]), od.field(pid, f)), (f, pid, d[f], od.field(pid, f))
