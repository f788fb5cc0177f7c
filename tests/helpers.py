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


from oracle.problems import global_rows, oracle_problem, problem_rows  # noqa: E402,F401


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
