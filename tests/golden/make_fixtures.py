#!/usr/bin/env python
"""Regenerates tests/golden/fixtures.json from the CPU oracle (run from the repo root: python tests/golden/make_fixtures.py).
The fixtures freeze, for seeded synthetic problems, (a) the integer structures that must be reproduced bit-exactly
(column maps, Import lists, distributor plans, CSR) for a 4-rank 3-D 7-point problem, and (b) SHA-256 digests of the
Float64 results that the CUDA path reproduces bit-exactly (stencil SpMV keeps the reference's summation order)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle.problems import oracle_problem  # noqa: E402
from jpetra_b200 import synthetic  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    fx = {"generator": "tests/golden/make_fixtures.py (CPU oracle, oracle/jp_oracle.cpp)"}
    # (a) 4-rank 7-point 6^3: structures
    P, N = 4, 6
    c, m, A = oracle_problem("7pt", N, P)
    ranks = {}
    for pid in range(1, P + 1):
        rp, ci, va = A.csr(pid)
        imp = A.importer
        d = imp.distributor
        ranks[str(pid)] = {
            "colMap_sha256": digest(A.col_map.my_gids(pid)), "colMap_len": int(len(A.col_map.my_gids(pid))),
            "colMap_tail": A.col_map.my_gids(pid)[-5:].tolist(),
            "numSame": imp.field(pid, "numSameIDs"), "remoteLIDs_sha256": digest(imp.field(pid, "remoteLIDs")),
            "exportLIDs_sha256": digest(imp.field(pid, "exportLIDs")), "exportPIDs": sorted(set(imp.field(pid, "exportPIDs").tolist())),
            "n_remote": int(len(imp.field(pid, "remoteLIDs"))), "n_export": int(len(imp.field(pid, "exportLIDs"))),
            "procs_to": d.field(pid, "procs_to").tolist(), "lengths_to": d.field(pid, "lengths_to").tolist(),
            "procs_from": d.field(pid, "procs_from").tolist(), "lengths_from": d.field(pid, "lengths_from").tolist(),
            "rowptr_sha256": digest(rp), "colind_sha256": digest(ci), "vals_sha256": digest(va), "nnz": int(len(va)),
        }
    fx["7pt_6_P4"] = ranks
    # (b) serial 7-point 64^3 and 5-point 300^2: digests of apply!(Y, A, X, NO_TRANS, 3, 0.5) and of A*X
    for kind, size in (("7pt", 64), ("5pt", 300), ("27pt", 20)):
        c, m, A = oracle_problem(kind, size, 1)
        n = m.info(1)["numMyElements"]
        rp, ci, va = A.csr(1)
        x = synthetic.vector_block(1, n, 1, seed=1)
        y0 = synthetic.vector_block(1, n, 1, seed=2)
        y = O.local_apply_raw(y0.copy(order="F"), rp, ci, va, x, O.NO_TRANS, 3.0, 0.5)
        ax = O.local_apply_raw(np.zeros((n, 1)), rp, ci, va, x, O.NO_TRANS, 1.0, 0.0)
        fx[f"{kind}_{size}_serial"] = {"n": int(n), "nnz": int(len(va)), "x_sha256": digest(x), "y0_sha256": digest(y0),
                                       "apply_3_05_sha256": digest(y), "ax_sha256": digest(ax), "dot_x_ax": float(O.dot_raw(x, ax)[0]),
                                       "norm2_ax": float(np.sqrt(O.norm_raw(ax, 2.0)[0]))}
    json.dump(fx, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.json"), "w"), indent=1, sort_keys=True)
    print("wrote fixtures.json")


if __name__ == "__main__":
    main()
