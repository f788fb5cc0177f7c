#!/usr/bin/env python
"""Regenerates tests/golden/bench_parity.json from the CPU oracle (run from the repo root:
python tests/golden/make_bench_fixtures.py).  For the bench workload (3-D 7-point Laplacian 256^3, X = vector_block(seed=1),
apply!(Y, A, X, NO_TRANS, 1, 0)) and every rank count the scaling run uses, the oracle's P-rank simulation of the
reference gives one SHA-256 per rank of that rank's slab of Y, plus the global dot(X, Y) and norm(Y, 2).  bench.py hashes
its own Y after the timed loop and compares: the stencil SpMV keeps the reference's summation order (column-map order
inside every row, which DEPENDS on P: remote columns sort last), so the digests must match bit for bit at every P."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle.problems import oracle_problem  # noqa: E402
from jpetra_b200 import synthetic  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


DEFAULT_CASES = (("7pt", 256, (1, 2, 4, 8)), ("7pt", 64, (1, 2, 3, 4, 8)), ("7pt", 512, (8,)))


def main():
    """no arguments: regenerate everything; `7pt-512/P8 ...`: regenerate only the named cases and merge them into the file
    (the 512^3 case, the matrix of BASELINE configs[4], needs ~40 GB of host memory and a few minutes)"""
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_parity.json")
    only = set(sys.argv[1:])
    fx = {"generator": "tests/golden/make_bench_fixtures.py (CPU oracle, oracle/jp_oracle.cpp)", "cases": {}}
    if only and os.path.exists(out_path):
        fx = json.load(open(out_path))
    for kind, N, ranks in DEFAULT_CASES:
        for P in ranks:
            name = f"{kind}-{N}/P{P}"
            if only and name not in only:
                continue
            t0 = time.time()
            c, m, A = oracle_problem(kind, N, P)
            xs = [synthetic.vector_block(m.info(p)["minMyGID"], m.info(p)["numMyElements"], 1, seed=1) for p in range(1, P + 1)]
            rs = [synthetic.vector_block(m.info(p)["minMyGID"], m.info(p)["numMyElements"], 1, seed=2) for p in range(1, P + 1)]
            X = O.OMV.from_arrays(m, xs)
            Y = O.OMV(m, 1)
            A.apply(Y, X, O.NO_TRANS, 1.0, 0.0)
            fx["cases"][name] = {
                "y_sha256": [digest(Y.get(p)) for p in range(1, P + 1)],
                "rows": [int(m.info(p)["numMyElements"]) for p in range(1, P + 1)],
                "dot_x_y": float(X.dot(Y)[0]), "norm2_y": float(Y.norm(2)[0]),
                "norm2_r": float(O.OMV.from_arrays(m, rs).norm(2)[0]),  # the residual vector of the CG-style loop (seed 2)
            }
            print(f"{name}: {time.time() - t0:.1f}s", flush=True)
            del A, X, Y
            json.dump(fx, open(out_path, "w"), indent=1, sort_keys=True)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
