#!/usr/bin/env python
"""A/B of the SpMM kernels on one GPU (BASELINE configs[2] family: 3-D 27-point stencil, k = 8 and 32): the matrix is
built once on the host and uploaded once per variant (the path is chosen per matrix from the environment at its first
k > 1 apply).  Device-event time per localApply, algorithmic GB/s (nnz*12 + (n+1)*4 + 16*n*k) and the maximum
difference from the gather kernel's result.  Prints one JSON object.
Usage: python tools/bench_spmm.py [--N 128] [--kind 27pt] [--reps 20]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpetra_b200 as jp  # noqa: E402
from jpetra_b200 import _lib, synthetic  # noqa: E402

VARIANTS = [
    ("gather", {"JP_SPMM_RUNS": "0", "JP_SPMM_DIA": "0"}),  # spmm_kernel: X gathered through L1 / L2
    ("runs", {"JP_SPMM_DIA": "0"}),                       # spmm_runs_kernel: CSR + X tiles through aligned column runs
    ("default", {}),                                      # spmm_dia_kernel for diagonal-structured matrices
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--kind", default="27pt")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--ks", default="8,32")
    ap.add_argument("--variants", default="")
    ap.add_argument("--direct", action="store_true", help="time the CSRMatrix's own handle (device fillComplete, default "
                    "path selection) instead of re-uploading the matrix per variant")
    args = ap.parse_args()
    lib = _lib.init(0)
    comm = jp.SerialComm()
    N, kind = args.N, args.kind
    n = N ** synthetic.STENCILS[kind][0]
    m = jp.BlockMap(n, comm)
    fill = "host"
    if args.direct:
        A = synthetic.build_matrix(jp, m, kind, N, fill=False)
        try:
            A.fillComplete(m, m, device=True)
            fill = "device"
        except Exception as e:  # noqa: BLE001 -- keep the measurement alive
            print(f"device fillComplete failed ({e}); falling back to the host path", file=sys.stderr)
            A = synthetic.build_matrix(jp, m, kind, N)
        nnz = A.getLocalNumEntries()
    else:
        A = synthetic.build_matrix(jp, m, kind, N)
        rp, ci, va = A.rowOffsets, A.localIndices1D, A.values
        nnz = len(va)
    peak = 6531.9
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    e0, e1 = C.c_int64(), C.c_int64()
    _lib.check(lib.jp_event_create(C.byref(e0)))
    _lib.check(lib.jp_event_create(C.byref(e1)))
    out = {"kind": kind, "N": N, "n": n, "nnz": nnz, "peak_gbs": peak, "fill": fill, "cases": []}
    ref = {}
    keep = [v for v in args.variants.split(",") if v]
    for label, env in ([("default_direct", {})] if args.direct else VARIANTS):
        if not args.direct and keep and label not in keep:
            continue
        for key in ("JP_SPMM_DIA", "JP_SPMM_RUNS"):
            os.environ.pop(key, None)
        os.environ.update(env)
        h = C.c_int64(A.h if args.direct else 0)
        if not args.direct:
            _lib.check(lib.jp_csr_create(n, n, n, nnz, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(va), 1, C.byref(h)))
        for k in (int(s) for s in args.ks.split(",")):
            X = jp.DenseMultiVector(m, synthetic.vector_block(1, n, k, seed=1))
            Y = jp.DenseMultiVector(m, k)
            try:
                for _ in range(3):
                    _lib.check(lib.jp_local_apply(Y.h, h.value, X.h, 1, 1.0, 0.0))
            except Exception as e:  # noqa: BLE001 -- a configuration that does not fit (shared memory) must not end the sweep
                out["cases"].append({"variant": label, "k": k, "error": str(e)[:200]})
                continue
            lib.jp_synchronize()
            _lib.check(lib.jp_event_record(e0.value))
            for _ in range(args.reps):
                _lib.check(lib.jp_local_apply(Y.h, h.value, X.h, 1, 1.0, 0.0))
            _lib.check(lib.jp_event_record(e1.value))
            ms = C.c_double()
            _lib.check(lib.jp_event_elapsed_ms(e0.value, e1.value, C.byref(ms)))
            t = ms.value / args.reps
            y = Y.data
            if label == "gather":
                ref[k] = y
            info = np.zeros(10, dtype=np.int64)
            _lib.check(lib.jp_csr_spmm_info(h.value, _lib.ptr(info)))
            gbs = (nnz * 12 + (n + 1) * 4 + 16 * n * k) / t / 1e6
            out["cases"].append({"variant": label, "k": k, "ms": t, "algorithmic_gbs": gbs, "frac_of_peak": gbs / peak,
                                 "gflops": 2.0 * nnz * k / t / 1e6, "runs_state": int(info[1]), "widest_tile": int(info[3]), "dia_state": int(info[4]), "dia_rows": int(info[6]),
                                 "max_abs_diff_vs_gather": float(np.abs(y - ref[k]).max()) if k in ref else None})
            del X, Y
        if not args.direct:
            _lib.check(lib.jp_csr_destroy(h.value))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
