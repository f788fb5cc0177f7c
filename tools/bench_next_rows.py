#!/usr/bin/env python
"""Measurements of the SURVEY.md 8f rows on one GPU (not the bench line): prints one JSON object.
  fill:  CSRMatrix.fillComplete on the host (numpy, then jp_csr_create) vs on the device (jp_csr_fill_complete),
         wall-clock seconds including the upload, same matrix, results compared.
  cg:    ms per CG iteration (device events) with the separate calls (apply!, dot, axpby, axpby, norm, axpby) and with
         the fused building blocks (apply_dot_, axpby, axpby_norm2_, axpby); algorithmic bytes per iteration are
         nnz*12 + (n+1)*4 + n*8*(14 separate | 11 fused).
Usage: python tools/bench_next_rows.py [--N 128] [--iters 50]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpetra_b200 as jp  # noqa: E402
from jpetra_b200 import _lib, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=128)
    ap.add_argument("--kind", default="7pt")
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    lib = _lib.init(0)
    comm = jp.SerialComm()
    N, kind = args.N, args.kind
    n = N if kind == "powerlaw" else N ** synthetic.STENCILS[kind][0]
    m = jp.BlockMap(n, comm)
    out = {"kind": kind, "N": N, "n": n}

    # ---- fillComplete: host vs device ----
    Ah = synthetic.build_matrix(jp, m, kind, N, fill=False)
    Ad = synthetic.build_matrix(jp, m, kind, N, fill=False)
    lib.jp_synchronize()
    t0 = time.perf_counter()
    Ah.fillComplete(m, m, device=False)
    lib.jp_synchronize()
    t1 = time.perf_counter()
    Ad.fillComplete(m, m, device=True)
    lib.jp_synchronize()
    t2 = time.perf_counter()
    nnz = Ah.getLocalNumEntries()
    same = (np.array_equal(Ah.rowOffsets, Ad.rowOffsets) and np.array_equal(Ah.localIndices1D, Ad.localIndices1D)
            and np.array_equal(Ah.values, Ad.values))
    out["fill"] = {"nnz": nnz, "host_s": t1 - t0, "device_s": t2 - t1, "speedup": (t1 - t0) / (t2 - t1), "identical": bool(same)}

    # ---- CG iteration: separate calls vs fused building blocks ----
    xs = synthetic.vector_block(1, n, 1, seed=7)
    B = Ah.apply(jp.DenseMultiVector(m, xs))
    e0, e1 = C.c_int64(), C.c_int64()
    _lib.check(lib.jp_event_create(C.byref(e0)))
    _lib.check(lib.jp_event_create(C.byref(e1)))
    cg = {}
    for fused in (False, True):
        X = jp.DenseMultiVector(m, 1)
        jp.cg(Ah, B, X, tol=0.0, maxiter=5, fused=fused)  # warm-up
        X = jp.DenseMultiVector(m, 1)
        lib.jp_synchronize()
        _lib.check(lib.jp_event_record(e0.value))
        its, rel = jp.cg(Ah, B, X, tol=0.0, maxiter=args.iters, fused=fused)
        _lib.check(lib.jp_event_record(e1.value))
        ms = C.c_double()
        _lib.check(lib.jp_event_elapsed_ms(e0.value, e1.value, C.byref(ms)))
        per_it = ms.value / its
        bytes_it = nnz * 12 + (n + 1) * 4 + n * 8 * (11 if fused else 14)
        cg["fused" if fused else "separate"] = {"iterations": its, "ms_per_iteration": per_it, "rel_residual": rel,
                                                "algorithmic_gbs": bytes_it / per_it / 1e6}
    try:  # device-resident scalars: no host round trip inside the iteration (jp_*_dev; round-2 candidate)
        X = jp.DenseMultiVector(m, 1)
        jp.cg_device(Ah, B, X, tol=0.0, maxiter=5, check_every=5)
        X = jp.DenseMultiVector(m, 1)
        lib.jp_synchronize()
        _lib.check(lib.jp_event_record(e0.value))
        its, rel = jp.cg_device(Ah, B, X, tol=0.0, maxiter=args.iters, check_every=args.iters)
        _lib.check(lib.jp_event_record(e1.value))
        ms = C.c_double()
        _lib.check(lib.jp_event_elapsed_ms(e0.value, e1.value, C.byref(ms)))
        cg["device_scalars"] = {"iterations": its, "ms_per_iteration": ms.value / its, "rel_residual": rel,
                                "algorithmic_gbs": (nnz * 12 + (n + 1) * 4 + n * 8 * 11) / (ms.value / its) / 1e6}
    except Exception as e:  # noqa: BLE001 -- keep the other measurements
        cg["device_scalars"] = {"error": str(e)}
    out["cg"] = cg

    # ---- the building blocks one by one (ms per call, device events; dot/norm include their host synchronisation) ----
    def timed(fn, reps=50):
        for _ in range(5):
            fn()
        lib.jp_synchronize()
        _lib.check(lib.jp_event_record(e0.value))
        for _ in range(reps):
            fn()
        _lib.check(lib.jp_event_record(e1.value))
        ms = C.c_double()
        _lib.check(lib.jp_event_elapsed_ms(e0.value, e1.value, C.byref(ms)))
        return ms.value / reps

    P = jp.DenseMultiVector(m, xs)
    Ap = jp.DenseMultiVector(m, 1)
    R = jp.DenseMultiVector(m, synthetic.vector_block(1, n, 1, seed=9))
    out["ops_ms"] = {
        "apply": timed(lambda: Ah.apply_(Ap, P)),
        "apply_then_dot": timed(lambda: (Ah.apply_(Ap, P), P.dot(Ap))),
        "apply_dot": timed(lambda: Ah.apply_dot_(Ap, P)),
        "dot": timed(lambda: P.dot(Ap)),
        "norm2": timed(lambda: R.norm(2)),
        "axpby": timed(lambda: R.axpby_(-1e-9, Ap, 1.0)),
        "axpby_then_norm2": timed(lambda: (R.axpby_(-1e-9, Ap, 1.0), R.norm(2))),
        "axpby_norm2": timed(lambda: R.axpby_norm2_(-1e-9, Ap, 1.0)),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
