#!/usr/bin/env python
"""Device-timed bandwidth of the MultiVector kernels (dot, norm2, axpby, scale!, fill!)
on one GPU.  Prints one JSON object.  dot/norm return host values, so each call includes the k-double D2H copy and
the stream synchronisation the API requires; the elementwise kernels are pure stream work."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jpetra_b200 as jp  # noqa: E402
from jpetra_b200 import _lib, synthetic  # noqa: E402


def timed(fn, reps, lib):
    e0, e1 = C.c_int64(), C.c_int64()
    _lib.check(lib.jp_event_create(C.byref(e0)))
    _lib.check(lib.jp_event_create(C.byref(e1)))
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


def main():
    lib = _lib.init(0)
    comm = jp.SerialComm()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = {"peak_gbs": peak, "cases": []}
    for n, k in ((1 << 24, 1), (1 << 24, 8), (1 << 27, 1)):
        m = jp.BlockMap(n, comm)
        X = jp.DenseMultiVector(m, synthetic.vector_block(1, n, k, seed=1))
        Y = jp.DenseMultiVector(m, synthetic.vector_block(1, n, k, seed=2))
        reps = 200 if n * k <= 1 << 27 else 50
        S = jp.Scalars(2)
        ops = {
            "dot (16 B/elem, host result: hand-over + host round trip per call)": (lambda: X.dot(Y), 16),
            "norm2 (8 B/elem, host result: hand-over + host round trip per call)": (lambda: X.norm(2), 8),
            "axpby y=a*x+b*y (24 B/elem)": (lambda: Y.axpby_(0.5, X, 0.5), 24),
            "scale! (16 B/elem)": (lambda: Y.scale_(1.0000001), 16),
            "fill! (8 B/elem)": (lambda: Y.fill_(1.5), 8),
        }
        if k == 1:  # the same reduction kernel with the result left on the device (jp_mv_dot_dev): kernel time, no host round trip
            ops["dot, device-resident result (16 B/elem, kernel only)"] = (lambda: _lib.check(lib.jp_mv_dot_dev(comm.handle, X.h, Y.h, S.h, 0)), 16)
            ops["norm2 as dot(x, x), device-resident result (8 B/elem of HBM traffic, kernel only)"] = (
                lambda: _lib.check(lib.jp_mv_dot_dev(comm.handle, X.h, X.h, S.h, 1)), 8)
        for name, (fn, bpe) in ops.items():
            ms = timed(fn, reps, lib)
            gbs = bpe * n * k / (ms * 1e-3) / 1e9
            out["cases"].append({"n": n, "k": k, "op": name, "ms": ms, "gbs": gbs, "frac_of_measured_peak": gbs / peak})
        del X, Y
    print(json.dumps(out))


if __name__ == "__main__":
    main()
