#!/usr/bin/env python
"""bench.py -- CSRMatrix apply! throughput on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's CPU algorithm on the host cores)

A "step" is one apply!(Y, A, X) (alpha=1, beta=0) on the 3-D 7-point Laplacian 256^3 (BASELINE configs[1]), rows
partitioned by BlockMap(N^3, comm) (strong scaling), halo exchange included.  Rank 0 prints ONE JSON line.
  value  : whole-job GFLOP/s (2*nnz per apply) with X, Y, A resident in HBM, CUDA-event timed, max over ranks
  e2e    : same metric through the host-buffer entry point jp_apply_host (pinned host X -> H2D -> kernels -> D2H Y)
  roofline: algorithmic bytes of one apply / measured duration vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline: the oracle's literal restatement of localApply on ONE host core, same matrix (rank 0, N=1 only)
  parity : after the timed loop every rank hashes ITS slab of Y and compares it with the committed digest of the CPU
           oracle's simulation of the reference on the same rank count (tests/golden/bench_parity.json), checks dot / norm
           (NCCL allreduce), the transposed apply, commReduce and the Comm collectives; any mismatch exits non-zero
Timed window: W warm-up steps, barrier, a few untimed ALIGNING steps (the ranks enter the window paced by the halo
handshake itself and the host's launch queue is ahead of the device), event, EXACTLY K steps, event; nothing else -- no
fork, no host synchronisation -- happens between the barrier and the second event.  A second pass of K steps with one
event pair per step gives the median / max step (diagnostics, not the reported value).
Other workloads (parity/profiling cases, not the bench line): --workload 5pt-1000 | 27pt-192 --k 8|32 | powerlaw --rows R
"""
import argparse
import ctypes as C
import datetime
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "CSRMatrix apply! GFLOP/s"
UNIT = "GFLOP/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="7pt-256")
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--rows", type=int, default=50_000_000, help="rows of the powerlaw workload")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-iters", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--fused", action="store_true", help="cg workloads: apply! + dot through the fused jp_apply_dot call")
    ap.add_argument("--device-scalars", action="store_true", help="cg workloads: dot / norm results stay on the device (jp_*_dev): no host sync per iteration")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--trans", action="store_true", help="time apply!(Y, A, X, TRANS) instead (profiling case: the cached transposed matrix)")
    ap.add_argument("--ref-ranks", type=int, default=0, help="--impl reference: simulated MPI ranks (default: host cores, max 32)")
    return ap.parse_args()


def workload_spec(args):
    w = args.workload
    if w == "powerlaw":
        return dict(kind="powerlaw", N=args.rows, n=args.rows, name=f"random power-law sparse matrix, {args.rows} rows, mean 16, max 10k")
    kind, size = w.split("-")
    N = int(size)
    if kind == "cg":  # BASELINE configs[4]: CG-style loop on the 3-D 7-point Laplacian
        return dict(kind="7pt", N=N, n=N ** 3, cg=True,
                    name=f"CG-style loop (apply! + dot + norm2 + axpy per iteration) on the 3-D 7-point Laplacian {N}^3")
    dim = 2 if kind == "5pt" else 3
    names = {"5pt": "2-D 5-point", "7pt": "3-D 7-point", "27pt": "3-D 27-point"}
    return dict(kind=kind, N=N, n=N ** dim, name=f"{names[kind]} Laplacian {N}^{dim}, rows partitioned by BlockMap, SpMV + Import halo")


def algorithmic_bytes(n, nnz, k, beta_zero=True):
    """nnz*(8+4) + (n+1)*4 + n*k*8*(1 or 2) + n*k*8"""
    return nnz * 12 + (n + 1) * 4 + n * k * 8 * (1 if beta_zero else 2) + n * k * 8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampling every 50 ms in a child process.  Started BEFORE the first barrier (a fork of a process with GBs
    of device mappings takes milliseconds: it must never sit between a barrier and the timed window); samples are
    time-stamped so that only those taken while the kernels ran are reported."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, windows):
        """windows: [(t0, t1)] wall-clock intervals (time.time()) during which the measured kernels were running"""
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                t = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if not any(t0 - 0.05 <= t <= t1 + 0.05 for t0, t1 in windows):
                    continue
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def bench_config(spec, nnz, k, n_gpus):
    """the `config` object of the JSON line: identical in the product arm and the reference arm"""
    per_gpu = algorithmic_bytes(spec["n"], nnz, k) / n_gpus
    flush = per_gpu < 1.5 * 126e6
    return {"workload": spec["name"], "n": spec["n"], "nnz": nnz, "k": k, "alpha": 1.0, "beta": 0.0,
            "parallelism": f"rows/{n_gpus} (BlockMap)",
            "l2": (f"L2 flushed (256 MB memset) between timed iterations, each iteration on its own event pair; {per_gpu / 1e6:.0f} MB per GPU per apply"
                   if flush else f"inputs larger than L2: {per_gpu / 1e6:.0f} MB streamed per GPU per apply vs 126 MB L2; no flush needed")}


def run_reference(args, spec):
    """--impl reference: the reference's CPU algorithm (oracle port: the Julia sources cannot run here) as P simulated MPI
    ranks on P host threads, one rank per core like `mpiexec -n P`, halo through the restated Import plan.  Honours
    --steps / --warmup; P = all host cores (at most 32) whatever --gpus says, which is generous to the CPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    from oracle.problems import oracle_problem
    from jpetra_b200 import synthetic
    cores = os.cpu_count() or 1
    P = args.ref_ranks or min(cores, 32)
    t0 = time.time()
    c, m, A = oracle_problem(spec["kind"], spec["N"], P)
    if spec["kind"] != "powerlaw":
        nnz = synthetic.stencil_nnz(spec["kind"], spec["N"])
    else:
        nnz = sum(len(A.csr(p)[2]) for p in range(1, P + 1))
    xs = []
    for pid in range(1, P + 1):
        info = m.info(pid)
        xs.append(synthetic.vector_block(info["minMyGID"], info["numMyElements"], args.k, seed=1))
    X = O.OMV.from_arrays(m, xs)
    Y = O.OMV(m, args.k)
    setup_s = time.time() - t0
    iters, warm = max(1, args.steps), max(0, args.warmup)
    secs = A.time_apply(Y, X, 1.0, 0.0, iters=iters, warmup=warm)
    val = 2.0 * nnz * args.k / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": iters, "warmup": warm,
        "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(spec, nnz, args.k, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": P, "kind": "port",
                         "sample": f"{iters} full apply! on the whole matrix ({warm} warm-up), {P} simulated MPI ranks on {P} threads of {cores} host cores "
                                   f"(all host cores, at most 32, at every --gpus); oracle setup {setup_s:.0f}s not timed"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hbm_gbs": algorithmic_bytes(spec["n"], nnz, args.k) / secs / 1e9,
    }
    print(json.dumps(line), flush=True)


def sha256_of(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parity_checks(jp, comm, A, X, Y, x_host, spec, k, world, rank, e2e_y, cg_scalars=None, R=None):
    """Run on every rank after the timed loop; returns (ok, report).  Everything here goes through the same C-ABI entry
    points as the timed loop: Y is what the LAST timed apply left on the device."""
    rep = {}
    ok = True
    fx = None
    try:
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_parity.json")))["cases"].get(f"{spec['kind']}-{spec['N']}/P{world}")
    except Exception:
        fx = None
    y = Y.data
    pid, P = comm.myPid(), comm.numProc()
    if fx is not None and k == 1:
        mine = sha256_of(y) == fx["y_sha256"][rank]
        rep["y_bit_exact_ranks"] = int(comm.sumAll([1 if mine else 0])[0])
        ok &= rep["y_bit_exact_ranks"] == P
        if e2e_y is not None:
            rep["e2e_y_bit_exact_ranks"] = int(comm.sumAll([1 if sha256_of(e2e_y) == fx["y_sha256"][rank] else 0])[0])
            ok &= rep["e2e_y_bit_exact_ranks"] == P
        d = float(X.dot(Y)[0, 0])
        nr = float(Y.norm(2)[0, 0])
        rep["dot_rel_err"] = abs(d - fx["dot_x_y"]) / abs(fx["dot_x_y"])
        rep["norm2_rel_err"] = abs(nr - fx["norm2_y"]) / fx["norm2_y"]
        rep["tolerance"] = 1e-12
        ok &= rep["dot_rel_err"] <= 1e-12 and rep["norm2_rel_err"] <= 1e-12
        if R is not None and "norm2_r" in fx:  # the residual vector of the CG-style loop
            rep["norm2_r_rel_err"] = abs(float(R.norm(2)[0, 0]) - fx["norm2_r"]) / fx["norm2_r"]
            ok &= rep["norm2_r_rel_err"] <= 1e-12
        if cg_scalars is not None and "norm2_r" in fx:  # what the device-resident-scalar loop left on the device: p'Ap, |r|^2
            rep["dev_pAp_rel_err"] = abs(cg_scalars[0] - fx["dot_x_y"]) / abs(fx["dot_x_y"])
            rep["dev_rr_rel_err"] = abs(cg_scalars[1] - fx["norm2_r"] ** 2) / fx["norm2_r"] ** 2
            ok &= rep["dev_pAp_rel_err"] <= 1e-12 and rep["dev_rr_rel_err"] <= 1e-12
    if spec["kind"] in ("5pt", "7pt", "27pt"):
        # A is symmetric: the transposed apply (scatter + reverse-mode halo exchange with ADD) must reproduce A*X within the
        # reduction-order tolerance, component-wise: |Yt - Y|_i <= 1e-12 * max(|Y_i|, diag*|x_i|) <= 1e-12 * (|A||x|)_i
        Yt = jp.DenseMultiVector(Y.getMap(), k)
        A.apply_(Yt, X, jp.TRANS, 1.0, 0.0)
        diag = float(jp.synthetic.STENCILS[spec["kind"]][1])
        bound = 1e-12 * np.maximum(np.abs(y), diag * np.abs(x_host))
        bad = int(comm.sumAll([int((np.abs(Yt.data - y) > bound).sum())])[0])
        rep["trans_symmetric_violations"] = bad
        ok &= bad == 0
    # commReduce on a replicated map (test/DenseMultiVectorTests.jl:96-99) and the Comm collectives (test/MPICommTests.jl:13-36)
    rep_map = jp.BlockMap(6, 6, comm)
    v = jp.DenseMultiVector(rep_map, (2.0 ** pid) * np.ones((6, 2)))
    v.commReduce()
    coll = np.array_equal(v.data, float(sum(2 ** i for i in range(1, P + 1))) * np.ones((6, 2)))
    everyone = list(range(1, P + 1))
    coll &= comm.gatherAll([pid, 2 * pid]).tolist() == [x for q in everyone for x in (q, 2 * q)]
    coll &= comm.sumAll([pid, 3]).tolist() == [P * (P + 1) // 2, 3 * P] and comm.maxAll([pid]).tolist() == [P] and comm.minAll([pid]).tolist() == [1]
    coll &= comm.scanSum([pid]).tolist() == [pid * (pid + 1) // 2] and comm.broadcastAll([pid, 7], P).tolist() == [P, 7]
    rep["collectives_ok_ranks"] = int(comm.sumAll([1 if coll else 0])[0])
    ok &= rep["collectives_ok_ranks"] == P
    rep["checked"] = True
    rep["bit_exact"] = bool(rep.get("y_bit_exact_ranks") == P) if "y_bit_exact_ranks" in rep else None
    rep["ok"] = bool(ok)
    rep["oracle"] = "tests/golden/bench_parity.json (CPU oracle's P-rank simulation of the reference)" if fx is not None else None
    return bool(ok), rep


def main():
    args = parse_args()
    spec = workload_spec(args)
    if args.impl == "reference":
        run_reference(args, spec)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import jpetra_b200 as jp
    from jpetra_b200 import _lib, synthetic
    lib = _lib.init(local)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = jp.NCCLComm.fromTorchDistributed()
    else:
        comm = jp.SerialComm()
    # the only fork of this process: long before any timed window
    sampler = ClockSampler(local) if rank == 0 else None
    windows = []

    def barrier():
        lib.jp_synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- problem setup (host mirror of the reference API; one-time, untimed) ----
    t_setup = time.time()
    k = args.k
    row_map = jp.BlockMap(spec["n"], comm)
    A = synthetic.build_matrix(jp, row_map, spec["kind"], spec["N"])
    n_my = row_map.numMyElements()
    nnz_my = len(A.values)
    nnz = int(sum_over_ranks(float(nnz_my)))
    n = spec["n"]
    x_host = synthetic.vector_block(row_map.minMyGID(), n_my, k, seed=1)
    X = jp.DenseMultiVector(row_map, x_host)
    Y = jp.DenseMultiVector(row_map, k)
    plan = A.importer.device_plan() if A.importer is not None else 0
    setup_s = time.time() - t_setup

    cg = bool(spec.get("cg"))
    if cg:
        # one iteration of the CG-style loop of BASELINE configs[4]: Ap = A*p; dot(p, Ap); norm(r, 2); x = x + a*p
        R = jp.DenseMultiVector(row_map, synthetic.vector_block(row_map.minMyGID(), n_my, k, seed=2))
        XS = jp.DenseMultiVector(row_map, k)
        args.no_e2e = True
        if args.device_scalars:
            S = jp.Scalars(2)  # [0] p'Ap, [1] |r|^2 -- never downloaded inside the loop

    def step():
        if cg and args.device_scalars:
            # the same operations with every scalar left on the device: no host synchronisation inside the iteration
            _lib.check(lib.jp_apply_dot_dev(comm.handle, Y.h, A.h, X.h, plan, 1.0, 0.0, X.h, S.h, 0))  # Ap = A p; s[0] = p'Ap (allreduce on the stream)
            _lib.check(lib.jp_mv_dot_dev(comm.handle, R.h, R.h, S.h, 1))                                # s[1] = |r|^2
            _lib.check(lib.jp_mv_axpby_dev(XS.h, X.h, S.h, 1e-3, 1, 0, 1.0, -1, -1))                    # x += (1e-3 * s[1] / s[0]) p
            return
        if cg and args.fused:
            pAp = A.apply_dot_(Y, X)  # jp_apply_dot: no host round trip between apply! and dot
        else:
            _lib.check(lib.jp_apply(comm.handle, Y.h, A.h, X.h, plan, jp.TRANS if args.trans else jp.NO_TRANS, 1.0, 0.0))
        if cg:
            if not args.fused:
                pAp = X.dot(Y)
            R.norm(2)
            XS.axpby_(1e-3 / (abs(pAp[0, 0]) + 1.0), X, 1.0)

    def new_event():
        e = C.c_int64()
        _lib.check(lib.jp_event_create(C.byref(e)))
        return e.value

    def elapsed(e0, e1):
        ms = C.c_double()
        _lib.check(lib.jp_event_elapsed_ms(e0, e1, C.byref(ms)))  # synchronises on e1
        return ms.value

    ev0, ev1 = new_event(), new_event()
    step_events = [new_event() for _ in range(args.steps + 1)]
    cnt0, cnt1 = C.c_int64(), C.c_int64()

    # L2 rule: either the per-GPU working set of one apply is well above the 126 MB L2 (no flush), or a 256 MB
    # scratch buffer is written between timed iterations and every iteration is timed on its own event pair
    bytes_rank_est = algorithmic_bytes(n_my, nnz_my, k)
    flush = algorithmic_bytes(n, nnz, k) / world < 1.5 * 126e6
    warmup = max(args.warmup, 3)
    _lib.check(lib.jp_event_record(ev0))
    for _ in range(warmup):
        step()
    _lib.check(lib.jp_event_record(ev1))
    est_ms = max(max_over_ranks(elapsed(ev0, ev1)) / warmup, 1e-3)  # the SAME number on every rank: applies are collective
    # aligning steps: ~20 ms of the same step (at least 10, at most 512 = well inside the launch queue), so the clocks are
    # up, every rank is paced by its neighbours' halo flags, and the host runs ahead of the device when the window opens
    align = 0 if flush else int(min(512, max(10, 20.0 / est_ms)))
    barrier()
    t_load0 = time.time()
    lib.jp_launch_count(C.byref(cnt0))
    host_enqueue_ms = None
    if not flush:
        for _ in range(align):
            step()
        lib.jp_launch_count(C.byref(cnt0))
        _lib.check(lib.jp_event_record(ev0))
        t_h = time.perf_counter()
        for _ in range(args.steps):
            step()
        host_enqueue_ms = (time.perf_counter() - t_h) * 1e3 / args.steps  # CPU time to enqueue one step (no sync inside)
        _lib.check(lib.jp_event_record(ev1))
        total_ms = elapsed(ev0, ev1)
    else:
        total_ms = 0.0
        for _ in range(args.steps):
            _lib.check(lib.jp_flush_l2(256 << 20))
            _lib.check(lib.jp_event_record(ev0))
            step()
            _lib.check(lib.jp_event_record(ev1))
            total_ms += elapsed(ev0, ev1)
    lib.jp_launch_count(C.byref(cnt1))
    barrier()
    windows.append((t_load0, time.time()))
    elapsed_ms = max_over_ranks(total_ms)
    ms_per_step = elapsed_ms / args.steps
    launches = int(sum_over_ranks(float(cnt1.value - cnt0.value)))

    # ---- diagnostics: the same K steps again, one event per step (not the reported value) ----
    per_step = None
    if not flush:
        for _ in range(min(align, 32)):
            step()
        _lib.check(lib.jp_event_record(step_events[0]))
        for i in range(args.steps):
            step()
            _lib.check(lib.jp_event_record(step_events[i + 1]))
        ts = [elapsed(step_events[i], step_events[i + 1]) for i in range(args.steps)]
        slow = int(np.argmax(ts))
        per_step = {"median_ms": max_over_ranks(statistics.median(ts)), "max_ms": max_over_ranks(max(ts)), "slowest_step_rank0": slow,
                    "total_ms": max_over_ranks(sum(ts)), "note": "second pass, one event pair per step (the events add launch gaps); diagnostics only"}
        barrier()

    # ---- clocks: the timed region is usually shorter than nvidia-smi's sampling period, so the same loop is repeated for
    #      ~1 s right behind it and the samples of both windows are reported ----
    clocks_window = "timed region"
    if elapsed_ms < 500.0:
        reps = max(1, int(1.0e3 / max(ms_per_step, 1e-3)))
        t_w = time.time()
        for i in range(reps):
            step()
            if i % 256 == 255:
                lib.jp_synchronize()
        barrier()
        windows.append((t_w, time.time()))
        clocks_window = "timed region + the same apply! loop repeated for ~1 s right after it (timed region shorter than the sampling period)"
    clocks = sampler.stop(windows) if sampler else None

    # pure CPU cost of enqueueing one step: a short burst into idle streams (well below the launch-queue depth)
    barrier()
    t_h = time.perf_counter()
    for _ in range(64):
        step()
    host_burst_ms = (time.perf_counter() - t_h) * 1e3 / 64
    barrier()

    flops_step = 2.0 * nnz * k + (6.0 * n * k if cg else 0.0)  # + dot, norm2, axpy: 2*n*k each
    value = flops_step / (ms_per_step * 1e-3) / 1e9
    bytes_apply = algorithmic_bytes(n, nnz, k) + (48 * n * k if cg else 0)  # dot 16, norm2 8, axpy 24 bytes per element
    peak, peak_src = measured_peak()
    # per-GPU roofline of the SpMV kernel: this rank's algorithmic bytes over the step duration
    bytes_rank = bytes_rank_est + (48 * n_my * k if cg else 0)
    achieved = bytes_rank / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region) ----
    e2e = None
    y_e2e = None
    if not args.no_e2e:
        px, py = C.c_void_p(), C.c_void_p()
        _lib.check(lib.jp_host_alloc(n_my * k * 8 + 64, C.byref(px)))
        _lib.check(lib.jp_host_alloc(n_my * k * 8 + 64, C.byref(py)))
        C.memmove(px, x_host.ctypes.data, n_my * k * 8)
        for _ in range(2):
            _lib.check(lib.jp_apply_host(comm.handle, A.h, plan, px, py, k, jp.NO_TRANS, 1.0, 0.0))
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            _lib.check(lib.jp_apply_host(comm.handle, A.h, plan, px, py, k, jp.NO_TRANS, 1.0, 0.0))  # synchronises: Y is on the host on return
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.e2e_steps
        y_e2e = np.ctypeslib.as_array(C.cast(py, C.POINTER(C.c_double)), shape=(n_my * k,)).copy().reshape((n_my, k), order="F")
        e2e = {"value": 2.0 * nnz * k / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * k * 8,
               "d2h_bytes_per_step": n * k * 8, "host_memory": "pinned (jp_host_alloc)", "steps": args.e2e_steps}
        lib.jp_host_free(px)
        lib.jp_host_free(py)

    # ---- parity of what the timed loop computed, at this rank count (every rank takes part) ----
    parity = {"checked": False}
    parity_ok = True
    if not args.no_parity:
        if not cg:
            _lib.check(lib.jp_apply(comm.handle, Y.h, A.h, X.h, plan, jp.NO_TRANS, 1.0, 0.0))  # Y of the very path that was timed
        cg_scalars = S.download() if (cg and args.device_scalars) else None
        parity_ok, parity = parity_checks(jp, comm, A, X, Y, x_host, spec, k, world, rank, y_e2e, cg_scalars, R if cg else None)

    # ---- CPU baseline: the oracle's localApply on one core, same matrix (rank 0 at N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        secs = O.time_local_apply_raw(A.rowOffsets, A.localIndices1D, A.values, x_host, 1.0, 0.0, iters=args.cpu_iters, warmup=1)
        cpu = {"value": 2.0 * nnz * k / secs / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "ms_per_step": secs * 1e3,
               "sample": f"{args.cpu_iters} full apply! of the same matrix (localApply restatement, src/CSRMatrix.jl:1131-1146), 1 thread = SerialComm",
               "host_cores": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(spec, nnz, k, world), "mode": "TRANS" if args.trans else "NO_TRANS",
            "setup_s": round(setup_s, 1), "aligning_steps": align, "per_step": per_step,
            "hbm_gbs": bytes_apply / (ms_per_step * 1e-3) / 1e9, "host_enqueue_ms_per_step": host_enqueue_ms, "host_enqueue_burst_ms_per_step": host_burst_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0, "bytes_per_launch": bytes_rank,
                         "kernel": "spmv_kernel (one launch per apply at N=1; at N>1 the same row-block code inside spmv_fused_halo_kernel, one launch per apply)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "parity": parity,
            "clocks": dict(clocks, window=clocks_window) if clocks else None,
        }
        print(json.dumps(line), flush=True)
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    if not parity_ok:
        sys.exit(3)


if __name__ == "__main__":
    main()
