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
Other workloads (parity/profiling cases, not the bench line): --workload 5pt-1000 | 27pt-192 --k 8|32 | powerlaw --rows R
"""
import argparse
import ctypes as C
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
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
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
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, spec):
    """--impl reference: the reference's CPU algorithm (oracle port: the Julia sources cannot run here) as P simulated MPI
    ranks on P host threads, one rank per core like `mpiexec -n P`, halo through the restated Import plan."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    from oracle.problems import oracle_problem
    cores = os.cpu_count() or 1
    P = args.ref_ranks or min(cores, 32)
    t0 = time.time()
    c, m, A = oracle_problem(spec["kind"], spec["N"], P)
    nnz = sum(len(A.csr(p)[2]) for p in range(1, P + 1)) if spec["n"] <= 1 << 22 else None
    if nnz is None:
        from jpetra_b200 import synthetic
        nnz = synthetic.stencil_nnz(spec["kind"], spec["N"]) if spec["kind"] != "powerlaw" else sum(len(A.csr(p)[2]) for p in range(1, P + 1))
    from jpetra_b200 import synthetic
    xs = []
    for pid in range(1, P + 1):
        info = m.info(pid)
        xs.append(synthetic.vector_block(info["minMyGID"], info["numMyElements"], args.k, seed=1))
    X = O.OMV.from_arrays(m, xs)
    Y = O.OMV(m, args.k)
    setup_s = time.time() - t0
    iters = max(1, min(args.steps, 5))
    secs = A.time_apply(Y, X, 1.0, 0.0, iters=iters, warmup=min(args.warmup, 1))
    val = 2.0 * nnz * args.k / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": iters, "warmup": min(args.warmup, 1),
        "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "n": spec["n"], "nnz": nnz, "k": args.k, "alpha": 1.0, "beta": 0.0},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": P, "kind": "port",
                         "sample": f"{iters} full apply! on the whole matrix, {P} simulated MPI ranks on {P} threads of {cores} host cores; "
                                   f"oracle setup {setup_s:.0f}s not timed"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hbm_gbs": algorithmic_bytes(spec["n"], nnz, args.k) / secs / 1e9,
    }
    print(json.dumps(line), flush=True)


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

    def step():
        if cg and args.fused:
            pAp = A.apply_dot_(Y, X)  # jp_apply_dot: no host round trip between apply! and dot
        else:
            _lib.check(lib.jp_apply(comm.handle, Y.h, A.h, X.h, plan, jp.NO_TRANS, 1.0, 0.0))
        if cg:
            if not args.fused:
                pAp = X.dot(Y)
            R.norm(2)
            XS.axpby_(1e-3 / (abs(pAp[0, 0]) + 1.0), X, 1.0)

    ev0, ev1 = C.c_int64(), C.c_int64()
    _lib.check(lib.jp_event_create(C.byref(ev0)))
    _lib.check(lib.jp_event_create(C.byref(ev1)))
    cnt0, cnt1 = C.c_int64(), C.c_int64()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    lib.jp_launch_count(C.byref(cnt0))
    # L2 rule: either the per-GPU working set of one apply is well above the 126 MB L2 (no flush), or a 256 MB
    # scratch buffer is written between timed iterations and every iteration is timed on its own event pair
    bytes_rank_est = algorithmic_bytes(n_my, nnz_my, k)
    flush = bytes_rank_est < 1.5 * 126e6
    ms = C.c_double()
    host_enqueue_ms = None
    if not flush:
        _lib.check(lib.jp_event_record(ev0.value))
        t_h = time.perf_counter()
        for _ in range(args.steps):
            step()
        host_enqueue_ms = (time.perf_counter() - t_h) * 1e3 / args.steps  # CPU time to enqueue one step (no sync inside)
        _lib.check(lib.jp_event_record(ev1.value))
        _lib.check(lib.jp_event_elapsed_ms(ev0.value, ev1.value, C.byref(ms)))
        total_ms = ms.value
    else:
        total_ms = 0.0
        for _ in range(args.steps):
            _lib.check(lib.jp_flush_l2(256 << 20))
            _lib.check(lib.jp_event_record(ev0.value))
            step()
            _lib.check(lib.jp_event_record(ev1.value))
            _lib.check(lib.jp_event_elapsed_ms(ev0.value, ev1.value, C.byref(ms)))
            total_ms += ms.value
    ms.value = total_ms
    lib.jp_launch_count(C.byref(cnt1))
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = max_over_ranks(ms.value)
    ms_per_step = elapsed_ms / args.steps
    clocks_window = "timed region"
    if rank == 0 and (clocks is None or clocks.get("samples", 0) < 5):
        # the timed region was too short for nvidia-smi's sampling period: sample the same kernel loop for ~1.5 s
        sampler = ClockSampler(local)
        t_end = time.time() + 1.5
    else:
        t_end = 0
    if world > 1:
        import torch
        flag = torch.tensor([1.0 if t_end else 0.0], device="cuda")
        dist.broadcast(flag, 0)
        loop = bool(flag.item())
    else:
        loop = bool(t_end)
    if loop:
        reps = max(1, int(1.5e3 / max(ms_per_step, 1e-3)))
        for _ in range(reps):
            step()
        barrier()
        if rank == 0:
            clocks = sampler.stop()
            clocks_window = "same apply! loop repeated for ~1.5 s right after the timed region (timed region shorter than the sampling period)"
    launches = int(sum_over_ranks(float(cnt1.value - cnt0.value)))
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
    bytes_rank = algorithmic_bytes(n_my, nnz_my, k) + (48 * n_my * k if cg else 0)
    achieved = bytes_rank / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region) ----
    e2e = None
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
        y_e2e = np.ctypeslib.as_array(C.cast(py, C.POINTER(C.c_double)), shape=(n_my * k,)).copy()
        assert np.array_equal(y_e2e.reshape((n_my, k), order="F"), Y.data), "host-buffer path disagrees with the device path"
        e2e = {"value": 2.0 * nnz * k / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": n * k * 8,
               "d2h_bytes_per_step": n * k * 8, "host_memory": "pinned (jp_host_alloc)"}
        lib.jp_host_free(px)
        lib.jp_host_free(py)

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
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["name"], "n": n, "nnz": nnz, "k": k, "alpha": 1.0, "beta": 0.0, "parallelism": f"rows/{world} (BlockMap)",
                       "l2": (f"L2 flushed (256 MB memset) between timed iterations, each iteration on its own event pair; {bytes_rank / 1e6:.0f} MB per GPU per apply"
                              if flush else f"inputs larger than L2: {bytes_rank / 1e6:.0f} MB streamed per GPU per apply vs 126 MB L2; no flush needed"),
                       "setup_s": round(setup_s, 1)},
            "hbm_gbs": bytes_apply / (ms_per_step * 1e-3) / 1e9, "host_enqueue_ms_per_step": host_enqueue_ms, "host_enqueue_burst_ms_per_step": host_burst_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0, "bytes_per_launch": bytes_rank,
                         "kernel": "spmv_kernel (one launch per apply at N=1; at N>1 the same row-block code inside spmv_fused_halo_kernel, one launch per apply)"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": dict(clocks or {}, window=clocks_window) if clocks else None,
        }
        print(json.dumps(line), flush=True)
    barrier()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
