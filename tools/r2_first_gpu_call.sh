#!/bin/bash
# The first GPU call of the next round (run through gpurun from the repo root, ONE GPU):
#   1. the device paths written after round 1's GPU budget was spent (exporter branch of apply!), opt-in tests
#   2. the whole single-GPU suite
#   3. ncu --set full of the run-staged SpMM kernel as it is now (TMA tile), of the gather kernel and of spmv_dot_kernel
#   4. SpMM A/B at the BASELINE size, fillComplete + CG building blocks at 256^3
#   5. power-law SpMV with three DIRECT-block rules
# Everything lands in gpurun_out/r2_first_*; budget ~6 GPU-minutes.
set -u
mkdir -p gpurun_out
JP_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_gpu_exporter.py tests/test_gpu_parity.py tests/test_c_abi_from_c.py tests/test_gpu_fused.py -k 'exporter or double_buffered or c_program or device_resident' -q > gpurun_out/r2_first_pytest_exporter.log 2>&1; echo "rc=$?" >> gpurun_out/r2_first_pytest_exporter.log
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r2_first_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_first_pytest_gpu.log
timeout 120 python tools/bench_spmm.py --N 192 --ks 8,32 --reps 10 --variants gather,default,runs_kt4_t256,runs_db,runs_db_rows64_t256 > gpurun_out/r2_first_spmm192.json 2> gpurun_out/r2_first_spmm192.err
timeout 90 python tools/bench_next_rows.py --N 256 --iters 50 > gpurun_out/r2_first_next_rows_256.json 2> gpurun_out/r2_first_next_rows_256.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:spmm_runs -s 3 -c 1 -f -o gpurun_out/r2_first_spmm_runs_tma \
    python tools/bench_spmm.py --N 128 --ks 8 --variants default --reps 3 > gpurun_out/r2_first_ncu_runs.log 2>&1
timeout 120 ncu --set full --clock-control none -k regex:spmv_dot -s 3 -c 1 -f -o gpurun_out/r2_first_spmv_dot \
    python tools/bench_next_rows.py --N 128 --iters 5 > gpurun_out/r2_first_ncu_dot.log 2>&1
# power-law SpMV: which blocks should go thread-per-row (DIRECT) instead of through the shared-memory two-phase path?
for rule in "32 2" "64 8" "2048 1000"; do
  set -- $rule
  JP_SPMV_DIRECT_MAXLEN=$1 JP_SPMV_DIRECT_RATIO=$2 timeout 120 python bench.py --workload powerlaw --rows 10000000 --steps 30 --warmup 5 \
      --no-cpu-baseline --no-e2e > gpurun_out/r2_first_powerlaw10M_direct_$1_$2.json 2> gpurun_out/r2_first_powerlaw10M_direct_$1_$2.err
done
tail -n 3 gpurun_out/r2_first_pytest_exporter.log gpurun_out/r2_first_pytest_gpu.log | cut -c1-200
