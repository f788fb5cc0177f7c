#!/bin/bash
# ONE GPU: ncu --set full of the two kernels written last (final diagonal SpMM kernel, column-slab SpMV kernel).
set -u
mkdir -p gpurun_out
T=${1:-r2x}
timeout 150 ncu --set full --clock-control none --import-source on -k regex:spmm_dia -s 4 -c 1 -f -o gpurun_out/${T}_spmm_dia_final \
    python tools/bench_spmm.py --N 128 --ks 8 --variants default --reps 3 > gpurun_out/${T}_ncu_dia.log 2>&1; echo "dia rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:spmv_cb_kernel -s 20 -c 3 -f -o gpurun_out/${T}_spmv_cb_10M \
    python bench.py --workload powerlaw --rows 10000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_cb.log 2>&1; echo "cb rc=$?"
tail -n 3 gpurun_out/${T}_ncu_*.log | cut -c1-200
