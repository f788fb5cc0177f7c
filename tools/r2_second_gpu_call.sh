#!/bin/bash
# The first MULTI-GPU call of the next round (gpurun --gpus 2, from the repo root): the multi-rank parity suite including
# the paths written after round 1's GPU budget was spent (exporter branch, fused halo + SpMV + dot kernel, device
# fillComplete on real ranks), then the CG-style bench with and without the fused call.
set -u
mkdir -p gpurun_out
JP_TEST_UNVALIDATED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    tests/multi_gpu_cases.py > gpurun_out/r2_second_multi_unvalidated.log 2>&1; echo "rc=$?" >> gpurun_out/r2_second_multi_unvalidated.log
timeout 300 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2_second_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2_second_pytest_multi.log
for fused in "" "--fused"; do
  JP_FUSED_DOT=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus 2 --workload cg-256 --steps 50 --warmup 5 $fused > gpurun_out/r2_second_cg256_n2${fused}.json 2> gpurun_out/r2_second_cg256_n2${fused}.err
done
tail -n 2 gpurun_out/r2_second_*.log | cut -c1-200
