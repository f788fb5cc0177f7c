#!/bin/bash
# gpurun --gpus 8: the multi-rank parity suite on 8 ranks, the driver's exact scaling command at N=8 (20 steps / 5 warm-up),
# and the CG-style loop of BASELINE configs[4] (512^3 over 8 GPUs) with fused apply+dot and with device-resident scalars.
set -u
mkdir -p gpurun_out
T=${1:-r2n8}
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $np "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  echo "$name rc=$?" >> gpurun_out/${T}_rc.log
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
    tests/multi_gpu_cases.py > gpurun_out/${T}_multi_cases.log 2>&1; echo "multi_cases rc=$?" >> gpurun_out/${T}_rc.log
run n8_20a 8 --steps 20 --warmup 5
run n8_20b 8 --steps 20 --warmup 5
run cg512_n8_dev 8 --workload cg-512 --steps 100 --warmup 5 --device-scalars
run cg512_n8_fused 8 --workload cg-512 --steps 100 --warmup 5 --fused
cat gpurun_out/${T}_rc.log
tail -n 3 gpurun_out/${T}_multi_cases.log | cut -c1-300
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    p=d.get('parity') or {}
    print(' ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],1), 'med/max', d['per_step'] and (round(d['per_step']['median_ms'],5), round(d['per_step']['max_ms'],5)), 'parity', p, 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 4 gpurun_out/${T}_*.err | grep -v "^\*\|OMP_NUM\|^$" | cut -c1-300 | tail -30
