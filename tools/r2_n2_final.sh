#!/bin/bash
# gpurun --gpus 2: the multi-rank parity suite on the final code, the driver's scaling command at N=2, and the CG-style
# loop with device-resident scalars (its parity block checks Y, p'Ap, |r|^2 against the oracle's 2-rank simulation).
set -u
mkdir -p gpurun_out
T=${1:-r2n2f}
run() { local name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus 2 "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  echo "$name rc=$?" >> gpurun_out/${T}_rc.log
}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    tests/multi_gpu_cases.py > gpurun_out/${T}_multi_cases.log 2>&1; echo "multi_cases rc=$?" >> gpurun_out/${T}_rc.log
run n2_20 --steps 20 --warmup 5
run cg256_n2_dev --workload cg-256 --steps 50 --warmup 5 --device-scalars
cat gpurun_out/${T}_rc.log
tail -n 4 gpurun_out/${T}_multi_cases.log | cut -c1-400
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],1), 'parity', d.get('parity'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 6 gpurun_out/${T}_*.err | grep -v "^\*\|OMP_NUM\|^$" | cut -c1-300 | tail -20
