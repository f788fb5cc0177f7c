#!/bin/bash
# gpurun --gpus 2: multi-rank parity suite (everything un-gated), then the driver's exact scaling command at N=1 and N=2
# (20 steps / 5 warm-up) twice, a long run, the two-stream variant, and the CG loop in its three forms.
set -u
mkdir -p gpurun_out
T=${1:-r2n2}
run() { # name, nproc, args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then
    timeout 150 python bench.py --gpus 1 "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  else
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus $np "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  fi
  echo "$name rc=$?" >> gpurun_out/${T}_rc.log
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    tests/multi_gpu_cases.py > gpurun_out/${T}_multi_cases.log 2>&1; echo "multi_cases rc=$?" >> gpurun_out/${T}_rc.log
run n1_20 1 --steps 20 --warmup 5
run n2_20a 2 --steps 20 --warmup 5
run n2_20b 2 --steps 20 --warmup 5
run n2_1000 2 --steps 1000 --warmup 20 --no-e2e
JP_FUSED=0 run n2_twostream 2 --steps 200 --warmup 20 --no-e2e
run cg256_n2_plain 2 --workload cg-256 --steps 50 --warmup 5
run cg256_n2_fused 2 --workload cg-256 --steps 50 --warmup 5 --fused
run cg256_n2_dev 2 --workload cg-256 --steps 50 --warmup 5 --device-scalars
cat gpurun_out/${T}_rc.log
tail -n 3 gpurun_out/${T}_multi_cases.log | cut -c1-300
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],1), 'per_step', d.get('per_step'), 'parity', d.get('parity'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 5 gpurun_out/${T}_*.err | cut -c1-300 | tail -40
