#!/bin/bash
# gpurun --gpus 8: the driver's exact scaling command at N=8 on the final code (boundary CTAs behind 75 % of the interior
# blocks, twice; at the tail of the grid as in the first 8-GPU call, once) and the CG-style loop of BASELINE configs[4].
set -u
mkdir -p gpurun_out
T=${1:-r2n8f}
run() { local name=$1; shift
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
      bench.py --gpus 8 "$@" > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
  echo "$name rc=$?" >> gpurun_out/${T}_rc.log
}
run n8_20a --steps 20 --warmup 5
run n8_20b --steps 20 --warmup 5
JP_FUSED_BOUNDARY_AT=100 run n8_20_tail --steps 20 --warmup 5 --no-e2e
run cg512_n8_dev --workload cg-512 --steps 100 --warmup 5 --device-scalars
cat gpurun_out/${T}_rc.log
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],1), 'med/max', d['per_step'] and (round(d['per_step']['median_ms'],5), round(d['per_step']['max_ms'],5)), 'parity', d.get('parity'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 6 gpurun_out/${T}_*.err | grep -v "^\*\|OMP_NUM\|^$" | cut -c1-300 | tail -20
