#!/bin/bash
set -u
mkdir -p gpurun_out
for ch in 4 8 12 16; do
  JP_HOST_CHUNKS=$ch timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 30 > gpurun_out/r2w_chunks$ch.json 2> gpurun_out/r2w_chunks$ch.err
  python - gpurun_out/r2w_chunks$ch.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print(sys.argv[1], 'e2e ms', d['e2e']['ms_per_step'], 'parity', d['parity']['ok'])
PY
done
