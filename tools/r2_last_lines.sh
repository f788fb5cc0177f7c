#!/bin/bash
# ONE GPU: bench lines of the remaining BASELINE configs on the final code (configs[0] 5-point 1000^2, configs[2] k = 32).
set -u
mkdir -p gpurun_out
T=${1:-r2y}
timeout 120 python bench.py --workload 5pt-1000 --steps 200 --warmup 10 > gpurun_out/${T}_5pt1000.json 2> gpurun_out/${T}_5pt1000.err; echo "5pt rc=$?"
timeout 200 python bench.py --workload 27pt-192 --k 32 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_27pt192_k32.json 2> gpurun_out/${T}_27pt192_k32.err; echo "27pt rc=$?"
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],1), 'frac', (d.get('roofline') or {}).get('frac'), 'parity', (d.get('parity') or {}).get('ok'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err | cut -c1-300 | tail -8
