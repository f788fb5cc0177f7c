#!/bin/bash
# ONE GPU, the round's last call: what the driver runs at round end (pytest -m gpu, smoke(), bench.py both arms at N=1)
# on the final tree, plus the SpMM bench line of BASELINE configs[2].
set -u
mkdir -p gpurun_out
T=${1:-r2z}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -n 5 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${T}_smoke.log
timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --workload 27pt-192 --k 8 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_27pt192_k8.json 2> gpurun_out/${T}_bench_27pt192_k8.err; echo "27pt rc=$?"
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'value', round(d['value'],1), 'frac', (d.get('roofline') or {}).get('frac'), 'parity', d.get('parity'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'config', d['config'])
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err | cut -c1-300 | tail -12
