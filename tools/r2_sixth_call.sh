#!/bin/bash
# ONE GPU: the -m gpu suite, the power-law SpMV at 10 M rows and at the BASELINE size (50 M rows), and an ncu capture at 50 M.
set -u
mkdir -p gpurun_out
T=${1:-r2h}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -n 8 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 200 python bench.py --workload powerlaw --rows 10000000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e \
    > gpurun_out/${T}_powerlaw10M.json 2> gpurun_out/${T}_powerlaw10M.err
timeout 500 python bench.py --workload powerlaw --rows 50000000 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
    > gpurun_out/${T}_powerlaw50M.json 2> gpurun_out/${T}_powerlaw50M.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -s 6 -c 1 -f -o gpurun_out/${T}_spmv_powerlaw50M \
    python bench.py --workload powerlaw --rows 50000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${T}_ncu_powerlaw50M.log 2>&1
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'parity', (d.get('parity') or {}).get('ok'), 'nnz', d['config']['nnz'])
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err gpurun_out/${T}_ncu_*.log | cut -c1-300 | tail -20
