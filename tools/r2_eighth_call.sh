#!/bin/bash
# ONE GPU: the -m gpu suite with the column-slab SpMV, its slab-size sweep at 10 M rows, the BASELINE size (50 M rows),
# and the transposed apply on the bench matrix.
set -u
mkdir -p gpurun_out
T=${1:-r2k}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -n 8 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
for mb in 0 16 24 32; do
  JP_CB_SLAB_MB=$mb timeout 200 python bench.py --workload powerlaw --rows 10000000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e \
      > gpurun_out/${T}_powerlaw10M_slab$mb.json 2> gpurun_out/${T}_powerlaw10M_slab$mb.err
done
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --trans > gpurun_out/${T}_bench_trans.json 2> gpurun_out/${T}_bench_trans.err
timeout 500 python bench.py --workload powerlaw --rows 50000000 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
    > gpurun_out/${T}_powerlaw50M.json 2> gpurun_out/${T}_powerlaw50M.err
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'parity', (d.get('parity') or {}).get('ok'), 'launches', d.get('gpu_launches'), d.get('mode'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err | cut -c1-300 | tail -20
