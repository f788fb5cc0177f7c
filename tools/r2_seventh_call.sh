#!/bin/bash
# ONE GPU: does the L2 fetch-granularity hint change the DRAM traffic of the random gathers (power-law SpMV)?
set -u
mkdir -p gpurun_out
T=${1:-r2i}
for g in 64 32; do
  JP_L2_FETCH=$g timeout 300 python bench.py --workload powerlaw --rows 20000000 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e \
      > gpurun_out/${T}_powerlaw20M_fetch$g.json 2> gpurun_out/${T}_powerlaw20M_fetch$g.err
done
JP_L2_FETCH=32 timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_fetch32.json 2> gpurun_out/${T}_bench_fetch32.err
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(' ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'parity', (d.get('parity') or {}).get('ok'), 'nnz', d['config']['nnz'], 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err | cut -c1-300 | tail -12
