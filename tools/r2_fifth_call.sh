#!/bin/bash
# ONE GPU: the -m gpu suite, SpMM A/B (diagonal kernel with immediate-offset tile reads), power-law SpMV (four gathers in
# flight per lane), reductions.
set -u
mkdir -p gpurun_out
T=${1:-r2g}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -n 12 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
timeout 300 python tools/bench_spmm.py --N 192 --ks 8,32 --reps 10 > gpurun_out/${T}_spmm192.json 2> gpurun_out/${T}_spmm192.err
timeout 200 python bench.py --workload powerlaw --rows 10000000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e \
    > gpurun_out/${T}_powerlaw10M.json 2> gpurun_out/${T}_powerlaw10M.err
timeout 120 python tools/bench_vecops.py > gpurun_out/${T}_vecops.json 2> gpurun_out/${T}_vecops.err
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    if 'ms_per_step' in d:
        print(' ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'parity', (d.get('parity') or {}).get('ok'))
    elif 'cases' in d:
        for c in d['cases']:
            print('  ', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('variant','k','ms','frac_of_peak','op','n','frac_of_measured_peak','error','dia_state')})
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err | cut -c1-300 | tail -20
