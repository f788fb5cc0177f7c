#!/bin/bash
# ONE GPU: the whole -m gpu suite (full-size parity included), then A/B measurements of this round's kernel changes.
set -u
mkdir -p gpurun_out
T=${1:-r2t}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -n 15 gpurun_out/${T}_pytest_gpu.log | cut -c1-300
# power-law SpMV: two-phase (round 1) vs sub-warp-per-row with 4 / 8 / 16 lanes per row at mean length 16
for div in 0 3 1.5 0.75; do
  JP_SPMV_SUBWARP_DIV=$div timeout 200 python bench.py --workload powerlaw --rows 10000000 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e \
      > gpurun_out/${T}_powerlaw10M_div$div.json 2> gpurun_out/${T}_powerlaw10M_div$div.err
done
timeout 120 python tools/bench_vecops.py > gpurun_out/${T}_vecops.json 2> gpurun_out/${T}_vecops.err
timeout 150 python tools/bench_next_rows.py --N 256 --iters 50 > gpurun_out/${T}_next_rows_256.json 2> gpurun_out/${T}_next_rows_256.err
for ch in 16 32 64; do
  JP_HOST_CHUNKS=$ch timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_chunks$ch.json 2> gpurun_out/${T}_bench_chunks$ch.err
done
timeout 300 python tools/bench_spmm.py --N 192 --ks 8,32 --reps 10 --variants gather,default,runs_kt8,runs_db,runs_kt4_t256 > gpurun_out/${T}_spmm192.json 2> gpurun_out/${T}_spmm192.err
for f in gpurun_out/${T}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    if 'ms_per_step' in d:
        print(' ms/step', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4), 'parity', (d.get('parity') or {}).get('ok'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
    elif 'cases' in d:
        for c in d['cases']:
            print('  ', {k:(round(v,4) if isinstance(v,float) else v) for k,v in c.items() if k in ('variant','k','ms','frac_of_peak','op','n','frac_of_measured_peak','error')})
    else:
        print(' ', json.dumps(d)[:1500])
except Exception as e:
    print(' no line', e)
PY
done
tail -n 3 gpurun_out/${T}_*.err | cut -c1-300 | tail -30
