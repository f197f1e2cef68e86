#!/bin/bash
# final-state measurements: tests, rgba A/B, launch list, other picture steps, ncu captures
OUT=gpurun_out/${1:-r1l}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/tests_all.log
MPEGB200_RGBA=wide timeout 600 python -m pytest tests -m gpu -x -q -k "rgba or RGBA or api" 2>&1 | tail -3 | tee $OUT/tests_rgba_wide.log
show() { python -c "
import sys,json
d=json.loads(open('$1').read().strip().splitlines()[-1])
r=d['roofline']; e=d.get('e2e') or {}
print('$1', 'step_ms', round(d['ms_per_step'],4), 'call_ms', round(r['launch_ms'],4), 'frac', round(r['frac'],4), 'fused_only', round(r['dominant_kernel']['launch_ms'],4), round(r['dominant_kernel']['frac'],4), 'plan', round(r['dominant_kernel']['plan_kernel_ms'],4), 'value', round(d['value']), 'e2e', round(e.get('value',0)), e.get('ms_per_step'))
"; }
for v in patch wide; do
  MPEGB200_RGBA=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench_rgba_$v.json 2> $OUT/bench_rgba_$v.err; show $OUT/bench_rgba_$v.json
  MPEGB200_RGBA=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file $OUT/launches_rgba_$v.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python tools/launch_list.py $OUT/launches_rgba_$v.csv
done
timeout 900 python tools/bench_steps.py 2>&1 | tail -6 | tee $OUT/steps.log; cp gpurun_out/steps.json $OUT/steps.json
