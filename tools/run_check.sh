#!/bin/bash
# usage (under gpurun): bash tools/run_check.sh TAG  -- GPU parity tests, the default bench line, a launch list incl. the e2e leg
OUT=gpurun_out/${1:-check}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/tests_all.log
timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
r=d['roofline']; e=d.get('e2e') or {}
print('step_ms', round(d['ms_per_step'],4), 'call_ms', round(r['launch_ms'],4), 'frac', round(r['frac'],4), 'fused_only', round(r['dominant_kernel']['launch_ms'],4), round(r['dominant_kernel']['frac'],4), 'plan', round(r['dominant_kernel']['plan_kernel_ms'],4), 'value', round(d['value']), 'e2e', round(e.get('value',0)), e.get('ms_per_step'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"expand|unpack" -c 6 --csv --log-file $OUT/launches_e2e.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_list.py $OUT/launches_e2e.csv
