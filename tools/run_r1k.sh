#!/bin/bash
OUT=gpurun_out/${1:-r1k}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/tests_all.log
show() { python -c "
import sys,json
d=json.loads(open('$1').read().strip().splitlines()[-1])
r=d['roofline']; e=d.get('e2e') or {}
print('$1', 'call_ms', round(r['launch_ms'],4), 'frac', round(r['frac'],4), 'fused_only', round(r['dominant_kernel']['launch_ms'],4), round(r['dominant_kernel']['frac'],4), 'plan', round(r['dominant_kernel']['plan_kernel_ms'],4), 'value', round(d['value']), 'e2e', round(e.get('value',0)), e.get('h2d_bytes_per_step'), e.get('ms_per_step'))
"; }
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_vlen.json 2> $OUT/bench_vlen.err; show $OUT/bench_vlen.json
