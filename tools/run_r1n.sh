#!/bin/bash
OUT=gpurun_out/${1:-r1n}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/tests_all.log
MPEGB200_STRIP=1 timeout 600 python -m pytest tests/test_gpu_video.py -m gpu -x -q 2>&1 | tail -2 | tee $OUT/tests_strip1.log
show() { python -c "
import sys,json
d=json.loads(open('$1').read().strip().splitlines()[-1])
r=d['roofline']; e=d.get('e2e') or {}
print('$1', 'step_ms', round(d['ms_per_step'],4), 'call_ms', round(r['launch_ms'],4), 'frac', round(r['frac'],4), 'fused_only', round(r['dominant_kernel']['launch_ms'],4), round(r['dominant_kernel']['frac'],4), 'plan', round(r['dominant_kernel']['plan_kernel_ms'],4), 'value', round(d['value']))
"; }
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/bench.json 2> $OUT/bench.err; show $OUT/bench.json
timeout 900 python tools/bench_steps.py 256 dense-P,natural-P,natural-B 2>&1 | tail -4 | tee $OUT/steps.log
MPEGB200_STRIP=1 timeout 900 python tools/bench_steps.py 256 natural-B 2>&1 | tail -1 | tee $OUT/steps_strip1.log
