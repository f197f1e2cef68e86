#!/bin/bash
# One gpurun call: GPU parity tests for every fused-kernel variant, then the bench of each, then ncu of the default.
# usage (under gpurun): bash tools/ab_run.sh [tag]
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== tests default (stream+strip)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/tests_default.log
echo "== tests oneshot+strip"; MPEGB200_FUSED=oneshot timeout 600 python -m pytest tests/test_gpu_video.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/tests_oneshot.log
echo "== tests stream boxes"; MPEGB200_STRIP=0 timeout 600 python -m pytest tests/test_gpu_video.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/tests_stream_box.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
run() { name=$1; shift; echo "== bench $name"; env "$@" timeout 600 $B 2> $OUT/bench_$name.err | tee $OUT/bench_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['fused_ms'], d['roofline']['frac'], d['value'], d['clocks'])"; }
run stream_strip X=1
run oneshot_strip MPEGB200_FUSED=oneshot
run stream_box MPEGB200_STRIP=0
run oneshot_box MPEGB200_FUSED=oneshot MPEGB200_STRIP=0
[ -f mpeg_b200/variants/libmpegb200_v42.so ] && run v42 MPEGB200_LIB=mpeg_b200/variants/libmpegb200_v42.so MPEGB200_FUSED=oneshot
echo "== steps"; timeout 900 python tools/bench_steps.py 2>&1 | tee $OUT/steps.log | cut -c1-200
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused|plan" -s 10 -c 4 -f -o $OUT/fused_full \
    python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
