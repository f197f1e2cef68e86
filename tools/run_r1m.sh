#!/bin/bash
# round-1 final captures: ncu --set full of the video kernels and the expansion kernel, then the two bench arms as the driver runs them
OUT=gpurun_out/${1:-r1m}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused|plan|rgba" -s 9 -c 3 -f -o $OUT/ncu_video \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_video.log 2>&1; tail -2 $OUT/ncu_video.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"expand" -s 1 -c 1 -f -o $OUT/ncu_expand \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/ncu_expand.log 2>&1; tail -2 $OUT/ncu_expand.log | cut -c1-200
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; cat $OUT/bench_default.json | cut -c1-6000
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json | cut -c1-3000
nproc
