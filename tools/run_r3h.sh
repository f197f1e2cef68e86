#!/bin/bash
# Round 2, GPU call 3H: ncu --set full of the final fused kernel (dense-P step of bench.py) and of the natural-P step.
cd "$(dirname "$0")/.."
O=gpurun_out/r3h; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tma -s 3 -c 1 -o $O/fused_dense -f python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-parity > $O/ncu_dense.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tma -s 4 -c 1 -o $O/fused_naturalP -f python tools/bench_steps.py 256 natural-P > $O/ncu_naturalP.log 2>&1
tail -1 $O/ncu_dense.log; tail -1 $O/ncu_naturalP.log
