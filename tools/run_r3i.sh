#!/bin/bash
# Round 2, GPU call 3I: L2 prefetch distance of the fused kernel after v5.4 (experiment build, MPEGB200_PREFETCH_DIST).
cd "$(dirname "$0")/.."
export MPEGB200_LIB=mpeg_b200/variants/libexp.so
for d in 444 296 592 888 148; do
  echo "prefetch distance $d"; MPEGB200_PREFETCH_DIST=$d timeout 200 python tools/bench_steps.py 256 dense-P,natural-P 2>/dev/null | grep -E "dense-P|natural-P" | cut -c1-40,150-230
done
