#!/bin/bash
# Round 2, GPU call A: sanity (GPU tests), TMA box-rate probe, A/B of kernel variants on the picture steps, ncu of natural-P/B.
cd "$(dirname "$0")/.."
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 120 tools/_build/tma_rate > $O/tma_rate.txt 2>&1
STEPS="dense-P,natural-P,natural-B"
for v in default exp occ7 coarse; do
  if [ $v = default ]; then unset MPEGB200_LIB; else export MPEGB200_LIB=$PWD/mpeg_b200/variants/lib$v.so; fi
  timeout 600 python tools/bench_steps.py 256 $STEPS > $O/steps_$v.log 2>&1
  cp gpurun_out/steps.json $O/steps_$v.json 2>/dev/null
done
export MPEGB200_LIB=$PWD/mpeg_b200/variants/libexp.so
for m in fetch math; do
  MPEGB200_MEASURE=$m timeout 600 python tools/bench_steps.py 256 $STEPS > $O/steps_exp_$m.log 2>&1
done
MPEGB200_STRIP=0 timeout 600 python tools/bench_steps.py 256 dense-P > $O/steps_exp_nostrip.log 2>&1
unset MPEGB200_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_tma|plan_kernel" -s 6 -c 2 -o $O/ncu_naturalP -f \
    python tools/bench_steps.py 256 natural-P > $O/ncu_naturalP.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_tma" -s 3 -c 1 -o $O/ncu_naturalB -f \
    python tools/bench_steps.py 256 natural-B > $O/ncu_naturalB.log 2>&1
tail -3 $O/pytest.log; cat $O/tma_rate.txt; grep -h decode_ms $O/steps_*.log | cut -c1-200
