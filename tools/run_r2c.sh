#!/bin/bash
# Round 2, GPU call C: where does the reference clip hang?  + ncu of the MP2 kernel.
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
timeout 90 python -u tools/debug_golden.py 400 > $O/golden.log 2>&1; echo "rc=$?" >> $O/golden.log
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -u tools/debug_golden.py 12 > $O/golden_memcheck.log 2>&1; echo "rc=$?" >> $O/golden_memcheck.log
MPEGB200_VALIDATE=0 CUDA_LAUNCH_BLOCKING=1 timeout 60 python -u tools/debug_golden.py 30 > $O/golden_blocking.log 2>&1; echo "rc=$?" >> $O/golden_blocking.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:audio_synth -s 4 -c 1 -o $O/ncu_audio -f python tools/bench_audio.py 4 > $O/ncu_audio.log 2>&1
tail -5 $O/golden.log; tail -15 $O/golden_memcheck.log; tail -4 $O/golden_blocking.log
