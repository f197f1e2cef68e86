#!/bin/bash
# Round 2, GPU call 3K: racecheck and synccheck over the video parity tests with the final fused kernel and plan pre-pass.
cd "$(dirname "$0")/.."
O=gpurun_out/r3k; mkdir -p $O
for tool in racecheck synccheck; do
timeout 500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_video.py -m gpu -x -q -p no:cacheprovider -k "golden or sweep or ragged or strip or mixed" > $O/$tool.log 2>&1; echo "$tool rc=$?" >> $O/$tool.log
tail -3 $O/$tool.log
done
