#!/bin/bash
# Round 2, GPU call 3J: compute-sanitizer memcheck over the video parity tests with the final fused kernel (v5.3 reads the row below and
# the byte to the right of every block's window whatever the mode).
cd "$(dirname "$0")/.."
O=gpurun_out/r3j; mkdir -p $O
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_video.py -m gpu -x -q -p no:cacheprovider -k "golden or sweep or ragged or wrap or strip or copy or mixed" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
tail -4 $O/memcheck.log
