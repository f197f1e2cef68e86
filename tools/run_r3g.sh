#!/bin/bash
# Round 2, GPU call 3G: plan pre-pass fast path for fully coded groups + two bins, against the sort for every group.
cd "$(dirname "$0")/.."
O=gpurun_out/r3g; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests/test_gpu_video.py tests/test_gpu_api.py -m gpu -q --timeout 400 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
echo "fast path for full groups (default)"; timeout 300 python tools/bench_steps.py 256 2>/dev/null | grep -E "dense-P|natural|I-only" | cut -c1-40,150-260
echo "sort for every group"; MPEGB200_LIB=mpeg_b200/variants/libnofull.so timeout 300 python tools/bench_steps.py 256 2>/dev/null | grep -E "dense-P|natural|I-only" | cut -c1-40,150-260
