#!/bin/bash
# Round 2, GPU call 3E: one interpolation path for the four half-pel modes (MPEGB200_GENERIC_INTERP) against one path per mode.
cd "$(dirname "$0")/.."
O=gpurun_out/r3e; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests/test_gpu_video.py tests/test_gpu_api.py -m gpu -q --timeout 400 --timeout-method=thread -p no:cacheprovider -k "not full_size" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
for rep in 1 2; do
echo "generic (default)"; timeout 300 python tools/bench_steps.py 256 2>/dev/null | grep -E "dense-P|natural|I-only" | head -8
echo "per mode"; MPEGB200_LIB=mpeg_b200/variants/libpermode.so timeout 300 python tools/bench_steps.py 256 2>/dev/null | grep -E "dense-P|natural|I-only" | head -8
done
