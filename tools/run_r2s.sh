#!/bin/bash
# Round 2, GPU call S: whole GPU suite (plain-C bitstream driver, single-stream device VLC, resident streams).
cd "$(dirname "$0")/.."
O=gpurun_out/r2s; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -12 $O/pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vlc.py tests/test_c_abi.py -m gpu -x -q -p no:cacheprovider -k "resident or tiny or single" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
tail -3 $O/memcheck.log
