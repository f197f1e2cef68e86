#!/bin/bash
# Round 2, GPU call K: the slice-parallel VLC stage on the device -- its tests, memcheck on them, bitstream -> frames at 720p.
cd "$(dirname "$0")/.."
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py -m gpu -v --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
timeout 600 python tools/bench_bitstream.py --streams 256 --mode natural --pictures 5 --distinct 2 --gpu --device-vlc > $O/bitstream_natural.json 2> $O/bitstream_natural.err
timeout 600 python tools/bench_bitstream.py --streams 256 --mode dense --pictures 4 --distinct 2 --gpu --device-vlc > $O/bitstream_dense.json 2> $O/bitstream_dense.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vlc.py -m gpu -x -q -p no:cacheprovider -k "golden or void or 352" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed|rc=" $O/pytest_vlc.log | tail -14; cat $O/bitstream_natural.json $O/bitstream_dense.json | cut -c1-2500; tail -3 $O/bitstream_natural.err; tail -4 $O/memcheck.log
