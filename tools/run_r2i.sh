#!/bin/bash
# Round 2, GPU call I: all GPU tests after the f3/f4 additions, bitstream -> frames at 720p, sanitizer on the new pieces.
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests -m gpu -v --timeout 180 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/bench_bitstream.py --streams 256 --mode dense --pictures 5 --distinct 2 --gpu > $O/bitstream_dense.json 2> $O/bitstream_dense.err
timeout 600 python tools/bench_bitstream.py --streams 256 --mode natural --pictures 5 --distinct 2 --gpu > $O/bitstream_natural.json 2> $O/bitstream_natural.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_api.py tests/test_gpu_audio.py tests/test_c_abi.py -m gpu -x -q -p no:cacheprovider -k "not full_size" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_audio.py -m gpu -x -q -p no:cacheprovider -k "golden or tiny or requant" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
grep -E "FAILED|ERROR|Timeout|passed|failed|rc=" $O/pytest.log | tail -12; cat $O/bitstream_dense.json $O/bitstream_natural.json | cut -c1-900; tail -4 $O/memcheck.log; tail -4 $O/racecheck.log
