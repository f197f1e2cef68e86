#!/bin/bash
# Round 2, GPU call E: GPU tests after the PDL fix, bench.py, ncu of the new MP2 kernel, compute-sanitizer on the API tests.
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests -m gpu -v --timeout 120 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:audio_synth -s 4 -c 1 -o $O/ncu_audio -f python tools/bench_audio.py 4 > $O/ncu_audio.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_api.py tests/test_gpu_audio.py -m gpu -x -q -p no:cacheprovider -k "not full_size" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed|rc=" $O/pytest.log | tail -80; tail -12 $O/bench.err; tail -5 $O/memcheck.log
