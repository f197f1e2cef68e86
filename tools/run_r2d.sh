#!/bin/bash
# Round 2, GPU call D: does round 1's code hang on the reference clip too?  Then the current tree, int16 and vlen.
cd "$(dirname "$0")/.."
O=gpurun_out/r2d; mkdir -p $O
(cd _r1 && timeout 120 python -u -m pytest tests/test_gpu_api.py -x -v --timeout 40 -p no:cacheprovider > ../$O/r1_pytest.log 2>&1; echo "rc=$?" >> ../$O/r1_pytest.log)
timeout 60 python -u tools/debug_golden.py 400 int16 > $O/golden_int16.log 2>&1; echo "rc=$?" >> $O/golden_int16.log
timeout 60 python -u tools/debug_golden.py 400 vlen > $O/golden_vlen.log 2>&1; echo "rc=$?" >> $O/golden_vlen.log
timeout 600 python -u -X faulthandler -m pytest tests -m gpu -v --timeout 60 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed|rc=" $O/r1_pytest.log | tail; tail -3 $O/golden_int16.log; tail -3 $O/golden_vlen.log
grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed|rc=" $O/pytest.log | tail -70
