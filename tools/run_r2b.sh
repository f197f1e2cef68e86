#!/bin/bash
# Round 2, GPU call B: new bench.py (parity, e2e forms, steps, audio, sustained), then the GPU tests verbose with per-test timeouts.
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 700 python -u -X faulthandler -m pytest tests -m gpu -v --timeout 90 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/bench.err; grep -E "PASSED|FAILED|ERROR|Timeout|passed|failed" $O/pytest.log | tail -60
