#!/bin/bash
# Round 2, GPU call O: whole GPU suite, racecheck + memcheck on the VLC stage, bench.py default line.
cd "$(dirname "$0")/.."
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_vlc.py -m gpu -x -q -p no:cacheprovider -k "golden or 352" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vlc.py -m gpu -x -q -p no:cacheprovider -k "golden or void or 352 or 576" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
tail -3 $O/pytest.log; tail -3 $O/racecheck.log; tail -3 $O/memcheck.log; tail -12 $O/bench.err; python - <<'PY'
import json
r=json.loads(open("gpurun_out/r2o/bench.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}); print("roofline",r["roofline"]["frac"],"e2e",r["e2e"]["value"],"cpu",r["cpu_baseline"]["value"] if r.get("cpu_baseline") else None)
print(json.dumps(r.get("bitstream"))[:1500])
PY
