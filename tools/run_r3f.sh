#!/bin/bash
# Round 2, GPU call 3F: final state with the generic interpolation path -- whole GPU suite, default bench line, reference arm, launch list.
cd "$(dirname "$0")/.."
O=gpurun_out/r3f; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
grep -E "step |audio|bitstream|cpu baseline" $O/bench.err; python - <<'PY'
import json
r=json.loads(open("gpurun_out/r3f/bench.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}); print("roofline",r["roofline"]["frac"],r["roofline"]["launch_ms"],"dominant",r["roofline"]["dominant_kernel"]["frac"],"e2e",r["e2e"]["value"],"sustained",r["sustained"]["roofline_frac"])
PY
