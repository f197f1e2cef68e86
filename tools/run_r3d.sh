#!/bin/bash
# Round 2, GPU call 3D: final state -- whole GPU suite, default bench line, reference arm.
cd "$(dirname "$0")/.."
O=gpurun_out/r3d; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
tail -8 $O/bench.err; python - <<'PY'
import json
r=json.loads(open("gpurun_out/r3d/bench.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}); print("roofline",r["roofline"]["frac"],"e2e",r["e2e"]["value"],"cpu",r["cpu_baseline"]["value"] if r.get("cpu_baseline") else None)
b=r["bitstream"]; print({k:(round(v["frames_per_sec"]),v["parity_ok"]) for k,v in b["paths"].items()}, round(b["cpu_decoder"]["frames_per_sec"]), b.get("to_host_vs_cpu_decoder"))
PY
