#!/bin/bash
# Round 2, GPU call Y: final default bench line (with the bitstream leg incl. the to-host path and the CPU decoder), reference arm, launch list.
cd "$(dirname "$0")/.."
O=gpurun_out/r2y; mkdir -p $O
timeout 1200 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
tail -14 $O/bench.err; tail -2 $O/bench_reference.err; cut -c1-400 $O/bench_reference.json; python - <<'PY'
import json
r=json.loads(open("gpurun_out/r2y/bench.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","gpu_launches")}); print("roofline",r["roofline"]["frac"],"e2e",r["e2e"]["value"],"cpu",r["cpu_baseline"]["value"] if r.get("cpu_baseline") else None)
b=r["bitstream"]; print({k:(round(v["frames_per_sec"]),v["parity_ok"]) for k,v in b["paths"].items()}, b.get("cpu_decoder"), b.get("to_host_vs_cpu_decoder"))
PY
