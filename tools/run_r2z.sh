#!/bin/bash
# Round 2, GPU call Z (8 GPUs): smoke, then bench.py under torchrun on all eight (the driver's scaling run does this at N = 1, 2, 4, 8).
cd "$(dirname "$0")/.."
O=gpurun_out/r2z; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "rc=$?" >> $O/bench_${N}gpu.err
tail -5 $O/bench_${N}gpu.err; python - $N <<'PY'
import json,sys
r=json.loads(open(f"gpurun_out/r2z/bench_{sys.argv[1]}gpu.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", r["e2e"]["value"], "gathered", r.get("value_gathered"))
b=r.get("bitstream") or {}
print({k:(round(v["frames_per_sec"]),v["parity_ok"]) for k,v in b.get("paths",{}).items()}, b.get("host_threads"))
PY
