#!/bin/bash
# Round 2, GPU call T (2 GPUs): bench.py under torchrun incl. the bitstream leg on every rank.
cd "$(dirname "$0")/.."
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
tail -6 $O/bench_2gpu.err; python - <<'PY'
import json
r=json.loads(open("gpurun_out/r2t/bench_2gpu.json").read().strip().splitlines()[-1])
print({k:r[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", r["e2e"]["value"], "gathered", r.get("value_gathered"))
print(json.dumps(r.get("bitstream"))[:900])
PY
