#!/bin/bash
# Round 2, GPU call H (2 GPUs): bench.py under torchrun like the driver launches it, both gather modes; reference arm.
cd "$(dirname "$0")/.."
O=gpurun_out/r2h; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --gather-mode gather --no-parity > $O/bench_2gpu_gather.json 2> $O/bench_2gpu_gather.err; echo "rc=$?" >> $O/bench_2gpu_gather.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_2gpu_ref.json 2> $O/bench_2gpu_ref.err; echo "rc=$?" >> $O/bench_2gpu_ref.err
tail -4 $O/bench_2gpu.err; tail -3 $O/bench_2gpu_gather.err; tail -2 $O/bench_2gpu_ref.err; cut -c1-600 $O/bench_2gpu_ref.json
