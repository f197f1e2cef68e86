#!/bin/bash
# usage (under gpurun): [TESTS=1] [LIST=1] [NCU="name|ENV;ENV|kernel-regex ..."] bash tools/ab_run3.sh tag "name|ENV=..;ENV=.." ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
for spec in "$@"; do
  name=${spec%%|*}; envs=${spec#*|}; envs=${envs//;/ }
  if [ -n "$TESTS" ]; then echo "== tests $name"; env $envs timeout 600 python -m pytest tests/test_gpu_video.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/tests_$name.log; fi
  echo "== bench $name"
  env $envs timeout 600 $B 2> $OUT/bench_$name.err | tee $OUT/bench_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['fused_ms'], d['roofline']['frac'], d['value'], d['clocks'])"
  grep mpegb200 $OUT/bench_$name.err
  if [ -n "$LIST" ]; then
    env $envs timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file $OUT/launches_$name.csv \
        python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
    python tools/launch_list.py $OUT/launches_$name.csv
  fi
done
for spec in $NCU; do
  IFS='|' read name envs kre <<< "$spec"; envs=${envs//;/ }
  echo "== ncu $name"
  env $envs timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s 3 -c 1 -f -o $OUT/ncu_$name \
      python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > $OUT/ncu_$name.log 2>&1
  tail -1 $OUT/ncu_$name.log
done
ls $OUT
