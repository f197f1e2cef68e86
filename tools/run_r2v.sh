#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2v; mkdir -p $O
for m in "natural 40" "dense 12"; do set -- $m
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:vlc_ -s 6 -c 9 --csv --log-file $O/launches_$1.csv python tools/bench_bitstream.py --streams 256 --mode $1 --pictures $2 --distinct 2 --gpu --device-vlc > /dev/null 2>&1
python - $1 <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f"gpurun_out/r2v/launches_{sys.argv[1]}.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(sys.argv[1], r[4].split('(')[0][-24:], r[-3], r[-1])
PY
done
