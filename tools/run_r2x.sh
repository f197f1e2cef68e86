#!/bin/bash
# Round 2, GPU call X: grid-stride dequantisation kernel.
cd "$(dirname "$0")/.."
O=gpurun_out/r2x; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py tests/test_c_abi.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
tail -3 $O/pytest_vlc.log
for m in "natural 40" "dense 12"; do set -- $m
timeout 300 python tools/bench_bitstream.py --streams 256 --mode $1 --pictures $2 --distinct 2 --gpu --device-vlc --resident 2> /dev/null | tee $O/bitstream_$1.json | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$1: parse+dequant+check ms', round(sorted(pm)[len(pm)//2],3), 'wave fps', round(d['frames_per_sec']), 'resident fps', round(r['device_vlc_resident']['frames_per_sec']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vlc_ -s 6 -c 6 --csv --log-file $O/launches_$1.csv python tools/bench_bitstream.py --streams 256 --mode $1 --pictures $2 --distinct 2 --gpu --device-vlc > /dev/null 2>&1
python - $1 <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f"gpurun_out/r2x/launches_{sys.argv[1]}.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(sys.argv[1], r[4].split('(')[0][-24:], r[-1])
PY
done
