#!/bin/bash
# Round 2, GPU call U: the VLC stage as two kernels (serial symbol walk + parallel dequantisation).
cd "$(dirname "$0")/.."
O=gpurun_out/r2u; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py tests/test_c_abi.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
tail -6 $O/pytest_vlc.log
for m in "natural 40" "dense 12"; do set -- $m
timeout 600 python tools/bench_bitstream.py --streams 256 --mode $1 --pictures $2 --distinct 2 --gpu --device-vlc --resident 2> $O/bitstream_$1.err | tee $O/bitstream_$1.json | python -c "
import json,sys; r=json.load(sys.stdin)
pm=r['device_vlc']['parse_kernel_ms_per_wave']; print('$1 parse+dequant+check ms', round(sorted(pm)[len(pm)//2],3))
for k in ('gpu','device_vlc','device_vlc_resident'):
    d=r[k]; print('$1',k,round(d['frames_per_sec']),{a:round(b,3) for a,b in d.get('seconds_in',{}).items()})"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python tools/bench_bitstream.py --streams 256 --mode natural --pictures 40 --distinct 2 --gpu --device-vlc > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2u/launches.csv")) if len(r)>10 and r[0].isdigit()]
t=collections.defaultdict(list)
for r in rows: t[r[4].split('(')[0]].append(float(r[-1].replace(',','')))
for k,v in t.items(): print(k, len(v), 'median us', sorted(v)[len(v)//2]/1e3 if max(v)>1e4 else sorted(v)[len(v)//2])
PY
