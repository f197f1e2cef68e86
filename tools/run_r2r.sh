#!/bin/bash
# Round 2, GPU call R: streams resident in HBM + start codes indexed on the device.
cd "$(dirname "$0")/.."
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
tail -15 $O/pytest_vlc.log
for m in "natural 40" "dense 12"; do set -- $m
timeout 600 python tools/bench_bitstream.py --streams 256 --mode $1 --pictures $2 --distinct 2 --gpu --device-vlc --resident 2> $O/bitstream_$1.err | tee $O/bitstream_$1.json | python -c "
import json,sys; r=json.load(sys.stdin)
for k in ('gpu','device_vlc','device_vlc_resident'):
    d=r[k]; print('$1',k,round(d['frames_per_sec']),{a:round(b,3) for a,b in d.get('seconds_in',{}).items()}, d.get('setup_seconds'), d.get('steps'))"
done
