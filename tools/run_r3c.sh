#!/bin/bash
# Round 2, GPU call 3C: native device stepper -- tests (both flows), bitstream bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r3c; mkdir -p $O
timeout 900 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py tests/test_c_abi.py tests/test_gpu_api.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for rep in 1 2; do
timeout 300 python tools/bench_bitstream.py --streams 256 --mode natural --pictures 40 --distinct 2 --gpu --device-vlc --resident 2> /dev/null | tee $O/bitstream_natural.json | python -c "
import json,sys; r=json.load(sys.stdin)
for k in ('device_vlc','device_vlc_resident'):
    d=r[k]; print(k, round(d['frames_per_sec']), {a:round(b,3) for a,b in d.get('seconds_in',{}).items()})"
done
