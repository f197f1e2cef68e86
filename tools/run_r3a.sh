#!/bin/bash
# Round 2, GPU call 3A: walker micro-optimisations (incremental row/col, 32-bit end-of-stream tests, one chroma window check).
cd "$(dirname "$0")/.."
O=gpurun_out/r3a; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py tests/test_c_abi.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
tail -3 $O/pytest_vlc.log
for m in "natural 40" "dense 12"; do set -- $m
for rep in 1 2; do
timeout 300 python tools/bench_bitstream.py --streams 256 --mode $1 --pictures $2 --distinct 2 --gpu --device-vlc --resident 2> /dev/null | tee $O/bitstream_$1.json | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$1: parse+check ms', round(sorted(pm)[len(pm)//2],3), 'wave fps', round(d['frames_per_sec']), 'resident fps', round(r['device_vlc_resident']['frames_per_sec']))"
done; done
