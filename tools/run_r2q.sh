#!/bin/bash
# Round 2, GPU call Q: A/B of the two parse kernel changes (coefficient loop per macroblock; branch-free word load).
cd "$(dirname "$0")/.."
run() {  # variant lanes mode pictures
  MPEGB200_LIB=mpeg_b200/variants/lib$1.so MPEGB200_VLC_LANES=$2 timeout 300 python tools/bench_bitstream.py --streams 256 --mode $3 --pictures $4 --distinct 2 --gpu --device-vlc 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$3 $1 lanes $2: parse ms', round(sorted(pm)[len(pm)//2],3), 'e2e fps', round(d['frames_per_sec']))"
}
for v in "exp 5" "mb 5" "bl 5" "mbbl 5" "mb 6" "mb 8" "exp 5"; do set -- $v; run $1 $2 natural 40; done
for v in "exp 5" "mb 5" "bl 5"; do set -- $v; run $1 $2 dense 12; done
