#!/bin/bash
# Round 2, GPU call N: vlc_parse_kernel variants -- warps per SM x slices per warp, refill form.
cd "$(dirname "$0")/.."
O=gpurun_out/r2n; mkdir -p $O
run() {  # variant lanes mode pictures
  MPEGB200_LIB=mpeg_b200/variants/lib$1.so MPEGB200_VLC_LANES=$2 timeout 300 python tools/bench_bitstream.py --streams 256 --mode $3 --pictures $4 --distinct 2 --gpu --device-vlc 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$3 $1 lanes $2: parse ms', round(sorted(pm)[len(pm)//2],3), 'e2e fps', round(d['frames_per_sec']), 'scan/submit/wait', [round(v,3) for v in d['seconds_in'].values()])"
}
for v in "exp 5" "tail 5" "w6 4" "w6 5" "w8 3" "w8 4" "w8 5" "w8c3 4" "w8c3 6" "exp 8"; do set -- $v; run $1 $2 natural 40; done
for v in "exp 5" "w6 4" "w8 3" "w8 5" "exp 8"; do set -- $v; run $1 $2 dense 12; done
