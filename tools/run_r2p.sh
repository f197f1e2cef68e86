#!/bin/bash
# Round 2, GPU call P: parse kernel with one coefficient loop per macroblock and a branch-free word load.
cd "$(dirname "$0")/.."
O=gpurun_out/r2p; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_vlc.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
run() {  # variant lanes mode pictures
  MPEGB200_LIB=mpeg_b200/variants/lib$1.so MPEGB200_VLC_LANES=$2 timeout 300 python tools/bench_bitstream.py --streams 256 --mode $3 --pictures $4 --distinct 2 --gpu --device-vlc 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$3 $1 lanes $2: parse ms', round(sorted(pm)[len(pm)//2],3), 'e2e fps', round(d['frames_per_sec']), 'scan/submit/wait', [round(v,3) for v in d['seconds_in'].values()])"
}
for v in "exp 4" "exp 5" "exp 6" "exp 8"; do set -- $v; run $1 $2 natural 40; done
for v in "exp 5" "exp 8"; do set -- $v; run $1 $2 dense 12; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vlc_parse -s 3 -c 1 -o $O/vlc_natural -f python tools/bench_bitstream.py --streams 256 --mode natural --pictures 40 --distinct 2 --gpu --device-vlc > $O/ncu.log 2>&1
tail -1 $O/ncu.log
