#!/bin/bash
# Round 2, GPU call M: what bounds vlc_parse_kernel -- ncu --set full on one natural wave, slices-per-warp sweep.
cd "$(dirname "$0")/.."
O=gpurun_out/r2m; mkdir -p $O
export MPEGB200_LIB=mpeg_b200/variants/libexp.so
for lanes in 1 2 3 5 8; do
  MPEGB200_VLC_LANES=$lanes timeout 300 python tools/bench_bitstream.py --streams 256 --mode natural --pictures 40 --distinct 2 --gpu --device-vlc 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('natural lanes $lanes parse ms', round(sorted(pm)[len(pm)//2],3), 'e2e fps', round(d['frames_per_sec']))"
done
for lanes in 2 5 8; do
  MPEGB200_VLC_LANES=$lanes timeout 300 python tools/bench_bitstream.py --streams 256 --mode dense --pictures 12 --distinct 2 --gpu --device-vlc 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('dense lanes $lanes parse ms', round(sorted(pm)[len(pm)//2],3), 'e2e fps', round(d['frames_per_sec']))"
done
unset MPEGB200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vlc_parse -s 3 -c 1 -o $O/vlc_natural -f python tools/bench_bitstream.py --streams 256 --mode natural --pictures 40 --distinct 2 --gpu --device-vlc > $O/ncu.log 2>&1
tail -2 $O/ncu.log
