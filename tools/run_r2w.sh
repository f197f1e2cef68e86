#!/bin/bash
# Round 2, GPU call W: two-kernel VLC stage with the scan-based dequantisation; warps per CTA A/B.
cd "$(dirname "$0")/.."
O=gpurun_out/r2w; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py tests/test_c_abi.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
tail -4 $O/pytest_vlc.log
run() {  # variant lanes mode pictures
  MPEGB200_LIB=mpeg_b200/variants/lib$1.so MPEGB200_VLC_LANES=$2 timeout 300 python tools/bench_bitstream.py --streams 256 --mode $3 --pictures $4 --distinct 2 --gpu --device-vlc --resident 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$3 $1 lanes $2: parse+dequant+check ms', round(sorted(pm)[len(pm)//2],3), 'wave fps', round(d['frames_per_sec']), 'resident fps', round(r['device_vlc_resident']['frames_per_sec']))"
}
for v in "exp 5" "exp 4" "exp 6" "w5 4" "w5 5" "w6 4" "w6 5"; do set -- $v; run $1 $2 natural 40; done
for v in "exp 5" "w5 4"; do set -- $v; run $1 $2 dense 12; done
