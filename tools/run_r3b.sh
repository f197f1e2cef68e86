#!/bin/bash
# Round 2, GPU call 3B: same-call A/B of the walker micro-optimisations (old = before them, new = all, div = all but the incremental row).
cd "$(dirname "$0")/.."
run() {
  MPEGB200_LIB=mpeg_b200/variants/lib$1.so timeout 300 python tools/bench_bitstream.py --streams 256 --mode $2 --pictures $3 --distinct 2 --gpu --device-vlc 2> /dev/null | python -c "
import json,sys; r=json.load(sys.stdin); d=r['device_vlc']; pm=d['parse_kernel_ms_per_wave']; print('$2 $1: parse+check ms', round(sorted(pm)[len(pm)//2],3))"
}
for v in old new div old new div; do run $v natural 40; done
for v in old new div; do run $v dense 12; done
