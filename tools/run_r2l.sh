#!/bin/bash
# Round 2, GPU call L: VLC stage after the kernel restructuring (few slices per warp, tables in shared memory): tests, steady-state bench.
cd "$(dirname "$0")/.."
O=gpurun_out/r2l; mkdir -p $O
timeout 600 python -u -X faulthandler -m pytest tests/test_gpu_vlc.py -m gpu -q --timeout 240 --timeout-method=thread -p no:cacheprovider > $O/pytest_vlc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_vlc.log
timeout 600 python tools/bench_bitstream.py --streams 256 --mode natural --pictures 40 --distinct 2 --gpu --device-vlc > $O/bitstream_natural.json 2> $O/bitstream_natural.err
timeout 600 python tools/bench_bitstream.py --streams 256 --mode dense --pictures 12 --distinct 2 --gpu --device-vlc > $O/bitstream_dense.json 2> $O/bitstream_dense.err
tail -3 $O/pytest_vlc.log; python - <<'PY'
import json
for m in ("natural","dense"):
    r=json.load(open(f"gpurun_out/r2l/bitstream_{m}.json"))
    d=r["device_vlc"]; pm=d.pop("parse_kernel_ms_per_wave"); dm=d.pop("decode_ms_per_wave"); d.pop("note")
    print(m,"host",r["host"]["pictures_per_sec"],"gpu(host parse)",r["gpu"]["frames_per_sec"]); print(d); print("parse ms",[round(x,3) for x in pm[:12]]); print("decode ms",[round(x,3) for x in dm[:6]])
PY
tail -2 $O/bitstream_natural.err
