#!/bin/bash
# Round 2, GPU call J: A/B of the dp2a prediction add (correctness first, then the picture steps).
cd "$(dirname "$0")/.."
O=gpurun_out/r2j; mkdir -p $O
MPEGB200_LIB=$PWD/mpeg_b200/variants/libdp2a.so timeout 600 python -m pytest tests/test_gpu_video.py tests/test_gpu_api.py -m gpu -x -q -p no:cacheprovider > $O/pytest_dp2a.log 2>&1; echo "rc=$?" >> $O/pytest_dp2a.log
for v in default dp2a default dp2a; do
  if [ $v = default ]; then unset MPEGB200_LIB; else export MPEGB200_LIB=$PWD/mpeg_b200/variants/lib$v.so; fi
  timeout 600 python tools/bench_steps.py 256 dense-P,natural-P,natural-B >> $O/steps_$v.log 2>&1
done
tail -3 $O/pytest_dp2a.log; for v in default dp2a; do echo $v; grep -h decode_ms $O/steps_$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  %-10s %.4f ms  frac %.3f'%(d['step'],d['decode_ms'],d['hbm_frac']))"; done
