#!/bin/bash
# Round 2, GPU call F: MP2 kernel after the overhead cuts.
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
timeout 300 python -u -m pytest tests/test_gpu_audio.py tests/test_gpu_api.py -m gpu -q --timeout 120 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 120 python tools/bench_audio.py 20 > $O/audio_unfused.json 2>&1
timeout 120 python tools/bench_audio.py 20 fma > $O/audio_fused.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:audio_synth -s 4 -c 1 -o $O/ncu_audio -f python tools/bench_audio.py 4 > $O/ncu_audio.log 2>&1
tail -4 $O/pytest.log; cat $O/audio_unfused.json $O/audio_fused.json
