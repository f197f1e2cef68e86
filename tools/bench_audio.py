#!/usr/bin/env python3
"""BASELINE config 4: MP2 synthesis, 1024 streams x 8 frames per launch on one B200.
Prints one JSON line (audio frames/s, GB/s against the algorithmic 20,352 B/frame, bit-exactness vs the CPU oracle
on a sample).  Not the headline bench (bench.py); used for profiles/."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import mpeg_b200
    import workload as wl
    S, F, steps, warmup = 1024, 8, 20, 3
    if len(sys.argv) > 1:
        steps = int(sys.argv[1])
    flag = 0x100 if len(sys.argv) > 2 and sys.argv[2] == "fma" else 0   # MPEGB200_AUDIO_WINDOW_FMA
    rng = wl.stream_rng(4, 0)
    samples = wl.audio_samples(rng, S * F)
    stream = torch.cuda.Stream()
    ctx = mpeg_b200.Context(0, S)
    ctx.set_stream(stream.cuda_stream)
    ids = np.arange(S, dtype=np.int32)
    for s in ids:
        ctx.audio_open(int(s))
    d_in = torch.from_numpy(samples.reshape(-1)).cuda()
    d_out = torch.empty(S * F * 2304, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(warmup):
        ctx.audio_synth_dev(ids, F, d_in.data_ptr(), flag, d_out.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(steps):
        ctx.audio_synth_dev(ids, F, d_in.data_ptr(), flag, d_out.data_ptr())
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    # parity on a sample of streams: replay the same number of launches through the oracle
    import oracle_lib as ol
    n_chk = 8
    st = ol.synth_states(n_chk)
    want = None
    for _ in range(warmup + steps):
        want = ol.synth_batch(st, n_chk, F, samples[: n_chk * F], 0, fma=bool(flag))
    got = d_out.cpu().numpy().reshape(S, F, 2304)[:n_chk]
    exact = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
    frames = S * F
    alg = 20352 * frames
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    print(json.dumps({"metric": "mp2_synthesis_frames_per_sec", "value": frames / (ms * 1e-3), "unit": "audio frames/s",
                      "ms_per_launch": ms, "window": "fused" if flag else "unfused", "streams": S, "frames_per_launch": F,
                      "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": alg},
                      "bit_exact_vs_oracle_sample": exact}))


if __name__ == "__main__":
    main()
