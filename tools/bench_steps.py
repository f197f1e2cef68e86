"""BASELINE config 3, the other picture steps SURVEY 8d asks for besides the headline dense-P step:
natural P (intra 10 %, cbp ~ U{0..63}, n ~ 1 + Geom), I-only and natural B, 256 synthetic 720p streams,
records resident in HBM, CUDA events around the decode call (plan + arithmetic kernel).  Each line carries the
algorithmic bytes of ITS records (workload.algorithmic_bytes) and the fraction of the measured HBM peak."""
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import mpeg_b200  # noqa: E402
import workload as wl  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ONLY = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None   # e.g. "dense-P,natural-B"
STEPS, WARM = 20, 3
peak = 6569.0
try:
    peak = float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
g = wl.HD720
stream = torch.cuda.Stream()
ctx = mpeg_b200.Context(device=0, max_streams=S)
ctx.set_stream(stream.cuda_stream)
rng0 = wl.stream_rng(3, 9999)
ref = wl.random_reference_frame(rng0, g)
for s in range(S):
    ctx.video_open(s, g.width, g.height)
    for b in range(3):
        ctx.video_write_frame(s, b, np.roll(ref, 4099 * (3 * s + b)))
out = []
for name, ptype, mode in [("dense-P", wl.PIC_P, "dense"), ("natural-P", wl.PIC_P, "natural"), ("I-only", wl.PIC_I, "natural"),
                          ("dense-I", wl.PIC_I, "dense"), ("natural-B", wl.PIC_B, "natural")]:
    if ONLY is not None and name not in ONLY:
        continue
    t0 = time.time()
    per = [wl.make_picture(wl.stream_rng(3, 5000 + s), g, ptype, mode) for s in range(S)]
    pics, mbs, coeffs = wl.batch_pictures(per, list(range(S)), ptype, [(0, 1, 2)] * S)
    alg, alg_read = wl.algorithmic_bytes(mbs, len(coeffs))
    ctx.video_validate(pics, mbs, len(coeffs))
    d = [torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda() for a in (pics, mbs, coeffs)]
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    torch.cuda.synchronize()
    for k in range(WARM + STEPS):
        with torch.cuda.stream(stream):
            if k >= WARM:
                evs[k - WARM][0].record(stream)
            ctx.video_decode_pictures_dev(len(pics), d[0].data_ptr(), len(mbs), d[1].data_ptr(), len(coeffs), d[2].data_ptr())
            if k >= WARM:
                evs[k - WARM][1].record(stream)
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    rec = {"step": name, "streams": S, "macroblocks": int(len(mbs)), "coded_blocks": int(len(coeffs)),
           "predicted_frac": round(float(((mbs["flags"] & wl.MB_PREDICT) != 0).mean()), 3),
           "algorithmic_bytes": int(alg), "decode_ms": round(ms, 4), "frames_per_sec": round(S / ms * 1e3),
           "gbps": round(alg / ms / 1e6, 1), "hbm_frac": round(alg / ms / 1e6 / peak, 3), "build_s": round(time.time() - t0, 1)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del d
(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out/steps.json", "w"), indent=1)
