"""Compressed bitstreams in, decoded frames in HBM out: N copies of the reference's test clip decoded in
lock-step through mpeg_b200.VideoBatch (host parse on a thread pool -> one kernel launch per wave).
Prints frames/s for a few (streams, threads) points.  This measures the whole product, host parser included;
the clip is 160x120, so a frame is only 80 macroblocks."""
import json
import pathlib
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import mpeg_b200  # noqa: E402

clip = (pathlib.Path(__file__).resolve().parents[1] / "tests/golden/test.mpeg1video").read_bytes()
out = []
for n, threads in [(64, 8), (256, 16), (256, 32), (1024, 32), (1024, 64), (4096, 64)]:
    with mpeg_b200.Context(device=0, max_streams=n) as c:
        b = mpeg_b200.VideoBatch(c, [clip] * n, threads=threads)
        frames, steps = 0, 0
        t0 = time.perf_counter()
        while True:
            has, buf, t = b.step()
            k = int(has.sum())
            if k == 0:
                break
            frames += k
            steps += 1
        c.sync()
        dt = time.perf_counter() - t0
        b.close()
    rec = {"streams": n, "threads": threads, "frames": frames, "steps": steps, "seconds": round(dt, 3),
           "frames_per_s": round(frames / dt), "macroblocks_per_s": round(frames * 80 / dt)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
pathlib.Path("gpurun_out").mkdir(exist_ok=True)
json.dump(out, open("gpurun_out/batch_bench.json", "w"), indent=1)
