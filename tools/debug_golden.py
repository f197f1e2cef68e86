#!/usr/bin/env python3
"""Debug aid: the reference clip through mpeg_b200.Video with a time stamp per Decode() (where does a slow / hung step sit?)."""
import faulthandler
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
faulthandler.dump_traceback_later(40, exit=True)
import mpeg_b200  # noqa: E402

ctx = mpeg_b200.Context(device=0, max_streams=8)
VLEN = len(sys.argv) > 2 and sys.argv[2] == "vlen"
v = mpeg_b200.Video((ROOT / "tests/golden/test.mpeg1video").read_bytes(), ctx, stream=0, vlen=VLEN)
t0 = time.time()
n = 0
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 400
while n < limit:
    t = time.time()
    f = v.decode()
    print(f"decode {n}: {'frame' if f is not None else 'end'} in {1e3 * (time.time() - t):.2f} ms", flush=True)
    if f is None:
        break
    n += 1
print(f"{n} frames in {time.time() - t0:.2f} s", flush=True)
