#!/usr/bin/env python3
"""Golden vectors for idct (video.go:801-928) and idct36 (audio.go:492-772).

There is no Go toolchain in this image, so the reference cannot be run.  Both functions
are straight-line integer / float32 arithmetic, though, and Go's expression syntax for
them is a subset of Python's: this script lifts the two function bodies out of the
reference's source text, rewrites the handful of Go-only tokens (``:=``, C-style ``for``,
braces, ``float32(...)``, untyped float literals) and executes the result with Python
ints (arbitrary precision, arithmetic ``>>`` like Go's) and numpy.float32 scalars (each
operation rounded to float32, like Go's float32 arithmetic on amd64 without FMA).
The vectors are therefore produced by the reference's own statements, not by the oracle.

Only runs where /root/reference exists (the build container).  Outputs are committed:
    tests/golden/go_idct_vectors.npz, tests/golden/go_idct36_vectors.npz
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent


def func_body(src: str, signature: str) -> str:
    start = src.index(signature)
    i = src.index("{", start)
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src[i + 1 : j]


def go_to_python(body: str, float_mode: bool) -> str:
    out, depth = [], 1
    for raw in body.splitlines():
        line = raw.strip()
        if not line or line.startswith("//"):
            continue
        line = re.sub(r"\s*//.*$", "", line)
        if line.startswith("var "):
            # declarations; multi-line var lists end on a line carrying the type name
            continue
        if re.fullmatch(r"[a-z0-9, ]+(int|float32)?", line) and "=" not in line:
            continue  # continuation lines of a var list
        if line == "}":
            depth -= 1
            continue
        if line == "} else {":
            out.append("    " * (depth - 1) + "else:")
            continue
        m = re.fullmatch(r"for (\w+) := (\d+); \1 < (\d+); \1(\+\+| \+= (\d+)) \{", line)
        if m:
            step = m.group(5) or "1"
            out.append("    " * depth + f"for {m.group(1)} in range({m.group(2)}, {m.group(3)}, {step}):")
            depth += 1
            continue
        m = re.fullmatch(r"if (.*) \{", line)
        if m:
            out.append("    " * depth + f"if {m.group(1)}:")
            depth += 1
            continue
        line = line.replace(":=", "=")
        if float_mode:
            line = line.replace("float32(", "F(")
            # untyped float constants take the float32 type of the other operand
            line = re.sub(r"(?<![\w.\[])(\d+\.\d+)", r"F(\1)", line)
        out.append("    " * depth + line)
    return "\n".join(out)


def build_idct(src):
    body = go_to_python(func_body(src, "func idct(block *[64]int, maxIndex int)"), float_mode=False)
    code = "def idct(block, maxIndex):\n" + body + "\n"
    ns = {}
    exec(code, ns)
    return ns["idct"], code


def build_idct36(src):
    body = go_to_python(func_body(src, "func idct36(s *[32][3]int, ss int, d *[1024]float32, dp int)"), float_mode=True)
    code = "def idct36(s, ss, d, dp):\n" + body + "\n"
    ns = {"F": np.float32}
    exec(code, ns)
    return ns["idct36"], code


PREMULT = None


def premult(src):
    body = re.search(r"var videoPremultiplierMatrix = \[64\]byte\{(.*?)\}", src, re.S).group(1)
    return np.array([int(x) for x in re.findall(r"\d+", body)], dtype=np.int64)


def zigzag(src):
    body = re.search(r"var videoZigZag = \[64\]byte\{(.*?)\}", src, re.S).group(1)
    return [int(x) for x in re.findall(r"\d+", body)]


def main():
    vsrc = (REF / "video.go").read_text()
    asrc = (REF / "audio.go").read_text()
    idct, idct_code = build_idct(vsrc)
    idct36, idct36_code = build_idct36(asrc)
    pm = premult(vsrc)
    zz = zigzag(vsrc)
    rng = np.random.default_rng(20260925)

    # ---- idct: blocks as decodeBlock produces them: level (clipped, video.go:737-741) * premultiplier,
    # non-zero only at zig-zag positions < n; n is passed as maxIndex (video.go:779,792).
    levels_in, n_in, outs = [], [], []
    cases = []
    for _ in range(1500):
        n = int(rng.integers(2, 65))
        lv = np.zeros(64, dtype=np.int64)
        scale = float(rng.choice([3.0, 30.0, 300.0, 3000.0]))
        for pos in range(n):
            if rng.random() < 0.7:
                lv[zz[pos]] = int(np.clip(np.rint(rng.laplace(0, scale / (1 + 0.2 * pos))), -2048, 2047))
        cases.append((lv, n))
    for sign_pattern in range(64):  # adversarial: every level at the clip limits (Q10 of SURVEY)
        bits = rng.integers(0, 2, 64)
        lv = np.where(bits == 1, 2047, -2048).astype(np.int64)
        if sign_pattern == 0:
            lv[:] = 2047
        if sign_pattern == 1:
            lv[:] = -2048
        cases.append((lv, 64))
    for dc in (0, 1, 128, 255, 1023, 2047, 4095, -4096, -1):  # intra DC as dc<<8 (video.go:672), n = 2..
        lv = np.zeros(64, dtype=np.int64)
        lv[0] = dc * 8
        lv[1] = 3
        cases.append((lv, 2))
    for lv, n in cases:
        block = [int(x) for x in (lv * pm)]
        idct(block, n)
        levels_in.append(lv)
        n_in.append(n)
        outs.append(block)
    np.savez_compressed(
        OUT / "go_idct_vectors.npz",
        levels=np.array(levels_in, dtype=np.int32),
        max_index=np.array(n_in, dtype=np.int32),
        out=np.array(outs, dtype=np.int64),
    )

    # ---- idct36
    s_in, ss_in, dp_in, d_out = [], [], [], []
    for k in range(400):
        amp = int(rng.choice([10, 1000, 30000, 65536]))
        s = rng.integers(-amp, amp + 1, size=(32, 3)).astype(np.int64)
        if k % 7 == 0:
            s[27:, :] = 0
        if k == 1:
            s[:] = 65536
        if k == 2:
            s[:] = -65536
        if k == 3:
            s[:] = 0
        ss = int(rng.integers(0, 3))
        dp = int(rng.integers(0, 16)) * 64
        d = [np.float32(7.5)] * 1024
        idct36([[int(v) for v in row] for row in s], ss, d, dp)
        s_in.append(s)
        ss_in.append(ss)
        dp_in.append(dp)
        d_out.append(np.array(d, dtype=np.float32))
    np.savez_compressed(
        OUT / "go_idct36_vectors.npz",
        s=np.array(s_in, dtype=np.int32),
        ss=np.array(ss_in, dtype=np.int32),
        dp=np.array(dp_in, dtype=np.int32),
        d=np.array(d_out, dtype=np.float32),
    )
    print("idct cases:", len(cases), " idct36 cases:", len(s_in))


if __name__ == "__main__":
    main()
