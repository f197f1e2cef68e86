"""CPU-only: the variable-width ("vlen") transfer form of the coefficient blocks (include/mpegb200.h).
The host packer (mpegb200_pack_coeffs_vlen, pure host code) is checked against an independent numpy reading of the
format as the header documents it; the device expansion is checked in tests/test_gpu_video.py."""
import ctypes as C

import numpy as np
import pytest

from mpeg_b200 import _lib
import workload as wl


def pack(coeffs):
    L = _lib.load()
    coeffs = np.ascontiguousarray(coeffs, dtype=np.int16).reshape(-1, 64)
    n = len(coeffs)
    headers = np.zeros(n, np.uint32)
    chunks = np.zeros((n + 31) // 32, np.uint64)
    cap = int(L.mpegb200_vlen_payload_bound(n))
    payload = np.zeros(cap, np.uint8)
    used = C.c_size_t(0)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    rc = L.mpegb200_pack_coeffs_vlen(vp(coeffs), n, vp(headers), vp(chunks), vp(payload), cap, C.byref(used))
    return rc, headers, chunks, payload[:used.value]


def unpack_reference(headers, chunks, payload):
    """Slow reading of the format, straight from the header's description."""
    n = len(headers)
    out = np.zeros((n, 64), np.int16)
    for b in range(n):
        if b % 32 == 0:
            off = int(chunks[b // 32])
        h = int(headers[b])
        for g in range(8):
            code = (h >> (4 * g)) & 15
            assert code <= 14
            if code == 0:
                continue
            w = 12 if code == 13 else 16 if code == 14 else code
            bits = int.from_bytes(payload[off:off + w].tobytes(), "little")
            off += w
            for i in range(8):
                c = (bits >> (i * w)) & ((1 << w) - 1)
                if c >= 1 << (w - 1):
                    c -= 1 << w
                x = c if code >= 13 else 2 * c - (c > 0) + (c < 0)
                out[b, wl.ZIGZAG[8 * g + i]] = x
    return out


def dense_blocks(n, seed):
    rng = np.random.default_rng(seed)
    return wl._draw_blocks(rng, n, np.zeros(n, bool), dense=True)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000])
def test_round_trip_dense(n):
    coeffs = dense_blocks(n, n)
    rc, h, c, p = pack(coeffs)
    assert rc == 0
    assert np.array_equal(unpack_reference(h, c, p), coeffs)
    assert len(p) >= 16 and not p[-16:].any()          # the padding the kernel's word reads rely on


def test_round_trip_natural_intra_and_extremes():
    rng = np.random.default_rng(7)
    n = 600
    intra = rng.random(n) < 0.3
    coeffs = wl._draw_blocks(rng, n, intra, dense=False)       # sparse blocks, even intra DC values
    coeffs[0] = 0                                              # an all-zero block: header 0, no payload
    coeffs[1, :] = 2047
    coeffs[2, :] = -2048                                       # even: raw groups
    coeffs[3, :] = -2047
    coeffs[4, ::2] = 1
    coeffs[4, 1::2] = -1
    coeffs[5, 63] = 2                                          # a lone even value in the last group
    rc, h, c, p = pack(coeffs)
    assert rc == 0
    assert h[0] == 0
    assert h[2] == 0xDDDDDDDD
    assert np.array_equal(unpack_reference(h, c, p), coeffs)


def test_sizes():
    """Dense blocks of BASELINE config 3: about 49 bytes instead of 128; sparse blocks: a few bytes."""
    coeffs = dense_blocks(4096, 3)
    rc, h, c, p = pack(coeffs)
    assert rc == 0
    per_block = 4 + (len(p) - 16) / len(coeffs) + 8 / 32
    assert per_block < 52, per_block
    rng = np.random.default_rng(9)
    sparse = wl._draw_blocks(rng, 4096, np.zeros(4096, bool), dense=False)
    rc, h, c, p = pack(sparse)
    assert rc == 0
    assert 4 + (len(p) - 16) / 4096 < 24


def test_out_of_range_and_capacity():
    L = _lib.load()
    wide = np.zeros((3, 64), np.int16)                         # values outside 12 bits travel as raw 16-bit groups (code 14)
    wide[0, 10] = 2048
    wide[1, 0] = 32760                                         # an intra DC of a damaged stream (dc * 8, |dc| <= 4095)
    wide[1, 63] = -32768
    wide[2, :] = np.arange(-32, 32) * 1000
    rc, h, c, p = pack(wide)
    assert rc == 0 and ((h[0] >> (4 * (int(np.flatnonzero(wl.ZIGZAG == 10)[0]) // 8))) & 15) == 14
    assert np.array_equal(unpack_reference(h, c, p), wide)
    coeffs = dense_blocks(64, 1)
    headers = np.zeros(64, np.uint32)
    chunks = np.zeros(2, np.uint64)
    payload = np.zeros(64, np.uint8)                           # far too small
    used = C.c_size_t(0)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    assert L.mpegb200_pack_coeffs_vlen(vp(coeffs), 64, vp(headers), vp(chunks), vp(payload), 64, C.byref(used)) != 0
    assert used.value > 64                                     # tells the caller what it needs


def test_round_trip_property():
    """Any int16 block content survives, whatever mixture of zeros, odd and even values the groups hold."""
    from hypothesis import given, settings, strategies as st
    import hypothesis.extra.numpy as hnp

    value = st.one_of(st.just(0), st.integers(-2048, 2047), st.integers(-32768, 32767), st.sampled_from([-2047, -1, 1, 2047, -2048, 2046]),
                      st.integers(-40, 40).map(lambda t: 2 * t + 1))

    @settings(max_examples=25, deadline=None)
    @given(hnp.arrays(np.int16, st.tuples(st.integers(1, 40), st.just(64)), elements=value))
    def check(coeffs):
        rc, h, c, p = pack(coeffs)
        assert rc == 0
        assert np.array_equal(unpack_reference(h, c, p), coeffs)
        # the size the header promises: 4-bit codes -> bytes
        size = lambda code: 12 if code == 13 else 16 if code == 14 else code
        want = sum(size((int(x) >> (4 * g)) & 15) for x in h for g in range(8))
        assert len(p) == want + 16

    check()


def test_validate_foreign_streams():
    """mpegb200_vlen_validate: accepts what the packer produces, rejects codes above 14, chunk offsets that are not back to
    back, and a payload size that does not match the headers."""
    L = _lib.load()
    vp = lambda a: C.c_void_p(a.ctypes.data)
    coeffs = dense_blocks(100, 5)
    rc, h, c, p = pack(coeffs)
    assert rc == 0
    ok = lambda hh, cc, nbytes: L.mpegb200_vlen_validate(vp(hh), vp(cc), len(hh), nbytes)
    assert ok(h, c, len(p)) == 0
    assert ok(h, c, len(p) - 1) != 0
    bad = h.copy(); bad[7] |= 0xF            # code 15
    assert ok(bad, c, len(p)) != 0
    bad = c.copy(); bad[1] += 1
    assert ok(h, bad, len(p)) != 0
    bad = h.copy(); bad[3] = 0               # a block that claims no payload: sizes no longer add up
    if h[3] != 0:
        assert ok(bad, c, len(p)) != 0
    assert L.mpegb200_vlen_validate(None, None, 0, 0) == 0


def test_portable_packer_agrees_with_the_simd_one():
    """The packer picks AVX2 / BMI2 forms at run time; MPEGB200_PACK_PORTABLE=1 forces the plain C++ forms.  Both must
    produce the same bytes (checked in a subprocess, the choice is made once per process)."""
    import os
    import subprocess
    import sys
    import hashlib
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import test_vlen_format as t\n"
        "import workload as wl\n"
        "rng = np.random.default_rng(11)\n"
        "a = t.dense_blocks(700, 2)\n"
        "b = wl._draw_blocks(rng, 700, rng.random(700) < 0.3, dense=False)\n"
        "b[5, :] = -2048; b[6, :] = 2047; b[7, :] = 0; b[8, :] = 30000; b[9, 3] = -32768; b[10, 60] = 2048\n"
        "h = hashlib.sha256()\n"
        "for c in (a, b):\n"
        "    rc, hd, ch, p = t.pack(c)\n"
        "    assert rc == 0\n"
        "    assert np.array_equal(t.unpack_reference(hd[:64], ch[:2], p), c[:64])\n"
        "    for x in (hd, ch, p): h.update(x.tobytes())\n"
        "print(h.hexdigest())\n"
    ) % (str(_lib.PKG.parent), str(_lib.PKG.parent / "tests"))
    outs = []
    for portable in ("0", "1"):
        env = dict(os.environ, MPEGB200_PACK_PORTABLE=portable)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and len(outs[0]) == 64
