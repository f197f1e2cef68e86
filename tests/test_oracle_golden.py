"""The CPU oracle against every golden vector the reference's own tests hold for the hot path.

Mirrors mpeg_test.go:164-231 (TestAudioGolden, TestVideoGolden), mpeg_test.go:233-274 (TestVideo),
mpeg_test.go:135-162 (TestAudio), video_test.go:63-103 (runParitySweep), audio_test.go:36-64
(runSynthWindowParity), plus vectors evaluated from the reference's own Go statements
(tests/golden/make_golden_from_go.py).  CPU only.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

VIDEO_GOLDEN = 0xEA6D7FCB1340BA3F          # mpeg_test.go:227
AUDIO_GOLDEN_NOFMA = 0xF1B76CDF8E6CDEA5    # mpeg_test.go:194 (amd64, no FMA)
AUDIO_GOLDEN_WINFMA = 0x50F3AB75F5FB0FB5   # mpeg_test.go:195 (amd64 AVX2: window FMA)


def read(golden_dir, name):
    return (golden_dir / name).read_bytes()


def test_video_golden_hash(golden_dir):
    v = ol.VideoOracle(read(golden_dir, "test.mpeg1video"))
    h, frames = ol.FNV_OFFSET, 0
    while True:
        f = v.decode()
        if f is None:
            break
        h = ol.fnv(h, f.plane("y"))
        h = ol.fnv(h, f.plane("cb"))
        h = ol.fnv(h, f.plane("cr"))
        frames += 1
    assert h == VIDEO_GOLDEN, f"video hash {h:#018x} frames={frames}"


def test_video_header_and_first_frame(golden_dir):
    # mpeg_test.go:233-274
    v = ol.VideoOracle(read(golden_dir, "test.mpeg1video"))
    assert v.has_header()
    assert (v.width, v.height, v.framerate) == (160, 120, 30.0)
    f = v.decode()
    assert f is not None and f.width == 160
    assert f.plane("y").size == 20480
    assert f.plane("cb").size == 20480 // 4


@pytest.mark.parametrize("fma,want", [(False, AUDIO_GOLDEN_NOFMA), (True, AUDIO_GOLDEN_WINFMA)])
def test_audio_golden_hash(golden_dir, fma, want):
    a = ol.AudioOracle(read(golden_dir, "test.mp2"), fma=fma)
    h, frames = ol.FNV_OFFSET, 0
    while True:
        s = a.decode()
        if s is None:
            break
        h = ol.fnv(h, s)
        frames += 1
    assert h == want, f"audio hash {h:#018x} frames={frames}"


def test_audio_header(golden_dir):
    # mpeg_test.go:135-162
    a = ol.AudioOracle(read(golden_dir, "test.mp2"))
    assert a.has_header() and a.samplerate == 44100 and a.channels == 1
    a.rewind()
    assert a.decode() is not None


def test_demuxed_program_stream_decodes(golden_dir):
    # BASELINE config 1: testdata/test.mpg through the CPU path.  The PS carries the same
    # 160x120 video; every frame must decode and the elementary streams must be non-trivial.
    video, audio, nv, na = ol.demux_split(read(golden_dir, "test.mpg"))
    assert nv == 143 and na == 37          # 143 video PES, 37 audio PES
    v = ol.VideoOracle(video)
    assert (v.width, v.height) == (160, 120)
    n = 0
    while v.decode() is not None:
        n += 1
    assert n > 200 and v.oob_count() == 0
    a = ol.AudioOracle(audio)
    m = 0
    while a.decode() is not None:
        m += 1
    assert m > 30


# ---------------------------------------------------------------------------------------------
# idct / idct36 against vectors evaluated from the reference's Go source text
# ---------------------------------------------------------------------------------------------
PREMULT = np.array([
    32, 44, 42, 38, 32, 25, 17, 9, 44, 62, 58, 52, 44, 35, 24, 12, 42, 58, 55, 49, 42, 33, 23, 12,
    38, 52, 49, 44, 38, 30, 20, 10, 32, 44, 42, 38, 32, 25, 17, 9, 25, 35, 33, 30, 25, 20, 14, 7,
    17, 24, 23, 20, 17, 14, 9, 5, 9, 12, 12, 10, 9, 7, 5, 2], dtype=np.int64)


def test_idct_matches_go_statements(golden_dir):
    z = np.load(golden_dir / "go_idct_vectors.npz")
    L = ol.lib()
    i64p = C.POINTER(C.c_int64)
    for lv, n, want in zip(z["levels"], z["max_index"], z["out"]):
        blk = (lv.astype(np.int64) * PREMULT).copy()
        L.orc_idct(blk.ctypes.data_as(i64p), int(n))
        assert np.array_equal(blk, want)
        # the sparse branch and the full transform agree on valid inputs
        blk2 = (lv.astype(np.int64) * PREMULT).copy()
        L.orc_idct_full(blk2.ctypes.data_as(i64p))
        assert np.array_equal(blk2, want)


def test_idct_int32_headroom(golden_dir):
    # with clipped levels every intermediate of the transform fits int32, so the CUDA
    # kernel may use 32-bit integers.  Check the outputs of the adversarial all-+-2047 blocks stay
    # far inside int32 (the intermediate bound itself, 1.897e9, is asserted in test_host_logic).
    z = np.load(golden_dir / "go_idct_vectors.npz")
    assert np.abs(z["out"]).max() < 2**31 // 256


def test_idct36_matches_go_statements(golden_dir):
    z = np.load(golden_dir / "go_idct36_vectors.npz")
    L = ol.lib()
    for s, ss, dp, want in zip(z["s"], z["ss"], z["dp"], z["d"]):
        s64 = np.ascontiguousarray(s, dtype=np.int64)
        d = np.full(1024, 7.5, dtype=np.float32)
        L.orc_idct36(s64.ctypes.data_as(C.POINTER(C.c_int64)), int(ss), d.ctypes.data_as(C.POINTER(C.c_float)), int(dp))
        assert np.array_equal(d.view(np.uint32), want.view(np.uint32))


# ---------------------------------------------------------------------------------------------
# motion-compensation sweep, video_test.go:45-103
# ---------------------------------------------------------------------------------------------
def fill_test_frame(fs: ol.FrameSet, stream, buf, fill):
    # newTestFrame, video_test.go:45-59
    f = fs.frame(stream, buf)
    y, cb, cr = f.plane("y"), f.plane("cb"), f.plane("cr")
    i = np.arange(y.size, dtype=np.int64)
    y.reshape(-1)[:] = ((i * 131 + fill * 7) & 0xFF).astype(np.uint8)
    i = np.arange(cb.size, dtype=np.int64)
    cb.reshape(-1)[:] = ((i * 197 + fill * 13) & 0xFF).astype(np.uint8)
    cr.reshape(-1)[:] = ((i * 251 + fill * 29) & 0xFF).astype(np.uint8)


def mc_reference_numpy(src, stride, size, mh, mv, mb_row, mb_col):
    """copyMacroblockRef's per-plane closure (video_test.go:11-37) in numpy, on a flat buffer."""
    hp, vp = mh >> 1, mv >> 1
    odd_h, odd_v = mh & 1, mv & 1
    out = np.empty((size, size), dtype=np.uint8)
    s = src.astype(np.int32)
    for y in range(size):
        for x in range(size):
            si = ((mb_row * size) + vp + y) * stride + mb_col * size + hp + x
            if not odd_h and not odd_v:
                v = s[si]
            elif odd_h and not odd_v:
                v = (s[si] + s[si + 1] + 1) >> 1
            elif not odd_h and odd_v:
                v = (s[si] + s[si + stride] + 1) >> 1
            else:
                v = (s[si] + s[si + 1] + s[si + stride] + s[si + stride + 1] + 2) >> 2
            out[y, x] = v
    return out


def go_div2(v):  # Go's / truncates toward zero (video_test.go:39-40)
    return int(v / 2)


@pytest.mark.parametrize("impl", ["orc_copy_macroblock", "orc_copy_macroblock_swar"])
def test_copy_macroblock_parity_sweep(impl):
    L = ol.lib()
    fn = getattr(L, impl)
    fs = ol.FrameSet(1, 64, 64)  # lumaWidth 64, chromaWidth 32 (video_test.go:74)
    fill_test_frame(fs, 0, 0, 1)
    src = fs.frame(0, 0)
    for mb_row in (1, 2):
        for mb_col in (1, 2):
            for mh in range(-3, 4):
                for mv in range(-3, 4):
                    fill_test_frame(fs, 0, 1, 0)
                    want_y = fs.frame(0, 1).plane("y").copy()
                    want_cb = fs.frame(0, 1).plane("cb").copy()
                    want_cr = fs.frame(0, 1).plane("cr").copy()
                    assert fn(mh, mv, mb_row, mb_col, C.byref(src), C.byref(fs.frame(0, 1))) == 0
                    want_y[mb_row * 16:mb_row * 16 + 16, mb_col * 16:mb_col * 16 + 16] = mc_reference_numpy(
                        src.plane("y").reshape(-1), 64, 16, mh, mv, mb_row, mb_col)
                    cmh, cmv = go_div2(mh), go_div2(mv)
                    want_cb[mb_row * 8:mb_row * 8 + 8, mb_col * 8:mb_col * 8 + 8] = mc_reference_numpy(
                        src.plane("cb").reshape(-1), 32, 8, cmh, cmv, mb_row, mb_col)
                    want_cr[mb_row * 8:mb_row * 8 + 8, mb_col * 8:mb_col * 8 + 8] = mc_reference_numpy(
                        src.plane("cr").reshape(-1), 32, 8, cmh, cmv, mb_row, mb_col)
                    got = fs.frame(0, 1)
                    assert np.array_equal(got.plane("y"), want_y), (mb_row, mb_col, mh, mv)
                    assert np.array_equal(got.plane("cb"), want_cb), (mb_row, mb_col, mh, mv)
                    assert np.array_equal(got.plane("cr"), want_cr), (mb_row, mb_col, mh, mv)


def test_copy_macroblock_rejects_windows_outside_the_buffer():
    # where the Go code panics (negative index / past cap), the oracle refuses and writes nothing
    L = ol.lib()
    fs = ol.FrameSet(1, 64, 64)
    fill_test_frame(fs, 0, 0, 1)
    before = fs.whole(0, 1).copy()
    assert L.orc_copy_macroblock(-2, -2, 0, 0, C.byref(fs.frame(0, 0)), C.byref(fs.frame(0, 1))) == -1
    assert np.array_equal(before, fs.whole(0, 1))
    # a window that runs off the bottom of Y reads into Cb (contiguous planes): allowed
    assert L.orc_copy_macroblock(0, 8, 3, 0, C.byref(fs.frame(0, 0)), C.byref(fs.frame(0, 1))) == 0
    src = fs.whole(0, 0)
    got = fs.frame(0, 1).plane("y")[48:64, 0:16]
    want = src[(48 + 4) * 64:(48 + 4 + 16) * 64].reshape(16, 64)[:, :16]
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------
# synthesis-window sweep, audio_test.go:36-64
# ---------------------------------------------------------------------------------------------
def window_inputs():
    i = np.arange(1024)
    d = ((i * 7) % 101 - 50).astype(np.float32) * np.float32(0.013)
    v = ((i * 13) % 97 - 48).astype(np.float32) * np.float32(0.011)
    return d, v


def window_reference_numpy(d, v, v_pos, fused=False):
    # synthWindowRef, audio_test.go:9-31
    u = np.zeros(32, dtype=np.float32)
    di = 512 - (v_pos >> 1)
    vi = (v_pos % 128) >> 1

    def tap(u, di, vi):
        if fused:
            return (d[di:di + 32].astype(np.float64) * v[vi:vi + 32].astype(np.float64) + u.astype(np.float64)).astype(np.float32)
        return u + d[di:di + 32] * v[vi:vi + 32]

    while vi < 1024:
        u = tap(u, di, vi)
        vi += 128
        di += 64
    di -= 512 - 32
    vi = (128 - 32 + 1024) - vi
    while vi < 1024:
        u = tap(u, di, vi)
        vi += 128
        di += 64
    return u


def test_synth_window_parity_sweep():
    L = ol.lib()
    d, v = window_inputs()
    f32p = C.POINTER(C.c_float)
    for v_pos in range(0, 1024, 64):
        got = np.zeros(32, dtype=np.float32)
        L.orc_synth_window(got.ctypes.data_as(f32p), d.ctypes.data_as(f32p), v.ctypes.data_as(f32p), v_pos)
        want = window_reference_numpy(d, v, v_pos)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), v_pos  # tol 0, audio_amd64_test.go:8
        fused = np.zeros(32, dtype=np.float32)
        L.orc_synth_window_fma(fused.ctypes.data_as(f32p), d.ctypes.data_as(f32p), v.ctypes.data_as(f32p), v_pos)
        want_f = window_reference_numpy(d, v, v_pos, fused=True)
        assert np.array_equal(fused.view(np.uint32), want_f.view(np.uint32)), v_pos
        # the 1e-5 rule the reference applies to its FMA kernel (audio_test.go:59, audio_amd64_test.go:15-16)
        assert np.all(np.abs(fused - want) <= 1e-5 * (1 + np.abs(want)))


def test_synthesis_window_table_is_duplicated():
    # audio.go:95-98
    w = np.ctypeslib.as_array(ol.lib().orc_synthesis_window_1024(), shape=(1024,))
    assert np.array_equal(w[:512], w[512:])
    assert w[0] == 0.0 and w[1] == -0.5 and w[256] == 37519.0


def test_synth_frame_equals_decoder(golden_dir):
    # the record-level synthesis entry (what the CUDA kernel is compared with) reproduces the
    # bitstream decoder: feed it the decoder's own requantised samples frame by frame
    a = ol.AudioOracle(read(golden_dir, "test.mp2"))
    st = ol.synth_states(1)
    h = ol.FNV_OFFSET
    for _ in range(40):
        want = a.decode()
        got = ol.synth_batch(st, 1, 1, a.last_samples())
        assert np.array_equal(got.reshape(-1).view(np.uint32), want.view(np.uint32))
    v, v_pos = a.state()
    assert v_pos == st[0].v_pos
    assert np.array_equal(np.ctypeslib.as_array(st[0].v).reshape(2, 1024), v)


def test_rgba_against_go_fixture(golden_dir):
    """Frame.RGBA() is Go standard-library arithmetic (image/draw of a 4:2:0 image.YCbCr) and the one function of the
    path the oracle restates without a reference vector ("parity unpinned").  go/rgba_fixture.go
    writes that vector on any machine with Go; when the file exists the oracle must reproduce it byte for byte."""
    import struct
    path = golden_dir / "go_rgba_fixture.bin"
    if not path.exists():
        pytest.skip("tests/golden/go_rgba_fixture.bin absent (no Go toolchain in the build image): run `go run go/rgba_fixture.go`")
    data = path.read_bytes()
    assert data[:8] == b"RGBAFIX1"
    (n,) = struct.unpack_from("<I", data, 8)
    off = 12
    for _ in range(n):
        w, h, lw, lh = struct.unpack_from("<4I", data, off)
        off += 16
        ny, nc = lw * lh, (lw // 2) * (lh // 2)
        fs = ol.FrameSet(1, w, h)
        assert (fs.luma_w, fs.luma_h) == (lw, lh)
        buf = fs.whole(0, 0)
        buf[:ny + 2 * nc] = np.frombuffer(data, np.uint8, ny + 2 * nc, off)
        off += ny + 2 * nc
        want = np.frombuffer(data, np.uint8, w * h * 4, off).reshape(h, w, 4)
        off += w * h * 4
        assert np.array_equal(fs.rgba(0, 0), want), f"{w}x{h}"
