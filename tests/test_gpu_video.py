"""Parity of the CUDA video path (through the C-ABI) with the CPU oracle.  Needs a B200.

Bit-exact everywhere: Y/Cb/Cr planes are integer work.  Reads like the reference's own tests:
TestVideoGolden (mpeg_test.go:203-231), runParitySweep (video_test.go:63-103)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
import workload as wl
from mpeg_b200.packing import resolve_rewrites

pytestmark = pytest.mark.gpu

VIDEO_GOLDEN = 0xEA6D7FCB1340BA3F


@pytest.fixture(scope="module")
def ctx():
    import mpeg_b200
    c = mpeg_b200.Context(device=0, max_streams=64)
    yield c
    c.close()


def fresh_stream(ctx, sid, w, h):
    try:
        ctx.video_close(sid)
    except Exception:
        pass
    ctx.video_open(sid, w, h)


def assert_frames_equal(ctx, fs, stream, oracle_stream=None, msg=""):
    o = stream if oracle_stream is None else oracle_stream
    for b in range(3):
        got = ctx.video_read_frame(stream, b)
        want = fs.whole(o, b)
        if not np.array_equal(got, want):
            bad = np.flatnonzero(got != want)
            raise AssertionError(f"{msg} stream {stream} buf {b}: {len(bad)} bytes differ, first at {bad[0]} "
                                 f"(got {got[bad[0]]}, want {want[bad[0]]})")


def test_golden_clip_through_gpu(ctx, golden_dir):
    """TestVideoGolden with the kernels swapped in: the oracle's parser emits the packed records,
    the GPU does MC + IDCT + add, the returned frames hash to the reference's golden value."""
    data = (golden_dir / "test.mpeg1video").read_bytes()
    v = ol.VideoOracle(data, tap=True)
    fresh_stream(ctx, 0, v.width, v.height)
    h, frames = ol.FNV_OFFSET, 0
    while True:
        f = v.decode()
        pics, mbs, coeffs = v.tap()
        if len(mbs):
            # one launch per picture: consecutive pictures of a stream depend on each other
            for i in range(len(pics)):
                p = pics[i:i + 1].copy()
                m = mbs[p["first_mb"][0]:p["first_mb"][0] + p["n_mb"][0]].copy()
                m["pic"] = 0
                p["first_mb"] = 0
                c = coeffs_slice(m, coeffs)
                # the clip revisits macroblocks inside a picture (overlapping slices): the packer resolves
                # that into launches without double writes, the kernels never see a duplicate
                for wp, wm, wc in resolve_rewrites(p, m, c):
                    ctx.video_validate(wp, wm, len(wc))
                    ctx.video_decode_pictures(wp, wm, wc)
        if f is None:
            break
        y, cb, cr = ctx.video_read_planes(0, v.last_buf())
        assert np.array_equal(y, f.plane("y")), f"frame {frames}"
        h = ol.fnv(h, y)
        h = ol.fnv(h, cb)
        h = ol.fnv(h, cr)
        frames += 1
    assert h == VIDEO_GOLDEN, f"{h:#018x} after {frames} frames"


def coeffs_slice(m, coeffs):
    """Coefficient blocks of a record slice, re-based so that coeff_block starts at 0."""
    if len(m) == 0:
        return np.zeros((0, 64), np.int16)
    first = int(m["coeff_block"][0])
    n = int(sum(bin(int(c)).count("1") for c in m["cbp"]))
    m["coeff_block"] -= first
    return coeffs[first:first + n]


def test_copy_macroblock_parity_sweep_gpu(ctx):
    """video_test.go:63-103 on the GPU: every half-pel mode, negative vectors, chroma rounding."""
    fs = ol.FrameSet(1, 64, 64)
    fresh_stream(ctx, 1, 64, 64)
    from test_oracle_golden import fill_test_frame
    fill_test_frame(fs, 0, 1, 1)  # source = forward buffer 1, fill 1
    ctx.video_write_frame(1, 1, fs.whole(0, 1))
    recs, want = [], []
    for mb_row in (1, 2):
        for mb_col in (1, 2):
            for mh in range(-3, 4):
                for mv in range(-3, 4):
                    fill_test_frame(fs, 0, 0, 0)
                    ctx.video_write_frame(1, 0, fs.whole(0, 0))
                    m = np.zeros(1, wl.MB_DTYPE)
                    m["mb_row"], m["mb_col"], m["mv_h"], m["mv_v"] = mb_row, mb_col, mh, mv
                    m["flags"] = wl.MB_PREDICT
                    p = np.zeros(1, wl.PICTURE_DTYPE)
                    p[0] = (1, wl.PIC_P, 0, 1, 2, 0, 1)
                    ctx.video_validate(p, m, 0)
                    ctx.video_decode_pictures(p, m, np.zeros((0, 64), np.int16))
                    assert ol.lib().orc_copy_macroblock(mh, mv, mb_row, mb_col, C.byref(fs.frame(0, 1)),
                                                        C.byref(fs.frame(0, 0))) == 0
                    got = ctx.video_read_frame(1, 0)
                    assert np.array_equal(got, fs.whole(0, 0)), (mb_row, mb_col, mh, mv)


@pytest.mark.parametrize("geometry,n_pictures", [(wl.CIF, 30), (wl.Geometry(176, 144), 12), (wl.Geometry(200, 120), 9)])
def test_synthetic_sequence_bit_exact(ctx, geometry, n_pictures):
    """BASELINE config 2 (ii): one stream, coding order I P B B P B B ..., natural distributions,
    all three buffers compared after every picture."""
    g = geometry
    rng = wl.stream_rng(2, 0)
    fs = ol.FrameSet(1, g.width, g.height)
    fresh_stream(ctx, 2, g.width, g.height)
    for b in range(3):  # start from arbitrary (stale) buffer contents, SURVEY Q6
        buf = wl.random_reference_frame(rng, g)
        fs.whole(0, b)[:] = buf
        ctx.video_write_frame(2, b, buf)
    rot = wl.BufferRotation()
    types = [wl.PIC_I] + [wl.PIC_P, wl.PIC_B, wl.PIC_B] * 20
    for i in range(n_pictures):
        t = types[i]
        dst, fwd, bwd = rot.begin(t)
        mbs, coeffs = wl.make_picture(rng, g, t, "natural")
        pics, mbs, coeffs = wl.batch_pictures([(mbs, coeffs)], [2], t, [(dst, fwd, bwd)])
        ctx.video_validate(pics, mbs, len(coeffs))
        ctx.video_decode_pictures(pics, mbs, coeffs)
        opics = pics.copy()
        opics["stream"] = 0
        assert fs.exec_pictures(opics, mbs, coeffs) == 0
        rot.end(t)
        assert_frames_equal(ctx, fs, 2, 0, msg=f"picture {i} type {t}")


def test_dense_720p_batch_bit_exact(ctx):
    """BASELINE config 3 at reduced stream count: dense-P 720p, several streams in one launch."""
    g = wl.HD720
    n = 6
    fs = ol.FrameSet(n, g.width, g.height)
    per, bufs = [], []
    for s in range(n):
        rng = wl.stream_rng(3, s)
        fresh_stream(ctx, 10 + s, g.width, g.height)
        ref = wl.random_reference_frame(rng, g)
        fs.whole(s, 1)[:] = ref
        ctx.video_write_frame(10 + s, 1, ref)
        per.append(wl.make_picture(rng, g, wl.PIC_P, "dense"))
        bufs.append((0, 1, 2))
    pics, mbs, coeffs = wl.batch_pictures(per, [10 + s for s in range(n)], wl.PIC_P, bufs)
    ctx.video_validate(pics, mbs, len(coeffs))
    ctx.video_decode_pictures(pics, mbs, coeffs)
    opics = pics.copy()
    opics["stream"] = np.arange(n)
    assert fs.exec_pictures(opics, mbs, coeffs, threads=ol.lib().orc_max_threads()) == 0
    for s in range(n):
        assert_frames_equal(ctx, fs, 10 + s, s, msg="dense 720p")
    # size-independent properties on the same batch: idempotence (same inputs -> same output)
    before = [ctx.video_read_frame(10 + s, 0) for s in range(n)]
    ctx.video_decode_pictures(pics, mbs, coeffs)
    for s in range(n):
        assert np.array_equal(before[s], ctx.video_read_frame(10 + s, 0))
    for s in range(n):
        ctx.video_close(10 + s)


@pytest.mark.parametrize("mode,pic_type", [("dense", wl.PIC_P), ("natural", wl.PIC_B)])
def test_full_size_config3_parity(mode, pic_type):
    """BASELINE configs[2] at its full size -- 256 streams of 1280x720 in ONE launch, 921,600 macroblocks -- dense P (the benchmarked
    step) and natural B (both references, the per-macroblock-box mode): every plane of every stream against the oracle executing
    the same records on all host threads, plus the size-independent property that a second decode of the same records into the
    same buffers changes nothing.  (bench.py repeats the dense-P comparison inside every run.)"""
    import mpeg_b200
    S, g = 256, wl.HD720
    per, refs = [], []
    for s in range(S):
        rng = wl.stream_rng(3, 5000 + s)
        refs.append((wl.random_reference_frame(rng, g), wl.random_reference_frame(rng, g)))
        per.append(wl.make_picture(rng, g, pic_type, mode))
    pics, mbs, coeffs = wl.batch_pictures(per, list(range(S)), pic_type, [(0, 1, 2)] * S)
    fs = ol.FrameSet(S, g.width, g.height)
    with mpeg_b200.Context(device=0, max_streams=S) as c:
        for s in range(S):
            c.video_open(s, g.width, g.height)
            for b in (1, 2):
                fs.whole(s, b)[:] = refs[s][b - 1]
                c.video_write_frame(s, b, refs[s][b - 1])
        c.video_validate(pics, mbs, len(coeffs))
        c.video_decode_pictures(pics, mbs, coeffs)
        assert fs.exec_pictures(pics, mbs, coeffs, threads=ol.lib().orc_max_threads()) == 0
        first = []
        for s in range(S):
            got = c.video_read_frame(s, 0)
            assert np.array_equal(got, fs.whole(s, 0)), f"{mode}: stream {s} differs from the oracle"
            if s % 37 == 0:
                first.append((s, got))
        c.video_decode_pictures(pics, mbs, coeffs)
        for s, before in first:
            assert np.array_equal(before, c.video_read_frame(s, 0))
    fs.close()


def test_zero_residual_zero_vector_picture_is_a_copy(ctx):
    """A P picture whose macroblocks all have a zero vector and no coded blocks reproduces the
    reference picture byte for byte (skipped macroblocks, video.go:503-510)."""
    g = wl.CIF
    rng = wl.stream_rng(2, 7)
    fresh_stream(ctx, 3, g.width, g.height)
    ref = wl.random_reference_frame(rng, g)
    ctx.video_write_frame(3, 1, ref)
    mbs = np.zeros(g.n_mb, wl.MB_DTYPE)
    mbs["mb_row"] = np.repeat(np.arange(g.mb_h), g.mb_w)
    mbs["mb_col"] = np.tile(np.arange(g.mb_w), g.mb_h)
    mbs["flags"] = wl.MB_PREDICT
    pics = np.zeros(1, wl.PICTURE_DTYPE)
    pics[0] = (3, wl.PIC_P, 0, 1, 2, 0, len(mbs))
    ctx.video_decode_pictures(pics, mbs, np.zeros((0, 64), np.int16))
    got = ctx.video_read_frame(3, 0)
    assert np.array_equal(got[:g.picture_bytes], ref[:g.picture_bytes])
    assert not got[g.picture_bytes:].any()  # the pad is never written


def test_ragged_and_edge_batches(ctx):
    g = wl.Geometry(64, 48)
    rng = wl.stream_rng(2, 11)
    fs = ol.FrameSet(2, g.width, g.height)
    for s in (4, 5):
        fresh_stream(ctx, s, g.width, g.height)
        for b in range(3):
            buf = wl.random_reference_frame(rng, g)
            fs.whole(s - 4, b)[:] = buf
            ctx.video_write_frame(s, b, buf)
    # empty batch: a no-op
    ctx.video_decode_pictures(np.zeros(0, wl.PICTURE_DTYPE), np.zeros(0, wl.MB_DTYPE), np.zeros((0, 64), np.int16))
    # ragged: stream 4 gets 5 scattered macroblocks, stream 5 a full B picture; untouched macroblocks keep stale pixels
    m4, c4 = wl.make_picture(rng, g, wl.PIC_P, "natural")
    keep = np.array([0, 3, 4, 7, 11])
    cnt = np.array([bin(int(c)).count("1") for c in m4["cbp"]])
    sel = np.concatenate([np.arange(m4["coeff_block"][k], m4["coeff_block"][k] + cnt[k]) for k in keep]).astype(int)
    m4k = m4[keep].copy()
    m4k["coeff_block"] = np.cumsum(cnt[keep]) - cnt[keep]
    c4k = c4[sel]
    m5, c5 = wl.make_picture(rng, g, wl.PIC_B, "natural")
    pics, mbs, coeffs = wl.batch_pictures([(m4k, c4k), (m5, c5)], [4, 5], wl.PIC_B, [(0, 1, 2), (2, 0, 1)])
    ctx.video_validate(pics, mbs, len(coeffs))
    ctx.video_decode_pictures(pics, mbs, coeffs)
    opics = pics.copy()
    opics["stream"] = [0, 1]
    assert fs.exec_pictures(opics, mbs, coeffs) == 0
    assert_frames_equal(ctx, fs, 4, 0, "ragged")
    assert_frames_equal(ctx, fs, 5, 1, "full B")
    # an intra macroblock with a hole in its cbp keeps the old pixels of the missing block (aborted block)
    m = np.zeros(1, wl.MB_DTYPE)
    m["mb_row"], m["mb_col"], m["flags"], m["cbp"] = 1, 2, wl.MB_INTRA, 0b101101
    co = wl._draw_blocks(rng, 4, np.ones(4, bool), dense=False)
    p = np.zeros(1, wl.PICTURE_DTYPE)
    p[0] = (4, wl.PIC_I, 1, 0, 2, 0, 1)
    ctx.video_decode_pictures(p, m, co)
    op = p.copy()
    op["stream"] = 0
    assert fs.exec_pictures(op, m, co) == 0
    assert_frames_equal(ctx, fs, 4, 0, "intra with holes")
    # windows running off the bottom of a plane read the next plane (contiguous planes)
    m = np.zeros(2, wl.MB_DTYPE)
    m["mb_row"], m["mb_col"] = g.mb_h - 1, [0, 1]
    m["mv_v"], m["mv_h"], m["flags"] = [9, 15], [1, 0], wl.MB_PREDICT
    p[0] = (4, wl.PIC_P, 2, 1, 0, 0, 2)
    ctx.video_validate(p, m, 0)
    ctx.video_decode_pictures(p, m, np.zeros((0, 64), np.int16))
    op = p.copy()
    op["stream"] = 0
    assert fs.exec_pictures(op, m, np.zeros((0, 64), np.int16)) == 0
    assert_frames_equal(ctx, fs, 4, 0, "over-read into next plane")


def test_validate_rejects_bad_records(ctx):
    import mpeg_b200
    g = wl.Geometry(64, 48)
    fresh_stream(ctx, 6, g.width, g.height)
    p = np.zeros(1, wl.PICTURE_DTYPE)
    p[0] = (6, wl.PIC_P, 0, 1, 2, 0, 1)
    m = np.zeros(1, wl.MB_DTYPE)
    m["flags"] = wl.MB_PREDICT
    ctx.video_validate(p, m, 0)

    def bad(**kw):
        mm, pp, nb = m.copy(), p.copy(), kw.pop("n_blocks", 0)
        for k, v in kw.items():
            (pp if k in pp.dtype.names else mm)[k] = v
        with pytest.raises(mpeg_b200.MpegB200Error) as e:
            ctx.video_validate(pp, mm, nb)
        assert e.value.code == -5

    bad(mv_h=-1)                      # reads left of the frame buffer: the Go code would panic
    bad(mb_row=g.mb_h)                # outside the picture
    bad(flags=0)                      # neither intra nor predicted
    bad(flags=wl.MB_INTRA | wl.MB_PREDICT)
    bad(cbp=1, n_blocks=0)            # coefficient block beyond the array
    bad(stream=7)                     # stream not open
    bad(dst_buf=3)
    bad(fwd_buf=0)                    # reference == destination
    two = np.concatenate([m, m])
    with pytest.raises(mpeg_b200.MpegB200Error):
        ctx.video_validate(p, two, 0)  # same macroblock twice in one picture
    with pytest.raises(mpeg_b200.MpegB200Error):
        ctx.video_validate(np.concatenate([p, p]), m, 0)  # two pictures of one stream
    with pytest.raises(mpeg_b200.MpegB200Error):
        ctx.video_open(6, 64, 48)      # already open
    with pytest.raises(mpeg_b200.MpegB200Error):
        ctx.video_read_frame(9, 0)     # not open


def test_rgba_matches_oracle(ctx):
    """Frame.RGBA(): GPU vs the CPU restatement of Go's image/draw arithmetic (parity unpinned
    against Go itself), including widths that are not multiples of 4 or 16."""
    for sid, (w, h) in enumerate([(160, 120), (352, 288), (50, 34), (61, 33)]):
        g = wl.Geometry(w, h)
        rng = wl.stream_rng(3, 100 + sid)
        fs = ol.FrameSet(1, w, h)
        fresh_stream(ctx, 20 + sid, w, h)
        buf = wl.random_reference_frame(rng, g)
        buf[:64] = [0, 255] * 32  # extremes
        fs.whole(0, 2)[:] = buf
        ctx.video_write_frame(20 + sid, 2, buf)
        got = ctx.video_rgba(20 + sid, 2, w, h)
        want = fs.rgba(0, 2)
        assert np.array_equal(got, want), (w, h)
        assert (got[..., 3] == 255).all()


def test_windows_wrap_across_row_ends_like_linear_indexing(ctx):
    """copyMacroblock indexes the plane linearly (si = y*stride + x, video_noasm.go:31): a vector that
    points left of column 0 reads the end of the row above, one that points past the last column reads
    the start of the next row.  The golden clip contains such vectors; pin the behaviour explicitly."""
    g = wl.Geometry(96, 64)
    rng = wl.stream_rng(2, 21)
    fs = ol.FrameSet(1, g.width, g.height)
    fresh_stream(ctx, 8, g.width, g.height)
    for b in range(3):
        buf = wl.random_reference_frame(rng, g)
        fs.whole(0, b)[:] = buf
        ctx.video_write_frame(8, b, buf)
    cases = [(1, 0, -3, 0), (1, 0, -31, 5), (2, 0, -1, -1), (1, g.mb_w - 1, 5, 0), (2, g.mb_w - 1, 31, 3),
             (g.mb_h - 1, g.mb_w - 1, 9, 2), (1, 0, -33, -7)]
    for row, col, mh, mv in cases:
        m = np.zeros(1, wl.MB_DTYPE)
        m["mb_row"], m["mb_col"], m["mv_h"], m["mv_v"], m["flags"] = row, col, mh, mv, wl.MB_PREDICT
        p = np.zeros(1, wl.PICTURE_DTYPE)
        p[0] = (8, wl.PIC_P, 0, 1, 2, 0, 1)
        ctx.video_validate(p, m, 0)
        ctx.video_decode_pictures(p, m, np.zeros((0, 64), np.int16))
        op = p.copy()
        op["stream"] = 0
        assert fs.exec_pictures(op, m, np.zeros((0, 64), np.int16)) == 0
        assert_frames_equal(ctx, fs, 8, 0, msg=f"wrap case {(row, col, mh, mv)}")


def test_mixed_geometries_slabs_and_kernel_selection(ctx):
    """Streams of different sizes live in different slabs (own tensor maps); a stream whose chroma pitch is not
    a multiple of 16 bytes (odd macroblock width) is served by the generic kernel, in the same batch as the streams
    the TMA kernel serves (each kernel skips the other's records).  Either way the output is the oracle's."""
    geos = [wl.Geometry(64, 48), wl.Geometry(96, 64), wl.Geometry(64, 48)]
    sids = [30, 31, 32]
    rng = wl.stream_rng(2, 77)
    oracles = []
    for sid, g in zip(sids, geos):
        fresh_stream(ctx, sid, g.width, g.height)
        fs = ol.FrameSet(1, g.width, g.height)
        for b in range(3):
            buf = wl.random_reference_frame(rng, g)
            fs.whole(0, b)[:] = buf
            ctx.video_write_frame(sid, b, buf)
        oracles.append(fs)

    def run_batch():
        per = [wl.make_picture(rng, g, wl.PIC_B, "natural") for g in geos]
        pics, mbs, coeffs = wl.batch_pictures(per, sids, wl.PIC_B, [(0, 1, 2)] * len(geos))
        ctx.video_validate(pics, mbs, len(coeffs))
        ctx.video_decode_pictures(pics, mbs, coeffs)
        for i, (fs, (m, c)) in enumerate(zip(oracles, per)):
            p1, m1, c1 = wl.batch_pictures([(m, c)], [0], wl.PIC_B, [(0, 1, 2)])
            assert fs.exec_pictures(p1, m1, c1) == 0
            assert_frames_equal(ctx, fs, sids[i], 0, msg="mixed geometries")

    run_batch()                                   # all streams TMA-capable: fast path, three slabs' worth of maps
    g_odd = wl.Geometry(72, 40)                   # mb_w = 5: chroma pitch 40 bytes
    fresh_stream(ctx, 33, g_odd.width, g_odd.height)
    fs_odd = ol.FrameSet(1, g_odd.width, g_odd.height)
    for b in range(3):
        buf = wl.random_reference_frame(rng, g_odd)
        fs_odd.whole(0, b)[:] = buf
        ctx.video_write_frame(33, b, buf)
    geos.append(g_odd)
    sids.append(33)
    oracles.append(fs_odd)
    run_batch()                                   # mixed batch: TMA kernel for three streams, generic kernel for the fourth
    ctx.video_close(33)
    geos.pop(); sids.pop(); oracles.pop()
    run_batch()                                   # back on the fast path
    # closing and re-opening a stream re-uses its slab slot and starts from zeroed buffers (video.go:340)
    ctx.video_close(31)
    ctx.video_open(31, 96, 64)
    assert not ctx.video_read_frame(31, 0).any()
    for s in (30, 31, 32):
        ctx.video_close(s)


def test_packed_12bit_transfer_form_is_equivalent(ctx):
    """mpegb200_video_decode_pictures_packed: the 96-byte transfer form expands on the device to the same int16 blocks."""
    import mpeg_b200
    g = wl.CIF
    rng = wl.stream_rng(2, 5)
    fs = ol.FrameSet(1, g.width, g.height)
    fresh_stream(ctx, 40, g.width, g.height)
    for b in range(3):
        buf = wl.random_reference_frame(rng, g)
        fs.whole(0, b)[:] = buf
        ctx.video_write_frame(40, b, buf)
    for t, bufs in [(wl.PIC_P, (0, 1, 2)), (wl.PIC_B, (2, 0, 1)), (wl.PIC_I, (1, 2, 0))]:
        mbs, coeffs = wl.make_picture(rng, g, t, "natural", adversarial=False)
        coeffs = np.clip(coeffs, -2048, 2047)          # intra dc*8 <= 2040 already; keep the extremes in
        coeffs[0, :4] = [2047, -2048, 1, -1]
        pics, mbs, coeffs = wl.batch_pictures([(mbs, coeffs)], [40], t, [bufs])
        packed = ctx.pack_coeffs12(coeffs)
        assert packed.shape == (len(coeffs), 96)
        ctx.video_decode_pictures_packed(pics, mbs, packed)
        op = pics.copy()
        op["stream"] = 0
        assert fs.exec_pictures(op, mbs, coeffs) == 0
        assert_frames_equal(ctx, fs, 40, 0, msg=f"packed path, picture type {t}")
    too_big = np.zeros((1, 64), np.int16)
    too_big[0, 0] = 2048
    with pytest.raises(mpeg_b200.MpegB200Error):
        ctx.pack_coeffs12(too_big)
    ctx.video_close(40)


def test_vlen_transfer_form_is_equivalent(ctx):
    """mpegb200_video_decode_pictures_vlen: the variable-width transfer form (eight groups per block in zig-zag order, each
    with its own bit width; coeff_vlen.cu) expands on the device to the same int16 blocks: natural P/B/I pictures (sparse
    blocks, even intra DCs -> raw groups), a dense-P picture, the extremes, and a block count that is no multiple of 32."""
    g = wl.CIF
    rng = wl.stream_rng(2, 6)
    fs = ol.FrameSet(1, g.width, g.height)
    fresh_stream(ctx, 41, g.width, g.height)
    for b in range(3):
        buf = wl.random_reference_frame(rng, g)
        fs.whole(0, b)[:] = buf
        ctx.video_write_frame(41, b, buf)
    steps = [(wl.PIC_P, (0, 1, 2), "natural"), (wl.PIC_B, (2, 0, 1), "natural"), (wl.PIC_I, (1, 2, 0), "natural"),
             (wl.PIC_P, (0, 1, 2), "dense")]
    for t, bufs, mode in steps:
        mbs, coeffs = wl.make_picture(rng, g, t, mode, adversarial=False)
        coeffs = np.clip(coeffs, -2048, 2047)
        coeffs[0, :4] = [2047, -2048, 1, -1]
        coeffs[1, :] = 0
        coeffs[2, :] = -2047
        pics, mbs, coeffs = wl.batch_pictures([(mbs, coeffs)], [41], t, [bufs])
        headers, chunks, payload = ctx.pack_coeffs_vlen(coeffs)
        assert len(headers) == len(coeffs) and len(chunks) == (len(coeffs) + 31) // 32
        ctx.video_decode_pictures_vlen(pics, mbs, headers, chunks, payload)
        op = pics.copy()
        op["stream"] = 0
        assert fs.exec_pictures(op, mbs, coeffs) == 0
        assert_frames_equal(ctx, fs, 41, 0, msg=f"vlen path, picture type {t}, {mode}")
    ctx.video_close(41)


def test_async_readback_is_ordered_against_later_decodes():
    """Host read-backs run on their own stream; a decode waits only for the read-backs of the buffers it writes
    (one event per physical buffer index).  Picture A goes into buffer 0 and is read back asynchronously; a decode
    into buffer 1 and then picture B into buffer 0 are enqueued right behind it.  The host copy must hold A, not B."""
    import mpeg_b200
    import torch
    g = wl.CIF
    n = 48
    c = mpeg_b200.Context(device=0, max_streams=n)
    try:
        rng = wl.stream_rng(2, 77)
        fs = ol.FrameSet(n, g.width, g.height)
        for s in range(n):
            c.video_open(s, g.width, g.height)
        ids = list(range(n))
        def intra_batch(dst):
            per = [wl.make_picture(rng, g, wl.PIC_I, "natural", adversarial=False) for _ in range(n)]
            return wl.batch_pictures(per, ids, wl.PIC_I, [(dst, (dst + 1) % 3, (dst + 2) % 3)] * n)
        a, mid, b = intra_batch(0), intra_batch(1), intra_batch(0)
        host = torch.empty(n * g.picture_bytes, dtype=torch.uint8, pin_memory=True)
        for rep in range(3):
            c.video_decode_pictures(*a)
            c.video_read_pictures(np.arange(n), np.zeros(n, np.uint8), host.data_ptr(), g.picture_bytes)
            c.video_decode_pictures(*mid)        # writes buffer 1: need not wait for the read-back
            c.video_decode_pictures(*b)          # overwrites buffer 0: must wait for it
            c.sync()
            assert fs.exec_pictures(a[0], a[1], a[2]) == 0
            got = host.numpy().reshape(n, g.picture_bytes)
            for s in range(n):
                assert np.array_equal(got[s], fs.whole(s, 0)[:g.picture_bytes]), f"rep {rep} stream {s}: read-back raced a later decode"
            assert fs.exec_pictures(mid[0], mid[1], mid[2]) == 0
            assert fs.exec_pictures(b[0], b[1], b[2]) == 0
            for s in (0, n - 1):
                assert np.array_equal(c.video_read_frame(s, 0), fs.whole(s, 0))
                assert np.array_equal(c.video_read_frame(s, 1), fs.whole(s, 1))
    finally:
        c.close()


def test_batched_readback_runs_and_singles():
    """mpegb200_video_read_pictures_host copies a run of equally spaced streams (consecutive slots of one slab, same
    buffer index) as ONE strided copy and everything else picture by picture.  Mixed order, two geometries, different
    buffer indices, a descending run: every picture must equal the stream's own read-back."""
    import mpeg_b200
    import torch
    ga, gb = wl.Geometry(96, 64), wl.Geometry(64, 48)
    c = mpeg_b200.Context(device=0, max_streams=8)
    try:
        rng = wl.stream_rng(2, 91)
        geo = {0: ga, 1: ga, 2: ga, 3: gb, 4: gb, 5: ga}
        for s, g in geo.items():
            c.video_open(s, g.width, g.height)
            for b in range(3):
                c.video_write_frame(s, b, wl.random_reference_frame(rng, g))
        stride = ga.picture_bytes + 64
        for streams, bufs in [([0, 1, 2, 5], [1, 1, 1, 1]),          # one run of three + (slot gap or not) the fourth
                              ([0, 1, 3, 2, 4], [0, 0, 2, 0, 2]),      # run of two, then singles of two slabs
                              ([2, 1, 0], [2, 2, 2]),                  # descending: no run
                              ([0, 1, 2], [0, 1, 0]),                  # same spacing, different buffers: no run
                              ([3, 4], [1, 1])]:
            host = torch.zeros(len(streams) * stride, dtype=torch.uint8, pin_memory=True)
            c.video_read_pictures(np.array(streams), np.array(bufs, np.uint8), host.data_ptr(), stride)
            c.sync()
            got = host.numpy().reshape(len(streams), stride)
            for i, (s, b) in enumerate(zip(streams, bufs)):
                n = geo[s].picture_bytes
                assert np.array_equal(got[i, :n], c.video_read_frame(s, b)[:n]), (streams, bufs, i)
                assert not got[i, n:].any()
    finally:
        c.close()


def test_kernel_timing_aid():
    """mpegb200_set_kernel_timing / mpegb200_kernel_times (bench.py's roofline of the arithmetic kernel alone): one
    (pre-pass, arithmetic kernel) pair of positive durations per decode call, oldest first, forgotten once read; the
    decoded pixels are unaffected."""
    import mpeg_b200
    g = wl.CIF
    rng = wl.stream_rng(2, 8)
    fs = ol.FrameSet(1, g.width, g.height)
    c = mpeg_b200.Context(device=0, max_streams=4)
    try:
        c.video_open(2, g.width, g.height)
        c.set_kernel_timing(True)
        for t, bufs in [(wl.PIC_I, (0, 1, 2)), (wl.PIC_P, (1, 0, 2)), (wl.PIC_P, (2, 1, 0))]:
            mbs, coeffs = wl.make_picture(rng, g, t, "natural", adversarial=False)
            pics, mbs, coeffs = wl.batch_pictures([(mbs, coeffs)], [2], t, [bufs])
            c.video_decode_pictures(pics, mbs, coeffs)
            op = pics.copy()
            op["stream"] = 0
            assert fs.exec_pictures(op, mbs, coeffs) == 0
        plan_ms, fused_ms = c.kernel_times()
        assert len(plan_ms) == len(fused_ms) == 3
        assert (plan_ms > 0).all() and (fused_ms > 0).all() and (fused_ms < 50).all()
        assert len(c.kernel_times()[0]) == 0          # read once
        c.set_kernel_timing(False)
        assert_frames_equal(c, fs, 2, 0, msg="decode with kernel timing on")
    finally:
        c.close()


def _run_pictures_and_compare(ctx, fs, sid, g, pictures, msg):
    """pictures: list of (type, (dst, fwd, bwd), mbs, coeffs); each decoded on the GPU and by the oracle, all three
    buffers compared after every one."""
    for i, (t, bufs, mbs, coeffs) in enumerate(pictures):
        pics, m, c = wl.batch_pictures([(mbs, coeffs)], [sid], t, [bufs])
        ctx.video_validate(pics, m, len(c))
        ctx.video_decode_pictures(pics, m, c)
        op = pics.copy()
        op["stream"] = 0
        assert fs.exec_pictures(op, m, c) == 0
        assert_frames_equal(ctx, fs, sid, 0, msg=f"{msg} picture {i}")


@pytest.mark.parametrize("width,height", [(256, 80), (512, 64), (1280, 48)])
def test_strip_staging_of_neighbouring_macroblocks(ctx, width, height):
    """Groups of 16 records whose windows fit one 304x48 / 160x24 rectangle in one reference buffer are staged as one
    luma and one chroma box (video_fused_tma.cu, strip mode); every other group with boxes per macroblock.  Widths
    that are multiples of 256 put whole groups on one macroblock row, so natural P pictures (vectors within +-16
    pixels, ~10 % intra macroblocks) run in strip mode, B pictures (two reference buffers inside a group) mix the
    two modes from group to group."""
    g = wl.Geometry(width, height)
    rng = wl.stream_rng(2, 31)
    fs = ol.FrameSet(1, g.width, g.height)
    fresh_stream(ctx, 9, g.width, g.height)
    for b in range(3):
        buf = wl.random_reference_frame(rng, g)
        fs.whole(0, b)[:] = buf
        ctx.video_write_frame(9, b, buf)
    rot = wl.BufferRotation()
    pictures = []
    for t in [wl.PIC_I, wl.PIC_P, wl.PIC_P, wl.PIC_B, wl.PIC_B, wl.PIC_P, wl.PIC_B]:
        bufs = rot.begin(t)
        mbs, coeffs = wl.make_picture(rng, g, t, "natural")
        pictures.append((t, bufs, mbs, coeffs))
        rot.end(t)
    _run_pictures_and_compare(ctx, fs, 9, g, pictures, f"natural {width}x{height}")
    ctx.video_close(9)


def test_strip_staging_edge_vectors(ctx):
    """The corner cases of the strip decision on a 256-wide picture (one group = one macroblock row):
    the same vector for a whole row pointing past the right edge (the strip uses the rows' 32-byte overlap), left of
    column 0 (x < 0: back to folded boxes), a row whose vertical spread exceeds the strip (boxes), a row with two
    predicted macroblocks only (below the strip threshold), dense residuals everywhere."""
    g = wl.Geometry(256, 96)
    rng = wl.stream_rng(2, 32)
    fs = ol.FrameSet(1, g.width, g.height)
    fresh_stream(ctx, 9, g.width, g.height)
    for b in range(3):
        buf = wl.random_reference_frame(rng, g)
        fs.whole(0, b)[:] = buf
        ctx.video_write_frame(9, b, buf)
    mbs, coeffs = wl.make_picture(rng, g, wl.PIC_P, "dense")
    mv_h = mbs["mv_h"].reshape(g.mb_h, g.mb_w)
    mv_v = mbs["mv_v"].reshape(g.mb_h, g.mb_w)
    flags = mbs["flags"].reshape(g.mb_h, g.mb_w)
    mv_h[1, :] = 31            # rightmost window ends at 256 + 15 + 17: inside the overlap
    mv_v[1, :] = 5
    mv_h[2, :] = -31           # leftmost window starts at -16: linear addressing reads the row above
    mv_v[2, :] = -3
    mv_v[3, ::2] = -32         # 16 rows up and 15 down in one group: 17 + 31 + 1 > 48
    mv_v[3, 1::2] = 31
    flags[4, :] = wl.MB_INTRA  # two predicted macroblocks in the group
    flags[4, 3] = flags[4, 12] = wl.MB_PREDICT
    mv_h[4, :] = np.where(flags[4] == wl.MB_PREDICT, mv_h[4], 0)
    mv_v[4, :] = np.where(flags[4] == wl.MB_PREDICT, mv_v[4], 0)
    mv_h[5, :] = np.arange(16) * 4 - 32   # a steady drift, half-pel phases of every kind
    mv_v[5, :] = 31 - np.arange(16) * 4
    mbs["mv_h"], mbs["mv_v"], mbs["flags"] = mv_h.ravel(), mv_v.ravel(), flags.ravel()
    intra_dc = np.repeat(mbs["flags"] == wl.MB_INTRA, 6)
    coeffs[intra_dc, 0] = (rng.integers(0, 256, int(intra_dc.sum())) * 8).astype(np.int16)
    _run_pictures_and_compare(ctx, fs, 9, g, [(wl.PIC_P, (0, 1, 2), mbs, coeffs)], "edge vectors")
    # the same records in reverse order: groups now run right to left, the rectangle is the same
    order = np.arange(len(mbs))[::-1]
    mbs_r = mbs[order].copy()
    coeffs_r = coeffs.reshape(len(mbs), 6, 64)[order].reshape(-1, 64).copy()
    mbs_r["coeff_block"] = np.arange(len(mbs)) * 6
    _run_pictures_and_compare(ctx, fs, 9, g, [(wl.PIC_P, (2, 1, 0), mbs_r, coeffs_r)], "reverse order")
    ctx.video_close(9)
