"""The slice-parallel VLC stage without a GPU: the device-side slice walker, compiled for the CPU
(tests/vlc_emu), against the product's host parser -- which itself equals the oracle's restatement of the reference
parser record for record (tests/test_host_parser.py).

Contract under test (include/mpegb200.h): for every picture the walker either FLAGS it (then the host parser decodes
it) or leaves exactly the records and coefficient blocks the host parser produces for that picture; streams written
by a conforming encoder never flag; the scan-mode parser walks the stream like the full parser does."""
import ctypes as C

import numpy as np
import pytest

import mpeg1_writer as mw
import oracle_lib as ol
import vlc_emu_lib as ve
from test_host_parser import parser_steps, video_streams
from test_mpeg1_writer import write_stream


def same_records(a, b):
    (am, ac), (bm, bc) = a, b
    if len(am) != len(bm) or ac.shape != bc.shape:
        return False
    for f in ("mb_row", "mb_col", "mv_h", "mv_v", "flags", "cbp", "coeff_block"):
        if not np.array_equal(am[f], bm[f]):
            return False
    return np.array_equal(ac, bc)


def step_through_the_walker(sb, st, i, tab, stats, resident=False):
    """Stream i of scan step `st` the way the product drives it (mpeg_b200.VideoBatch): device pictures wave by wave; at the first
    flagged picture the host parses the rest of the step.  Returns (has_frame, frame_buf, time, launches)."""
    hosted = ve.ScanBatch.host_steps(st)
    if i in hosted:
        stats["host_steps"] += 1
        return st.has_frame[i], st.frame_buf[i], st.time[i], hosted[i]
    mb_w, mb_h = sb.sizes[i]
    launches = []
    for w in range(st.n_waves):
        wave = st.waves[w]
        ks = [k for k in range(wave.n_pictures) if wave.pics[k].stream == i]
        if not ks:
            break
        k = ks[0]
        assert wave.step_picture[k] == w
        members = [wave.pics[j].stream for j in range(wave.n_pictures)]
        mbs, coeffs, flags = ve.emulate_wave(wave, [sb.sizes[j][0] for j in members], [sb.sizes[j][1] for j in members], tab,
                                             resident_bytes=sb.datas[i] if resident else None)
        P = wave.pics[k]
        stats["pictures"] += 1
        if flags[k]:
            stats["flagged"] += 1
            has, buf, t, tail = sb.redo(i, w)
            return has, buf, t, launches + tail
        gm, gc = ve.picture_records(mbs, coeffs, P)
        launches.append(((P.type, P.dst_buf, P.fwd_buf, P.bwd_buf, len(gm)), gm, gc))
    return st.has_frame[i], st.frame_buf[i], st.time[i], launches


def run_stream(data, expect_clean, label="", resident=False):
    """Walk `data` in scan mode next to the full host parser: every step must come out the same.  Returns the statistics.
    resident: the batch works like with streams resident in device memory (start codes handed in, waves without bytes)."""
    full = parser_steps(data)
    sb = ve.ScanBatch([data], resident=resident)
    tab = ve.tables()
    stats = {"pictures": 0, "flagged": 0, "host_steps": 0, "steps": 0}
    try:
        while True:
            st = sb.next()
            want = next(full, None)
            has, buf, t, launches = step_through_the_walker(sb, st, 0, tab, stats, resident)
            if not has:
                assert want is None, f"{label}: the scan parser ends before the full parser (step {stats['steps']})"
                break
            assert want is not None, f"{label}: the scan parser goes on behind the full parser's end"
            assert (buf, t) == (want[0], want[1]), f"{label}: step {stats['steps']} returns another frame"
            # a launch without records is no work; the full parser may split a picture with rewrites into several launches, the
            # device path never carries such a picture (it is flagged), so launch lists compare one to one
            got = [l for l in launches if l[0][4]]
            ref = [l for l in want[2] if l[0][4]]
            assert len(got) == len(ref), f"{label}: step {stats['steps']}: {len(got)} launches, the host parser has {len(ref)}"
            for (h1, m1, c1), (h2, m2, c2) in zip(got, ref):
                assert h1 == h2 and same_records((m1, c1), (m2, c2)), f"{label}: step {stats['steps']} differs from the host parser"
            stats["steps"] += 1
    finally:
        sb.close()
    if expect_clean:
        assert stats["flagged"] == 0 and stats["host_steps"] == 0, f"{label}: {stats} in a conforming stream"
    return stats


@pytest.mark.parametrize("size,pictures,mode", [
    ((64, 48), [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_B, mw.PIC_P], "natural"),
    ((96, 64), [mw.PIC_I, mw.PIC_P], "dense"),
    ((352, 288), [mw.PIC_I, mw.PIC_P, mw.PIC_B], "natural"),
    ((1280, 720), [mw.PIC_I, mw.PIC_P], "natural"),
])
def test_walker_equals_host_parser_on_written_streams(size, pictures, mode):
    w, _ = write_stream(size[0], size[1], pictures, seed=size[0] + len(pictures), mode=mode)
    stats = run_stream(w.tobytes(), expect_clean=True, label=f"{size} {mode}")
    assert stats["pictures"] == len(pictures)


def test_walker_wide_vectors_and_quantiser_scales():
    for f_code, mv_range, scale in ((4, 128, 3), (1, 16, 31), (3, 64, 1)):
        w, _ = write_stream(160, 128, [mw.PIC_I, mw.PIC_P, mw.PIC_B], seed=f_code, mode="natural", mv_range=mv_range, f_code=f_code, scale=scale)
        run_stream(w.tobytes(), expect_clean=True, label=f"f_code {f_code}")


@pytest.mark.parametrize("which", ["test.mpeg1video", "test.mpg video"])
def test_walker_on_the_reference_clips(golden_dir, which):
    """The video of test.mpg is a clean stream: no picture may flag.  testdata/test.mpeg1video is damaged (slices that overlap,
    run into the next start code, leave the picture: DESIGN section 2): those pictures must flag and the host finishes their
    steps, all others must equal the host parser -- every step compares equal either way."""
    stats = run_stream(video_streams(golden_dir)[which], expect_clean=which == "test.mpg video", label=which)
    assert stats["steps"] > 250
    if which == "test.mpeg1video":
        assert 0 < stats["flagged"] < stats["pictures"] // 2, stats   # about a quarter of its pictures


def test_walker_on_damaged_streams(golden_dir):
    """Flipped bits and truncation: every picture is either flagged or identical to the host parser's; nothing crashes."""
    data = (golden_dir / "test.mpeg1video").read_bytes()[:80000]
    rng = np.random.default_rng(11)
    total = flagged = 0
    for trial in range(12):
        d = bytearray(data)
        for pos in rng.integers(200, len(d), 30):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        d = bytes(d[: len(d) - int(rng.integers(0, 3000))])
        stats = run_stream(d, expect_clean=False, label=f"damaged {trial}")
        total += stats["pictures"]
        flagged += stats["flagged"]
    assert total > 100 and 0 < flagged < total


def test_walker_on_damaged_written_streams():
    w, _ = write_stream(176, 144, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_P], seed=5, mode="natural")
    data = w.tobytes()
    rng = np.random.default_rng(3)
    for trial in range(20):
        d = bytearray(data)
        for pos in rng.integers(20, len(d), 6):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        run_stream(bytes(d), expect_clean=False, label=f"damaged written {trial}")


def test_batch_scan_waves_hold_each_streams_picture(golden_dir):
    """Several streams of different lengths and sizes in one scan batch: wave w holds the w-th picture of every stream that
    has one, slots laid out back to back, and every stream's steps equal its own full parse."""
    es = video_streams(golden_dir)["test.mpeg1video"]
    w1, _ = write_stream(64, 48, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_P], seed=1, mode="natural")
    w2, _ = write_stream(352, 288, [mw.PIC_I, mw.PIC_P], seed=2, mode="natural")
    datas = [es[:40000], w1.tobytes(), w2.tobytes()]
    fulls = [parser_steps(d) for d in datas]
    sb = ve.ScanBatch(datas, threads=3)
    tab = ve.tables()
    stats = {"pictures": 0, "flagged": 0, "host_steps": 0, "steps": 0}
    alive = [True] * len(datas)
    try:
        while any(alive):
            st = sb.next()
            for w in range(st.n_waves):
                wave = st.waves[w]
                members = [wave.pics[k].stream for k in range(wave.n_pictures)]
                assert members == sorted(members) and len(set(members)) == len(members)
                slot = 0
                for k in range(wave.n_pictures):
                    P = wave.pics[k]
                    assert P.mb_slot == slot and P.quant == k
                    slot += P.n_mb_slots
                assert slot == wave.n_mb_slots
            for i in range(len(datas)):
                if not alive[i]:
                    assert not st.has_frame[i]
                    continue
                want = next(fulls[i], None)
                has, buf, t, launches = step_through_the_walker(sb, st, i, tab, stats)
                if not has:
                    assert want is None
                    alive[i] = False
                    continue
                assert want is not None and (buf, t) == (want[0], want[1])
                got = [l for l in launches if l[0][4]]
                ref = [l for l in want[2] if l[0][4]]
                assert len(got) == len(ref)
                for (h1, m1, c1), (h2, m2, c2) in zip(got, ref):
                    assert h1 == h2 and same_records((m1, c1), (m2, c2))
            stats["steps"] += 1
        assert stats["steps"] >= 4 and stats["pictures"] > 20
    finally:
        sb.close()


def run_stream_scanning_ahead(data, label=""):
    """The product's pipelined flow (mpeg_b200.VideoBatch, scan_ahead): step k + 1 is scanned BEFORE step k's flags are known; a
    flag withdraws that scan (mpegb200_video_batch_unscan), the host finishes step k, and step k + 1 is scanned again.  Every step
    must still equal the full host parser's, and the wave arrays of step k must survive the scan of step k + 1."""
    full = parser_steps(data)
    sb = ve.ScanBatch([data])
    tab = ve.tables()
    mb_w, mb_h = sb.sizes[0]
    ahead, steps, withdrawn = None, 0, 0
    try:
        while True:
            st = ahead if ahead is not None else sb.next()
            ahead = None
            has, buf, t = st.has_frame[0], st.frame_buf[0], st.time[0]
            hosted = ve.ScanBatch.host_steps(st)
            launches = hosted.get(0, [])
            for w in range(st.n_waves if 0 not in hosted else 0):
                wave = st.waves[w]
                if w == st.n_waves - 1:
                    ahead = sb.next()                      # the guess, made while "the device" still holds wave w
                mbs, coeffs, flags = ve.emulate_wave(wave, mb_w, mb_h, tab)
                P = wave.pics[0]
                if flags[0]:
                    if ahead is not None:
                        sb.unscan()
                        ahead = None
                        withdrawn += 1
                    has, buf, t, tail = sb.redo(0, w)
                    launches = launches + tail
                    break
                gm, gc = ve.picture_records(mbs, coeffs, P)
                launches.append(((P.type, P.dst_buf, P.fwd_buf, P.bwd_buf, len(gm)), gm, gc))
            want = next(full, None)
            if not has:
                assert want is None, f"{label}: ends early at step {steps}"
                break
            assert want is not None and (buf, t) == (want[0], want[1]), f"{label}: step {steps}"
            got = [l for l in launches if l[0][4]]
            ref = [l for l in want[2] if l[0][4]]
            assert len(got) == len(ref), f"{label}: step {steps}"
            for (h1, m1, c1), (h2, m2, c2) in zip(got, ref):
                assert h1 == h2 and same_records((m1, c1), (m2, c2)), f"{label}: step {steps} differs from the host parser"
            steps += 1
    finally:
        sb.close()
    return steps, withdrawn


def test_scanning_ahead_and_withdrawing(golden_dir):
    steps, withdrawn = run_stream_scanning_ahead(video_streams(golden_dir)["test.mpeg1video"], "test.mpeg1video")
    assert steps > 250 and withdrawn > 20
    steps, withdrawn = run_stream_scanning_ahead(video_streams(golden_dir)["test.mpg video"], "test.mpg video")
    assert steps > 250 and withdrawn == 0
    data = (golden_dir / "test.mpeg1video").read_bytes()[:80000]
    rng = np.random.default_rng(21)
    for trial in range(6):
        d = bytearray(data)
        for pos in rng.integers(200, len(d), 30):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        run_stream_scanning_ahead(bytes(d), f"damaged {trial}")


def test_resident_mode_with_handed_in_start_codes(golden_dir):
    """Streams resident in device memory: the parser takes the start codes from an index (found on the device in the product,
    by bytes.find here) instead of searching, the waves carry no bytes and the slices' offsets count from the first byte of
    their stream.  Every step must still equal the full host parser's -- clean, damaged and truncated streams."""
    for which in ("test.mpg video", "test.mpeg1video"):
        a = run_stream(video_streams(golden_dir)[which], expect_clean=which == "test.mpg video", label=which)
        b = run_stream(video_streams(golden_dir)[which], expect_clean=which == "test.mpg video", label=which + " resident", resident=True)
        assert a == b
    w, _ = write_stream(352, 288, [mw.PIC_I, mw.PIC_P, mw.PIC_B], seed=8, mode="natural")
    run_stream(w.tobytes(), expect_clean=True, label="written resident", resident=True)
    data = (golden_dir / "test.mpeg1video").read_bytes()[:60000]
    rng = np.random.default_rng(31)
    for trial in range(6):
        d = bytearray(data)
        for pos in rng.integers(200, len(d), 30):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        d = bytes(d[: len(d) - int(rng.integers(0, 3000))])
        assert run_stream(d, False, f"damaged {trial}") == run_stream(d, False, f"damaged {trial} resident", resident=True)
    # an index that does not belong to the stream is refused
    from mpeg_b200 import _lib
    L = _lib.load()
    h = L.mpegb200_video_parser_new(data, len(data))
    bad = np.array([5, 17], np.uint64)
    assert L.mpegb200_video_parser_set_start_codes(h, C.c_void_p(bad.ctypes.data), 2) == -1
    L.mpegb200_video_parser_free(h)


def test_stale_coefficients_make_the_host_parse_the_next_step():
    """An invalid run in the last block a picture decodes leaves coefficients behind that the serial reference leaks into the next
    block it decodes -- a picture later (video.go:712-714, 774-777).  Only the host parser carries that state: the flagged picture's
    step ends on the host, and the NEXT step is parsed by the host as a whole (mpegb200_video_scan_step.host_step) until the stale
    coefficients are gone.  Two mutations of a written stream that do exactly that (found by search, seeds fixed)."""
    w, _ = write_stream(176, 144, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_P, mw.PIC_P], seed=5, mode="natural")
    data = w.tobytes()
    for trial in (106, 368):
        rng = np.random.default_rng(1000 + trial)
        d = bytearray(data)
        for pos in rng.integers(20, len(d), 4):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        stats = run_stream(bytes(d), expect_clean=False, label=f"stale {trial}")
        assert stats["host_steps"] >= 1 and stats["flagged"] >= 1, stats
        assert run_stream(bytes(d), expect_clean=False, label=f"stale {trial} resident", resident=True) == stats
        run_stream_scanning_ahead(bytes(d), f"stale {trial} ahead")
