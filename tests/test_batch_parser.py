"""mpegb200_video_batch_* (lock-step parse of many streams on a host thread pool, merged into waves)
against the single-stream parser: every wave must hold exactly the w-th launch of each stream's step.  CPU only."""
import ctypes as C

import numpy as np

import oracle_lib as ol
from mpeg_b200 import _lib
from mpeg_b200.batch import BatchStep
from test_host_parser import parser_steps, video_streams

PIC_DTYPE = np.dtype([("stream", "<i4"), ("type", "u1"), ("dst_buf", "u1"), ("fwd_buf", "u1"), ("bwd_buf", "u1"),
                      ("first_mb", "<u4"), ("n_mb", "<u4")])


def cut_at_picture(data: bytes, n_pictures: int) -> bytes:
    """The stream up to (not including) its (n_pictures+1)-th picture start code."""
    pos, seen = 0, 0
    while True:
        pos = data.find(b"\x00\x00\x01\x00", pos)
        if pos < 0:
            return data
        if seen == n_pictures:
            return data[:pos]
        seen += 1
        pos += 4


def test_pic_dtype_matches_abi():
    assert PIC_DTYPE.itemsize == 16 == ol.PIC_DTYPE.itemsize


def test_batch_waves_equal_per_stream_launches(golden_dir):
    src = video_streams(golden_dir)
    es, ps = src["test.mpeg1video"], src["test.mpg video"]
    datas = [es, ps, cut_at_picture(es, 7), es, cut_at_picture(ps, 20)]
    L = _lib.load()
    b = L.mpegb200_video_batch_new(len(datas), 3, None, None)
    assert b
    try:
        w, h = C.c_int(), C.c_int()
        for i, d in enumerate(datas):
            assert L.mpegb200_video_batch_set_stream(b, i, d, len(d)) == 0
            assert L.mpegb200_video_batch_stream_size(b, i, C.byref(w), C.byref(h)) == 0
            assert (w.value, h.value) == (160, 120)
        singles = [parser_steps(d) for d in datas]
        alive = [True] * len(datas)
        n_steps = n_waves = 0
        while any(alive):
            st = BatchStep()
            assert L.mpegb200_video_batch_next(b, C.byref(st)) == 0
            assert st.n_streams == len(datas)
            want = []
            for i, it in enumerate(singles):
                step = next(it, None) if alive[i] else None
                if step is None:
                    alive[i] = False
                    assert st.has_frame[i] == 0
                else:
                    assert st.has_frame[i] == 1 and st.frame_buf[i] == step[0] and st.time[i] == step[1]
                want.append(step)
            depth = max([len(s[2]) for s in want if s is not None] or [0])
            assert st.n_waves == depth
            for wv in range(depth):
                wave = st.waves[wv]
                # a launch without macroblocks (the clip has such a picture) is no work for the GPU and is left out
                members = [i for i, s in enumerate(want) if s is not None and len(s[2]) > wv and s[2][wv][0][4] > 0]
                assert wave.n_pictures == len(members)
                pics = np.frombuffer(C.string_at(wave.pics, 16 * wave.n_pictures), dtype=PIC_DTYPE)
                mbs = np.frombuffer(C.string_at(wave.mbs, 16 * wave.n_mb), dtype=ol.MB_DTYPE) if wave.n_mb else np.zeros(0, ol.MB_DTYPE)
                co = np.frombuffer(C.string_at(wave.coeffs, 128 * wave.n_blocks), dtype=np.int16).reshape(-1, 64) if wave.n_blocks else np.zeros((0, 64), np.int16)
                assert list(pics["stream"]) == members
                assert int(pics["n_mb"].sum()) == wave.n_mb
                for k, i in enumerate(members):
                    hdr, wm, wc = want[i][2][wv]
                    p = pics[k]
                    assert (p["type"], p["dst_buf"], p["fwd_buf"], p["bwd_buf"], p["n_mb"]) == hdr
                    gm = mbs[p["first_mb"]:p["first_mb"] + p["n_mb"]].copy()
                    assert np.all(gm["pic"] == k)
                    nb = int(sum(bin(int(c)).count("1") for c in gm["cbp"]))
                    first = int(gm["coeff_block"][0]) if len(gm) else 0
                    gm["pic"] = 0
                    gm["coeff_block"] -= first
                    assert np.array_equal(gm, wm), f"step {n_steps} wave {wv} stream {i}: macroblock records"
                    assert np.array_equal(co[first:first + nb], wc), f"step {n_steps} wave {wv} stream {i}: coefficients"
                n_waves += 1
            n_steps += 1
        assert n_steps > 10 and n_waves >= n_steps - 1
        # after the end every further step is empty
        st = BatchStep()
        assert L.mpegb200_video_batch_next(b, C.byref(st)) == 0
        assert st.n_waves == 0 and not any(st.has_frame[i] for i in range(len(datas)))
    finally:
        L.mpegb200_video_batch_free(b)


def test_batch_waves_in_vlen_form_match_the_converter(golden_dir):
    """Lock-step batch with the parsers emitting the variable-width form: every wave's headers / chunk offsets / payload equal
    mpegb200_pack_coeffs_vlen of the int16 wave the same batch produces without the switch (streams of different lengths and
    phases, so waves merge launches whose blocks do not start on chunk boundaries)."""
    import ctypes as C
    from mpeg_b200 import _lib
    from mpeg_b200.batch import BatchStep
    from test_vlen_format import pack
    L = _lib.load()
    es = (golden_dir / "test.mpeg1video").read_bytes()
    streams = [es, es[:50000], es[:20000], es]
    hs = []
    for vlen in (0, 1):
        h = L.mpegb200_video_batch_new(len(streams), 3, None, None)
        assert L.mpegb200_video_batch_set_vlen(h, vlen) == 0
        for i, d in enumerate(streams):
            assert L.mpegb200_video_batch_set_stream(h, i, d, len(d)) == 0
        hs.append(h)
    waves = blocks = 0
    for step in range(60):
        a, b = BatchStep(), BatchStep()
        assert L.mpegb200_video_batch_next(hs[0], C.byref(a)) == 0 and L.mpegb200_video_batch_next(hs[1], C.byref(b)) == 0
        assert a.n_waves == b.n_waves
        for w in range(a.n_waves):
            wa, wb = a.waves[w], b.waves[w]
            assert (wa.n_pictures, wa.n_mb, wa.n_blocks) == (wb.n_pictures, wb.n_mb, wb.n_blocks)
            assert C.string_at(wa.mbs, 16 * wa.n_mb) == C.string_at(wb.mbs, 16 * wb.n_mb)
            if wa.n_blocks == 0:
                continue
            assert wa.coeffs and not wb.coeffs and wb.vlen_headers
            co = np.frombuffer(C.string_at(wa.coeffs, 128 * wa.n_blocks), dtype=np.int16).reshape(-1, 64)
            rc, hd, ch, pl = pack(co)
            assert rc == 0 and wb.vlen_payload_bytes == len(pl)
            assert C.string_at(wb.vlen_headers, 4 * wb.n_blocks) == hd.tobytes()
            assert C.string_at(wb.vlen_chunk_offsets, 8 * len(ch)) == ch.tobytes()
            assert C.string_at(wb.vlen_payload, len(pl)) == pl.tobytes()
            waves += 1
            blocks += wa.n_blocks
    assert waves > 50 and blocks > 10000
    for h in hs:
        L.mpegb200_video_batch_free(h)


def test_audio_batch_matches_the_single_stream_parser(golden_dir):
    """mpegb200_audio_batch_*: streams of different lengths, five frames per step; the rectangular part and the tails together
    hold exactly the frames the single-stream parser yields, in order."""
    import ctypes as C
    from mpeg_b200 import _lib
    from mpeg_b200.batch import AudioBatchStep
    L = _lib.load()
    mp2 = (golden_dir / "test.mp2").read_bytes()
    datas = [mp2, mp2[: len(mp2) // 2], mp2[: len(mp2) // 3], mp2, b"junk" * 100]
    want = []
    for d in datas:
        h = L.mpegb200_audio_parser_new(d, len(d))
        frames, s, t = [], np.zeros((2, 36, 32), np.int32), C.c_double()
        while L.mpegb200_audio_parser_next(h, C.c_void_p(s.ctypes.data), C.byref(t)):
            frames.append((s.copy(), t.value))
        L.mpegb200_audio_parser_free(h)
        want.append(frames)
    assert len(want[0]) > len(want[1]) > len(want[2]) > 5 and len(want[4]) == 0
    b = L.mpegb200_audio_batch_new(len(datas), 3, None, None)
    for i, d in enumerate(datas):
        assert L.mpegb200_audio_batch_set_stream(b, i, d, len(d)) == 0
    F, got = 5, [[] for _ in datas]
    saw_tail = False
    while True:
        st = AudioBatchStep()
        assert L.mpegb200_audio_batch_next(b, F, C.byref(st)) == 0
        if st.n_full == 0 and st.n_tail == 0:
            break
        full = np.frombuffer(C.string_at(st.full_samples, st.n_full * F * 2304 * 4), np.int32).reshape(st.n_full, F, 2, 36, 32) if st.n_full else None
        for j in range(st.n_full):
            i = st.full_index[j]
            assert st.n_frames[i] == F
            assert st.time[i] == want[i][len(got[i])][1]
            got[i].extend(full[j])
        at = 0
        for j in range(st.n_tail):
            i, k = st.tail_index[j], st.tail_frames[j]
            assert 0 < k < F and st.n_frames[i] == k
            tail = np.frombuffer(C.string_at(st.tail_samples + at * 2304 * 4, k * 2304 * 4), np.int32).reshape(k, 2, 36, 32)
            got[i].extend(tail)
            at += k
            saw_tail = True
    assert saw_tail
    for i in range(len(datas)):
        assert len(got[i]) == len(want[i])
        for a, (w, _) in zip(got[i], want[i]):
            assert np.array_equal(a, w)
    L.mpegb200_audio_batch_free(b)
