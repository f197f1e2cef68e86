"""The slice-parallel VLC stage on the GPU: compressed slices go to the device as they are, vlc_parse_kernel writes
the records, the decode kernels execute them.  Checked against the CPU run of the same walker (tests/vlc_emu), against the host
parser path, and against the reference's golden hash.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest

import mpeg1_writer as mw
import oracle_lib as ol
import vlc_emu_lib as ve
from test_mpeg1_writer import write_stream

pytestmark = pytest.mark.gpu

VIDEO_GOLDEN = 0xEA6D7FCB1340BA3F          # mpeg_test.go:227


def oracle_hash(data):
    o, h, n = ol.VideoOracle(data), ol.FNV_OFFSET, 0
    while (f := o.decode()) is not None:
        for which in ("y", "cb", "cr"):
            h = ol.fnv(h, f.plane(which))
        n += 1
    return h, n


def batch_hashes(c, datas, first=0, **kw):
    import mpeg_b200
    n = len(datas)
    batch = mpeg_b200.VideoBatch(c, datas, threads=4, first_stream=first, **kw)
    hashes, frames = [ol.FNV_OFFSET] * n, [0] * n
    hosts = {}
    while True:
        has, buf, t = batch.step()
        if not has.any():
            break
        for i in np.nonzero(has)[0]:
            geo = c.video_geometry(first + int(i))
            pic_bytes = geo[0] * geo[1] + 2 * geo[2] * geo[3]
            host = hosts.setdefault(pic_bytes, np.empty((1, pic_bytes), np.uint8))
            c.video_read_pictures(np.array([first + int(i)]), buf[i:i + 1], host.ctypes.data, pic_bytes)
            c.sync()
            hashes[i] = ol.fnv(hashes[i], host[0])
            frames[i] += 1
    stats = (batch.flagged, batch.host_steps)
    batch.close()
    return hashes, frames, stats


def test_device_vlc_reproduces_the_golden_hash(golden_dir):
    """TestVideoGolden (mpeg_test.go:203-231) with the slices parsed on the GPU.  testdata/test.mpeg1video is damaged: about a
    quarter of its pictures flag and finish on the host parser -- the hash must come out all the same; the video of test.mpg is
    clean and must not flag at all."""
    import mpeg_b200
    from test_batch_parser import cut_at_picture
    es = (golden_dir / "test.mpeg1video").read_bytes()
    ps_video = ol.demux_split((golden_dir / "test.mpg").read_bytes())[0]
    with mpeg_b200.Context(device=0, max_streams=16) as c:
        hashes, frames, (flagged, _) = batch_hashes(c, [es], device_vlc=True)
        assert hashes[0] == VIDEO_GOLDEN, f"{hashes[0]:#018x} after {frames[0]} frames"
        assert 0 < flagged < frames[0] // 2
    with mpeg_b200.Context(device=0, max_streams=16) as c:
        kinds = [ps_video, cut_at_picture(ps_video, 33), es, cut_at_picture(es, 7)]
        want = [oracle_hash(d) for d in kinds]
        datas = [kinds[i % 4] for i in range(12)]
        hashes, frames, _ = batch_hashes(c, datas, first=2, device_vlc=True)
        for i in range(12):
            assert (hashes[i], frames[i]) == want[i % 4], f"stream {i}"
    with mpeg_b200.Context(device=0, max_streams=4) as c:
        hashes, frames, (flagged, host_steps) = batch_hashes(c, [ps_video], device_vlc=True)
        assert (hashes[0], frames[0]) == oracle_hash(ps_video) and flagged == 0 and host_steps == 0


@pytest.mark.parametrize("size,pictures,mode", [
    ((352, 288), [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_B, mw.PIC_P], "natural"),
    ((1280, 720), [mw.PIC_I, mw.PIC_P, mw.PIC_B], "natural"),
    ((1280, 720), [mw.PIC_I, mw.PIC_P], "dense"),
    ((720, 576), [mw.PIC_I, mw.PIC_P, mw.PIC_B], "natural"),      # odd macroblock width: the generic decode kernel executes the records
])
def test_device_records_equal_the_cpu_run_of_the_walker(size, pictures, mode):
    """vlc_parse_kernel against the same walker compiled for the CPU (which the no-GPU suite holds against the host parser):
    every record slot, every coefficient block the records name, every flag; and the frames the decode kernels make of the
    device records against the host-parser path."""
    import mpeg_b200
    w, _ = write_stream(size[0], size[1], pictures, seed=size[0] + len(pictures), mode=mode)
    data = w.tobytes()
    L = mpeg_b200._lib.load()
    tab = ve.tables()
    with mpeg_b200.Context(device=0, max_streams=4) as c:
        c.video_open(0, *size)
        sb = ve.ScanBatch([data])
        ref = mpeg_b200.Video(data, c, stream=1)            # the host-parser path decodes the same stream next to it
        mb_w, mb_h = sb.sizes[0]
        n = 0
        try:
            while True:
                st = sb.next()
                if not st.has_frame[0]:
                    break
                for wv in range(st.n_waves):
                    wave = st.waves[wv]
                    c._ck(L.mpegb200_video_decode_bitstream(c.h, wave.n_pictures, wave.pics, wave.n_slices, wave.slices, C.c_void_p(wave.bitstream),
                                                            wave.bitstream_bytes, C.c_void_p(wave.quant), wave.n_quant, wave.n_mb_slots))
                    flags = np.zeros(wave.n_pictures, np.int32)
                    assert L.mpegb200_video_bitstream_flags(c.h, C.c_void_p(flags.ctypes.data), wave.n_pictures) == 0
                    gm = np.zeros(wave.n_mb_slots, ol.MB_DTYPE)
                    gc = np.zeros((6 * wave.n_mb_slots, 64), np.int16)
                    c._ck(L.mpegb200_video_bitstream_records(c.h, C.c_void_p(gm.ctypes.data), C.c_void_p(gc.ctypes.data)))
                    em, ec, ef = ve.emulate_wave(wave, mb_w, mb_h, tab)
                    assert not ef.any()
                    assert np.array_equal(gm, em), "record slots differ from the CPU run of the walker"
                    for m in em[em["pic"] != 0xffff]:
                        b0, nc = int(m["coeff_block"]), bin(int(m["cbp"])).count("1")
                        assert np.array_equal(gc[b0:b0 + nc], ec[b0:b0 + nc])
                    n += 1
                f = ref.decode()
                assert f is not None and f._buf == st.frame_buf[0]
                for b in range(3):
                    assert np.array_equal(c.video_read_frame(0, b), c.video_read_frame(1, b)), f"frame buffer {b} after step {n}"
        finally:
            sb.close()
            ref.close()
        assert n == len(pictures)


def test_device_vlc_batch_720p_equals_host_parser_batch():
    """48 streams of 720p (three distinct ones, I P B P) through both front ends of the lock-step batch: the frames every step
    returns must be identical, and the device path must not have flagged anything."""
    import mpeg_b200
    distinct = []
    for d in range(3):
        w, _ = write_stream(1280, 720, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_P], seed=40 + d, mode="natural")
        distinct.append(w.tobytes())
    datas = [distinct[i % 3] for i in range(48)]
    with mpeg_b200.Context(device=0, max_streams=48) as c:
        h_host, f_host, _ = batch_hashes(c, datas)
    with mpeg_b200.Context(device=0, max_streams=48) as c:
        h_dev, f_dev, (flagged, host_steps) = batch_hashes(c, datas, device_vlc=True)
    assert f_host == f_dev and f_host[0] == 4
    assert h_host == h_dev
    assert flagged == 0 and host_steps == 0
    want = [oracle_hash(d)[0] for d in distinct]
    assert all(h_dev[i] == want[i % 3] for i in range(48))


def test_void_and_foreign_tables_are_refused_or_skipped(golden_dir):
    """The slot tables are checked on the host before anything is enqueued; a picture withdrawn by the caller (type 0) leaves
    null records and its flag, and touches no frame buffer."""
    import mpeg_b200
    w, _ = write_stream(64, 48, [mw.PIC_I, mw.PIC_P], seed=3, mode="natural")
    L = mpeg_b200._lib.load()
    with mpeg_b200.Context(device=0, max_streams=2) as c:
        c.video_open(0, 64, 48)
        sb = ve.ScanBatch([w.tobytes()])
        try:
            st = sb.next()
            wave = st.waves[0]
            args = lambda: (c.h, wave.n_pictures, wave.pics, wave.n_slices, wave.slices, C.c_void_p(wave.bitstream), wave.bitstream_bytes,
                            C.c_void_p(wave.quant), wave.n_quant, wave.n_mb_slots)
            wave.slices[1].mb_slot += 16
            assert L.mpegb200_video_decode_bitstream(*args()) == -1      # MPEGB200_EINVAL: a hole in the record slots
            wave.slices[1].mb_slot -= 16
            before = [c.video_read_frame(0, b) for b in range(3)]
            wave.pics[0].type = 0
            c._ck(L.mpegb200_video_decode_bitstream(*args()))
            flags = np.zeros(1, np.int32)
            assert L.mpegb200_video_bitstream_flags(c.h, C.c_void_p(flags.ctypes.data), 1) == 1 and flags[0] == 0x80
            gm = np.zeros(wave.n_mb_slots, ol.MB_DTYPE)
            c._ck(L.mpegb200_video_bitstream_records(c.h, C.c_void_p(gm.ctypes.data), None))
            assert (gm["pic"] == 0xffff).all()
            for b in range(3):
                assert np.array_equal(c.video_read_frame(0, b), before[b])
        finally:
            sb.close()


def test_resident_streams_and_the_device_start_code_index(golden_dir):
    """Streams uploaded to HBM once, start codes indexed by startcode_index_kernel: the index must equal a plain search of the
    bytes, and the batch that runs from it (waves without bytes) must return the same frames as every other path -- the damaged
    reference clip with its flagged pictures included."""
    import mpeg_b200
    from test_batch_parser import cut_at_picture
    es = (golden_dir / "test.mpeg1video").read_bytes()
    ps_video = ol.demux_split((golden_dir / "test.mpg").read_bytes())[0]
    L = mpeg_b200._lib.load()
    with mpeg_b200.Context(device=0, max_streams=16) as c:
        for sid, d in ((5, es), (6, ps_video), (7, b"\x00\x00\x01" * 40 + b"\x00"), (8, b"\x00\x00\x01\xb3"), (9, b"\x00\x00\x01\xb3\x00")):
            c._ck(L.mpegb200_video_stream_upload(c.h, sid, d, len(d)))
            want = ve.start_code_positions(d)
            want = want[want + 5 <= len(d)]          # buffer.go:284: a start code needs its code byte and one more
            got = np.zeros(len(want) + 8, np.uint64)
            n = C.c_size_t()
            c._ck(L.mpegb200_video_stream_index(c.h, sid, C.c_void_p(got.ctypes.data), len(got), C.byref(n)))
            assert n.value == len(want) and np.array_equal(got[:n.value], want), f"stream {sid}"
            if len(want) > 2:                        # too little room: the count still comes back
                assert L.mpegb200_video_stream_index(c.h, sid, C.c_void_p(got.ctypes.data), 2, C.byref(n)) == -1 and n.value == len(want)
            c._ck(L.mpegb200_video_stream_upload(c.h, sid, None, 0))
    with mpeg_b200.Context(device=0, max_streams=16) as c:
        kinds = [ps_video, cut_at_picture(ps_video, 33), es, cut_at_picture(es, 7)]
        want = [oracle_hash(d) for d in kinds]
        datas = [kinds[i % 4] for i in range(12)]
        hashes, frames, (flagged, _) = batch_hashes(c, datas, first=2, device_vlc=True, resident=True)
        for i in range(12):
            assert (hashes[i], frames[i]) == want[i % 4], f"stream {i}"
        assert flagged > 0 and want[2][0] == VIDEO_GOLDEN
    distinct = []
    for d in range(2):
        w, _ = write_stream(1280, 720, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_P], seed=60 + d, mode="natural")
        distinct.append(w.tobytes())
    datas = [distinct[i % 2] for i in range(32)]
    with mpeg_b200.Context(device=0, max_streams=32) as c:
        h_dev, f_dev, (flagged, host_steps) = batch_hashes(c, datas, device_vlc=True, resident=True)
    want = [oracle_hash(d)[0] for d in distinct]
    assert all(h_dev[i] == want[i % 2] for i in range(32)) and f_dev[0] == 4 and flagged == 0 and host_steps == 0


def test_single_stream_video_with_device_vlc(golden_dir):
    """mpeg_b200.Video(device_vlc=True): TestVideoGolden (mpeg_test.go:203-231) frame by frame through the public mirror."""
    import mpeg_b200
    with mpeg_b200.Context(device=0, max_streams=4) as c:
        video = mpeg_b200.Video((golden_dir / "test.mpeg1video").read_bytes(), c, stream=1, device_vlc=True)
        assert (video.width, video.height, video.framerate) == (160, 120, 30.0)
        h, frames, last = ol.FNV_OFFSET, 0, -1.0
        while (frame := video.decode()) is not None:
            for plane in (frame.y, frame.cb, frame.cr):
                h = ol.fnv(h, plane)
            assert frame.time > last
            last = frame.time
            frames += 1
        assert h == VIDEO_GOLDEN, f"{h:#018x} after {frames} frames"
        video.close()


def test_damaged_streams_through_the_device_parser():
    """Forty mutations of a written stream (four flipped bits each; two of them leave stale coefficients behind a dropped block, which
    makes the host parse the following step as a whole) in one lock-step batch, wave form and resident form: every stream's frames
    must equal the oracle's decode of the same damaged bytes -- whatever mix of device-parsed, flagged and host-parsed pictures it
    takes.  (Mutations whose vectors leave the frame buffer are left out: the reference panics there, this path raises.)"""
    import mpeg_b200
    w, _ = write_stream(176, 144, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_P, mw.PIC_P], seed=5, mode="natural")
    data = w.tobytes()
    datas, want = [], []
    for trial in list(range(100, 140)) + [368]:
        rng = np.random.default_rng(1000 + trial)
        d = bytearray(data)
        for pos in rng.integers(20, len(d), 4):
            d[pos] ^= 1 << int(rng.integers(0, 8))
        o = ol.VideoOracle(bytes(d))
        h, n = ol.FNV_OFFSET, 0
        while (f := o.decode()) is not None:
            for which in ("y", "cb", "cr"):
                h = ol.fnv(h, f.plane(which))
            n += 1
        if o.oob_count() == 0:
            datas.append(bytes(d))
            want.append((h, n))
    assert len(datas) >= 36
    for kw in ({"device_vlc": True}, {"device_vlc": True, "resident": True}, {"device_vlc": True, "scan_ahead": False},
               {"device_vlc": True, "native_step": False}, {"device_vlc": True, "resident": True, "native_step": False, "scan_ahead": False}):
        with mpeg_b200.Context(device=0, max_streams=64) as c:
            hashes, frames, (flagged, host_steps) = batch_hashes(c, datas, **kw)
        assert list(zip(hashes, frames)) == want, kw
        assert flagged > 5 and host_steps >= 2, (flagged, host_steps)


def test_rewind_and_no_delay_in_device_mode(golden_dir):
    """Video.Rewind (video.go:195-201: the frame buffers are NOT cleared) and SetNoDelay reach the parser that walks the stream
    when the slices are parsed on the device; a scan made ahead of time is withdrawn first."""
    import mpeg_b200
    data = ol.demux_split((golden_dir / "test.mpg").read_bytes())[0]
    for no_delay in (False, True):
        with mpeg_b200.Context(device=0, max_streams=2) as c:
            video, oracle = mpeg_b200.Video(data, c, stream=0, device_vlc=True), ol.VideoOracle(data)
            video.set_no_delay(no_delay)
            oracle.set_no_delay(no_delay)
            for rounds in (7, 25):
                for k in range(rounds):
                    f, want = video.decode(), oracle.decode()
                    assert (f is None) == (want is None) and f.time == want.time
                    assert np.array_equal(f.y.reshape(-1), want.plane("y").reshape(-1)) and np.array_equal(f.cr.reshape(-1), want.plane("cr").reshape(-1))
                assert not video.has_ended()
                video.rewind()
                oracle.rewind()
            video.close()


def test_mixed_geometries_in_one_wave(golden_dir):
    """Streams of different picture sizes (one of them with an odd macroblock width, which the generic decode kernel executes) and
    different lengths in one lock-step batch with the slices parsed on the device: per-picture geometry, slot tables and quantiser
    matrices in one wave."""
    import mpeg_b200
    ps_video = ol.demux_split((golden_dir / "test.mpg").read_bytes())[0]
    w1, _ = write_stream(352, 288, [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_B, mw.PIC_P], seed=71, mode="natural")
    w2, _ = write_stream(720, 576, [mw.PIC_I, mw.PIC_P, mw.PIC_B], seed=72, mode="natural")
    w3, _ = write_stream(64, 48, [mw.PIC_I] + [mw.PIC_P] * 9, seed=73, mode="dense")
    kinds = [ps_video[:60000], w1.tobytes(), w2.tobytes(), w3.tobytes()]
    want = [oracle_hash(d) for d in kinds]
    datas = [kinds[i % 4] for i in range(8)]
    for kw in ({"device_vlc": True}, {"device_vlc": True, "resident": True}):
        with mpeg_b200.Context(device=0, max_streams=16) as c:
            hashes, frames, _ = batch_hashes(c, datas, first=1, **kw)
        for i in range(8):
            assert (hashes[i], frames[i]) == want[i % 4], (i, kw)
