"""The reference's own API tests (mpeg_test.go) against the Python mirror of its public surface, with every
pixel and sample produced on the GPU: product host parser -> C-ABI -> sm_100a kernels.  Needs a B200."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

VIDEO_GOLDEN = 0xEA6D7FCB1340BA3F          # mpeg_test.go:227
AUDIO_GOLDEN_NOFMA = 0xF1B76CDF8E6CDEA5    # mpeg_test.go:194


@pytest.fixture(scope="module")
def ctx():
    import mpeg_b200
    c = mpeg_b200.Context(device=0, max_streams=8)
    yield c
    c.close()


def test_video_golden(ctx, golden_dir):
    # TestVideoGolden, mpeg_test.go:203-231
    import mpeg_b200
    video = mpeg_b200.Video((golden_dir / "test.mpeg1video").read_bytes(), ctx, stream=0)
    h, frames = ol.FNV_OFFSET, 0
    while True:
        frame = video.decode()
        if frame is None:
            break
        h = ol.fnv(h, frame.y)
        h = ol.fnv(h, frame.cb)
        h = ol.fnv(h, frame.cr)
        frames += 1
    assert h == VIDEO_GOLDEN, f"video output hash: got {h:#018x} (frames={frames})"
    video.close()


def test_video_header_first_frame_and_rgba(ctx, golden_dir):
    # TestVideo, mpeg_test.go:233-274 (+ Frame.RGBA against the CPU restatement)
    import mpeg_b200
    data = (golden_dir / "test.mpeg1video").read_bytes()
    video = mpeg_b200.Video(data, ctx, stream=1)
    assert video.has_header()
    assert (video.width, video.height, video.framerate) == (160, 120, 30.0)
    frame = video.decode()
    assert frame is not None and frame.width == video.width
    assert frame.y.size == 20480 and frame.cb.size == frame.y.size // 4
    o = ol.VideoOracle(data)
    f = o.decode()
    assert np.array_equal(frame.y, f.plane("y")) and np.array_equal(frame.cr, f.plane("cr"))
    fs = ol.FrameSet(1, 160, 120)
    fs.whole(0, 0)[:] = f.whole()
    rgba = frame.rgba()
    assert rgba.shape == (120, 160, 4) and np.array_equal(rgba, fs.rgba(0, 0))
    video.close()


def test_audio_golden(ctx, golden_dir):
    # TestAudioGolden, mpeg_test.go:164-201: hash of Float32bits(Interleaved)
    import mpeg_b200
    audio = mpeg_b200.Audio((golden_dir / "test.mp2").read_bytes(), ctx, stream=0)
    assert audio.has_header() and audio.samplerate == 44100 and audio.channels == 1   # TestAudio, :135-162
    h, frames, t_prev = ol.FNV_OFFSET, 0, -1.0
    while True:
        s = audio.decode()
        if s is None:
            break
        assert s.time > t_prev
        t_prev = s.time
        h = ol.fnv(h, s.interleaved)
        frames += 1
    assert h == AUDIO_GOLDEN_NOFMA, f"audio output hash: got {h:#018x} (frames={frames})"
    audio.close()


def test_mpeg_program_stream(ctx, golden_dir):
    # TestMpeg / TestDemux, mpeg_test.go:41-82, 276-398 (the parts that touch the decode path)
    import mpeg_b200
    data = (golden_dir / "test.mpg").read_bytes()
    with pytest.raises(mpeg_b200.ErrInvalidMPEG):
        mpeg_b200.MPEG(b"not an mpeg stream", ctx)
    m = mpeg_b200.MPEG(data, ctx, video_stream=2, audio_stream=2)
    assert (m.num_video_packets, m.num_audio_packets) == (143, 37)
    assert (m.video.width, m.video.height) == (160, 120)
    es_video, es_audio, _, _ = ol.demux_split(data)
    ov, oa = ol.VideoOracle(es_video), ol.AudioOracle(es_audio)
    n = 0
    while True:
        f, of = m.decode_video(), ov.decode()
        assert (f is None) == (of is None)
        if f is None:
            break
        assert f.time == of.time
        if n % 7 == 0:
            assert np.array_equal(f.y, of.plane("y")) and np.array_equal(f.cb, of.plane("cb"))
        n += 1
    assert n > 200
    k = 0
    while True:
        s, os_ = m.decode_audio(), oa.decode()
        assert (s is None) == (os_ is None)
        if s is None:
            break
        assert np.array_equal(s.interleaved.view(np.uint32), os_.view(np.uint32))
        k += 1
    assert k > 30
    m.close()


def test_video_batch_lockstep_golden(golden_dir):
    # many streams decoded in lock-step (one launch per wave for all of them): every stream must reproduce
    # TestVideoGolden's hash (mpeg_test.go:203-231) / the oracle's hash for its own bitstream
    import mpeg_b200
    from test_batch_parser import cut_at_picture
    es = (golden_dir / "test.mpeg1video").read_bytes()
    ps_video = ol.demux_split((golden_dir / "test.mpg").read_bytes())[0]
    kinds = [es, ps_video, cut_at_picture(es, 7), cut_at_picture(ps_video, 33)]
    want = []
    for k, d in enumerate(kinds):
        o, h = ol.VideoOracle(d), ol.FNV_OFFSET
        while (f := o.decode()) is not None:
            for which in ("y", "cb", "cr"):
                h = ol.fnv(h, f.plane(which))
        want.append(h)
    n = 24
    datas = [kinds[i % len(kinds)] for i in range(n)]
    with mpeg_b200.Context(device=0, max_streams=32) as c:
        batch = mpeg_b200.VideoBatch(c, datas, threads=4, first_stream=3)
        geo = c.video_geometry(3)
        pic_bytes = geo[0] * geo[1] + 2 * geo[2] * geo[3]
        hashes, frames = [ol.FNV_OFFSET] * n, [0] * n
        host = np.empty((n, pic_bytes), np.uint8)
        while True:
            has, buf, t = batch.step()
            if not has.any():
                break
            live = np.nonzero(has)[0]
            c.video_read_pictures(live + 3, buf[live], host.ctypes.data, pic_bytes)
            c.sync()
            for k, i in enumerate(live):
                hashes[i] = ol.fnv(hashes[i], host[k])
                frames[i] += 1
        batch.close()
    assert frames[0] > frames[2] > 0
    for i in range(n):
        assert hashes[i] == want[i % len(kinds)], f"stream {i}: {hashes[i]:#018x}"
    # the oracle's hash of the elementary stream is the reference's golden value
    assert want[0] == VIDEO_GOLDEN


@pytest.mark.parametrize("device_vlc", [False, True])
def test_program_stream_batch_with_display_ring_and_audio_batch(golden_dir, device_vlc):
    """the batched front end.  Program streams are demultiplexed, their video decodes in lock-step with the
    returned frames kept in a device display ring (read back LATE, two steps behind the decoder), their audio goes through the
    lock-step audio batch (eight frames per stream and launch, streams of different lengths).  Every stream must reproduce
    the oracle's hashes for its own bitstream."""
    import mpeg_b200
    full = (golden_dir / "test.mpg").read_bytes()
    es_v, es_a, _, _ = ol.demux_split(full)

    def oracle_hashes(video, audio):
        o, hv = ol.VideoOracle(video), ol.FNV_OFFSET
        while (f := o.decode()) is not None:
            for which in ("y", "cb", "cr"):
                hv = ol.fnv(hv, f.plane(which))
        a, ha, k = ol.AudioOracle(audio), ol.FNV_OFFSET, 0
        while (s := a.decode()) is not None:
            ha = ol.fnv(ha, s)
            k += 1
        return hv, ha, k

    # program streams cut at different packet boundaries: different lengths in both elementary streams
    cuts = [len(full), full.rfind(b"\x00\x00\x01\xe0", 0, 300000), full.rfind(b"\x00\x00\x01\xc0", 0, 200000)]
    datas = [full[:c] for c in cuts]
    want = []
    for d in datas:
        v, a, _, _ = ol.demux_split(d)
        want.append(oracle_hashes(v, a))
    n = 9
    streams = [datas[i % 3] for i in range(n)]
    with mpeg_b200.Context(device=0, max_streams=16) as c:
        mb = mpeg_b200.MPEGBatch(c, streams, threads=4, ring_depth=4, frames_per_step=8, device_vlc=device_vlc)   # True: video slices parsed on the device
        assert mb.packets[0] == (143, 37)
        geo = c.video_geometry(0)
        pic_bytes = geo[0] * geo[1] + 2 * geo[2] * geo[3]
        hv = [ol.FNV_OFFSET] * n
        pending = []                       # (slot, has) of the steps not yet read back
        while True:
            has, t, slot = mb.decode_video()
            if has.any():
                pending.append((slot, has.copy()))
            while pending and (len(pending) > 2 or not has.any()):   # consume two steps behind the decoder
                s, h = pending.pop(0)
                pics = mb.ring.read(s, pic_bytes)
                for i in np.nonzero(h)[0]:
                    hv[i] = ol.fnv(hv[i], pics[i])
            if not has.any():
                break
        ha, frames = [ol.FNV_OFFSET] * n, [0] * n
        while True:
            nf, t, out = mb.decode_audio()
            if not nf.any():
                break
            for i in range(n):
                for k in range(int(nf[i])):
                    ha[i] = ol.fnv(ha[i], out[i][k])
                frames[i] += int(nf[i])
        mb.close()
    for i in range(n):
        assert hv[i] == want[i % 3][0], f"video of stream {i}"
        assert ha[i] == want[i % 3][1] and frames[i] == want[i % 3][2], f"audio of stream {i}: {frames[i]} frames"
    assert want[0][2] > want[2][2] > 8    # different lengths: the shorter streams end inside a step (tail launches)
