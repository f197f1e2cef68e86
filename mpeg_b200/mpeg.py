"""Python mirror of the reference's public surface over the B200 path.

    mpeg.New(io.Reader) -> *MPEG            mpeg.go:85         -> MPEG(data)
    (*Video).Decode() -> *Frame             video.go:209       -> Video.decode() -> Frame | None
    (*Audio).Decode() -> *Samples           audio.go:163       -> Audio.decode() -> Samples | None
    (*Frame).RGBA() / YCbCr planes          video.go:26-43     -> Frame.rgba(), Frame.y / .cb / .cr

The serial half (parse, dequantise) runs in the host library (include/mpegb200_host.h); every pixel and
sample is produced by the CUDA kernels through the C-ABI.  Decode calls return None at the end of the
stream, like the reference returns nil; constructors raise on invalid input, like it returns errors.
"""
import ctypes as C

import numpy as np

from . import _lib
from .context import MB_DTYPE, Context


class Launch(C.Structure):
    _fields_ = [("stream", C.c_int32), ("type", C.c_uint8), ("dst_buf", C.c_uint8), ("fwd_buf", C.c_uint8),
                ("bwd_buf", C.c_uint8), ("pic_first_mb", C.c_uint32), ("pic_n_mb", C.c_uint32),
                ("first_mb", C.c_uint32), ("n_mb", C.c_uint32), ("first_block", C.c_uint32), ("n_blocks", C.c_uint32)]


class LaunchVlen(C.Structure):
    _fields_ = [("first_chunk", C.c_uint32), ("reserved", C.c_uint32), ("payload_offset", C.c_uint64), ("payload_bytes", C.c_uint64)]


class VideoStep(C.Structure):
    _fields_ = [("has_frame", C.c_int), ("frame_buf", C.c_int), ("time", C.c_double), ("n_launches", C.c_int),
                ("launches", C.POINTER(Launch)), ("mbs", C.c_void_p), ("coeffs", C.c_void_p),
                ("vlen_launches", C.POINTER(LaunchVlen)), ("vlen_headers", C.c_void_p), ("vlen_chunk_offsets", C.c_void_p),
                ("vlen_payload", C.c_void_p)]


assert C.sizeof(Launch) == 32 and C.sizeof(LaunchVlen) == 24


class ErrInvalidMPEG(ValueError):
    """mpeg.ErrInvalidMPEG (mpeg.go:55) / demux.ErrInvalidHeader (demux.go:32)."""


class Frame:
    """Decoded frame, valid until the next decode() of its Video (mpeg.go:413-415).  Planes are the
    macroblock-padded planes of the reference's Plane.Data (video.go:45-54), fetched on first use."""

    def __init__(self, video, buf, time):
        self._v, self._buf, self.time = video, buf, time
        self.width, self.height = video.width, video.height
        self._planes = None

    def _fetch(self):
        if self._planes is None:
            self._planes = self._v.ctx.video_read_planes(self._v.stream, self._buf)
        return self._planes

    @property
    def y(self):
        return self._fetch()[0]

    @property
    def cb(self):
        return self._fetch()[1]

    @property
    def cr(self):
        return self._fetch()[2]

    def rgba(self) -> np.ndarray:
        """Frame.RGBA(), video.go:31-36: H x W x 4 uint8."""
        return self._v.ctx.video_rgba(self._v.stream, self._buf, self.width, self.height)


class Video:
    """mpeg.Video (video.go:57): MPEG-1 video elementary stream -> frames."""

    def __init__(self, data: bytes, ctx: Context, stream: int = 0, vlen: bool = True, device_vlc: bool = False):
        """vlen: the parser emits the coefficients in the variable-width transfer form (a header and a few bytes per block
        instead of int16[64]) and the launches go through mpegb200_video_decode_pictures_vlen; False: int16 blocks.
        device_vlc: the stream is uploaded to device memory, its start codes are indexed there and the slices are parsed on the
        GPU (mpegb200_video_decode_bitstream); the host reads headers only (a one-stream VideoBatch does the work)."""
        self.L = _lib.load()
        self.ctx, self.stream = ctx, stream
        self._data = bytes(data)
        self._batch = None
        self._device_vlc = device_vlc
        self.h = self.L.mpegb200_video_parser_new(self._data, len(self._data))
        if not self.h:
            raise MemoryError
        self.L.mpegb200_video_parser_set_vlen(self.h, int(vlen))
        self._opened = False

    def has_header(self) -> bool:
        return bool(self.L.mpegb200_video_parser_has_header(self.h))

    @property
    def width(self):
        return self.L.mpegb200_video_parser_width(self.h)

    @property
    def height(self):
        return self.L.mpegb200_video_parser_height(self.h)

    @property
    def framerate(self):
        return self.L.mpegb200_video_parser_framerate(self.h)

    def _parsers(self):
        """The parser that answers header queries, and (device_vlc) the one inside the one-stream batch that walks the stream."""
        if self._device_vlc and self._batch is None:
            from .batch import VideoBatch
            self._batch = VideoBatch(self.ctx, [self._data], threads=1, first_stream=self.stream, device_vlc=True, resident=True)
        return [self.h] + ([self._batch.parser(0)] if self._batch is not None else [])

    def set_no_delay(self, on: bool):
        for h in self._parsers():
            self.L.mpegb200_video_parser_set_no_delay(h, int(on))

    def rewind(self):
        if self._batch is not None:
            self._batch.drop_scan_ahead()
        for h in self._parsers():
            self.L.mpegb200_video_parser_rewind(h)

    def has_ended(self) -> bool:
        return bool(self.L.mpegb200_video_parser_has_ended(self._parsers()[-1]))

    def decode(self):
        """Video.Decode(): parse up to the picture that makes a frame due, run its launches, return the Frame."""
        if not self.has_header():
            return None
        if self._device_vlc:
            self._parsers()
            has, buf, t = self._batch.step()
            if not has[0]:
                return None
            self.ctx.sync()
            return Frame(self, int(buf[0]), float(t[0]))
        if not self._opened:
            self.ctx.video_open(self.stream, self.width, self.height)
            self.ctx.set_validate(True)   # records parsed from a bitstream are untrusted: malformed ones raise (the reference panics)
            self._opened = True
        step = VideoStep()
        rc = self.L.mpegb200_video_parser_next(self.h, C.byref(step))
        if rc != 0 or not step.has_frame:
            return None
        for i in range(step.n_launches):
            ln = step.launches[i]
            if ln.n_mb == 0:
                continue
            ln.stream = self.stream
            mb_ptr = step.mbs + 16 * ln.first_mb
            if step.vlen_launches:
                lv = step.vlen_launches[i]
                self.ctx._ck(self.L.mpegb200_video_decode_pictures_vlen(
                    self.ctx.h, 1, C.byref(ln), ln.n_mb, C.c_void_p(mb_ptr), ln.n_blocks,
                    C.c_void_p((step.vlen_headers or 0) + 4 * ln.first_block), C.c_void_p((step.vlen_chunk_offsets or 0) + 8 * lv.first_chunk),
                    C.c_void_p((step.vlen_payload or 0) + lv.payload_offset), lv.payload_bytes))
                continue
            co_ptr = step.coeffs + 128 * ln.first_block
            self.ctx._ck(self.L.mpegb200_video_decode_pictures(self.ctx.h, 1, C.byref(ln), ln.n_mb, C.c_void_p(mb_ptr),
                                                               ln.n_blocks, C.c_void_p(co_ptr)))
        # the parser re-uses its arrays on the next call: the copies to the device must have been issued from them
        self.ctx.sync()
        return Frame(self, step.frame_buf, step.time)

    def close(self):
        if getattr(self, "_batch", None) is not None:
            self._batch.close()
            self._batch = None
        if getattr(self, "h", None):
            if self._opened:
                try:
                    self.ctx.video_close(self.stream)
                except Exception:
                    pass
            self.L.mpegb200_video_parser_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Samples:
    """mpeg.Samples (audio.go:27-36) for the default AudioF32N format."""

    def __init__(self, interleaved, time):
        self.time = time
        self.interleaved = interleaved          # 2304 float32, L R L R ...
        self.left = interleaved[0::2]
        self.right = interleaved[1::2]

    def bytes(self) -> bytes:
        return self.interleaved.tobytes()


class Audio:
    """mpeg.Audio (audio.go:53): MP2 elementary stream -> 1152-sample frames."""

    def __init__(self, data: bytes, ctx: Context, stream: int = 0, fmt: int = 0, coded: bool = True):
        """coded: the host parser stops before the requantisation (audio.go:476-489) and the device does it
        (mpegb200_audio_synth_coded: half the bytes over PCIe); False: requantised int32 samples from the host."""
        self.L = _lib.load()
        self.ctx, self.stream, self.fmt, self.coded = ctx, stream, fmt, coded
        self._info, self._codes = np.zeros(256, np.uint8), np.zeros((2, 36, 32), np.uint16)
        self._data = bytes(data)
        self.h = self.L.mpegb200_audio_parser_new(self._data, len(self._data))
        if not self.h:
            raise MemoryError
        self.ctx.audio_open(stream)
        self._samples = np.zeros((2, 36, 32), np.int32)

    def has_header(self) -> bool:
        return bool(self.L.mpegb200_audio_parser_has_header(self.h))

    @property
    def samplerate(self):
        return self.L.mpegb200_audio_parser_samplerate(self.h)

    @property
    def channels(self):
        return self.L.mpegb200_audio_parser_channels(self.h)

    def rewind(self):
        self.L.mpegb200_audio_parser_rewind(self.h)

    def decode(self):
        t = C.c_double()
        if self.coded:
            if not self.L.mpegb200_audio_parser_next_coded(self.h, C.c_void_p(self._info.ctypes.data), C.c_void_p(self._codes.ctypes.data), C.byref(t)):
                return None
            out = self.ctx.audio_synth_coded([self.stream], 1, self._info, self._codes, self.fmt)
        else:
            if not self.L.mpegb200_audio_parser_next(self.h, C.c_void_p(self._samples.ctypes.data), C.byref(t)):
                return None
            out = self.ctx.audio_synth([self.stream], 1, self._samples, self.fmt)
        return Samples(out[0, 0], t.value) if self.fmt != 3 else out[0, 0]

    def close(self):
        if getattr(self, "h", None):
            try:
                self.ctx.audio_close(self.stream)
            except Exception:
                pass
            self.L.mpegb200_audio_parser_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def demux_split(data: bytes):
    """Video (0xE0) and first audio (0xC0) elementary streams of a program stream (demux.go)."""
    L = _lib.load()
    v, a = C.c_void_p(), C.c_void_p()
    vl, al = C.c_size_t(), C.c_size_t()
    nv, na = C.c_int(), C.c_int()
    rc = L.mpegb200_demux_split(data, len(data), C.byref(v), C.byref(vl), C.byref(a), C.byref(al), C.byref(nv), C.byref(na))
    if rc != 0:
        raise ErrInvalidMPEG("invalid MPEG-PS header")
    video, audio = C.string_at(v, vl.value), C.string_at(a, al.value)
    L.mpegb200_buffer_free(v)
    L.mpegb200_buffer_free(a)
    return video, audio, nv.value, na.value


class MPEG:
    """mpeg.MPEG (mpeg.go:58): a program stream with one video and one audio stream."""

    def __init__(self, data: bytes, ctx: Context, video_stream: int = 0, audio_stream: int = 0):
        if len(data) < 4 or data[:4] != b"\x00\x00\x01\xba":   # mpeg.go:95-100
            raise ErrInvalidMPEG("invalid MPEG")
        video, audio, self.num_video_packets, self.num_audio_packets = demux_split(data)
        self.video = Video(video, ctx, video_stream) if video else None
        self.audio = Audio(audio, ctx, audio_stream) if audio else None

    def decode_video(self):          # MPEG.DecodeVideo, mpeg.go:416
        return self.video.decode() if self.video else None

    def decode_audio(self):          # MPEG.DecodeAudio, mpeg.go:438
        return self.audio.decode() if self.audio else None

    def close(self):
        for s in (self.video, self.audio):
            if s:
                s.close()
