"""ctypes binding of the CPU oracle (oracle/, TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (mpeg_b200) never does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "_build" / "liboracle_mpeg.so"
GOLDEN = ROOT / "tests" / "golden"

FNV_OFFSET = 0xCBF29CE484222325


def build_oracle(force: bool = False) -> Path:
    """(Re)build oracle/_build/liboracle_mpeg.so with gcc when missing or stale."""
    srcs = [p for p in ORACLE_DIR.iterdir() if p.suffix in (".c", ".h", ".inc")] + [ROOT / "include" / "mpegb200.h"]
    stale = (not ORACLE_SO.exists()) or any(p.stat().st_mtime > ORACLE_SO.stat().st_mtime for p in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B"], check=True, env=env, capture_output=True)
    return ORACLE_SO


class Frame(C.Structure):
    _fields_ = [
        ("width", C.c_int), ("height", C.c_int),
        ("luma_w", C.c_int), ("luma_h", C.c_int), ("chroma_w", C.c_int), ("chroma_h", C.c_int),
        ("buf_bytes", C.c_size_t),
        ("base", C.c_void_p), ("y", C.c_void_p), ("cb", C.c_void_p), ("cr", C.c_void_p),
        ("time", C.c_double),
    ]

    def plane(self, which: str) -> np.ndarray:
        ptr = {"y": self.y, "cb": self.cb, "cr": self.cr}[which]
        w, h = (self.luma_w, self.luma_h) if which == "y" else (self.chroma_w, self.chroma_h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h, w))

    def whole(self) -> np.ndarray:
        return np.ctypeslib.as_array(C.cast(self.base, C.POINTER(C.c_uint8)), shape=(self.buf_bytes,))


# packed record dtypes (include/mpegb200.h)
MB_DTYPE = np.dtype([
    ("mb_row", "<u2"), ("mb_col", "<u2"), ("mv_h", "<i2"), ("mv_v", "<i2"),
    ("flags", "u1"), ("cbp", "u1"), ("pic", "<u2"), ("coeff_block", "<u4"),
])
PIC_DTYPE = np.dtype([
    ("stream", "<i4"), ("type", "u1"), ("dst_buf", "u1"), ("fwd_buf", "u1"), ("bwd_buf", "u1"),
    ("first_mb", "<u4"), ("n_mb", "<u4"),
])
assert MB_DTYPE.itemsize == 16 and PIC_DTYPE.itemsize == 16

MB_INTRA, MB_PREDICT, MB_REF_BWD = 1, 2, 4
PIC_I, PIC_P, PIC_B = 1, 2, 3


class SynthState(C.Structure):
    _fields_ = [("v", C.c_float * 2048), ("v_pos", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build_oracle()
    L = C.CDLL(str(ORACLE_SO))
    vp, u8p, i64p, f32p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.orc_fnv1a64.restype = C.c_uint64
    L.orc_fnv1a64.argtypes = [C.c_uint64, vp, C.c_size_t]
    L.orc_idct.argtypes = [i64p, C.c_int]
    L.orc_idct_full.argtypes = [i64p]
    L.orc_frame_init.argtypes = [C.POINTER(Frame), C.c_int, C.c_int]
    L.orc_frame_free.argtypes = [C.POINTER(Frame)]
    L.orc_copy_macroblock.argtypes = [C.c_int] * 4 + [C.POINTER(Frame), C.POINTER(Frame)]
    L.orc_copy_macroblock_swar.argtypes = [C.c_int] * 4 + [C.POINTER(Frame), C.POINTER(Frame)]
    L.orc_rgba.argtypes = [C.POINTER(Frame), vp]
    L.orc_rgba_batch.argtypes = [C.POINTER(Frame), C.c_int, vp, vp, vp, C.c_size_t, C.c_int]
    L.orc_use_swar_mc.argtypes = [C.c_int]
    L.orc_exec_pictures.argtypes = [C.POINTER(Frame), C.c_int, vp, C.c_size_t, vp, vp, C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_video_open.restype = vp
    L.orc_video_open.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_video_close.argtypes = [vp]
    L.orc_video_decode.restype = C.POINTER(Frame)
    L.orc_video_decode.argtypes = [vp]
    for name in ("orc_video_has_header", "orc_video_width", "orc_video_height", "orc_video_last_buf", "orc_video_oob_count"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.orc_video_framerate.argtypes = [vp]
    L.orc_video_framerate.restype = C.c_double
    L.orc_video_set_no_delay.argtypes = [vp, C.c_int]
    L.orc_video_rewind.argtypes = [vp]
    L.orc_video_tap_enable.argtypes = [vp, C.c_int]
    L.orc_video_tap_pictures.argtypes = [vp, C.POINTER(vp)]
    L.orc_video_tap_pictures.restype = C.c_int
    L.orc_video_tap_mbs.argtypes = [vp, C.POINTER(vp)]
    L.orc_video_tap_mbs.restype = C.c_size_t
    L.orc_video_tap_blocks.argtypes = [vp, C.POINTER(vp)]
    L.orc_video_tap_blocks.restype = C.c_size_t
    L.orc_idct36.argtypes = [i64p, C.c_int, f32p, C.c_int]
    L.orc_synth_window.argtypes = [f32p, f32p, f32p, C.c_int]
    L.orc_synth_window_fma.argtypes = [f32p, f32p, f32p, C.c_int]
    L.orc_synthesis_window_1024.restype = f32p
    L.orc_synth_frame.argtypes = [C.POINTER(SynthState), vp, C.c_int, vp, C.c_int]
    L.orc_synth_batch.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int]
    L.orc_audio_open.restype = vp
    L.orc_audio_open.argtypes = [C.c_char_p, C.c_size_t]
    L.orc_audio_close.argtypes = [vp]
    L.orc_audio_decode.restype = vp
    L.orc_audio_decode.argtypes = [vp, C.POINTER(C.c_double)]
    for name in ("orc_audio_has_header", "orc_audio_samplerate", "orc_audio_channels"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = C.c_int
    L.orc_audio_set_format.argtypes = [vp, C.c_int]
    L.orc_audio_set_fma.argtypes = [vp, C.c_int]
    L.orc_audio_rewind.argtypes = [vp]
    L.orc_audio_last_samples.restype = i32p
    L.orc_audio_last_samples.argtypes = [vp]
    L.orc_audio_state.restype = C.POINTER(SynthState)
    L.orc_audio_state.argtypes = [vp]
    L.orc_demux_split.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(vp),
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_free.argtypes = [vp]
    _lib = L
    return L


def fnv(h: int, arr) -> int:
    a = np.ascontiguousarray(arr)
    return lib().orc_fnv1a64(h, a.ctypes.data, a.nbytes)


# ------------------------------------------------------------------------------------------
# video
# ------------------------------------------------------------------------------------------
class FrameSet:
    """frames[stream*3 + buf]: the three physical buffers of n streams of one geometry."""

    def __init__(self, n_streams: int, width: int, height: int):
        self.n_streams, self.width, self.height = n_streams, width, height
        self.arr = (Frame * (3 * n_streams))()
        for i in range(3 * n_streams):
            if lib().orc_frame_init(C.byref(self.arr[i]), width, height) != 0:
                raise MemoryError
        f = self.arr[0]
        self.luma_w, self.luma_h, self.chroma_w, self.chroma_h = f.luma_w, f.luma_h, f.chroma_w, f.chroma_h
        self.buf_bytes = f.buf_bytes

    def frame(self, stream: int, buf: int) -> Frame:
        return self.arr[stream * 3 + buf]

    def whole(self, stream: int, buf: int) -> np.ndarray:
        return self.frame(stream, buf).whole()

    def exec_pictures(self, pics: np.ndarray, mbs: np.ndarray, coeffs: np.ndarray, threads: int = 1) -> int:
        pics = np.ascontiguousarray(pics)
        mbs = np.ascontiguousarray(mbs)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.int16)
        return lib().orc_exec_pictures(self.arr, len(pics), pics.ctypes.data, len(mbs), mbs.ctypes.data,
                                       coeffs.ctypes.data, threads)

    def rgba(self, stream: int, buf: int) -> np.ndarray:
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        lib().orc_rgba(C.byref(self.frame(stream, buf)), out.ctypes.data)
        return out

    def rgba_batch(self, streams, bufs, out: np.ndarray, threads: int = 1):
        """Frame.RGBA() of frames (streams[i], bufs[i]) into out[i] (uint8 [n, height, width, 4], preallocated)."""
        streams = np.ascontiguousarray(streams, np.int32)
        bufs = np.ascontiguousarray(bufs, np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous and out.shape[0] >= len(streams)
        lib().orc_rgba_batch(self.arr, len(streams), streams.ctypes.data, bufs.ctypes.data, out.ctypes.data,
                             self.width * self.height * 4, threads)

    def close(self):
        for i in range(3 * self.n_streams):
            lib().orc_frame_free(C.byref(self.arr[i]))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VideoOracle:
    """orc_video: restatement of mpeg.NewVideo / Video.Decode (video.go)."""

    def __init__(self, data: bytes, tap: bool = False):
        self._data = data
        self.h = lib().orc_video_open(data, len(data))
        if tap:
            lib().orc_video_tap_enable(self.h, 1)

    def has_header(self):
        return bool(lib().orc_video_has_header(self.h))

    @property
    def width(self):
        return lib().orc_video_width(self.h)

    @property
    def height(self):
        return lib().orc_video_height(self.h)

    @property
    def framerate(self):
        return lib().orc_video_framerate(self.h)

    def set_no_delay(self, on: bool):
        lib().orc_video_set_no_delay(self.h, int(on))

    def rewind(self):
        lib().orc_video_rewind(self.h)

    def decode(self):
        f = lib().orc_video_decode(self.h)
        return f.contents if f else None

    def last_buf(self) -> int:
        return lib().orc_video_last_buf(self.h)

    def oob_count(self) -> int:
        return lib().orc_video_oob_count(self.h)

    def tap(self):
        """Packed records of the pictures decoded by the last decode() call (copies)."""
        p = C.c_void_p()
        n_p = lib().orc_video_tap_pictures(self.h, C.byref(p))
        m = C.c_void_p()
        n_m = lib().orc_video_tap_mbs(self.h, C.byref(m))
        c = C.c_void_p()
        n_b = lib().orc_video_tap_blocks(self.h, C.byref(c))
        pics = np.frombuffer(C.string_at(p, n_p * 16), dtype=PIC_DTYPE).copy() if n_p else np.zeros(0, PIC_DTYPE)
        mbs = np.frombuffer(C.string_at(m, n_m * 16), dtype=MB_DTYPE).copy() if n_m else np.zeros(0, MB_DTYPE)
        coeffs = (np.frombuffer(C.string_at(c, n_b * 128), dtype=np.int16).reshape(n_b, 64).copy()
                  if n_b else np.zeros((0, 64), np.int16))
        return pics, mbs, coeffs

    def close(self):
        if self.h:
            lib().orc_video_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------
# audio
# ------------------------------------------------------------------------------------------
class AudioOracle:
    """orc_audio: restatement of mpeg.NewAudio / Audio.Decode (audio.go)."""

    def __init__(self, data: bytes, fmt: int = 0, fma: bool = False):
        self._data = data
        self.fmt = fmt
        self.h = lib().orc_audio_open(data, len(data))
        lib().orc_audio_set_format(self.h, fmt)
        lib().orc_audio_set_fma(self.h, int(fma))

    def has_header(self):
        return bool(lib().orc_audio_has_header(self.h))

    @property
    def samplerate(self):
        return lib().orc_audio_samplerate(self.h)

    @property
    def channels(self):
        return lib().orc_audio_channels(self.h)

    def rewind(self):
        lib().orc_audio_rewind(self.h)

    def decode(self):
        t = C.c_double()
        p = lib().orc_audio_decode(self.h, C.byref(t))
        if not p:
            return None
        if self.fmt == 3:
            return np.frombuffer(C.string_at(p, 2304 * 2), dtype=np.int16).copy()
        return np.frombuffer(C.string_at(p, 2304 * 4), dtype=np.float32).copy()

    def last_samples(self) -> np.ndarray:
        p = lib().orc_audio_last_samples(self.h)
        return np.ctypeslib.as_array(p, shape=(2, 36, 32)).copy()

    def state(self):
        st = lib().orc_audio_state(self.h).contents
        return np.ctypeslib.as_array(st.v).reshape(2, 1024).copy(), st.v_pos

    def close(self):
        if self.h:
            lib().orc_audio_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_states(n: int):
    return (SynthState * n)()


def synth_batch(states, n_streams: int, frames_per_stream: int, samples: np.ndarray, fmt: int = 0, fma: bool = False,
                threads: int = 1) -> np.ndarray:
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    assert samples.size == n_streams * frames_per_stream * 2 * 36 * 32
    out = np.empty((n_streams, frames_per_stream, 2304), dtype=np.int16 if fmt == 3 else np.float32)
    lib().orc_synth_batch(states, n_streams, frames_per_stream, samples.ctypes.data, fmt, out.ctypes.data, int(fma), threads)
    return out


def demux_split(data: bytes):
    v, a = C.c_void_p(), C.c_void_p()
    vl, al = C.c_size_t(), C.c_size_t()
    nv, na = C.c_int(), C.c_int()
    rc = lib().orc_demux_split(data, len(data), C.byref(v), C.byref(vl), C.byref(a), C.byref(al), C.byref(nv), C.byref(na))
    if rc != 0:
        raise ValueError("not an MPEG-PS stream")
    video = C.string_at(v, vl.value)
    audio = C.string_at(a, al.value)
    lib().orc_free(v)
    lib().orc_free(a)
    return video, audio, nv.value, na.value
