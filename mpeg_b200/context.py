"""Context: thin, typed wrapper over the C-ABI (one per GPU / host thread)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MpegB200Error

# record dtypes of include/mpegb200.h
MB_DTYPE = np.dtype([
    ("mb_row", "<u2"), ("mb_col", "<u2"), ("mv_h", "<i2"), ("mv_v", "<i2"),
    ("flags", "u1"), ("cbp", "u1"), ("pic", "<u2"), ("coeff_block", "<u4"),
])
PICTURE_DTYPE = np.dtype([
    ("stream", "<i4"), ("type", "u1"), ("dst_buf", "u1"), ("fwd_buf", "u1"), ("bwd_buf", "u1"),
    ("first_mb", "<u4"), ("n_mb", "<u4"),
])
MB_INTRA, MB_PREDICT, MB_REF_BWD = 0x01, 0x02, 0x04
PIC_I, PIC_P, PIC_B = 1, 2, 3
AUDIO_F32N, AUDIO_F32NLR, AUDIO_F32, AUDIO_S16 = 0, 1, 2, 3
AUDIO_WINDOW_FMA = 0x100   # OR into the format: fused multiply-add window (the reference's AVX2 / NEON back-end)
SAMPLES_PER_FRAME = 1152


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class Context:
    def __init__(self, device: int = 0, max_streams: int = 256):
        self.L = _lib.load()
        err = C.c_int(0)
        self.h = self.L.mpegb200_create(device, max_streams, C.byref(err))
        if not self.h:
            raise MpegB200Error(err.value, "mpegb200_create failed: an sm_100 (B200) CUDA device is required; "
                                           "there is no CPU fallback")
        self.device, self.max_streams = device, max_streams

    def _ck(self, rc):
        if rc != 0:
            raise MpegB200Error(rc, (self.L.mpegb200_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mpegb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- plumbing
    def set_stream(self, cuda_stream_handle: int):
        self._ck(self.L.mpegb200_set_stream(self.h, C.c_void_p(cuda_stream_handle)))

    def sync(self):
        self._ck(self.L.mpegb200_sync(self.h))

    def join_readbacks(self):
        """The compute stream waits for the asynchronous read-backs enqueued so far (events on it then cover them)."""
        self._ck(self.L.mpegb200_join_readbacks(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.L.mpegb200_launch_count(self.h))

    def set_validate(self, on: bool):
        """Host-pointer decode calls validate their records first and raise instead of decoding malformed ones."""
        self._ck(self.L.mpegb200_set_validate(self.h, int(on)))

    def set_kernel_timing(self, on: bool):
        """Measurement aid: bracket the kernels of every decode call with CUDA events (mpegb200_set_kernel_timing)."""
        self._ck(self.L.mpegb200_set_kernel_timing(self.h, int(on)))

    def kernel_times(self, cap: int = 4096):
        """(plan_ms, fused_ms) arrays of the decode calls since the last read; synchronises the stream."""
        a, b = np.empty(cap, np.float32), np.empty(cap, np.float32)
        n = self.L.mpegb200_kernel_times(self.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), cap)
        if n < 0:
            self._ck(n)
        return a[:n].copy(), b[:n].copy()

    # ---- video
    def video_open(self, stream: int, width: int, height: int):
        self._ck(self.L.mpegb200_video_open(self.h, stream, width, height))

    def video_close(self, stream: int):
        self._ck(self.L.mpegb200_video_close(self.h, stream))

    def video_geometry(self, stream: int):
        lw, lh, cw, ch = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        fb = C.c_size_t()
        self._ck(self.L.mpegb200_video_geometry(self.h, stream, C.byref(lw), C.byref(lh), C.byref(cw), C.byref(ch), C.byref(fb)))
        return lw.value, lh.value, cw.value, ch.value, fb.value

    @staticmethod
    def _records(pics, mbs, coeffs):
        pics = np.ascontiguousarray(pics, dtype=PICTURE_DTYPE)
        mbs = np.ascontiguousarray(mbs, dtype=MB_DTYPE)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.int16).reshape(-1, 64)
        return pics, mbs, coeffs

    def video_validate(self, pics, mbs, n_blocks: int):
        pics = np.ascontiguousarray(pics, dtype=PICTURE_DTYPE)
        mbs = np.ascontiguousarray(mbs, dtype=MB_DTYPE)
        self._ck(self.L.mpegb200_video_validate(self.h, len(pics), _ptr(pics), len(mbs), _ptr(mbs), n_blocks))

    def video_decode_pictures(self, pics, mbs, coeffs):
        """Host arrays -> H2D copy + fused MC/IDCT/add launch (asynchronous on the context stream)."""
        pics, mbs, coeffs = self._records(pics, mbs, coeffs)
        self._keep = (pics, mbs, coeffs)  # keep alive until the stream has consumed them
        self._ck(self.L.mpegb200_video_decode_pictures(self.h, len(pics), _ptr(pics), len(mbs), _ptr(mbs), len(coeffs), _ptr(coeffs)))

    def pack_coeffs12(self, coeffs) -> np.ndarray:
        """int16 blocks -> the 96-byte transfer form (raises if a value does not fit 12 bits)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.int16).reshape(-1, 64)
        out = np.empty((len(coeffs), 96), np.uint8)
        self._ck(self.L.mpegb200_pack_coeffs12(_ptr(coeffs), len(coeffs), _ptr(out)))
        return out

    def video_decode_pictures_packed(self, pics, mbs, coeffs12):
        """Like video_decode_pictures, with the coefficients in the 12-bit transfer form."""
        pics = np.ascontiguousarray(pics, dtype=PICTURE_DTYPE)
        mbs = np.ascontiguousarray(mbs, dtype=MB_DTYPE)
        coeffs12 = np.ascontiguousarray(coeffs12, dtype=np.uint8).reshape(-1, 96)
        self._keep = (pics, mbs, coeffs12)
        self._ck(self.L.mpegb200_video_decode_pictures_packed(self.h, len(pics), _ptr(pics), len(mbs), _ptr(mbs), len(coeffs12), _ptr(coeffs12)))

    def pack_coeffs_vlen(self, coeffs):
        """int16 blocks -> the variable-width transfer form: (headers u32[n], chunk_offsets u64[ceil(n/32)], payload u8[...])."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.int16).reshape(-1, 64)
        n = len(coeffs)
        headers = np.empty(n, np.uint32)
        chunks = np.empty((n + 31) // 32, np.uint64)
        cap = int(self.L.mpegb200_vlen_payload_bound(n))
        payload = np.empty(cap, np.uint8)
        used = C.c_size_t(0)
        self._ck(self.L.mpegb200_pack_coeffs_vlen(_ptr(coeffs), n, _ptr(headers), _ptr(chunks), _ptr(payload), cap, C.byref(used)))
        return headers, chunks, payload[:used.value].copy()

    def video_decode_pictures_vlen(self, pics, mbs, headers, chunks, payload):
        """Like video_decode_pictures, with the coefficients in the variable-width transfer form."""
        pics = np.ascontiguousarray(pics, dtype=PICTURE_DTYPE)
        mbs = np.ascontiguousarray(mbs, dtype=MB_DTYPE)
        headers = np.ascontiguousarray(headers, dtype=np.uint32)
        chunks = np.ascontiguousarray(chunks, dtype=np.uint64)
        payload = np.ascontiguousarray(payload, dtype=np.uint8)
        self._keep = (pics, mbs, headers, chunks, payload)
        self._ck(self.L.mpegb200_video_decode_pictures_vlen(self.h, len(pics), _ptr(pics), len(mbs), _ptr(mbs), len(headers),
                                                            _ptr(headers), _ptr(chunks), _ptr(payload), len(payload)))

    def video_decode_pictures_dev(self, n_pics: int, d_pics: int, n_mb: int, d_mbs: int, n_blocks: int, d_coeffs: int):
        """Device pointers (ints) of arrays already resident in HBM."""
        self._ck(self.L.mpegb200_video_decode_pictures_dev(self.h, n_pics, C.c_void_p(d_pics), n_mb, C.c_void_p(d_mbs), n_blocks, C.c_void_p(d_coeffs)))

    def video_read_planes(self, stream: int, buf: int):
        lw, lh, cw, ch, _ = self.video_geometry(stream)
        y = np.empty((lh, lw), np.uint8)
        cb = np.empty((ch, cw), np.uint8)
        cr = np.empty((ch, cw), np.uint8)
        self._ck(self.L.mpegb200_video_read_planes(self.h, stream, buf, _ptr(y), _ptr(cb), _ptr(cr)))
        return y, cb, cr

    def video_write_planes(self, stream: int, buf: int, y=None, cb=None, cr=None):
        y = None if y is None else np.ascontiguousarray(y, np.uint8)
        cb = None if cb is None else np.ascontiguousarray(cb, np.uint8)
        cr = None if cr is None else np.ascontiguousarray(cr, np.uint8)
        self._ck(self.L.mpegb200_video_write_planes(self.h, stream, buf, _ptr(y), _ptr(cb), _ptr(cr)))

    def video_read_frame(self, stream: int, buf: int) -> np.ndarray:
        fb = self.video_geometry(stream)[4]
        out = np.empty(fb, np.uint8)
        self._ck(self.L.mpegb200_video_read_frame(self.h, stream, buf, _ptr(out), out.nbytes))
        return out

    def video_write_frame(self, stream: int, buf: int, data):
        data = np.ascontiguousarray(data, np.uint8).reshape(-1)
        self._ck(self.L.mpegb200_video_write_frame(self.h, stream, buf, _ptr(data), data.nbytes))

    def video_read_pictures(self, streams, bufs, dst: int, stride: int, device: bool = False):
        """Asynchronous batched read-back of Y|Cb|Cr into host (pinned) or device memory at `dst`."""
        streams = np.ascontiguousarray(streams, np.int32)
        bufs = np.ascontiguousarray(bufs, np.uint8)
        fn = self.L.mpegb200_video_read_pictures_dev if device else self.L.mpegb200_video_read_pictures_host
        self._ck(fn(self.h, len(streams), _ptr(streams), _ptr(bufs), C.c_void_p(dst), stride))

    def video_frame_dev(self, stream: int, buf: int) -> int:
        p = self.L.mpegb200_video_frame_dev(self.h, stream, buf)
        if not p:
            raise MpegB200Error(-4, "stream not open")
        return int(p)

    def video_rgba(self, stream: int, buf: int, width: int, height: int) -> np.ndarray:
        out = np.empty((height, width, 4), np.uint8)
        self._ck(self.L.mpegb200_video_rgba(self.h, stream, buf, _ptr(out)))
        return out

    def video_rgba_batch_dev(self, streams, bufs, d_rgba: int, stride_bytes: int):
        streams = np.ascontiguousarray(streams, np.int32)
        bufs = np.ascontiguousarray(bufs, np.uint8)
        self._ck(self.L.mpegb200_video_rgba_batch_dev(self.h, len(streams), _ptr(streams), _ptr(bufs), C.c_void_p(d_rgba), stride_bytes))

    # ---- audio
    def audio_open(self, stream: int):
        self._ck(self.L.mpegb200_audio_open(self.h, stream))

    def audio_close(self, stream: int):
        self._ck(self.L.mpegb200_audio_close(self.h, stream))

    def audio_synth(self, stream_ids, frames_per_stream: int, samples, fmt: int = AUDIO_F32N) -> np.ndarray:
        ids = np.ascontiguousarray(stream_ids, np.int32)
        samples = np.ascontiguousarray(samples, np.int32)
        assert samples.size == len(ids) * frames_per_stream * 2 * 36 * 32
        out = np.empty((len(ids), frames_per_stream, 2 * SAMPLES_PER_FRAME), np.int16 if (fmt & 0xff) == AUDIO_S16 else np.float32)
        self._ck(self.L.mpegb200_audio_synth(self.h, len(ids), _ptr(ids), frames_per_stream, _ptr(samples), fmt, _ptr(out)))
        return out

    def audio_synth_coded(self, stream_ids, frames_per_stream: int, info, codes, fmt: int = AUDIO_F32N) -> np.ndarray:
        """Like audio_synth with the requantisation on the device: info uint8 [frames, 256] (quantiser + scale-factor indices),
        codes uint16 [frames, 2, 36, 32] (mpegb200_audio_parser_next_coded)."""
        ids = np.ascontiguousarray(stream_ids, np.int32)
        info = np.ascontiguousarray(info, np.uint8)
        codes = np.ascontiguousarray(codes, np.uint16)
        n = len(ids) * frames_per_stream
        assert info.size == n * 256 and codes.size == n * 2304
        out = np.empty((len(ids), frames_per_stream, 2 * SAMPLES_PER_FRAME), np.int16 if (fmt & 0xff) == AUDIO_S16 else np.float32)
        self._ck(self.L.mpegb200_audio_synth_coded(self.h, len(ids), _ptr(ids), frames_per_stream, _ptr(info), _ptr(codes), fmt, _ptr(out)))
        return out

    def audio_synth_dev(self, stream_ids, frames_per_stream: int, d_samples: int, fmt: int, d_out: int):
        ids = np.ascontiguousarray(stream_ids, np.int32)
        self._ck(self.L.mpegb200_audio_synth_dev(self.h, len(ids), _ptr(ids), frames_per_stream, C.c_void_p(d_samples), fmt, C.c_void_p(d_out)))

    def audio_read_state(self, stream: int):
        v = np.empty((2, 1024), np.float32)
        vp = C.c_int()
        self._ck(self.L.mpegb200_audio_read_state(self.h, stream, _ptr(v), C.byref(vp)))
        return v, vp.value

    def audio_write_state(self, stream: int, v, v_pos: int):
        v = np.ascontiguousarray(v, np.float32).reshape(2, 1024)
        self._ck(self.L.mpegb200_audio_write_state(self.h, stream, _ptr(v), v_pos))
