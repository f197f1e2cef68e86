"""Many video streams decoded in lock-step: one kernel launch per picture step for all of them.

This is the deployment the kernels are built for: every stream's host parser
produces its next Decode() step on a pool of host threads (mpegb200_video_batch_*), the per-stream launches
are merged into waves, and each wave is one mpegb200_video_decode_pictures call.  Stream i of the batch is
stream id `first_stream + i` of the context.
"""
import ctypes as C

import numpy as np

from . import _lib
from .context import Context


class Wave(C.Structure):
    _fields_ = [("n_pictures", C.c_int), ("pics", C.c_void_p), ("n_mb", C.c_size_t), ("mbs", C.c_void_p),
                ("n_blocks", C.c_size_t), ("coeffs", C.c_void_p), ("vlen_headers", C.c_void_p), ("vlen_chunk_offsets", C.c_void_p),
                ("vlen_payload", C.c_void_p), ("vlen_payload_bytes", C.c_size_t)]


class BatchStep(C.Structure):
    _fields_ = [("n_streams", C.c_int), ("has_frame", C.POINTER(C.c_int)), ("frame_buf", C.POINTER(C.c_int)),
                ("time", C.POINTER(C.c_double)), ("n_waves", C.c_int), ("waves", C.POINTER(Wave))]


class VideoBatch:
    def __init__(self, ctx: Context, streams, threads: int = 8, first_stream: int = 0, pinned: bool = True, validate: bool = True,
                 vlen: bool = True):
        self.L = _lib.load()
        ctx.set_validate(validate)   # bitstream-derived records: a malformed wave raises instead of decoding (pass False for trusted input)
        self.ctx, self.n, self.first = ctx, len(streams), first_stream
        alloc = C.cast(self.L.mpegb200_host_alloc, C.c_void_p) if pinned else None
        free = C.cast(self.L.mpegb200_host_free, C.c_void_p) if pinned else None
        self.h = self.L.mpegb200_video_batch_new(self.n, threads, alloc, free)
        if not self.h:
            raise MemoryError
        self.L.mpegb200_video_batch_set_vlen(self.h, int(vlen))   # coefficients leave the parsers in the variable-width transfer form
        self._data = [bytes(s) for s in streams]
        w, h = C.c_int(), C.c_int()
        self.sizes = []
        for i, d in enumerate(self._data):
            if self.L.mpegb200_video_batch_set_stream(self.h, i, d, len(d)) != 0:
                raise MemoryError
            self.L.mpegb200_video_batch_stream_size(self.h, i, C.byref(w), C.byref(h))
            if w.value <= 0 or h.value <= 0:
                raise ValueError(f"stream {i}: no MPEG-1 sequence header")
            self.sizes.append((w.value, h.value))
            ctx.video_open(first_stream + i, w.value, h.value)
        self._pic_dtype = np.dtype([("stream", "<i4"), ("rest", "V12")])
        self.steps = 0

    def step(self):
        """One Video.Decode() of every stream.  Returns (has_frame[n] bool, frame_buf[n] uint8, time[n]).
        The kernels run asynchronously; call ctx.sync() (or read frames back) before touching results."""
        if self.steps >= 2:
            self.ctx._ck(self.L.mpegb200_sync_uploads(self.ctx.h))  # the arrays about to be re-used were uploaded
        st = BatchStep()
        self.ctx._ck(self.L.mpegb200_video_batch_next(self.h, C.byref(st)))
        for w in range(st.n_waves):
            wave = st.waves[w]
            if wave.n_mb == 0:
                continue
            if self.first:  # stream index in the batch -> stream id in the context
                pics = np.ctypeslib.as_array(C.cast(wave.pics, C.POINTER(C.c_int32)), shape=(wave.n_pictures, 4))
                pics[:, 0] += self.first
            if wave.vlen_headers:
                self.ctx._ck(self.L.mpegb200_video_decode_pictures_vlen(
                    self.ctx.h, wave.n_pictures, C.c_void_p(wave.pics), wave.n_mb, C.c_void_p(wave.mbs), wave.n_blocks,
                    C.c_void_p(wave.vlen_headers), C.c_void_p(wave.vlen_chunk_offsets), C.c_void_p(wave.vlen_payload), wave.vlen_payload_bytes))
                continue
            self.ctx._ck(self.L.mpegb200_video_decode_pictures(self.ctx.h, wave.n_pictures, C.c_void_p(wave.pics), wave.n_mb,
                                                               C.c_void_p(wave.mbs), wave.n_blocks, C.c_void_p(wave.coeffs)))
        self.steps += 1
        has = np.ctypeslib.as_array(st.has_frame, shape=(self.n,)).astype(bool)
        buf = np.ctypeslib.as_array(st.frame_buf, shape=(self.n,)).astype(np.uint8)
        t = np.ctypeslib.as_array(st.time, shape=(self.n,)).copy()
        return has, buf, t

    def close(self):
        if getattr(self, "h", None):
            self.ctx.sync()
            for i in range(self.n):
                try:
                    self.ctx.video_close(self.first + i)
                except Exception:
                    pass
            self.L.mpegb200_video_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
