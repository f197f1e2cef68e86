"""Many streams decoded in lock-step: one kernel launch per picture step (video) / per frame step (audio) for all of them,
a display ring that keeps the returned frames in device memory, and the batched program-stream front end.

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


class DisplayRing:
    """The frames Video.Decode() returns, kept in device memory in display order for `depth` steps (mpegb200_video_ring_*):
    consumers may lag behind the decoder.  Stream i of the ring is context stream streams[i]."""

    def __init__(self, ctx: Context, streams, depth: int = 4):
        self.L = _lib.load()
        self.ctx = ctx
        self.streams = np.ascontiguousarray(streams, np.int32)
        self.n, self.depth = len(self.streams), depth
        self.h = self.L.mpegb200_video_ring_new(ctx.h, self.n, C.c_void_p(self.streams.ctypes.data), depth)
        if not self.h:
            raise MemoryError((self.L.mpegb200_last_error(ctx.h) or b"").decode())
        stride = C.c_size_t()
        self.L.mpegb200_video_ring_slot_dev(self.h, 0, C.byref(stride))
        self.stride = stride.value

    def push(self, has_frame, frame_buf) -> int:
        """After a lock-step Decode(): copy every returned frame into the next slot; returns the slot index."""
        bufs = np.where(np.asarray(has_frame, bool), np.asarray(frame_buf, np.uint8), 255).astype(np.uint8)
        slot = self.L.mpegb200_video_ring_push(self.h, C.c_void_p(bufs.ctypes.data))
        if slot < 0:
            self.ctx._ck(slot)
        return slot

    def slot_dev(self, slot: int) -> int:
        return int(self.L.mpegb200_video_ring_slot_dev(self.h, slot, None))

    def read(self, slot: int, picture_bytes: int = None) -> np.ndarray:
        """The slot's pictures as a host array [n, picture_bytes] (synchronises)."""
        width = picture_bytes or self.stride
        out = np.empty((self.n, width), np.uint8)
        self.ctx._ck(self.L.mpegb200_video_ring_read_host(self.h, slot, C.c_void_p(out.ctypes.data), width))
        self.ctx.sync()
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.mpegb200_video_ring_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AudioBatchStep(C.Structure):
    _fields_ = [("n_streams", C.c_int), ("n_frames", C.POINTER(C.c_int)), ("time", C.POINTER(C.c_double)),
                ("frames_per_stream", C.c_int), ("n_full", C.c_int), ("full_index", C.POINTER(C.c_int32)), ("full_samples", C.c_void_p),
                ("n_tail", C.c_int), ("tail_index", C.POINTER(C.c_int32)), ("tail_frames", C.POINTER(C.c_int32)), ("tail_samples", C.c_void_p)]


class AudioBatch:
    """Many MP2 streams in lock-step: every step parses up to `frames_per_step` frames of every stream on the host threads and
    synthesises them with ONE launch (the rectangular part) plus one small launch per stream that ends inside the step.
    Stream i of the batch is audio stream id first_stream + i of the context."""

    def __init__(self, ctx: Context, streams, threads: int = 8, first_stream: int = 0, fmt: int = 0, frames_per_step: int = 8,
                 pinned: bool = True):
        self.L = _lib.load()
        self.ctx, self.n, self.first, self.fmt, self.F = ctx, len(streams), first_stream, fmt, frames_per_step
        alloc = C.cast(self.L.mpegb200_host_alloc, C.c_void_p) if pinned else None
        free = C.cast(self.L.mpegb200_host_free, C.c_void_p) if pinned else None
        self.h = self.L.mpegb200_audio_batch_new(self.n, threads, alloc, free)
        if not self.h:
            raise MemoryError
        self._data = [bytes(s) for s in streams]
        for i, d in enumerate(self._data):
            if self.L.mpegb200_audio_batch_set_stream(self.h, i, d, len(d)) != 0:
                raise MemoryError
            ctx.audio_open(first_stream + i)
        self._dtype = np.int16 if (fmt & 0xff) == 3 else np.float32

    def step(self):
        """Returns (n_frames[n], time[n], samples): samples[i] is an array [n_frames[i], 2304] in the batch's format (or None)."""
        st = AudioBatchStep()
        self.ctx._ck(self.L.mpegb200_audio_batch_next(self.h, self.F, C.byref(st)))
        n_frames = np.ctypeslib.as_array(st.n_frames, shape=(self.n,)).copy()
        times = np.ctypeslib.as_array(st.time, shape=(self.n,)).copy()
        out = [None] * self.n
        if st.n_full:
            idx = np.ctypeslib.as_array(st.full_index, shape=(st.n_full,)).copy()
            ids = (idx + self.first).astype(np.int32)
            res = np.empty((st.n_full, self.F, 2304), self._dtype)
            self.ctx._ck(self.L.mpegb200_audio_synth(self.ctx.h, st.n_full, C.c_void_p(ids.ctypes.data), self.F, C.c_void_p(st.full_samples),
                                                     self.fmt, C.c_void_p(res.ctypes.data)))
            for j, i in enumerate(idx):
                out[int(i)] = res[j]
        at = 0
        for j in range(st.n_tail):
            i, k = int(st.tail_index[j]), int(st.tail_frames[j])
            ids = np.array([i + self.first], np.int32)
            res = np.empty((1, k, 2304), self._dtype)
            self.ctx._ck(self.L.mpegb200_audio_synth(self.ctx.h, 1, C.c_void_p(ids.ctypes.data), k, C.c_void_p(st.tail_samples + at * 2 * 36 * 32 * 4),
                                                     self.fmt, C.c_void_p(res.ctypes.data)))
            out[i] = res[0]
            at += k
        return n_frames, times, out

    def close(self):
        if getattr(self, "h", None):
            for i in range(self.n):
                try:
                    self.ctx.audio_close(self.first + i)
                except Exception:
                    pass
            self.L.mpegb200_audio_batch_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MPEGBatch:
    """The batched front end for program streams (demux.go:473-584 per stream, mpeg.go:356-411 without the player clock): every
    .mpg is split into its video and audio elementary streams (host threads), the video streams decode in lock-step through a
    VideoBatch with a display ring behind it, the audio streams through an AudioBatch."""

    def __init__(self, ctx: Context, program_streams, threads: int = 8, ring_depth: int = 4, audio_fmt: int = 0, frames_per_step: int = 8):
        from concurrent.futures import ThreadPoolExecutor
        from .mpeg import demux_split
        with ThreadPoolExecutor(max(1, threads)) as ex:   # ctypes releases the GIL inside mpegb200_demux_split
            parts = list(ex.map(demux_split, [bytes(p) for p in program_streams]))
        self.packets = [(p[2], p[3]) for p in parts]
        self.video = VideoBatch(ctx, [p[0] for p in parts], threads=threads)
        self.audio = AudioBatch(ctx, [p[1] for p in parts], threads=threads, fmt=audio_fmt, frames_per_step=frames_per_step)
        self.ring = DisplayRing(ctx, np.arange(len(parts), dtype=np.int32), ring_depth)

    def decode_video(self):
        """One DecodeVideo() of every stream: (has_frame, time, ring slot holding the returned frames in device memory)."""
        has, buf, t = self.video.step()
        slot = self.ring.push(has, buf) if has.any() else -1
        return has, t, slot

    def decode_audio(self):
        return self.audio.step()

    def close(self):
        self.ring.close()
        self.video.close()
        self.audio.close()
