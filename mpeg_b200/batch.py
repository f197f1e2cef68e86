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


class VlcPicture(C.Structure):      # mpegb200_vlc_picture
    _fields_ = [("stream", C.c_int32), ("type", C.c_uint8), ("dst_buf", C.c_uint8), ("fwd_buf", C.c_uint8), ("bwd_buf", C.c_uint8),
                ("fwd_full_px", C.c_uint8), ("fwd_r_size", C.c_uint8), ("bwd_full_px", C.c_uint8), ("bwd_r_size", C.c_uint8),
                ("first_slice", C.c_uint32), ("n_slices", C.c_uint32), ("mb_slot", C.c_uint32), ("n_mb_slots", C.c_uint32),
                ("quant", C.c_uint32)]


class VlcSlice(C.Structure):        # mpegb200_vlc_slice
    _fields_ = [("data_offset", C.c_uint64), ("next_code", C.c_uint32), ("stream_left", C.c_uint32), ("pic", C.c_uint32),
                ("vpos", C.c_uint32), ("mb_slot", C.c_uint32), ("mb_cap", C.c_uint32)]


class VlcWave(C.Structure):         # mpegb200_vlc_wave
    _fields_ = [("n_pictures", C.c_int), ("pics", C.POINTER(VlcPicture)), ("step_picture", C.POINTER(C.c_int32)),
                ("n_slices", C.c_size_t), ("slices", C.POINTER(VlcSlice)), ("bitstream", C.c_void_p), ("bitstream_bytes", C.c_size_t),
                ("quant", C.c_void_p), ("n_quant", C.c_size_t), ("n_mb_slots", C.c_size_t)]


class BatchScanStep(C.Structure):   # mpegb200_batch_scan_step
    _fields_ = [("n_streams", C.c_int), ("has_frame", C.POINTER(C.c_int)), ("frame_buf", C.POINTER(C.c_int)),
                ("time", C.POINTER(C.c_double)), ("n_waves", C.c_int), ("waves", C.POINTER(VlcWave)),
                ("n_host", C.c_int), ("host_index", C.POINTER(C.c_int)), ("host_steps", C.c_void_p)]


assert C.sizeof(VlcPicture) == 32 and C.sizeof(VlcSlice) == 32


class StepperStats(C.Structure):   # mpegb200_device_stepper_stats
    _fields_ = [("steps", C.c_uint64), ("waves", C.c_uint64), ("flagged_pictures", C.c_uint64), ("host_steps", C.c_uint64),
                ("withdrawn_scans", C.c_uint64), ("seconds_host_scan", C.c_double), ("seconds_submit", C.c_double), ("seconds_waiting", C.c_double)]


class VideoBatch:
    def __init__(self, ctx: Context, streams, threads: int = 8, first_stream: int = 0, pinned: bool = True, validate: bool = True,
                 vlen: bool = True, device_vlc: bool = False, scan_ahead: bool = True, resident: bool = False, native_step: bool = True):
        """device_vlc: the host only scans headers and start codes; the slices are parsed on the GPU, one thread per slice
        (mpegb200_video_decode_bitstream), and pictures the device flags are re-parsed by the host parser.
        resident (with device_vlc): every stream is uploaded to device memory once and its start codes are indexed there
        (mpegb200_video_stream_upload / _index): per step the host touches headers only and the waves carry tables only.
        native_step (with device_vlc): the step's control flow (waves, scan-ahead, flags, re-parses) runs in the library
        (mpegb200_device_stepper_*, C++); False: the same flow written out in Python below (_step_device_vlc), kept as the
        readable reference of the call sequence and for A/B."""
        self.L = _lib.load()
        self.device_vlc = device_vlc
        self.flagged = 0    # pictures the device flagged so far: their step's tail took the host path (device_vlc)
        self.host_steps = 0  # steps the host parsed itself because stale coefficients were pending (device_vlc)
        self.t_scan = self.t_submit = self.t_wait = 0.0   # device_vlc: seconds in the host scan, in the submission, waiting for the flags
        self.scan_ahead = scan_ahead   # device_vlc: scan step k + 1 on the host while the device works on step k
        self._ahead = None
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
        self.resident = bool(resident and device_vlc)
        if self.resident:
            self.L.mpegb200_video_batch_set_resident(self.h, 1)
            n_codes = C.c_size_t()
            for i, d in enumerate(self._data):
                ctx._ck(self.L.mpegb200_video_stream_upload(ctx.h, first_stream + i, d, len(d)))
                pos = np.empty(len(d) // 64 + 4096, np.uint64)
                rc = self.L.mpegb200_video_stream_index(ctx.h, first_stream + i, C.c_void_p(pos.ctypes.data), len(pos), C.byref(n_codes))
                if rc != 0 and n_codes.value > len(pos):   # a stream of (almost) nothing but start codes
                    pos = np.empty(n_codes.value, np.uint64)
                    rc = self.L.mpegb200_video_stream_index(ctx.h, first_stream + i, C.c_void_p(pos.ctypes.data), len(pos), C.byref(n_codes))
                ctx._ck(rc)
                if self.L.mpegb200_video_batch_set_start_codes(self.h, i, C.c_void_p(pos.ctypes.data), n_codes.value) != 0:
                    raise RuntimeError(f"stream {i}: the device's start-code index does not fit the stream")
        self._stepper = None
        if device_vlc and native_step:
            self._stepper = self.L.mpegb200_device_stepper_new(ctx.h, self.h, first_stream, int(scan_ahead))
            if not self._stepper:
                raise MemoryError
            self._has, self._buf, self._time = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32), np.zeros(self.n, np.float64)
        self._pic_dtype = np.dtype([("stream", "<i4"), ("rest", "V12")])
        self.steps = 0

    def step(self):
        """One Video.Decode() of every stream.  Returns (has_frame[n] bool, frame_buf[n] uint8, time[n]).
        The kernels run asynchronously; call ctx.sync() (or read frames back) before touching results."""
        if self._stepper:
            self.ctx._ck(self.L.mpegb200_device_stepper_step(self._stepper, C.c_void_p(self._has.ctypes.data), C.c_void_p(self._buf.ctypes.data),
                                                              C.c_void_p(self._time.ctypes.data)))
            st = StepperStats()
            self.L.mpegb200_device_stepper_get_stats(self._stepper, C.byref(st))
            self.flagged, self.host_steps, self.steps = int(st.flagged_pictures), int(st.host_steps), int(st.steps)
            self.t_scan, self.t_submit, self.t_wait = st.seconds_host_scan, st.seconds_submit, st.seconds_waiting
            return self._has.astype(bool), self._buf.astype(np.uint8), self._time.copy()
        if self.device_vlc:
            return self._step_device_vlc()
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

    def parser(self, i: int):
        """Handle of stream i's host parser (set_no_delay / rewind / has_ended of one stream of the batch)."""
        return self.L.mpegb200_video_batch_parser(self.h, i)

    def drop_scan_ahead(self):
        """Withdraw a scan made ahead of time (before a control call that changes where a parser stands)."""
        if self._stepper:
            self.ctx._ck(self.L.mpegb200_device_stepper_drop_scan_ahead(self._stepper))
        if self._ahead is not None:
            self.ctx._ck(self.L.mpegb200_video_batch_unscan(self.h))
            self._ahead = None

    def _scan(self):
        """One scan step of every stream (host): the step structure plus copies of its per-stream results."""
        import time
        st = BatchScanStep()
        t0 = time.perf_counter()
        self.ctx._ck(self.L.mpegb200_video_batch_next_scan(self.h, C.byref(st)))
        self.t_scan += time.perf_counter() - t0
        has = np.ctypeslib.as_array(st.has_frame, shape=(self.n,)).astype(bool)
        buf = np.ctypeslib.as_array(st.frame_buf, shape=(self.n,)).astype(np.uint8)
        t = np.ctypeslib.as_array(st.time, shape=(self.n,)).copy()
        return st, has, buf, t

    def _step_device_vlc(self):
        import time
        from .mpeg import VideoStep
        st, has, buf, t = self._ahead if self._ahead is not None else self._scan()
        self._ahead = None
        # streams whose step the host parsed itself (stale coefficients pending): plain launches, no part in the waves
        host_steps = C.cast(st.host_steps, C.POINTER(VideoStep))
        for j in range(st.n_host):
            self._run_launches(host_steps[j], st.host_index[j] + self.first)
            self.host_steps += 1
        done = set()    # streams whose step was finished by the host parser after a flag: their later pictures are void
        for w in range(st.n_waves):
            wave = st.waves[w]
            if wave.n_pictures == 0:
                continue
            if done or self.first:
                for k in range(wave.n_pictures):
                    if wave.pics[k].stream in done:
                        wave.pics[k].type = 0
                    wave.pics[k].stream += self.first
            t0 = time.perf_counter()
            self.ctx._ck(self.L.mpegb200_video_decode_bitstream(
                self.ctx.h, wave.n_pictures, wave.pics, wave.n_slices, wave.slices, C.c_void_p(wave.bitstream), wave.bitstream_bytes,
                C.c_void_p(wave.quant), wave.n_quant, wave.n_mb_slots))
            self.t_submit += time.perf_counter() - t0
            # While the device parses the step's last wave the host scans the next step.  That is a guess -- it assumes no
            # picture of this wave flags; if one does the guess is withdrawn below (mpegb200_video_batch_unscan).
            if self.scan_ahead and w == st.n_waves - 1:
                self._ahead = self._scan()
            # The flags come back with the wave: what the serial reference resolves by order of arrival goes through the host
            # parser, before anything builds on this wave.
            t1 = time.perf_counter()
            flags = np.zeros(wave.n_pictures, np.int32)
            bad = self.L.mpegb200_video_bitstream_flags(self.ctx.h, C.c_void_p(flags.ctypes.data), wave.n_pictures)
            self.t_wait += time.perf_counter() - t1
            if bad < 0:
                self.ctx._ck(bad)
            for k in np.nonzero(flags)[0] if bad else ():
                index = wave.pics[k].stream - self.first
                if index in done:
                    continue
                if self._ahead is not None:    # the parsers go back to where they stood after this step's scan
                    self.ctx._ck(self.L.mpegb200_video_batch_unscan(self.h))
                    self._ahead = None
                redo = VideoStep()
                self.ctx._ck(self.L.mpegb200_video_batch_redo(self.h, index, wave.step_picture[k], C.byref(redo)))
                self._run_launches(redo, index + self.first)
                has[index], buf[index], t[index] = bool(redo.has_frame), redo.frame_buf, redo.time
                done.add(index)
                self.flagged += 1
        self.steps += 1
        return has, buf, t

    def _run_launches(self, step, stream_id: int):
        """The launches of one host-parsed picture (mpegb200_video_step) for context stream `stream_id`."""
        for i in range(step.n_launches):
            ln = step.launches[i]
            if ln.n_mb == 0:
                continue
            ln.stream = stream_id
            mb_ptr = step.mbs + 16 * ln.first_mb
            if step.vlen_launches:
                lv = step.vlen_launches[i]
                self.ctx._ck(self.L.mpegb200_video_decode_pictures_vlen(
                    self.ctx.h, 1, C.byref(ln), ln.n_mb, C.c_void_p(mb_ptr), ln.n_blocks,
                    C.c_void_p((step.vlen_headers or 0) + 4 * ln.first_block), C.c_void_p((step.vlen_chunk_offsets or 0) + 8 * lv.first_chunk),
                    C.c_void_p((step.vlen_payload or 0) + lv.payload_offset), lv.payload_bytes))
            else:
                self.ctx._ck(self.L.mpegb200_video_decode_pictures(self.ctx.h, 1, C.byref(ln), ln.n_mb, C.c_void_p(mb_ptr), ln.n_blocks,
                                                                   C.c_void_p(step.coeffs + 128 * ln.first_block)))
        self.ctx._ck(self.L.mpegb200_sync_uploads(self.ctx.h))   # the parser re-uses these arrays on its next call

    def close(self):
        if getattr(self, "h", None):
            self.ctx.sync()
            if getattr(self, "_stepper", None):
                self.L.mpegb200_device_stepper_free(self._stepper)
                self._stepper = None
            for i in range(self.n):
                try:
                    self.ctx.video_close(self.first + i)
                    if self.resident:
                        self.L.mpegb200_video_stream_upload(self.ctx.h, self.first + i, None, 0)
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

    def __init__(self, ctx: Context, program_streams, threads: int = 8, ring_depth: int = 4, audio_fmt: int = 0, frames_per_step: int = 8,
                 device_vlc: bool = False):
        from concurrent.futures import ThreadPoolExecutor
        from .mpeg import demux_split
        with ThreadPoolExecutor(max(1, threads)) as ex:   # ctypes releases the GIL inside mpegb200_demux_split
            parts = list(ex.map(demux_split, [bytes(p) for p in program_streams]))
        self.packets = [(p[2], p[3]) for p in parts]
        # device_vlc: the video elementary streams stay in HBM and their slices are parsed there (VideoBatch, resident form)
        self.video = VideoBatch(ctx, [p[0] for p in parts], threads=threads, device_vlc=device_vlc, resident=device_vlc)
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
