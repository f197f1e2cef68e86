"""TEST INFRASTRUCTURE: the device-side slice walker (mpeg_b200/csrc/vlc_slice_walk.h) compiled for the CPU
(tests/vlc_emu/vlc_emu.cpp) so that the no-GPU suite can hold it against the host parser.  Not used by the product."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

import oracle_lib as ol
from mpeg_b200 import _lib
from mpeg_b200.batch import BatchScanStep
from mpeg_b200.mpeg import VideoStep

HERE = Path(__file__).resolve().parent
SRC = HERE / "vlc_emu" / "vlc_emu.cpp"
OUT = HERE / "vlc_emu" / "_build" / "libvlc_emu.so"
_emu = None


def emu():
    global _emu
    if _emu is None:
        deps = [SRC, HERE.parent / "mpeg_b200" / "csrc" / "vlc_slice_walk.h", HERE.parent / "mpeg_b200" / "csrc" / "vlc_device_tables.h",
                HERE.parent / "include" / "mpegb200.h"]
        if not OUT.exists() or OUT.stat().st_mtime < max(d.stat().st_mtime for d in deps):
            OUT.parent.mkdir(parents=True, exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-o", str(OUT), str(SRC)], check=True)
        _emu = C.CDLL(str(OUT))
        _emu.vlc_emu_wave.restype = C.c_int
    return _emu


def tables():
    L = _lib.load()
    L.mpegb200_internal_vlc_tables.restype = C.c_size_t
    L.mpegb200_internal_vlc_tables.argtypes = [C.c_void_p, C.c_size_t]
    buf = C.create_string_buffer(1 << 17)
    n = L.mpegb200_internal_vlc_tables(buf, len(buf))
    assert n > 0, "the host tables do not have the shape the device walker assumes"
    return buf


def start_code_positions(data: bytes) -> np.ndarray:
    """Byte offsets of every 00 00 01 of `data`, ascending (what mpegb200_video_stream_index finds on the device)."""
    out, at = [], data.find(b"\x00\x00\x01")
    while at >= 0:
        out.append(at)
        at = data.find(b"\x00\x00\x01", at + 1)
    return np.array(out, np.uint64)


def emulate_wave(wave, mb_w, mb_h, tab=None, resident_bytes: bytes = None):
    """Run the walker over one mpegb200_vlc_wave.  Returns (mbs [n_mb_slots], coeffs [6 * n_mb_slots, 64], flags [n_pictures]).
    resident_bytes: the (single) stream a wave of a resident batch reads from (its bitstream pointer is NULL)."""
    tab = tab or tables()
    n = wave.n_pictures
    bits, n_bits = C.c_void_p(wave.bitstream), wave.bitstream_bytes
    if resident_bytes is not None:
        assert not wave.bitstream
        keep = C.create_string_buffer(resident_bytes, len(resident_bytes) + 64)
        bits, n_bits = C.cast(keep, C.c_void_p), len(resident_bytes)
    mbs = np.zeros(max(1, wave.n_mb_slots), ol.MB_DTYPE)
    coeffs = np.full((max(1, 6 * wave.n_mb_slots), 64), 0x5555, np.int16)     # blocks the walker does not write keep the pattern
    flags = np.zeros(max(1, n), np.int32)
    w = (C.c_int * max(1, n))(*([mb_w] * n if np.isscalar(mb_w) else mb_w))
    h = (C.c_int * max(1, n))(*([mb_h] * n if np.isscalar(mb_h) else mb_h))
    rc = emu().vlc_emu_wave(tab, C.c_int(n), wave.pics, C.c_size_t(wave.n_slices), wave.slices, bits,
                            C.c_size_t(n_bits), C.c_void_p(wave.quant), C.c_size_t(wave.n_quant), C.c_size_t(wave.n_mb_slots),
                            w, h, C.c_void_p(mbs.ctypes.data), C.c_void_p(coeffs.ctypes.data), C.c_void_p(flags.ctypes.data))
    assert rc == 0, f"emulator refused the wave: {rc}"
    return mbs[:wave.n_mb_slots], coeffs[:6 * wave.n_mb_slots], flags[:n]


def picture_records(mbs, coeffs, pic):
    """The non-null records of picture `pic` in slot order with their blocks gathered: (mbs with coeff_block renumbered from 0, blocks)."""
    sel = mbs[(mbs["pic"] == pic.index) if hasattr(pic, "index") else slice(pic.mb_slot, pic.mb_slot + pic.n_mb_slots)]
    sel = sel[sel["pic"] != 0xffff].copy()
    blocks = []
    at = 0
    for m in sel:
        nc = bin(int(m["cbp"])).count("1")
        blocks.append(coeffs[int(m["coeff_block"]):int(m["coeff_block"]) + nc])
        m["coeff_block"] = at
        at += nc
    return sel, (np.concatenate(blocks) if blocks else np.zeros((0, 64), np.int16))


def host_records(step: VideoStep):
    """All launches of a host-parsed step as (header, mbs, coeffs) like test_host_parser.parser_steps."""
    out = []
    for i in range(step.n_launches):
        ln = step.launches[i]
        mbs = np.frombuffer(C.string_at(step.mbs + 16 * ln.first_mb, 16 * ln.n_mb), dtype=ol.MB_DTYPE).copy() if ln.n_mb else np.zeros(0, ol.MB_DTYPE)
        co = (np.frombuffer(C.string_at(step.coeffs + 128 * ln.first_block, 128 * ln.n_blocks), dtype=np.int16).reshape(-1, 64).copy()
              if ln.n_blocks else np.zeros((0, 64), np.int16))
        out.append(((ln.type, ln.dst_buf, ln.fwd_buf, ln.bwd_buf, ln.n_mb), mbs, co))
    return out


class ScanBatch:
    """mpegb200_video_batch_* in scan mode over a list of elementary streams (host side only)."""

    def __init__(self, datas, threads=2, resident=False):
        self.L = _lib.load()
        self.datas = [bytes(d) for d in datas]
        self.h = self.L.mpegb200_video_batch_new(len(self.datas), threads, None, None)
        assert self.h
        self.resident = resident
        self.sizes = []
        w, h = C.c_int(), C.c_int()
        for i, d in enumerate(self.datas):
            assert self.L.mpegb200_video_batch_set_stream(self.h, i, d, len(d)) == 0
            self.L.mpegb200_video_batch_stream_size(self.h, i, C.byref(w), C.byref(h))
            self.sizes.append(((w.value + 15) >> 4, (h.value + 15) >> 4))
        if resident:   # as if the streams sat in device memory with their start codes indexed there
            assert self.L.mpegb200_video_batch_set_resident(self.h, 1) == 0
            for i, d in enumerate(self.datas):
                pos = start_code_positions(d)
                assert self.L.mpegb200_video_batch_set_start_codes(self.h, i, C.c_void_p(pos.ctypes.data), len(pos)) == 0

    def next(self):
        st = BatchScanStep()
        assert self.L.mpegb200_video_batch_next_scan(self.h, C.byref(st)) == 0
        return st

    def unscan(self):
        assert self.L.mpegb200_video_batch_unscan(self.h) == 0

    def redo(self, index, step_picture):
        """The tail of stream `index`'s step from `step_picture` on, parsed by the host: (has_frame, frame_buf, time, launches)."""
        st = VideoStep()
        assert self.L.mpegb200_video_batch_redo(self.h, index, step_picture, C.byref(st)) == 0
        return st.has_frame, st.frame_buf, st.time, host_records(st)

    @staticmethod
    def host_steps(st):
        """{stream index: launches} of the streams whose step the host parsed itself."""
        steps = C.cast(st.host_steps, C.POINTER(VideoStep))
        return {st.host_index[j]: host_records(steps[j]) for j in range(st.n_host)}

    def close(self):
        if self.h:
            self.L.mpegb200_video_batch_free(self.h)
            self.h = None
