"""TEST / BENCH INFRASTRUCTURE: an MPEG-1 video elementary-stream WRITER for synthetic pictures.

The inverse of the reference's parse half (video.go:270-331 sequence header, :374-434 picture, :436-460 slice, :462-562
macroblock, :564-606 motion vectors, :639-746 block), so that the host parser can be fed pictures of any size and
density -- in particular 720p pictures shaped like BASELINE configs[2], for which no clip exists.  It works in the
QUANTISED domain: a picture is described by macroblock types, motion vectors, coded block patterns and quantised levels
in zig-zag order; `expected_records` computes, independently of any parser, the packed records (include/mpegb200.h) that
decoding the stream must produce (dequantise, oddify, clip: video.go:719-741), and tests/test_mpeg1_writer.py checks the
product parser and the oracle's parser against that.

VLC code words come from oracle/vlc_codes.inc (ISO/IEC 11172-2 Annex B, machine-derived from the reference's tables).
"""
import re
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PIC_I, PIC_P, PIC_B = 1, 2, 3
MB_INTRA, MB_PREDICT, MB_REF_BWD = 1, 2, 4

INTRA_QUANT = np.array([8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26, 27, 29, 34, 37, 40,
                        22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83],
                       dtype=np.int64)   # ISO 11172-2 default intra matrix, natural order (video.go:1055-1064)


def _zigzag():
    out, r, c, up = [], 0, 0, True
    for _ in range(64):
        out.append(r * 8 + c)
        if up:
            if c == 7:
                r += 1; up = False
            elif r == 0:
                c += 1; up = False
            else:
                r -= 1; c += 1
        else:
            if r == 7:
                c += 1; up = True
            elif c == 0:
                r += 1; up = True
            else:
                r += 1; c -= 1
    return np.array(out, dtype=np.int64)


ZIGZAG = _zigzag()
_TABLES = None


def tables():
    """{table name: {value: code word}} parsed from oracle/vlc_codes.inc."""
    global _TABLES
    if _TABLES is None:
        text = (ROOT / "oracle" / "vlc_codes.inc").read_text()
        _TABLES = {}
        for name, body in re.findall(r"static const vlc_code (\w+)\[\] = \{(.*?)\n\};", text, re.S):
            t = {}
            for bits, value in re.findall(r'\{"([01]+)",\s*([^}]+)\}', body):
                value = value.strip()
                if value == "VLC_INVALID":
                    continue
                t[int(value, 0)] = bits
            _TABLES[name] = t
    return _TABLES


@dataclass
class Macroblock:
    """One macroblock in the quantised domain.  `blocks[k]` is None (not coded) or an int array of 64 quantised levels in
    zig-zag order; for an intra macroblock blocks[k][0] is ignored and `dc[k]` (0..255) is the reconstructed DC."""
    intra: bool = False
    fwd: tuple = None          # (h, v) in half-pels, or None
    bwd: tuple = None
    blocks: list = field(default_factory=lambda: [None] * 6)
    dc: list = field(default_factory=lambda: [128] * 6)
    skipped: bool = False      # P pictures only: not transmitted (prediction with a zero vector)


class BitWriter:
    def __init__(self):
        self.parts = []

    def put(self, bits: str):
        self.parts.append(bits)

    def put_uint(self, value: int, n: int):
        if n:
            self.parts.append(format(value & ((1 << n) - 1), "0%db" % n))

    def tobytes(self) -> bytes:
        s = "".join(self.parts)
        s += "0" * (-len(s) % 8)
        return int(s, 2).to_bytes(len(s) // 8, "big") if s else b""


class StreamWriter:
    """Writes a sequence header and pictures; one slice per macroblock row, a fixed quantiser scale per stream."""

    def __init__(self, width: int, height: int, quantizer_scale: int = 8, f_code: int = 2, frame_rate_code: int = 5):
        assert 1 <= quantizer_scale <= 31 and 1 <= f_code <= 7
        self.width, self.height, self.q, self.f_code = width, height, quantizer_scale, f_code
        self.mb_w, self.mb_h = (width + 15) >> 4, (height + 15) >> 4
        assert self.mb_h <= 175, "one slice per macroblock row needs at most 175 rows"
        self.chunks = []
        bw = BitWriter()
        bw.put_uint(width, 12); bw.put_uint(height, 12); bw.put_uint(1, 4); bw.put_uint(frame_rate_code, 4)
        bw.put_uint(0x3ffff, 18); bw.put_uint(1, 1); bw.put_uint(20, 10); bw.put_uint(0, 1)
        bw.put_uint(0, 1); bw.put_uint(0, 1)       # default quantiser matrices
        self.chunks.append(b"\x00\x00\x01\xb3" + bw.tobytes())
        self.temporal = 0

    # ---- motion vectors (video.go:583-606, inverse)
    def _mv_delta(self, bw, target: int, pred: int):
        t = tables()["VLC_MOTION"]
        r_size = self.f_code - 1
        f = 1 << r_size
        d = target - pred
        if d > 16 * f - 1:
            d -= 32 * f
        elif d < -16 * f:
            d += 32 * f
        assert -16 * f <= d <= 16 * f - 1, "vector outside the range of f_code"
        if d == 0:
            bw.put(t[0])
            return
        a = abs(d) - 1
        code = (a >> r_size) + 1
        bw.put(t[code if d > 0 else -code])
        if f != 1:
            bw.put_uint(a & (f - 1), r_size)

    def _block(self, bw, mb: Macroblock, k: int, dc_pred: list):
        tb = tables()
        coef = tb["VLC_DCT_COEFF"]
        q = mb.blocks[k]
        start = 0
        if mb.intra:
            plane = k - 3 if k > 3 else 0
            diff = int(mb.dc[k]) - dc_pred[plane]
            dc_pred[plane] = int(mb.dc[k])
            size = abs(diff).bit_length()
            bw.put(tb["VLC_DC_SIZE_LUMA" if plane == 0 else "VLC_DC_SIZE_CHROMA"][size])
            if size:
                bw.put_uint(diff if diff > 0 else diff + (1 << size) - 1, size)
            start = 1
        run, first = 0, not mb.intra
        for p in range(start, 64):
            lv = int(q[p])
            if lv == 0:
                run += 1
                continue
            key = (run << 8) | abs(lv)
            if abs(lv) <= 255 and key in coef and key != 0x0001:
                bw.put(coef[key])
                bw.put("1" if lv < 0 else "0")
            elif key == 0x0001:     # run 0, level +-1: "1s" as a block's first coefficient, "11s" afterwards (video.go:686-691)
                bw.put("1" if first else "11")
                bw.put("1" if lv < 0 else "0")
            else:                   # escape: 6 bits of run, 8 (or 8 + 8) bits of level (video.go:693-708)
                assert -255 <= lv <= 255
                bw.put(coef[0xffff])
                bw.put_uint(run, 6)
                if -127 <= lv <= 127:
                    bw.put_uint(lv, 8)
                elif lv > 0:
                    bw.put_uint(0, 8); bw.put_uint(lv, 8)
                else:
                    bw.put_uint(128, 8); bw.put_uint(lv + 256, 8)
            run, first = 0, False
        assert not first, "a coded non-intra block needs at least one coefficient"
        bw.put("10")                # end of block

    def picture(self, pic_type: int, mbs: list):
        """mbs: mb_w * mb_h Macroblocks in raster order."""
        assert len(mbs) == self.mb_w * self.mb_h
        tb = tables()
        bw = BitWriter()
        bw.put_uint(self.temporal & 1023, 10); bw.put_uint(pic_type, 3); bw.put_uint(0xffff, 16)
        self.temporal += 1
        if pic_type in (PIC_P, PIC_B):
            bw.put_uint(0, 1); bw.put_uint(self.f_code, 3)      # half-pel vectors
        if pic_type == PIC_B:
            bw.put_uint(0, 1); bw.put_uint(self.f_code, 3)
        bw.put_uint(0, 1)                                        # no extra information
        self.chunks.append(b"\x00\x00\x01\x00" + bw.tobytes())
        type_table = tb["VLC_MB_TYPE_I" if pic_type == PIC_I else "VLC_MB_TYPE_P" if pic_type == PIC_P else "VLC_MB_TYPE_B"]
        for row in range(self.mb_h):
            bw = BitWriter()
            bw.put_uint(self.q, 5); bw.put_uint(0, 1)
            dc_pred = [128, 128, 128]
            pf, pb = [0, 0], [0, 0]          # vector predictors (reset at the slice start, video.go:443-446)
            inc = 1
            for col in range(self.mb_w):
                mb = mbs[row * self.mb_w + col]
                if mb.skipped:
                    assert pic_type == PIC_P and 0 < col < self.mb_w - 1, "skipped macroblocks: P pictures, not first / last of a slice"
                    inc += 1
                    continue
                while inc > 33:
                    bw.put(tb["VLC_MB_ADDR_INC"][35]); inc -= 33
                bw.put(tb["VLC_MB_ADDR_INC"][inc])
                if inc > 1:                   # video.go:496-501
                    dc_pred = [128, 128, 128]
                    if pic_type == PIC_P:
                        pf = [0, 0]
                inc = 1
                cbp = sum((0x20 >> k) for k in range(6) if mb.blocks[k] is not None)
                if mb.intra:
                    assert cbp == 0x3f
                    mtype = 0x01
                else:
                    mtype = (0x08 if mb.fwd is not None else 0) | (0x04 if mb.bwd is not None else 0) | (0x02 if cbp else 0)
                    if pic_type == PIC_P and mtype == 0:
                        raise ValueError("a P macroblock without vector and without residual must be skipped")
                    if pic_type == PIC_B:
                        assert mtype & 0x0c, "B macroblocks carry at least one vector"
                bw.put(type_table[mtype])
                if mb.intra:
                    pf, pb = [0, 0], [0, 0]   # video.go:525-529
                else:
                    dc_pred = [128, 128, 128]
                    if mb.fwd is not None:
                        self._mv_delta(bw, mb.fwd[0], pf[0]); self._mv_delta(bw, mb.fwd[1], pf[1])
                        pf = list(mb.fwd)
                    elif pic_type == PIC_P:
                        pf = [0, 0]
                    if mb.bwd is not None:
                        self._mv_delta(bw, mb.bwd[0], pb[0]); self._mv_delta(bw, mb.bwd[1], pb[1])
                        pb = list(mb.bwd)
                if not mb.intra and cbp:
                    bw.put(tb["VLC_CBP"][cbp])
                for k in range(6):
                    if mb.blocks[k] is not None:
                        self._block(bw, mb, k, dc_pred)
            self.chunks.append(bytes([0, 0, 1, row + 1]) + bw.tobytes())

    def tobytes(self) -> bytes:
        return b"".join(self.chunks)


# ------------------------------------------------------------------------------------------------
# what decoding must produce
# ------------------------------------------------------------------------------------------------
def dequantise(q_zigzag, intra: bool, scale: int):
    """Quantised levels in zig-zag order -> int16[64] levels in natural order as video.go:719-741 leaves them (before the
    premultiply): doubled, +-1 for non-intra, times scale and matrix, >> 4, made odd toward zero, clipped."""
    out = np.zeros(64, np.int64)
    for p in np.flatnonzero(q_zigzag):
        dz = int(ZIGZAG[p])
        lv = 2 * int(q_zigzag[p])
        if not intra:
            lv += -1 if lv < 0 else 1
        lv = (lv * scale * (int(INTRA_QUANT[dz]) if intra else 16)) >> 4
        if (lv & 1) == 0:
            lv -= 1 if lv > 0 else -1
        out[dz] = min(2047, max(-2048, lv))
    return out


def expected_records(mbs, mb_w: int, pic_type: int, scale: int):
    """(mb records as tuples (row, col, mv_h, mv_v, flags, cbp), coefficient blocks int16[n][64]) of one picture."""
    recs, blocks = [], []
    for i, mb in enumerate(mbs):
        row, col = divmod(i, mb_w)
        if mb.skipped:
            recs.append((row, col, 0, 0, MB_PREDICT, 0))
            continue
        cbp = sum((0x20 >> k) for k in range(6) if mb.blocks[k] is not None)
        if mb.intra:
            flags, mv = MB_INTRA, (0, 0)
        elif pic_type == PIC_B and (mb.fwd is None or mb.bwd is not None):
            flags, mv = MB_PREDICT | MB_REF_BWD, mb.bwd      # the backward copy is the one that stays (video.go:626-630)
        else:
            flags, mv = MB_PREDICT, mb.fwd if mb.fwd is not None else (0, 0)
        recs.append((row, col, mv[0], mv[1], flags, cbp))
        for k in range(6):
            if mb.blocks[k] is None:
                continue
            q = np.array(mb.blocks[k], dtype=np.int64)
            if mb.intra:
                q = q.copy(); q[0] = 0
            lv = dequantise(q, mb.intra, scale)
            if mb.intra:
                lv[0] = int(mb.dc[k]) * 8
            blocks.append(lv.astype(np.int16))
    return recs, (np.stack(blocks) if blocks else np.zeros((0, 64), np.int16))


# ------------------------------------------------------------------------------------------------
# synthetic pictures (the distributions of tests/workload.py, in the quantised domain)
# ------------------------------------------------------------------------------------------------
def random_picture(rng, mb_w: int, mb_h: int, pic_type: int, mode: str = "natural", mv_range: int = 32, luma_w=None, luma_h=None):
    """mode 'dense': every macroblock predicted (intra in an I picture) with six blocks of 64 coefficients (BASELINE configs[2]);
    'natural': 10 % intra, cbp ~ U{1..63} (or none), n ~ Geom coefficients, some skipped macroblocks in P pictures."""
    luma_w = luma_w or mb_w * 16
    luma_h = luma_h or mb_h * 16
    mbs = []
    for i in range(mb_w * mb_h):
        row, col = divmod(i, mb_w)
        intra = pic_type == PIC_I or (mode != "dense" and rng.random() < 0.1)
        mb = Macroblock(intra=intra)
        if not intra:
            def vec():
                while True:
                    h, v = int(rng.integers(-mv_range, mv_range)), int(rng.integers(-mv_range, mv_range))
                    x0, y0 = col * 16 + (h >> 1), row * 16 + (v >> 1)
                    if x0 >= 0 and y0 >= 0 and x0 + 16 + (h & 1) <= luma_w and y0 + 16 + (v & 1) <= luma_h:
                        return (h, v)
            if pic_type == PIC_P:
                if mode != "dense" and 0 < col < mb_w - 1 and rng.random() < 0.08:
                    mb.skipped = True
                    mbs.append(mb)
                    continue
                mb.fwd = vec()
            else:
                kind = int(rng.integers(0, 3))
                if kind != 1:
                    mb.fwd = vec()
                if kind != 0:
                    mb.bwd = vec()
        cbp = 0x3f if (intra or mode == "dense") else int(rng.integers(0, 64))
        for k in range(6):
            if not cbp & (0x20 >> k):
                continue
            n = 64 if mode == "dense" else int(min(64, rng.geometric(0.15)))
            q = np.zeros(64, np.int64)
            mag = np.minimum(rng.geometric(0.45, n), 40)            # mostly 1..3, now and then an escape
            q[:n] = mag * rng.choice([-1, 1], n)
            if mode != "dense" and n > 2:
                q[rng.integers(1, n, max(1, n // 4))] = 0           # runs of zeros
            if not intra and not q.any():
                q[0] = 1
            if rng.random() < 0.02:
                q[int(rng.integers(0 if not intra else 1, 64))] = int(rng.choice([-255, -128, 128, 255, 127, -127]))
            mb.blocks[k] = q
            if intra:
                mb.dc[k] = int(rng.integers(0, 256))
        mbs.append(mb)
    return mbs
