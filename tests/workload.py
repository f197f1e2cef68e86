"""TEST / BENCH INFRASTRUCTURE (not part of the product package).

Synthetic pre-parsed block batches (BASELINE.json configs 2-5).

Pure host-side numpy: produces the packed records of include/mpegb200.h that a bitstream
parser would produce, with the distributions SURVEY 8d fixes.  Used by bench.py and by the
parity tests (which feed the same arrays to the CUDA path and to the CPU oracle).
Seeds: numpy PCG64, seed = 20260925 + 1000*config + stream_id.
"""
from dataclasses import dataclass

import numpy as np

from mpeg_b200.context import MB_DTYPE, MB_INTRA, MB_PREDICT, MB_REF_BWD, PIC_B, PIC_I, PIC_P, PICTURE_DTYPE

BASE_SEED = 20260925

# zig-zag scan order (video.go:1044-1053), generated
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


@dataclass(frozen=True)
class Geometry:
    width: int
    height: int

    @property
    def mb_w(self):
        return (self.width + 15) >> 4

    @property
    def mb_h(self):
        return (self.height + 15) >> 4

    @property
    def luma_w(self):
        return self.mb_w << 4

    @property
    def luma_h(self):
        return self.mb_h << 4

    @property
    def chroma_w(self):
        return self.mb_w << 3

    @property
    def chroma_h(self):
        return self.mb_h << 3

    @property
    def n_mb(self):
        return self.mb_w * self.mb_h

    @property
    def luma_bytes(self):
        return self.luma_w * self.luma_h

    @property
    def chroma_bytes(self):
        return self.chroma_w * self.chroma_h

    @property
    def frame_bytes(self):  # video.go:336-340
        return self.luma_bytes + 2 * self.chroma_bytes + self.luma_w * 16

    @property
    def picture_bytes(self):
        return self.luma_bytes + 2 * self.chroma_bytes


CIF = Geometry(352, 288)
HD720 = Geometry(1280, 720)


def stream_rng(config: int, stream_id: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(BASE_SEED + 1000 * config + stream_id))


def _draw_motion_vectors(rng, g: Geometry, rows, cols, mv_range: int = 32):
    """mvH, mvV ~ U{-mv_range..mv_range-1} half-pel (default 32: +-16 pixels), re-drawn until the 17x17 luma and 9x9
    chroma windows lie inside their planes."""
    n = len(rows)
    mvh = np.zeros(n, np.int64)
    mvv = np.zeros(n, np.int64)
    todo = np.ones(n, bool)
    while todo.any():
        k = int(todo.sum())
        h = rng.integers(-mv_range, mv_range, k)
        v = rng.integers(-mv_range, mv_range, k)
        r, c = rows[todo], cols[todo]
        x0 = c * 16 + (h >> 1)
        y0 = r * 16 + (v >> 1)
        ok = (x0 >= 0) & (x0 + 15 + (h & 1) <= g.luma_w - 1) & (y0 >= 0) & (y0 + 15 + (v & 1) <= g.luma_h - 1)
        ch = np.trunc(h / 2).astype(np.int64)  # toward zero, video_noasm.go:35-36
        cv = np.trunc(v / 2).astype(np.int64)
        cx0 = c * 8 + (ch >> 1)
        cy0 = r * 8 + (cv >> 1)
        ok &= (cx0 >= 0) & (cx0 + 7 + (ch & 1) <= g.chroma_w - 1) & (cy0 >= 0) & (cy0 + 7 + (cv & 1) <= g.chroma_h - 1)
        idx = np.flatnonzero(todo)
        mvh[idx[ok]] = h[ok]
        mvv[idx[ok]] = v[ok]
        todo[idx[ok]] = False
    return mvh, mvv


def _oddify(level):
    """video.go:730-736: even levels move one step toward zero; zero becomes +1."""
    even = (level & 1) == 0
    return np.where(even, np.where(level > 0, level - 1, level + 1), level)


def _draw_blocks(rng, n_blocks: int, intra_block, dense: bool):
    """Coefficient levels (natural order) for n_blocks coded blocks."""
    if n_blocks == 0:
        return np.zeros((0, 64), np.int16)
    if dense:
        n = np.full(n_blocks, 64)
    else:
        n = np.minimum(1 + rng.geometric(0.15, n_blocks) - 1, 64)  # n ~ 1 + Geom(0.15), capped
        n = np.maximum(n, 1)
    pos = np.arange(64)
    scale = 300.0 / (1.0 + pos)
    lv = np.rint(rng.laplace(0.0, 1.0, (n_blocks, 64)) * scale).astype(np.int64)
    lv = np.clip(lv, -2047, 2047)
    lv = _oddify(lv)
    lv = np.clip(lv, -2048, 2047)
    lv[pos[None, :] >= n[:, None]] = 0
    dc = rng.integers(0, 256, n_blocks) * 8  # intra DC travels as dc*8 (video.go:672)
    lv[:, 0] = np.where(intra_block, dc, lv[:, 0])
    out = np.zeros((n_blocks, 64), np.int16)
    out[:, ZIGZAG] = lv.astype(np.int16)  # zig-zag position p -> natural index ZIGZAG[p]
    return out


def make_picture(rng, g: Geometry, pic_type: int, mode: str = "natural", adversarial: bool = True, mv_range: int = 32):
    """Records of one picture of one stream: (mbs, coeffs) with pic = 0, coeff_block from 0.

    mode 'natural': intra w.p. 0.1 in P/B, cbp ~ U{0..63}, n ~ 1+Geom(0.15);
    mode 'dense'  : every macroblock predicted, cbp = 63, n = 64 (the headline dense-P step)."""
    n_mb = g.n_mb
    rows = np.repeat(np.arange(g.mb_h), g.mb_w)
    cols = np.tile(np.arange(g.mb_w), g.mb_h)
    if pic_type == PIC_I:
        intra = np.ones(n_mb, bool)
    elif mode == "dense":
        intra = np.zeros(n_mb, bool)
    else:
        intra = rng.random(n_mb) < 0.1
    flags = np.where(intra, MB_INTRA, MB_PREDICT).astype(np.uint8)
    if pic_type == PIC_B:
        # {fwd, bwd, both} w.p. 1/3 each; 'both' resolves to the backward reference (video.go:626-630)
        kind = rng.integers(0, 3, n_mb)
        flags = np.where(~intra & (kind != 0), flags | MB_REF_BWD, flags).astype(np.uint8)
    mvh, mvv = _draw_motion_vectors(rng, g, rows, cols, mv_range)
    mvh = np.where(intra, 0, mvh)
    mvv = np.where(intra, 0, mvv)
    if mode == "dense":
        cbp = np.full(n_mb, 63)
    else:
        cbp = rng.integers(0, 64, n_mb)
    cbp = np.where(intra, 63, cbp).astype(np.uint8)
    ncoded = np.array([bin(int(c)).count("1") for c in range(64)])[cbp]
    coeff_block = np.cumsum(ncoded) - ncoded
    n_blocks = int(ncoded.sum())
    coeffs = _draw_blocks(rng, n_blocks, np.repeat(intra, ncoded), dense=(mode == "dense"))
    if adversarial and n_blocks:
        # one all-+-2047 block per picture: the int32 headroom case
        b = int(rng.integers(0, n_blocks))
        sign = rng.integers(0, 2, 64) * 2 - 1
        coeffs[b] = (2047 * sign).astype(np.int16)
        if np.repeat(intra, ncoded)[b]:
            coeffs[b, 0] = 2047 * 8
    mbs = np.zeros(n_mb, MB_DTYPE)
    mbs["mb_row"], mbs["mb_col"] = rows, cols
    mbs["mv_h"], mbs["mv_v"] = mvh, mvv
    mbs["flags"], mbs["cbp"] = flags, cbp
    mbs["pic"] = 0
    mbs["coeff_block"] = coeff_block
    return mbs, coeffs


class BufferRotation:
    """Host mirror of the reference's frame-buffer rotation (video.go:406-409, 430-433)."""

    def __init__(self):
        self.cur, self.fwd, self.bwd = 0, 1, 2

    def begin(self, pic_type: int):
        """Returns (dst, fwd, bwd) physical buffer indices for a picture of this type."""
        self._temp = self.fwd
        if pic_type in (PIC_I, PIC_P):
            self.fwd = self.bwd
        return self.cur, self.fwd, self.bwd

    def end(self, pic_type: int):
        if pic_type in (PIC_I, PIC_P):
            self.bwd = self.cur
            self.cur = self._temp


def batch_pictures(per_stream, stream_ids, pic_type, bufs):
    """Concatenate per-stream (mbs, coeffs) of ONE picture step into a launch batch.

    per_stream: list of (mbs, coeffs); bufs: list of (dst, fwd, bwd) per stream."""
    pics = np.zeros(len(per_stream), PICTURE_DTYPE)
    all_mbs, all_coeffs = [], []
    mb_base = blk_base = 0
    for i, ((mbs, coeffs), sid, (dst, fwd, bwd)) in enumerate(zip(per_stream, stream_ids, bufs)):
        m = mbs.copy()
        m["pic"] = i
        m["coeff_block"] += blk_base
        pics[i] = (sid, pic_type, dst, fwd, bwd, mb_base, len(m))
        all_mbs.append(m)
        all_coeffs.append(coeffs)
        mb_base += len(m)
        blk_base += len(coeffs)
    return pics, np.concatenate(all_mbs), np.concatenate(all_coeffs)


def algorithmic_bytes(mbs, n_blocks: int):
    """per macroblock 16 B record + 384 B written (+ 384 B of reference if predicted)
    + 128 B per coded block.  Returns (total, read_only)."""
    n_mb = len(mbs)
    n_pred = int(((mbs["flags"] & MB_PREDICT) != 0).sum())
    read = 16 * n_mb + 384 * n_pred + 128 * n_blocks
    write = 384 * n_mb
    return read + write, read


def random_reference_frame(rng, g: Geometry) -> np.ndarray:
    """Whole frame buffer (Y|Cb|Cr|pad) with U{0..255} planes and a zero pad, like a decoded frame."""
    buf = np.zeros(g.frame_bytes, np.uint8)
    buf[: g.picture_bytes] = rng.integers(0, 256, g.picture_bytes, dtype=np.uint8)
    return buf


def audio_samples(rng, n_frames: int) -> np.ndarray:
    """Config 4: samples[frame][ch][36][32] ~ round(N(0, 8000)) clipped +-65536, sb >= 27 zero."""
    s = np.clip(np.rint(rng.normal(0.0, 8000.0, (n_frames, 2, 36, 32))), -65536, 65536).astype(np.int32)
    s[..., 27:] = 0
    return s
