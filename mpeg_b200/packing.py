"""Host-side packing rules that go with the record format of include/mpegb200.h.

The kernels process the macroblock records of a launch in parallel, so a launch must not
write the same macroblock of the same picture twice.  The reference decodes serially and
real streams do revisit macroblocks (a slice start code whose vertical position overlaps
macroblocks an earlier, multi-row slice already covered -- the reference's own test clip
does it, mpeg_test.go:203-231 hashes the result).  Serial semantics: the later record wins.
`resolve_rewrites` turns one picture batch into one or more launches ("waves") that
reproduce exactly that.
"""
import numpy as np

from .context import MB_DTYPE, MB_INTRA, MB_PREDICT

_POPC = np.array([bin(i).count("1") for i in range(64)], dtype=np.int64)


def _compact(pics, mbs, coeffs, keep):
    """Records `keep` (indices into mbs, ascending) with their coefficient blocks re-packed."""
    m = mbs[keep].copy()
    cnt = _POPC[m["cbp"]]
    starts = mbs["coeff_block"][keep].astype(np.int64)
    if len(keep):
        idx = np.concatenate([np.arange(s, s + c) for s, c in zip(starts, cnt)]) if cnt.sum() else np.zeros(0, np.int64)
    else:
        idx = np.zeros(0, np.int64)
    c = coeffs[idx] if len(idx) else np.zeros((0, 64), np.int16)
    m["coeff_block"] = np.cumsum(cnt) - cnt
    p = pics.copy()
    for i in range(len(p)):
        sel = np.flatnonzero(m["pic"] == i)
        p["first_mb"][i] = sel[0] if len(sel) else 0
        p["n_mb"][i] = len(sel)
    return p, m, c


def resolve_rewrites(pics, mbs, coeffs):
    """Split a batch into launches without double writes.  Returns a list of (pics, mbs, coeffs).

    A later record that defines all six blocks of its macroblock (predicted, or intra with cbp 63)
    simply replaces the earlier one.  A later record that defines only some blocks (an intra
    macroblock with an aborted block) must see the earlier result underneath, so the
    batch is cut there and the remainder goes into the next launch."""
    mbs = np.ascontiguousarray(mbs, dtype=MB_DTYPE)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.int16).reshape(-1, 64)
    key = (mbs["pic"].astype(np.int64) << 32) | (mbs["mb_row"].astype(np.int64) << 16) | mbs["mb_col"].astype(np.int64)
    if len(np.unique(key)) == len(key):
        return [(pics, mbs, coeffs)]
    waves, start = [], 0
    while start < len(mbs):
        last = {}
        dead = set()
        end = len(mbs)
        for i in range(start, len(mbs)):
            k = int(key[i])
            if k in last:
                complete = bool(mbs["flags"][i] & MB_PREDICT) or (bool(mbs["flags"][i] & MB_INTRA) and mbs["cbp"][i] == 63)
                if not complete:
                    end = i
                    break
                dead.add(last[k])
            last[k] = i
        keep = np.array([i for i in range(start, end) if i not in dead], dtype=np.int64)
        waves.append(_compact(pics, mbs, coeffs, keep))
        start = end
    return waves
