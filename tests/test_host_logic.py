"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host-side
workload generator and buffer-rotation mirror behave, the int32 bound of the IDCT holds."""
import ctypes as C

import numpy as np
from pathlib import Path
import pytest

import oracle_lib as ol


def test_library_exports_every_declared_symbol():
    from mpeg_b200 import _lib
    L = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mpegb200.h but not exported"
    assert set(L._signatures) == set(names)
    assert L.mpegb200_abi_version() == 1


def test_no_silent_cpu_fallback_without_a_gpu():
    import torch
    import mpeg_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mpeg_b200.MpegB200Error):
        mpeg_b200.Context()


def test_product_does_not_import_the_oracle():
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for p in list((root / "mpeg_b200").rglob("*.py")) + list((root / "mpeg_b200" / "csrc").glob("*.c*")):
        text = p.read_text()
        for needle in ("oracle_lib", "liboracle", "import oracle", "from oracle", "oracle/", "mpeg_oracle", "orc_"):
            assert needle not in text, f"{p} reaches into the oracle ({needle})"


def test_record_struct_layouts_match_the_header():
    from mpeg_b200 import context as cx
    assert cx.MB_DTYPE.itemsize == 16 and cx.PICTURE_DTYPE.itemsize == 16
    assert cx.MB_DTYPE.fields["coeff_block"][1] == 12 and cx.MB_DTYPE.fields["pic"][1] == 10
    assert cx.PICTURE_DTYPE.fields["first_mb"][1] == 8
    assert cx.MB_DTYPE == ol.MB_DTYPE and cx.PICTURE_DTYPE == ol.PIC_DTYPE


def test_zigzag_and_geometry():
    import workload as wl
    assert wl.ZIGZAG[:10].tolist() == [0, 1, 8, 16, 9, 2, 3, 10, 17, 24]  # video.go:1044-1046
    assert sorted(wl.ZIGZAG.tolist()) == list(range(64))
    g = wl.HD720
    assert (g.mb_w, g.mb_h, g.n_mb) == (80, 45, 3600)
    assert g.picture_bytes == 1382400 and g.frame_bytes == 1382400 + 20480
    assert wl.CIF.n_mb == 396 and wl.CIF.picture_bytes == 152064


def test_workload_generator_obeys_the_packing_rules():
    import workload as wl
    g = wl.CIF
    rng = wl.stream_rng(2, 0)
    for t, mode in [(wl.PIC_I, "natural"), (wl.PIC_P, "natural"), (wl.PIC_B, "natural"), (wl.PIC_P, "dense")]:
        mbs, coeffs = wl.make_picture(rng, g, t, mode)
        cnt = np.array([bin(int(c)).count("1") for c in mbs["cbp"]])
        assert np.array_equal(mbs["coeff_block"], np.cumsum(cnt) - cnt) and cnt.sum() == len(coeffs)
        intra = (mbs["flags"] & wl.MB_INTRA) != 0
        assert ((mbs["flags"] & wl.MB_PREDICT) != 0).sum() + intra.sum() == len(mbs)
        assert (mbs["cbp"][intra] == 63).all()
        assert np.abs(coeffs.astype(int)).max() <= 2047 * 8
        if t == wl.PIC_I:
            assert intra.all()
        if mode == "dense":
            assert (mbs["cbp"] == 63).all() and (coeffs != 0).mean() > 0.99
            total, read = wl.algorithmic_bytes(mbs, len(coeffs))
            assert total == 1552 * g.n_mb and read == 1168 * g.n_mb   # 1552 B / dense-P macroblock
        # the oracle accepts every generated record (windows inside the planes)
        fs = ol.FrameSet(1, g.width, g.height)
        pics, m, c = wl.batch_pictures([(mbs, coeffs)], [0], t, [(0, 1, 2)])
        assert fs.exec_pictures(pics, m, c) == 0


def test_buffer_rotation_mirrors_the_decoder(golden_dir):
    """The host-side rotation mirror (video.go:406-409, 430-433) reproduces the buffer roles the
    oracle's decoder reports for the reference clip."""
    import workload as wl
    v = ol.VideoOracle((golden_dir / "test.mpeg1video").read_bytes(), tap=True)
    rot = wl.BufferRotation()
    seen = 0
    while v.decode() is not None and seen < 60:
        pics, _, _ = v.tap()
        for p in pics:
            dst, fwd, bwd = rot.begin(int(p["type"]))
            assert (dst, fwd, bwd) == (int(p["dst_buf"]), int(p["fwd_buf"]), int(p["bwd_buf"]))
            rot.end(int(p["type"]))
            seen += 1
    assert seen >= 60


def test_idct_intermediates_fit_int32():
    """propagate |.| bounds through both passes of the transform for clipped levels
    (|level| <= 2048, intra DC <= 2047*8 in level form): every intermediate stays below 2^31, so
    the kernel's int32 arithmetic is exact.  Triangle-inequality bound, not sampling."""
    pm = np.array([32, 44, 42, 38, 32, 25, 17, 9, 44, 62, 58, 52, 44, 35, 24, 12, 42, 58, 55, 49, 42, 33, 23, 12,
                   38, 52, 49, 44, 38, 30, 20, 10, 32, 44, 42, 38, 32, 25, 17, 9, 25, 35, 33, 30, 25, 20, 14, 7,
                   17, 24, 23, 20, 17, 14, 9, 5, 9, 12, 12, 10, 9, 7, 5, 2], dtype=np.int64).reshape(8, 8)
    bound = 2048 * pm
    bound[0, 0] = max(bound[0, 0], 2047 * 8 * 32) + 128  # the kernel carries the final +128 on the DC term
    worst = 0

    def pass8(s):
        nonlocal worst
        s0, s1, s2, s3, s4, s5, s6, s7 = s
        b1, b3, b4 = s4, s2 + s6, s5 + s3
        t1, t2, b6 = s1 + s7, s3 + s5, s1 + s7
        b7 = t1 + t2
        p1 = b6 * 473 + b4 * 196 + 128
        x4 = (p1 >> 8) + 1 + b7
        p2 = (t1 + t2) * 362 + 128
        x0 = x4 + (p2 >> 8) + 1
        x1 = s0 + b1
        p3 = (s2 + s6) * 362 + 128
        x2 = (p3 >> 8) + 1 + b3
        x3 = s0 + b1
        y3, y4, y5, y6 = x1 + x2, x3 + b3, x1 + x2, x3 + b3
        p4 = b4 * 473 + b6 * 196 + 128
        y7 = x0 + (p4 >> 8) + 1
        outs = [b7 + y4, x4 + y3, y5 + x0, y6 + y7, y6 + y7, x0 + y5, y3 + x4, y4 + b7]
        worst = max(worst, p1, p2, p3, p4, *outs)
        return outs

    cols = [pass8([int(bound[r, c]) for r in range(8)]) for c in range(8)]
    after = np.array(cols, dtype=object).T  # after[r][c]
    for r in range(8):
        outs = pass8([int(after[r][c]) for c in range(8)])
        worst = max(worst, *[o + 128 + 255 * 256 for o in outs])  # + prediction << 8 (block_finish)
    assert worst < 2**31, worst
    assert worst > 1.5e9  # the bound is tight-ish: int32 has ~12 % headroom (1.897e9)


def test_bench_rank_binding_to_gpu_local_cpus(tmp_path, monkeypatch):
    """bench.py, multi-rank runs: a rank moves onto the CPUs sysfs lists as local to its GPU (so that its pinned buffers
    land on that NUMA node), and leaves its affinity alone when the list is missing, tiny, or covers everything."""
    import importlib.util
    import os
    import sys
    import types
    spec = importlib.util.spec_from_file_location("bench_under_test", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert bench.parse_cpulist("") == set()

    import torch
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    before = os.sched_getaffinity(0)
    msgs = []
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    try:
        bench.bind_to_gpu_cpus(0, msgs.append, sysfs_root=str(tmp_path))          # no local_cpulist: skipped
        assert os.sched_getaffinity(0) == before and "skipped" in msgs[-1]
        (dev / "local_cpulist").write_text(",".join(map(str, sorted(before))) + "\n")
        bench.bind_to_gpu_cpus(0, msgs.append, sysfs_root=str(tmp_path))          # everything is local: nothing to do
        assert os.sched_getaffinity(0) == before
        if len(before) >= 5:
            some = set(sorted(before)[:4])
            (dev / "local_cpulist").write_text(",".join(map(str, sorted(some))) + ",100000\n")
            bench.bind_to_gpu_cpus(0, msgs.append, sysfs_root=str(tmp_path))
            assert os.sched_getaffinity(0) == some and "bound to 4" in msgs[-1]
            os.sched_setaffinity(0, before)
            (dev / "local_cpulist").write_text(str(sorted(before)[0]) + "\n")     # a single CPU: left alone
            bench.bind_to_gpu_cpus(0, msgs.append, sysfs_root=str(tmp_path))
            assert os.sched_getaffinity(0) == before
    finally:
        os.sched_setaffinity(0, before)


def test_one_interpolation_path_gives_all_four_half_pel_modes():
    """The arithmetic of block_load's single interpolation path (video_fused_tma.cu, v5.3), lane for lane in Python: row sums
    a[x] + a[x + hx] and their vertical sums with row y + vy in 16-bit lanes, the missing neighbour replaced by the pixel itself
    through a mask, then (s + 2) >> 2 and the byte pack -- against the four modes of video_noasm.go:44-80 (copy, (a+b+1)>>1
    horizontally / vertically, (a+b+c+d+2)>>2), all byte alignments, saturated inputs included."""
    def funnel_rc(lo, hi, sh):
        return (((hi << 32) | lo) >> min(sh, 32)) & 0xffffffff

    def funnel_r(lo, hi, sh):
        return (((hi << 32) | lo) >> (sh & 31)) & 0xffffffff

    def byte_perm(a, b, sel):
        by = [(a >> (8 * i)) & 0xff for i in range(4)] + [(b >> (8 * i)) & 0xff for i in range(4)]
        return sum(by[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))

    def sel32(m, a, b):
        return (a & m) | (b & ~m & 0xffffffff)

    def hsum_row_sel(w0, w1, w2, sh, mh):
        u0, u1, s1 = funnel_rc(w0, w1, sh), funnel_rc(w1, w2, sh), funnel_rc(w1, w2, sh + 8)
        e0, o0, e1, o1 = u0 & 0x00ff00ff, byte_perm(u0, 0, 0x4341), u1 & 0x00ff00ff, byte_perm(u1, 0, 0x4341)
        m0, m1 = funnel_r(e0, e1, 16), byte_perm(s1, 0, 0x4341)
        return [(e0 + sel32(mh, o0, e0)) & 0xffffffff, (o0 + sel32(mh, m0, o0)) & 0xffffffff,
                (e1 + sel32(mh, o1, e1)) & 0xffffffff, (o1 + sel32(mh, m1, o1)) & 0xffffffff]

    def vsum_pack(ae, be, ao, bo):
        return byte_perm(((ae + be + 0x00020002) & 0xffffffff) >> 2, ((ao + bo + 0x00020002) & 0xffffffff) >> 2, 0x6240)

    rng = np.random.default_rng(53)
    for trial in range(400):
        win = rng.integers(0, 256, (9, 16), dtype=np.uint8)
        if trial < 8:
            win[:] = 255 if trial % 2 else 0
        words = win.view("<u4")
        for x0 in range(4):
            for mode in range(4):
                mh, mv = (0xffffffff if mode & 1 else 0), (0xffffffff if mode & 2 else 0)
                rows = [hsum_row_sel(int(words[r, 0]), int(words[r, 1]), int(words[r, 2]), 8 * x0, mh) for r in range(9)]
                for r in range(8):
                    up, dn = rows[r], [sel32(mv, rows[r + 1][i], rows[r][i]) for i in range(4)]
                    p0, p1 = vsum_pack(up[0], dn[0], up[1], dn[1]), vsum_pack(up[2], dn[2], up[3], dn[3])
                    got = [(p0 >> (8 * i)) & 0xff for i in range(4)] + [(p1 >> (8 * i)) & 0xff for i in range(4)]
                    a = win[r, x0:x0 + 8].astype(int)
                    b, c, d = win[r, x0 + 1:x0 + 9].astype(int), win[r + 1, x0:x0 + 8].astype(int), win[r + 1, x0 + 1:x0 + 9].astype(int)
                    want = [a, (a + b + 1) >> 1, (a + c + 1) >> 1, (a + b + c + d + 2) >> 2][mode]
                    assert got == list(want), (trial, x0, mode, r)
