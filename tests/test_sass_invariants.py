"""Static checks on the SASS of the built kernels (cuobjdump, no GPU needed).

1. Programmatic dependent launch: fused_tma_kernel starts while the plan pre-pass may still be running and must not read a
   plan before griddepcontrol.wait (SASS: ACQBULK).  The compiler once hoisted the plan-head loads (invariant
   LDG.E.CONSTANT through a const __restrict__ pointer) above the wait; the CTA then armed its mbarrier with a stale byte
   count and never woke up -- a hang that depended on timing.  No global load, TMA load or bulk copy may precede ACQBULK.
2. The unfused MP2 window must not contain fused multiply-adds (bit-exactness with the reference's Go / SSE path), the
   fused one must (the AVX2 / NEON path).
"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "mpeg_b200" / "csrc" / "_build"

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")


def kernels(obj):
    """{mangled function name: [opcode, ...]} of one object file."""
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = res.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            cur.append(m.group(1))
    return res


def test_no_plan_read_before_the_dependency_wait():
    obj = BUILD / "video_fused_tma.o"
    if not obj.exists():
        pytest.skip("library not built")
    ks = {k: v for k, v in kernels(obj).items() if "fused_tma_kernel" in k}
    assert ks, "fused_tma_kernel not found"
    for name, ops in ks.items():
        assert "ACQBULK" in ops, "griddepcontrol.wait is gone"
        before = ops[: ops.index("ACQBULK")]
        early = [o for o in before if o.startswith(("LDG", "LD.", "UTMALDG", "UBLKCP", "UTMAPF", "LDGSTS"))]
        assert not early, f"{name}: memory reads ahead of griddepcontrol.wait: {early}"


def test_window_arithmetic_of_the_two_audio_modes():
    obj = BUILD / "audio_kernels.o"
    if not obj.exists():
        pytest.skip("library not built")
    ks = {k: v for k, v in kernels(obj).items() if "audio_synth_kernel" in k}
    unfused = {k: v for k, v in ks.items() if "ILb0E" in k}
    fused = {k: v for k, v in ks.items() if "ILb1E" in k}
    assert len(unfused) == 4 and len(fused) == 4      # four output formats each
    for name, ops in unfused.items():
        # 16 unrolled slot bodies x 16 taps: 256 multiplies and 256 adds of their own, next to the 80 + 209 of the DCT;
        # fused multiply-adds only in the output scaling (exact division, a handful per body)
        n_fma = sum(o.startswith("FFMA") for o in ops)
        n_mul, n_add = sum(o.startswith("FMUL") for o in ops), sum(o.startswith("FADD") for o in ops)
        assert n_fma <= 16 * 10, f"{name}: {n_fma} fused multiply-adds (only the output scaling may use them)"
        assert n_mul >= 256 + 80 and n_add >= 256 + 200, f"{name}: {n_mul} FMUL, {n_add} FADD"
    for name, ops in fused.items():
        assert sum(o.startswith("FFMA") for o in ops) >= 256 + 16, name
