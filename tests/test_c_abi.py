"""include/mpegb200.h as a plain C compiler sees it: tests/c_abi/driver.c is built with gcc -std=c11 -Wall -Wextra -Werror
against the headers and libmpegb200.so (struct layouts are _Static_asserts), runs its host-only checks here and one
tiny decode with hand-checkable values on the GPU box."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "c_abi" / "driver.c"
EXE = ROOT / "tests" / "c_abi" / "_build" / "driver"
LIBDIR = ROOT / "mpeg_b200"


def build():
    EXE.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(SRC),
                    "-o", str(EXE), f"-L{LIBDIR}", "-lmpegb200", f"-Wl,-rpath,{LIBDIR}"], check=True, capture_output=True, text=True)


def test_header_compiles_as_c11_and_host_entry_points_work():
    try:
        build()
    except subprocess.CalledProcessError as e:
        pytest.fail("gcc -std=c11 -Wall -Wextra -Werror -pedantic rejects the C-ABI:\n" + e.stderr)
    r = subprocess.run([str(EXE), "abi"], capture_output=True, text=True)
    assert r.returncode == 0 and "abi ok" in r.stdout, r.stderr


@pytest.mark.gpu
def test_tiny_decode_from_plain_c():
    build()
    r = subprocess.run([str(EXE), "gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "gpu ok" in r.stdout, r.stderr
