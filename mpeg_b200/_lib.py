"""Loader for libmpegb200.so, the C-ABI of include/mpegb200.h.

The shared library is built in tree (mpeg_b200/csrc/Makefile, nvcc -gencode arch=compute_100a,code=sm_100a)
and is the only compute path of this package: there is no CPU or PyTorch fallback.  If the library
is missing, or no sm_100 device is present when a context is created, the package fails loudly.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
# MPEGB200_LIB selects another build of the same library (A/B runs of kernel variants, tools/build_variants.sh)
LIB_PATH = Path(os.environ["MPEGB200_LIB"]).resolve() if os.environ.get("MPEGB200_LIB") else PKG / "libmpegb200.so"


class MpegB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"mpegb200 error {code}: {text}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile the CUDA extension for sm_100a (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(["make", "-C", str(PKG / "csrc")], env=env, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libmpegb200.so failed")
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the decode kernels)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32p, u8p, szp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
    ip = C.POINTER(C.c_int)
    sig = {
        "mpegb200_abi_version": (C.c_int, []),
        "mpegb200_create": (vp, [C.c_int, C.c_int, ip]),
        "mpegb200_destroy": (None, [vp]),
        "mpegb200_last_error": (C.c_char_p, [vp]),
        "mpegb200_set_stream": (C.c_int, [vp, vp]),
        "mpegb200_get_stream": (vp, [vp]),
        "mpegb200_sync": (C.c_int, [vp]),
        "mpegb200_sync_uploads": (C.c_int, [vp]),
        "mpegb200_join_readbacks": (C.c_int, [vp]),
        "mpegb200_launch_count": (C.c_uint64, [vp]),
        "mpegb200_set_validate": (C.c_int, [vp, C.c_int]),
        "mpegb200_set_kernel_timing": (C.c_int, [vp, C.c_int]),
        "mpegb200_kernel_times": (C.c_int, [vp, vp, vp, C.c_int]),
        "mpegb200_video_open": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "mpegb200_video_close": (C.c_int, [vp, C.c_int]),
        "mpegb200_video_geometry": (C.c_int, [vp, C.c_int, ip, ip, ip, ip, szp]),
        "mpegb200_video_validate": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t]),
        "mpegb200_video_decode_pictures": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp]),
        "mpegb200_video_decode_pictures_dev": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp]),
        "mpegb200_pack_coeffs12": (C.c_int, [vp, C.c_size_t, vp]),
        "mpegb200_video_decode_pictures_packed": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp]),
        "mpegb200_vlen_payload_bound": (C.c_size_t, [C.c_size_t]),
        "mpegb200_vlen_validate": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t]),
        "mpegb200_pack_coeffs_vlen": (C.c_int, [vp, C.c_size_t, vp, vp, vp, C.c_size_t, szp]),
        "mpegb200_video_decode_pictures_vlen": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, vp, vp, vp, C.c_size_t]),
        "mpegb200_video_decode_bitstream": (C.c_int, [vp, C.c_int, vp, C.c_size_t, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_size_t]),
        "mpegb200_video_stream_upload": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_size_t]),
        "mpegb200_video_stream_index": (C.c_int, [vp, C.c_int, vp, C.c_size_t, szp]),
        "mpegb200_video_bitstream_flags": (C.c_int, [vp, vp, C.c_int]),
        "mpegb200_video_bitstream_records": (C.c_int, [vp, vp, vp]),
        "mpegb200_video_bitstream_parse_ms": (C.c_int, [vp, C.POINTER(C.c_float)]),
        "mpegb200_video_read_planes": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
        "mpegb200_video_write_planes": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
        "mpegb200_video_read_frame": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
        "mpegb200_video_write_frame": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
        "mpegb200_video_frame_dev": (vp, [vp, C.c_int, C.c_int]),
        "mpegb200_video_read_pictures_host": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t]),
        "mpegb200_video_read_pictures_dev": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t]),
        "mpegb200_video_ring_new": (vp, [vp, C.c_int, vp, C.c_int]),
        "mpegb200_video_ring_free": (None, [vp]),
        "mpegb200_video_ring_push": (C.c_int, [vp, vp]),
        "mpegb200_video_ring_slot_dev": (vp, [vp, C.c_int, szp]),
        "mpegb200_video_ring_read_host": (C.c_int, [vp, C.c_int, vp, C.c_size_t]),
        "mpegb200_video_rgba": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "mpegb200_video_rgba_batch_dev": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t]),
        "mpegb200_audio_open": (C.c_int, [vp, C.c_int]),
        "mpegb200_audio_close": (C.c_int, [vp, C.c_int]),
        "mpegb200_audio_synth": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]),
        "mpegb200_audio_synth_dev": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]),
        "mpegb200_audio_synth_coded": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp]),
        "mpegb200_audio_read_state": (C.c_int, [vp, C.c_int, vp, ip]),
        "mpegb200_audio_write_state": (C.c_int, [vp, C.c_int, vp, C.c_int]),
        "mpegb200_host_alloc": (vp, [C.c_size_t]),
        "mpegb200_host_free": (None, [vp]),
        # host half (include/mpegb200_host.h)
        "mpegb200_video_parser_new": (vp, [C.c_char_p, C.c_size_t]),
        "mpegb200_video_parser_free": (None, [vp]),
        "mpegb200_video_parser_has_header": (C.c_int, [vp]),
        "mpegb200_video_parser_width": (C.c_int, [vp]),
        "mpegb200_video_parser_height": (C.c_int, [vp]),
        "mpegb200_video_parser_framerate": (C.c_double, [vp]),
        "mpegb200_video_parser_set_no_delay": (None, [vp, C.c_int]),
        "mpegb200_video_parser_set_vlen": (None, [vp, C.c_int]),
        "mpegb200_video_parser_rewind": (None, [vp]),
        "mpegb200_video_parser_has_ended": (C.c_int, [vp]),
        "mpegb200_video_parser_next": (C.c_int, [vp, vp]),
        "mpegb200_video_parser_next_scan": (C.c_int, [vp, vp]),
        "mpegb200_video_parser_redo": (C.c_int, [vp, C.c_int, vp]),
        "mpegb200_video_parser_unscan": (C.c_int, [vp]),
        "mpegb200_video_parser_set_start_codes": (C.c_int, [vp, vp, C.c_size_t]),
        "mpegb200_video_batch_parser": (vp, [vp, C.c_int]),
        "mpegb200_video_batch_set_resident": (C.c_int, [vp, C.c_int]),
        "mpegb200_video_batch_set_start_codes": (C.c_int, [vp, C.c_int, vp, C.c_size_t]),
        "mpegb200_video_batch_unscan": (C.c_int, [vp]),
        "mpegb200_video_batch_next_scan": (C.c_int, [vp, vp]),
        "mpegb200_video_batch_redo": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "mpegb200_video_batch_new": (vp, [C.c_int, C.c_int, vp, vp]),
        "mpegb200_video_batch_free": (None, [vp]),
        "mpegb200_video_batch_set_stream": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_size_t]),
        "mpegb200_video_batch_stream_size": (C.c_int, [vp, C.c_int, ip, ip]),
        "mpegb200_video_batch_set_vlen": (C.c_int, [vp, C.c_int]),
        "mpegb200_video_batch_next": (C.c_int, [vp, vp]),
        "mpegb200_device_stepper_new": (vp, [vp, vp, C.c_int, C.c_int]),
        "mpegb200_device_stepper_free": (None, [vp]),
        "mpegb200_device_stepper_step": (C.c_int, [vp, vp, vp, vp]),
        "mpegb200_device_stepper_drop_scan_ahead": (C.c_int, [vp]),
        "mpegb200_device_stepper_get_stats": (C.c_int, [vp, vp]),
        "mpegb200_audio_parser_new": (vp, [C.c_char_p, C.c_size_t]),
        "mpegb200_audio_parser_free": (None, [vp]),
        "mpegb200_audio_parser_has_header": (C.c_int, [vp]),
        "mpegb200_audio_parser_samplerate": (C.c_int, [vp]),
        "mpegb200_audio_parser_channels": (C.c_int, [vp]),
        "mpegb200_audio_parser_rewind": (None, [vp]),
        "mpegb200_audio_parser_next": (C.c_int, [vp, vp, C.POINTER(C.c_double)]),
        "mpegb200_audio_parser_next_coded": (C.c_int, [vp, vp, vp, C.POINTER(C.c_double)]),
        "mpegb200_audio_batch_new": (vp, [C.c_int, C.c_int, vp, vp]),
        "mpegb200_audio_batch_free": (None, [vp]),
        "mpegb200_audio_batch_set_stream": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_size_t]),
        "mpegb200_audio_batch_stream_info": (C.c_int, [vp, C.c_int, ip, ip]),
        "mpegb200_audio_batch_next": (C.c_int, [vp, C.c_int, vp]),
        "mpegb200_demux_split": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(vp), szp, C.POINTER(vp), szp, ip, ip]),
        "mpegb200_buffer_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


EXPORTED_SYMBOLS = None


def exported_symbols():
    """Names include/mpegb200.h declares (parsed from the header), for the no-GPU ABI test."""
    import re
    names = set()
    for h in ("mpegb200.h", "mpegb200_host.h"):
        hdr = (PKG.parent / "include" / h).read_text()
        names |= set(re.findall(r"\b(mpegb200_[a-z0-9_]+)\s*\(", hdr))
    return sorted(names)
