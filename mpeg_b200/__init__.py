"""mpeg_b200 -- B200-native (sm_100a) decode hot path for MPEG-1 video and MP2 audio.

A drop-in for the data-parallel kernels of gen2brain/mpeg: 8x8 integer IDCT, half-pel motion
compensation with residual add, YCbCr->RGBA and the MP2 synthesis filterbank run as hand-written
CUDA behind the C-ABI of include/mpegb200.h; the serial demux / VLC parse stays on the host.
"""
from ._lib import MpegB200Error, build, load  # noqa: F401
from .context import (AUDIO_F32, AUDIO_F32N, AUDIO_F32NLR, AUDIO_S16, AUDIO_WINDOW_FMA, MB_DTYPE, MB_INTRA, MB_PREDICT,  # noqa: F401
                      MB_REF_BWD, PIC_B, PIC_I, PIC_P, PICTURE_DTYPE, SAMPLES_PER_FRAME, Context)

from .batch import AudioBatch, DisplayRing, MPEGBatch, VideoBatch  # noqa: F401
from .mpeg import MPEG, Audio, ErrInvalidMPEG, Frame, Samples, Video, demux_split  # noqa: F401

__all__ = ["Context", "MpegB200Error", "build", "load", "MPEG", "Video", "Audio", "Frame", "Samples"]
