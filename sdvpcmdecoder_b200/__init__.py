"""sdvpcmdecoder_b200 -- B200-native (sm_100a CUDA) per-frame decode path of SDVPCMdecoder.

Only the hot path lives here: video-line binarization + CRCC (VideoToDigital / Binarizer), frame assembly with preset
geometry and deinterleave + P/Q correction (STC007DataStitcher / STC007Deinterleaver).  The package is a thin host
layer over the C ABI in include/sdvpcm.h; it has no CPU implementation of the path.
"""
from . import capi  # noqa: F401
from .capi import Handle, SdvError  # noqa: F401
