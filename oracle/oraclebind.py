"""ctypes binding of oracle/_build/liboracle.so (the plain-C restatement, oracle/*.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .refbind import LINE_REC, BLOCK_REC  # identical record layouts

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        for name in ("sdvo_crc_stc007", "sdvo_crc_pcm1", "sdvo_crc_pcm16x0"):
            getattr(_lib, name).restype = C.c_uint16
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def crc_stc007(w):
    w = np.ascontiguousarray(w, dtype=np.uint16)
    return int(lib().sdvo_crc_stc007(_p(w)))


def crc_pcm1(w):
    w = np.ascontiguousarray(w, dtype=np.uint16)
    return int(lib().sdvo_crc_pcm1(_p(w)))


def crc_pcm16x0(w):
    w = np.ascontiguousarray(w, dtype=np.uint16)
    return int(lib().sdvo_crc_pcm16x0(_p(w)))


def binarize_lines_stc007(mode, lines, ref=0, black=0, white=0, start=0, stop=0):
    lines = np.ascontiguousarray(lines, dtype=np.uint8)
    n, w = lines.shape
    out = np.zeros(n, dtype=LINE_REC)
    lib().sdvo_binarize_lines_stc007(mode, _p(lines), n, w, w, ref, black, white, start, stop, _p(out))
    return out


def v2d_stc007(mode, luma, line_dup=True):
    luma = np.ascontiguousarray(luma, dtype=np.uint8)
    f, h, w = luma.shape
    out = np.zeros(f * h, dtype=LINE_REC)
    n = lib().sdvo_v2d_stc007(mode, int(line_dup), _p(luma), f, h, w, _p(out))
    return out[:n]


def deint_stc007(words, crc_ok, res_mode=0, ignore_crc=False, force_check=True, p_corr=True, q_corr=True):
    words = np.ascontiguousarray(words, dtype=np.uint16)
    crc_ok = np.ascontiguousarray(crc_ok, dtype=np.uint8)
    n = words.shape[0]
    out = np.zeros(max(n - 112, 1), dtype=BLOCK_REC)
    got = lib().sdvo_deint_stc007(_p(words), _p(crc_ok), n, res_mode, int(ignore_crc), int(force_check),
                                  int(p_corr), int(q_corr), _p(out))
    return out[:got].copy()


def deint_pcm1(lr, flags, ignore_crc=False):
    """PCM1Deinterleaver over whole fields: lr [n_fields*735, 2] u16, flags [n_fields*735] u8 -> (samples i16, flags u8), 1470 per field."""
    lr = np.ascontiguousarray(lr, dtype=np.uint16)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    n_fields = lr.shape[0] // 735
    s = np.zeros(n_fields * 1470, dtype=np.int16)
    f = np.zeros(n_fields * 1470, dtype=np.uint8)
    got = lib().sdvo_deint_pcm1(_p(lr), _p(flags), n_fields, int(ignore_crc), _p(s), _p(f))
    assert got == n_fields * 1470, got
    return s, f


def deint_pcm16x0(words, flags, picked_left, ignore_crc=False, force_check=True, p_corr=True, ei=False):
    """PCM16X0Deinterleaver (SI) over interleave blocks of 105 sub-lines: words [n, 3] u16, flags [n] u8, picked_left [n] u8
    -> (samples i16 [nb, 6], flags u8 [nb, 6], audio_state u8 [nb, 3]), nb = 35 per interleave block."""
    words = np.ascontiguousarray(words, dtype=np.uint16)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    picked_left = np.ascontiguousarray(picked_left, dtype=np.uint8)
    n_itl = words.shape[0] // (1470 if ei else 105)          # EI: units of one frame (1470 sub-lines, 490 data blocks)
    nb = n_itl * (490 if ei else 35)
    s = np.zeros((nb, 6), dtype=np.int16)
    f = np.zeros((nb, 6), dtype=np.uint8)
    st = np.zeros((nb, 3), dtype=np.uint8)
    got = getattr(lib(), "sdvo_deint_pcm16x0_ei" if ei else "sdvo_deint_pcm16x0")(_p(words), _p(flags), _p(picked_left), n_itl, int(ignore_crc), int(force_check), int(p_corr),
                                   _p(s), _p(f), _p(st))
    assert got == nb, got
    return s, f, st


def find_padding(w1, ok1, w2, ok2, video_std=1, resolution_16bit=False, res_mode=0, ignore_crc=False, p_corr=True, q_corr=True,
                 lim14=0x40, lim16=0x20):
    """STC007DataStitcher::findPadding -> (padding, DS_RET_* code, last_pad_counter)."""
    w1 = np.ascontiguousarray(w1, dtype=np.uint16); w2 = np.ascontiguousarray(w2, dtype=np.uint16)
    ok1 = np.ascontiguousarray(ok1, dtype=np.uint8); ok2 = np.ascontiguousarray(ok2, dtype=np.uint8)
    out = np.zeros(3, dtype=np.uint16)
    lib().sdvo_find_padding(_p(w1), _p(ok1), len(w1), _p(w2), _p(ok2), len(w2), int(video_std), int(resolution_16bit), res_mode,
                            int(ignore_crc), int(p_corr), int(q_corr), lim14, lim16, _p(out))
    return tuple(int(x) for x in out)


def try_padding(w1, ok1, w2, ok2, n_pad=32, res_mode=0, ignore_crc=False, p_corr=True, q_corr=True, lim14=0x40, lim16=0x20):
    """STC007DataStitcher::tryPadding for paddings 0..n_pad-1 -> u16 [n_pad, 6] (index, valid, silent, unchecked, broken, code)."""
    w1 = np.ascontiguousarray(w1, dtype=np.uint16); w2 = np.ascontiguousarray(w2, dtype=np.uint16)
    ok1 = np.ascontiguousarray(ok1, dtype=np.uint8); ok2 = np.ascontiguousarray(ok2, dtype=np.uint8)
    out = np.zeros((n_pad, 6), dtype=np.uint16)
    lib().sdvo_try_padding(_p(w1), _p(ok1), len(w1), _p(w2), _p(ok2), len(w2), n_pad, res_mode, int(ignore_crc), int(p_corr),
                           int(q_corr), lim14, lim16, _p(out))
    return out
