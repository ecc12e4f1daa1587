"""ctypes binding of oracle/_ref/libsdvref.so (the UNMODIFIED reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm.
Never imported by the sdvpcmdecoder_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsdvref.so")

# Enumerations of the reference (videotodigital.h:75-82, binarizer.h:209-216, pcmline.h:78-108,
# frametrimset.h FrameAsmDescriptor, stc007datastitcher.h:213-220, stc007deinterleaver.h:108-116).
TYPE_PCM1, TYPE_PCM16X0, TYPE_STC007 = 0, 1, 2
MODE_DRAFT, MODE_FAST, MODE_NORMAL, MODE_INSANE = 0, 1, 2, 3
SRV_NO, SRV_NEW_FILE, SRV_END_FILE, SRV_FILLER, SRV_END_FIELD, SRV_END_FRAME, SRV_HEADER, SRV_CTRL_BLOCK = range(8)
RES_MODE_14BIT, RES_MODE_14BIT_AUTO, RES_MODE_16BIT_AUTO, RES_MODE_16BIT = 0, 1, 2, 3

LINE_REC = np.dtype([
    ("frame", "<u4"), ("line", "<u2"), ("words", "<u2", (9,)), ("data_start", "<i2"), ("data_stop", "<i2"),
    ("black", "u1"), ("white", "u1"), ("ref_low", "u1"), ("ref", "u1"), ("ref_high", "u1"), ("hyst", "u1"),
    ("shift", "u1"), ("service_type", "u1"), ("flags", "<u2"), ("mark_st_stage", "u1"), ("mark_ed_stage", "u1"),
    ("marker_start_bg", "<u2"), ("marker_start_ed", "<u2"), ("marker_stop_ed", "<u2"), ("word_crc_mask", "<u2"),
    ("word_valid_mask", "<u2"), ("line_part", "u1"), ("pcm_type", "u1"), ("queue_order", "<u2"), ("pad", "u1", (8,)),
], align=True)
PAIR_REC = np.dtype([("l", "<i2"), ("r", "<i2"), ("flags_l", "u1"), ("flags_r", "u1"), ("service_type", "u1"),
                     ("emphasis", "u1"), ("sample_rate", "<u2"), ("pad", "<u2")], align=True)
BLOCK_REC = np.dtype([("words", "<u2", (8,)), ("line_crc", "u1"), ("word_valid", "u1"), ("cwd_fixed", "u1"),
                      ("audio_state", "u1"), ("resolution", "u1"), ("flags", "u1"), ("start_line", "<u2"),
                      ("stop_line", "<u2"), ("start_frame", "<u4"), ("stop_frame", "<u4"), ("samples", "<i2", (6,)),
                      ("pad", "u1", (2,))], align=True)

# flags bits of LINE_REC
F_CRC_OK, F_CRC_OK_IGN, F_FORCED_BAD, F_BW_SET, F_COORDS_SET, F_REF_SWEEP, F_BY_EXT, F_COORD_SWEEP = (1 << i for i in range(8))
F_MARKERS, F_START_MARK, F_STOP_MARK, F_CTRL_BIT, F_ALMOST_SILENT = (1 << i for i in range(8, 13))


class StitchCfg(C.Structure):
    _fields_ = [("video_std", C.c_int), ("field_order", C.c_int), ("resolution", C.c_int), ("p_corr", C.c_int),
                ("q_corr", C.c_int), ("cwd", C.c_int), ("sample_rate", C.c_int), ("pcm16x0_format", C.c_int),
                ("auto_line_offset", C.c_int), ("reserved", C.c_int * 7)]


_lib = None


TYPE_M2 = 3          # VideoToDigital::TYPE_M2: STC-007 lines, M2 sample format


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        assert _lib.sdvref_sizeof_line_rec() == LINE_REC.itemsize, (_lib.sdvref_sizeof_line_rec(), LINE_REC.itemsize)
        assert _lib.sdvref_sizeof_pair_rec() == PAIR_REC.itemsize, (_lib.sdvref_sizeof_pair_rec(), PAIR_REC.itemsize)
        assert _lib.sdvref_sizeof_block_rec() == BLOCK_REC.itemsize, (_lib.sdvref_sizeof_block_rec(), BLOCK_REC.itemsize)
        for name in ("sdvref_crc_stc007", "sdvref_crc_pcm1", "sdvref_crc_pcm16x0"):
            getattr(_lib, name).restype = C.c_uint16
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def crc_stc007(words8):
    w = np.ascontiguousarray(words8, dtype=np.uint16)
    return int(lib().sdvref_crc_stc007(_p(w)))


def crc_pcm1(words6):
    w = np.ascontiguousarray(words6, dtype=np.uint16)
    return int(lib().sdvref_crc_pcm1(_p(w)))


def crc_pcm16x0(words3):
    w = np.ascontiguousarray(words3, dtype=np.uint16)
    return int(lib().sdvref_crc_pcm16x0(_p(w)))


def binarize_lines(pcm_type, mode, lines, part=0, ref=0, black=0, white=0, start=0, stop=0):
    """Binarizer::processLine on independent lines (u8 [n][W]) with explicit presets (0 = none)."""
    lines = np.ascontiguousarray(lines, dtype=np.uint8)
    n, w = lines.shape
    out = np.zeros(n, dtype=LINE_REC)
    lib().sdvref_binarize_lines(pcm_type, mode, part, _p(lines), n, w, w, ref, black, white, start, stop, _p(out))
    return out


FINE_FIELDS = ("max_black_lvl", "min_white_lvl", "min_contrast", "min_ref_lvl", "max_ref_lvl", "min_valid_crcs", "mark_max_dist",
               "left_bit_pick", "right_bit_pick", "en_coord_search", "en_first_line_dup")
FINE_DEFAULTS = dict(zip(FINE_FIELDS, (160, 28, 10, 7, 240, 5, 6, 4, 2, 1, 1)))


def set_fine_settings(**fields):
    """bin_preset_t for the following runs (VideoToDigital::setFineSettings); no arguments = the defaults."""
    if not fields:
        lib().sdvref_set_fine_settings(None)
        return
    v = dict(FINE_DEFAULTS)
    v.update(fields)
    lib().sdvref_set_fine_settings((C.c_int * len(FINE_FIELDS))(*[int(v[k]) for k in FINE_FIELDS]))


def v2d_run(pcm_type, mode, luma, line_dup=True, eof_mode=0):
    """VideoToDigital::doBinarize over u8 [F][H][W]; returns all emitted line records (service lines included)."""
    luma = np.ascontiguousarray(luma, dtype=np.uint8)
    f, h, w = luma.shape
    mult = 3 if pcm_type == TYPE_PCM16X0 else 1
    cap = (f + 2) * (h * mult + 8) + 16
    out = np.zeros(cap, dtype=LINE_REC)
    n = lib().sdvref_v2d_run(pcm_type, mode, int(line_dup), eof_mode, _p(luma), f, h, w, _p(out), cap)
    assert n >= 0, "record buffer overflow"
    return out[:n].copy()


def pipeline_run(pcm_type, mode, luma, cfg: StitchCfg, line_dup=True, eof_mode=0, taps=True):
    """Full reference pipeline (V2D + stitcher).  Returns (pairs, assembled_lines, blocks)."""
    luma = np.ascontiguousarray(luma, dtype=np.uint8)
    f, h, w = luma.shape
    cap_pairs = (f + 3) * (h + 32) * 3 + 4096      # 3 pairs per assembled line, 588 (PAL) / 490 (NTSC) lines per frame
    pairs = np.zeros(cap_pairs, dtype=PAIR_REC)
    cap_lines = (f + 3) * (h + 64) * 2
    cap_blocks = (f + 3) * (h + 64)
    lines = np.zeros(cap_lines if taps else 1, dtype=LINE_REC)
    blocks = np.zeros(cap_blocks if taps else 1, dtype=BLOCK_REC)
    n_lines = C.c_int(0)
    n_blocks = C.c_int(0)
    n = lib().sdvref_pipeline_run(pcm_type, mode, int(line_dup), eof_mode, C.byref(cfg), _p(luma), f, h, w,
                                  _p(pairs), cap_pairs,
                                  _p(lines) if taps else None, cap_lines, C.byref(n_lines),
                                  _p(blocks) if taps else None, cap_blocks, C.byref(n_blocks))
    assert n >= 0, "pair buffer overflow"
    assert n_lines.value <= cap_lines and n_blocks.value <= cap_blocks
    return pairs[:n].copy(), lines[:n_lines.value].copy(), blocks[:n_blocks.value].copy()


def deint_stc007(words, crc_ok, res_mode=RES_MODE_14BIT, ignore_crc=False, force_check=True, p_corr=True, q_corr=True):
    """STC007Deinterleaver::processBlock for every start line s in [0, n-112)."""
    words = np.ascontiguousarray(words, dtype=np.uint16)
    crc_ok = np.ascontiguousarray(crc_ok, dtype=np.uint8)
    n = words.shape[0]
    nb = max(n - 112, 0)
    out = np.zeros(max(nb, 1), dtype=BLOCK_REC)
    got = lib().sdvref_deint_stc007(_p(words), _p(crc_ok), n, res_mode, int(ignore_crc), int(force_check),
                                    int(p_corr), int(q_corr), _p(out))
    return out[:got].copy()


def deint_pcm1(lr, flags, ignore_crc=False):
    """PCM1Deinterleaver over whole fields: lr [n_fields*735, 2] u16, flags [n_fields*735] u8 -> (samples i16, flags u8), 1470 per field."""
    lr = np.ascontiguousarray(lr, dtype=np.uint16)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    n_fields = lr.shape[0] // 735
    s = np.zeros(n_fields * 1470, dtype=np.int16)
    f = np.zeros(n_fields * 1470, dtype=np.uint8)
    got = lib().sdvref_deint_pcm1(_p(lr), _p(flags), n_fields, int(ignore_crc), _p(s), _p(f))
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
    got = getattr(lib(), "sdvref_deint_pcm16x0_ei" if ei else "sdvref_deint_pcm16x0")(_p(words), _p(flags), _p(picked_left), n_itl, int(ignore_crc), int(force_check), int(p_corr),
                                   _p(s), _p(f), _p(st))
    assert got == nb, got
    return s, f, st


def find_padding(w1, ok1, w2, ok2, video_std=1, resolution_16bit=False, p_corr=True, q_corr=True):
    """STC007DataStitcher::findPadding (private member) -> (padding, DS_RET_* code, last_pad_counter)."""
    w1 = np.ascontiguousarray(w1, dtype=np.uint16); w2 = np.ascontiguousarray(w2, dtype=np.uint16)
    ok1 = np.ascontiguousarray(ok1, dtype=np.uint8); ok2 = np.ascontiguousarray(ok2, dtype=np.uint8)
    out = np.zeros(3, dtype=np.uint16)
    lib().sdvref_find_padding(_p(w1), _p(ok1), len(w1), _p(w2), _p(ok2), len(w2), int(video_std), int(resolution_16bit),
                              int(p_corr), int(q_corr), _p(out))
    return tuple(int(x) for x in out)


def try_padding(w1, ok1, w2, ok2, n_pad=32, p_corr=True, q_corr=True):
    """STC007DataStitcher::tryPadding (private member, reached through the test harness) for paddings 0..n_pad-1."""
    w1 = np.ascontiguousarray(w1, dtype=np.uint16); w2 = np.ascontiguousarray(w2, dtype=np.uint16)
    ok1 = np.ascontiguousarray(ok1, dtype=np.uint8); ok2 = np.ascontiguousarray(ok2, dtype=np.uint8)
    out = np.zeros((n_pad, 6), dtype=np.uint16)
    lib().sdvref_try_padding(_p(w1), _p(ok1), len(w1), _p(w2), _p(ok2), len(w2), n_pad, int(p_corr), int(q_corr), _p(out))
    return out
