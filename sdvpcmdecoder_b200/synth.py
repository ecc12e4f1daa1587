"""Synthetic PCM-on-video tape generator (STC-007, PCM-1, PCM-16x0).

The reference is decode-only, so the encoder side is derived from the decode-side
structure: line layouts from stc007line.h:72-152 / pcm1line.h:62-99 / pcm16x0subline.h:74-124, interleave
from stc007datablock.h:24-46, pcm1deinterleaver.cpp:150-278 and pcm16x0datablock.cpp:1029-1157, P/Q codes from
stc007deinterleaver.cpp:4-75,1297-1317.  Everything is seeded and deterministic (numpy only).

Frame layout follows VideoInFFMPEG::spliceFrame (vin_ffmpeg.cpp:213-364): odd field = rows 0,2,4.., even
field = rows 1,3,5.. (TFF), u8 luma [F][H][W].
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------- CRC-16 CCITT-FALSE, n-bit words
CRC_POLY = 0x1021
CRC_INIT = 0xFFFF


def crc16_words(words: np.ndarray, bits: int) -> np.ndarray:
    """CRC over words[..., k] taken MSB-first, `bits` bits each (pcmline.cpp:461-487)."""
    words = np.asarray(words, dtype=np.uint32)
    crc = np.full(words.shape[:-1], CRC_INIT, dtype=np.uint32)
    for k in range(words.shape[-1]):
        w = words[..., k]
        for b in range(bits - 1, -1, -1):
            inbit = (w >> b) & 1
            msb = (crc >> 15) & 1
            crc = (crc << 1) & 0xFFFF
            crc = np.where((msb ^ inbit) != 0, crc ^ CRC_POLY, crc)
    return crc.astype(np.uint16)


# ----------------------------------------------------------------------------- STC-007 P/Q
def _t_mul(v: np.ndarray) -> np.ndarray:
    """Multiply a 14-bit vector by T (x mod x^14 + x^8 + 1), cf. TP1_MATRIX stc007deinterleaver.cpp:8-11."""
    v = np.asarray(v, dtype=np.uint32)
    fb = (v >> 13) & 1
    return (((v << 1) & 0x3FFF) ^ fb ^ (fb << 8)).astype(np.uint32)


def t_pow(v: np.ndarray, k: int) -> np.ndarray:
    for _ in range(k):
        v = _t_mul(v)
    return v


def stc007_pq(audio: np.ndarray):
    """audio[..., 6] 14-bit words (L0 R0 L1 R1 L2 R2) -> (P, Q)."""
    a = np.asarray(audio, dtype=np.uint32)
    p = a[..., 0] ^ a[..., 1] ^ a[..., 2] ^ a[..., 3] ^ a[..., 4] ^ a[..., 5]
    q = np.zeros_like(p)
    for i in range(6):
        q ^= t_pow(a[..., i], 6 - i)
    return p.astype(np.uint16), q.astype(np.uint16)


# ----------------------------------------------------------------------------- painting bits into luma
def paint_lines(bits: np.ndarray, width: int, x0: int, x1: int, black: int, white: int) -> np.ndarray:
    """bits[n, nbits] (0/1) -> u8 [n, width]; pixel x in [x0,x1) shows bit floor((x-x0)*nbits/(x1-x0)).
    x0 < 0 or x1 > width paints a line whose first/last bit cells are cut off by the capture."""
    n, nbits = bits.shape
    x = np.arange(max(x0, 0), min(x1, width))
    idx = ((x - x0) * nbits) // (x1 - x0)
    out = np.full((n, width), black, dtype=np.uint8)
    seg = bits[:, idx]
    out[:, x[0]:x[-1] + 1] = np.where(seg != 0, np.uint8(white), np.uint8(black))
    return out


def _words_to_bits(words: np.ndarray, nbits: int) -> np.ndarray:
    """words[n, k] -> bits[n, k*nbits] MSB first."""
    w = np.asarray(words, dtype=np.uint32)
    sh = np.arange(nbits - 1, -1, -1, dtype=np.uint32)
    b = (w[..., None] >> sh) & 1
    return b.reshape(w.shape[0], -1).astype(np.uint8)


# ----------------------------------------------------------------------------- STC-007
STC007_CTRL_BLOCK = (0x3333, 0x0CCC, 0x3333, 0x0CCC, 0x0000, 0x0000, 0x0000, 0x0000)


def stc007_line_words(audio_blocks: np.ndarray, n_lines: int, periodic: bool = False) -> np.ndarray:
    """Interleave: line n carries word k of block n-16k (0 when n-16k < 0; block (n-16k) mod nb on a periodic tape,
    which can then be repeated end to end as one continuous valid tape).  Returns u16 [n_lines, 8]."""
    nb = audio_blocks.shape[0]
    p, q = stc007_pq(audio_blocks)
    blk = np.concatenate([audio_blocks.astype(np.uint16), p[:, None], q[:, None]], axis=1)  # [nb, 8]
    lines = np.zeros((n_lines, 8), dtype=np.uint16)
    n = np.arange(n_lines)
    for k in range(8):
        src = n - 16 * k
        if periodic:
            src = src % nb
        ok = (src >= 0) & (src < nb)
        lines[ok, k] = blk[src[ok], k]
    return lines


def stc007_bits(line_words: np.ndarray) -> np.ndarray:
    """u16 [n, 8] -> bits [n, 137] = 1010 + 8x14 + CRC16 + 01111."""
    crc = crc16_words(line_words, 14)
    n = line_words.shape[0]
    data = _words_to_bits(line_words, 14)
    crcb = _words_to_bits(crc[:, None], 16)
    start = np.tile(np.array([1, 0, 1, 0], dtype=np.uint8), (n, 1))
    stop = np.tile(np.array([0, 1, 1, 1, 1], dtype=np.uint8), (n, 1))
    return np.concatenate([start, data, crcb, stop], axis=1)


def make_stc007(n_frames: int, seed: int = 1234, pal: bool = True, width: int = 720, x0: int = 14, x1: int = 706,
                black: int = 16, white: int = 200, control_block: bool = False, field_start_line: int | None = None,
                periodic: bool = False, quiet_frac: float = 0.0, f1_16bit: bool = False):
    """Config-1 style tape.  Returns dict(luma u8[F][H][W], audio u16[nblocks][6], ...).

    The continuous PCM line stream has 294 (PAL) / 245 (NTSC) lines per field; the captured rows of a field
    are stream lines j0..j0+H/2-1 of that field (top j0 lines are not captured), j0 = 6 (PAL) / 5 (NTSC).
    """
    lpf = 294 if pal else 245
    height = 576 if pal else 480
    rows_pf = height // 2
    j0 = (lpf - rows_pf) if field_start_line is None else field_start_line
    n_fields = 2 * n_frames
    n_stream = n_fields * lpf
    rng = np.random.RandomState(seed)
    audio = rng.randint(0, 1 << 14, size=(n_stream, 6)).astype(np.uint16)
    if quiet_frac > 0:
        # stretches of near-silent audio (words around zero, incl. the M2 range bit): exercises the "almost silent" rules
        quiet = rng.rand(n_stream) < quiet_frac
        small = rng.randint(-6, 7, size=(n_stream, 6))
        audio[quiet] = (np.where(rng.rand(n_stream, 6) < 0.5, small & 0x3FFF, (small & 0x1FFF) | 0x2000).astype(np.uint16))[quiet]
    if f1_16bit:
        # PCM-F1 16-bit mode (stc007datablock.h:83-91, stc007deinterleaver.cpp 16-bit fill): the seven words L0..R2, P are 16 bit wide
        # (P = XOR of the six samples); a line carries their upper 14 bits in its first seven words and, in place of Q, the S word =
        # the two low bits of each of those seven words (L0 at bits 13..12 down to P at bits 1..0).  Returned audio is 16 bit.
        audio16 = rng.randint(0, 1 << 16, size=(n_stream, 6)).astype(np.uint16)
        p16 = audio16[:, 0] ^ audio16[:, 1] ^ audio16[:, 2] ^ audio16[:, 3] ^ audio16[:, 4] ^ audio16[:, 5]
        blk16 = np.concatenate([audio16, p16[:, None]], axis=1)
        words = np.zeros((n_stream, 8), dtype=np.uint16)
        n = np.arange(n_stream)
        for k in range(7):
            src = (n - 16 * k) % n_stream if periodic else n - 16 * k
            ok = (src >= 0) & (src < n_stream)
            w = np.zeros(n_stream, dtype=np.uint16)
            w[ok] = blk16[src[ok], k]
            words[:, k] = w >> 2
            words[:, 7] |= ((w & 3) << (12 - 2 * k)).astype(np.uint16)
        audio = audio16
    else:
        words = stc007_line_words(audio, n_stream, periodic=periodic)
    if control_block:
        # The first captured line of every field carries a Control Block instead of audio words.
        cb = np.array(STC007_CTRL_BLOCK, dtype=np.uint16)
        words[np.arange(n_fields) * lpf + j0] = cb
    bits = stc007_bits(words)
    stream_luma = paint_lines(bits, width, x0, x1, black, white)
    luma = np.empty((n_frames, height, width), dtype=np.uint8)
    line_index = np.empty((n_frames, height), dtype=np.int64)   # stream line shown in each row
    for f in range(n_frames):
        for fld in range(2):
            base = (2 * f + fld) * lpf + j0
            luma[f, fld::2, :] = stream_luma[base:base + rows_pf]
            line_index[f, fld::2] = np.arange(base, base + rows_pf)
    return dict(luma=luma, audio=audio, line_words=words, line_index=line_index, lines_per_field=lpf, j0=j0,
                x0=x0, x1=x1, black=black, white=white)


def damage_stc007(luma: np.ndarray, seed: int = 4567, sigma: float = 12.0, jitter: bool = True, blur: bool = True,
                  dropout_frac: float = 0.02, marker_kill_frac: float = 0.005, src_black: int = 16, src_white: int = 200):
    """Config-4 damage: per-line gain/offset jitter, Gaussian noise, 3-5 px box blur, dropouts, killed markers."""
    rng = np.random.RandomState(seed)
    f, h, w = luma.shape
    x = luma.reshape(f * h, w).astype(np.float32)
    n = x.shape[0]
    if blur:
        ks = rng.randint(3, 6, size=n)
        out = np.empty_like(x)
        for k in (3, 4, 5):
            sel = np.nonzero(ks == k)[0]
            if sel.size == 0:
                continue
            pad = np.pad(x[sel], ((0, 0), (k // 2, k - 1 - k // 2)), mode="edge")
            cs = np.cumsum(np.pad(pad, ((0, 0), (1, 0))), axis=1)
            out[sel] = (cs[:, k:] - cs[:, :-k]) / k
        x = out
    if jitter:
        nb = rng.randint(8, 49, size=n).astype(np.float32)
        nw = rng.randint(120, 236, size=n).astype(np.float32)
        x = (x - src_black) / float(src_white - src_black) * (nw - nb)[:, None] + nb[:, None]
    if sigma > 0:
        x = x + rng.normal(0.0, sigma, size=x.shape).astype(np.float32)
    x = np.clip(np.rint(x), 0, 255).astype(np.uint8)
    nd = int(n * dropout_frac)
    if nd:
        rows = rng.choice(n, nd, replace=False)
        for r in rows:
            ln = min(rng.randint(40, 401), w - 1)
            st = rng.randint(0, w - ln)
            x[r, st:st + ln] = 255 if rng.randint(2) else 0
    nk = int(n * marker_kill_frac)
    if nk:
        rows = rng.choice(n, nk, replace=False)
        for r in rows:
            if rng.randint(2):
                x[r, :40] = x[r, 40]
            else:
                x[r, w - 60:] = x[r, w - 61]
    return x.reshape(f, h, w)


# ----------------------------------------------------------------------------- PCM-1
def pcm1_expand(w13: np.ndarray) -> np.ndarray:
    """13 -> 16 bit sample expansion (pcm1line.cpp:196-233)."""
    w = np.asarray(w13, dtype=np.uint16)
    hi = (w & 0x1000) == 0
    a = (w << 4).astype(np.uint16)
    low = (w & 0x0FFF).astype(np.uint16)
    b = (low << 2).astype(np.uint16)
    b = np.where((low & 0x0800) != 0, b | 0xC000, b)
    return np.where(hi, a, b).astype(np.uint16).view(np.int16)


def make_pcm1(n_frames: int, seed: int = 2345, width: int = 720, x0: int = 8, x1: int = 712,
              black: int = 16, white: int = 200, header: bool = False):
    """Config-2 tape: NTSC 720x480, 245 lines per field, rows = lines 5..244; header=True (variant B) shows the header
    line (pcm1line.cpp:314-323) as the first captured line of every field."""
    lpf, height, rows_pf, j0 = 245, 480, 240, 5
    n_fields = 2 * n_frames
    rng = np.random.RandomState(seed)
    pairs = rng.randint(0, 1 << 13, size=(n_fields * 735, 2)).astype(np.uint16)   # source sample pairs (L,R)
    s = np.arange(735)
    nblk, r = s // 92, s % 92
    half, j = r // 46, r % 46
    odd = np.where(nblk % 2 == 0, half == 1, half == 0)
    t = np.where(odd, 2 * j, 2 * j + 1)
    q_in_field = 92 * nblk + t                                  # sub-line s carries source pair q
    words = np.zeros((n_fields * lpf, 6), dtype=np.uint16)
    for fld in range(n_fields):
        q = fld * 735 + q_in_field
        sub = pairs[np.minimum(q, pairs.shape[0] - 1)]        # [735, 2]
        words[fld * lpf:(fld + 1) * lpf] = sub.reshape(lpf, 6)
    crc = (~crc16_words((~words) & 0x1FFF, 13)).astype(np.uint16)
    if header:
        rows = np.arange(n_fields) * lpf + j0
        words[rows] = np.array([0x0666, 0x0CCC, 0x1999, 0x1333, 0x0666, 0x0CCC], dtype=np.uint16)
        crc[rows] = 0xCCCC
    bits = np.concatenate([_words_to_bits(words, 13), _words_to_bits(crc[:, None], 16)], axis=1)
    stream_luma = paint_lines(bits, width, x0, x1, black, white)
    luma = np.empty((n_frames, height, width), dtype=np.uint8)
    for f in range(n_frames):
        for fld in range(2):
            base = (2 * f + fld) * lpf + j0
            luma[f, fld::2, :] = stream_luma[base:base + rows_pf]
    return dict(luma=luma, pairs=pairs, line_words=words, lines_per_field=lpf, j0=j0)


# ----------------------------------------------------------------------------- PCM-16x0 (SI format)
def make_pcm16x0(n_frames: int, seed: int = 3456, width: int = 720, black: int = 16, white: int = 200,
                 x0: int | None = None, x1: int | None = None, ctrl_lines=(1,), ei: bool = False, silent_frames=()):
    """Config-3 tape: NTSC 720x480, SI format, 44.1 kHz (control bit 0 on line 1 of each 35-line interleave block).
    x0 / x1: data coordinates (default width/90 from either edge; off-screen values cut bit cells off).
    ei: the EI format -- one interleave unit per frame, data block b = sub-lines b, b+490, b+980 of the frame's 1470
    (pcm16x0datablock.h:41,70-72); pass ctrl_lines=(1, 2) to flag it in the control bits.  silent_frames: frames of zero samples."""
    lpf, height, rows_pf, j0 = 245, 480, 240, 5
    x0 = width // 90 if x0 is None else x0
    x1 = width - width // 90 if x1 is None else x1
    n_fields = 2 * n_frames
    rng = np.random.RandomState(seed)
    pairs = rng.randint(0, 1 << 16, size=(n_fields * 735, 2)).astype(np.uint16)
    for f in silent_frames:                      # digital silence: every sample of the frame zero ...
        if isinstance(f, tuple):                 # ... or (frame, first pair, last pair + 1) of its 1470 pairs
            pairs[f[0] * 1470 + f[1]:f[0] * 1470 + f[2]] = 0
        else:
            pairs[f * 1470:(f + 1) * 1470] = 0
    if ei:
        s = np.arange(1470)
        g, i = s // 490, s % 490
    else:
        s = np.arange(735)
        m, r = s // 105, s % 105
        g, i = r // 35, r % 35
    sub_words = np.zeros((n_fields, 735, 3), dtype=np.uint16)
    for fld in range(n_frames if ei else n_fields):
        for jw in range(3):
            q = ((fld * 490 + i) * 3 + jw) if ei else (((fld * 7 + m) * 35 + i) * 3 + jw)
            l_, r_ = pairs[q, 0], pairs[q, 1]
            takes_l = np.where(i % 2 == 0, jw == 1, jw != 1)
            w0 = np.where(takes_l, l_, r_)
            w2 = np.where(takes_l, r_, l_)
            v = np.where(g == 1, l_ ^ r_, np.where(g == 0, w0, w2))
            if ei:
                sub_words[2 * fld:2 * fld + 2, :, jw] = v.reshape(2, 735)
            else:
                sub_words[fld, :, jw] = v
    sw = sub_words.reshape(n_fields * 735, 3)
    crc = crc16_words(sw, 16)
    part_bits = np.concatenate([_words_to_bits(sw, 16), _words_to_bits(crc[:, None], 16)], axis=1)  # [n*735, 64]
    pb = part_bits.reshape(n_fields * lpf, 3, 64)
    line_in_field = np.tile(np.arange(lpf), n_fields)
    # control bit (active = 0) on the lines [ctrl_lines] of every 35-line interleave block: 0 emphasis, 1 44.1 kHz, 2 EI format, 3 code
    ctrl = np.where(np.isin(line_in_field % 35, list(ctrl_lines)), 0, 1).astype(np.uint8)
    bits = np.concatenate([pb[:, 0], pb[:, 1], ctrl[:, None], pb[:, 2]], axis=1)   # 193 bits
    stream_luma = paint_lines(bits, width, x0, x1, black, white)
    luma = np.empty((n_frames, height, width), dtype=np.uint8)
    for f in range(n_frames):
        for fld in range(2):
            base = (2 * f + fld) * lpf + j0
            luma[f, fld::2, :] = stream_luma[base:base + rows_pf]
    return dict(luma=luma, pairs=pairs, sub_words=sw, lines_per_field=lpf, j0=j0)
