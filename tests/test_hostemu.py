"""The integer logic of the CUDA path, compiled as host code with a one-thread "block" (tests/hostemu), against the
oracle.  This validates stc007_line.cuh / stc007_chain.cuh / stc007_deint.cuh and the bulk<->chain hand-off rules on a
machine without a GPU; the kernels themselves are checked by the -m gpu tests."""
import numpy as np
import pytest

from oracle import oraclebind as O
from sdvpcmdecoder_b200 import synth
from sdvpcmdecoder_b200.capi import LINE_REC
from tests import util


def _check(luma, mode=2, dup=True, hybrid=True):
    o = O.v2d_stc007(mode, luma, dup)
    rec, aux, st = util.emu_v2d(luma, mode, dup, hybrid=hybrid)
    bad = util.compare_line_records(o, rec, aux)
    assert not bad, bad
    return o, st


def test_clean_tape_chain_and_hybrid():
    luma = synth.make_stc007(4)["luma"]
    _check(luma, hybrid=False)
    o, st = _check(luma, hybrid=True)
    assert st[2] == 3 and st[3] == 1            # frames 1..3 come from one bulk pass
    assert (o["flags"] & 1).mean() > 0.99


def test_damaged_tape_all_modes():
    luma = synth.make_stc007(2)["luma"]
    for mode in (0, 1, 2, 3):
        _check(synth.damage_stc007(luma, seed=100 + mode), mode=mode)


def test_control_block_and_no_dup():
    _check(synth.make_stc007(3, control_block=True)["luma"])
    _check(synth.damage_stc007(synth.make_stc007(2)["luma"], seed=9), dup=False)


def test_sparse_damage_hands_back_to_bulk():
    """A clean tape with a few bad lines: the chain takes the dirty frames, the bulk records serve the rest and the
    black/white levels that changed after a bad line are patched in."""
    luma = synth.make_stc007(12)["luma"].copy()
    luma[3, 200:203] = synth.damage_stc007(luma[3:4, 200:203].copy(), seed=1, jitter=True)[0]
    luma[7, 50] = 0
    luma[7, 301, 300:500] = 255
    # one line with a different gain: its black/white levels replace the preset ones for everything after it
    x = luma[9, 400].astype(np.float32)
    luma[9, 400] = np.clip((x - 16) * 0.8 + 30, 0, 255).astype(np.uint8)
    luma[9, 398] = 0
    o, st = _check(luma)
    assert st[2] >= 6


def test_ntsc_geometry_and_odd_width():
    t = synth.make_stc007(3, pal=False, width=722, x0=15, x1=709)
    _check(t["luma"])
    _check(synth.damage_stc007(t["luma"], seed=5))


def _random_lines(n, seed, p_bad=0.05, burst=False):
    rng = np.random.RandomState(seed)
    audio = rng.randint(0, 1 << 14, size=(n, 6)).astype(np.uint16)
    words = synth.stc007_line_words(audio, n)
    lines = np.zeros(n, LINE_REC)
    lines["words"][:, :8] = words
    ok = rng.rand(n) >= p_bad
    if burst:
        s = rng.randint(200, n - 200)
        ok[s:s + 40] = False
    # some "valid" lines carry wrong words: forces BROKEN blocks
    liar = rng.rand(n) < 0.002
    lines["words"][liar, 2] ^= 0x155
    corrupt = ~ok
    lines["words"][corrupt, :8] ^= rng.randint(1, 1 << 14, size=(int(corrupt.sum()), 8)).astype(np.uint16)
    lines["flags"] = np.where(ok, 3, 0).astype(np.uint16)
    return lines


@pytest.mark.parametrize("res_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("pq", [(True, True), (True, False), (False, False)])
def test_deint_block_logic(res_mode, pq):
    lines = _random_lines(3000, seed=res_mode * 7 + pq[0] * 2 + pq[1], p_bad=0.06, burst=True)
    ob = O.deint_stc007(lines["words"][:, :8], (lines["flags"] & 3).astype(np.uint8), res_mode, False, True, pq[0], pq[1])
    blocks, samples, flags = util.emu_deint(lines, res_mode, False, True, pq[0], pq[1], broken_mask_dur=0)
    bad = util.compare_blocks(ob, blocks, samples, flags)
    assert not bad, bad
    assert (ob["flags"] & 2).any() or not pq[0]                 # the lying lines produced BROKEN blocks


def _lines_with_data_flags(n, seed):
    """Lines for the CRC-ignoring mode: 'has data' = valid coordinates and black/white levels set."""
    lines = _random_lines(n, seed=seed, p_bad=0.2)
    rng = np.random.RandomState(seed + 1)
    has_data = rng.rand(n) < 0.9
    lines["data_start"] = np.where(has_data, 14, -32768).astype(np.int16)
    lines["data_stop"] = np.where(has_data, 706, 32767).astype(np.int16)
    lines["flags"] |= np.where(has_data, 8, 0).astype(np.uint16)
    crc_ok = ((lines["flags"] & 1) | (has_data.astype(np.uint16) << 1)).astype(np.uint8)
    return lines, crc_ok


def test_deint_ignore_crc():
    lines, crc_ok = _lines_with_data_flags(2000, 17)
    ob = O.deint_stc007(lines["words"][:, :8], crc_ok, 0, True, False, True, True)
    blocks, samples, flags = util.emu_deint(lines, 0, True, False, True, True)
    bad = util.compare_blocks(ob, blocks, samples, flags)
    assert not bad, bad


@pytest.mark.parametrize("p_bad", [0.02, 0.1, 0.25])
def test_deint_standard_setting_many_patterns(p_bad):
    """The straight-line processBlock of the standard setting (14-bit, forced check, P+Q) over many erasure patterns,
    including valid-looking lines with wrong words (BROKEN blocks) and broken-block windows."""
    lines = _random_lines(20000, seed=int(p_bad * 1000), p_bad=p_bad, burst=True)
    ob = O.deint_stc007(lines["words"][:, :8], (lines["flags"] & 3).astype(np.uint8), 0, False, True, True, True)
    blocks, samples, flags = util.emu_deint(lines, 0, False, True, True, True, broken_mask_dur=0)
    bad = util.compare_blocks(ob, blocks, samples, flags)
    assert not bad, bad
    assert (ob["audio_state"] == 1).any() and (ob["audio_state"] == 2).any() and (ob["audio_state"] == 3).any()
