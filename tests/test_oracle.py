"""Pins the oracle (oracle/*.c, test infrastructure): PCMTester known answers, golden fixtures generated from the
unmodified reference (tests/golden/make_golden.py) and -- where oracle/_ref exists -- live differential runs."""
import os

import numpy as np
import pytest

from oracle import oraclebind as O, refbind as R
from sdvpcmdecoder_b200 import synth
from tests import util
from tests.test_hostemu import _random_lines

GOLD = os.path.join(os.path.dirname(__file__), "golden")
have_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (reference tree absent)")


def test_crc_known_answers():
    # pcmtester.cpp:14-21, 45-49, 73-82
    assert O.crc_pcm1([0x1A35, 0x1248, 0x0DD9, 0x13FB, 0x1C0E, 0x09CB]) == 0x9EB9
    assert O.crc_pcm16x0([0xD527, 0x9C36, 0x02A5]) == 0xFB40
    assert O.crc_stc007([0x2D4B, 0x18EE, 0x152B, 0x3A7F, 0x04AB, 0x301B, 0x22F6, 0x0DD6]) == 0xB2ED
    # silent lines: stc007line.h:120, pcm1line.h:98, pcm16x0subline.h:102
    assert O.crc_stc007([0] * 8) == 0xA96A
    assert O.crc_pcm1([0x1000] * 6) == 0xECBF or O.crc_pcm1([0] * 6) == 0xECBF
    assert O.crc_pcm16x0([0] * 3) == 0x0E10


def test_pq_known_block_and_erasure_property():
    # pcmtester.cpp:119-126: L0..R2 = 3B43 3FDB 3B52 3FDA 3B5F 3FDA, P = 0495, Q = 1DB7
    audio = np.array([[0x3B43, 0x3FDB, 0x3B52, 0x3FDA, 0x3B5F, 0x3FDA]], dtype=np.uint16)
    p, q = synth.stc007_pq(audio)
    assert (int(p[0]), int(q[0])) == (0x0495, 0x1DB7)
    # pcmtester.cpp:110-369: kill 1 / kill 2 words -> exact recovery; more -> block not valid
    block = np.array([0x3B43, 0x3FDB, 0x3B52, 0x3FDA, 0x3B5F, 0x3FDA, 0x0495, 0x1DB7], dtype=np.uint16)
    rng = np.random.RandomState(5)
    for trial in range(600):
        nkill = 1 + trial % 4
        words = np.tile(block, (113, 1))
        ok = np.full(113, 3, np.uint8)
        killed = rng.choice(8, nkill, replace=False)
        for k in killed:
            words[16 * k, k] ^= rng.randint(1, 1 << 14)
            ok[16 * k] = 0
        b = O.deint_stc007(words, ok, 0, False, True, True, True)[0]
        if nkill <= 2:
            assert np.array_equal(b["words"], block) and (b["flags"] & 1)
        else:
            assert (not (b["flags"] & 1)) or not any(k < 6 for k in killed)


def _golden_lines(name):
    g = np.load(os.path.join(GOLD, f"stc007_lines_{name}.npz"))
    return g["recs"].reshape(-1).view(R.LINE_REC), int(g["mode"])


@pytest.mark.parametrize("name,make", [("clean", lambda: synth.make_stc007(3, seed=11)["luma"]),
                                       ("damaged", lambda: synth.damage_stc007(synth.make_stc007(2, seed=12)["luma"], seed=4567))])
def test_v2d_against_golden(name, make):
    ref, mode = _golden_lines(name)
    got = O.v2d_stc007(mode, make())
    for f in R.LINE_REC.names:
        if f != "pad":
            assert np.array_equal(ref[f], got[f]), f


def test_deint_against_golden():
    g = np.load(os.path.join(GOLD, "stc007_deint.npz"))
    for res_mode in range(4):
        ref = g[f"blocks_{res_mode}"].reshape(-1).view(R.BLOCK_REC)
        got = O.deint_stc007(g["words"], g["crc_ok"], res_mode, False, True, True, True)
        for f in ("words", "line_crc", "word_valid", "audio_state", "resolution", "samples"):
            assert np.array_equal(ref[f], got[f]), (res_mode, f)
        assert np.array_equal(ref["flags"] & 0x1F, got["flags"] & 0x1F), res_mode


def test_pipeline_golden_against_oracle_chain():
    """The product's fixed-geometry assembly (lead-in 80 lines, 294 lines per field) reproduces the reference's
    PCMSamplePair stream exactly -- checked here with oracle line records + the host-compiled deinterleave logic."""
    from sdvpcmdecoder_b200.capi import LINE_REC
    g = np.load(os.path.join(GOLD, "stc007_pipeline_pal.npz"))
    n_frames = int(g["n_frames"])
    luma = synth.make_stc007(n_frames, seed=int(g["seed"]))["luma"]
    recs = util.lines_from_oracle(O.v2d_stc007(2, luma, True))
    lpf, hf, height = 294, 288, 576
    nb = 80 + n_frames * 2 * lpf
    asm = np.zeros(nb + 112, LINE_REC)
    for fld in range(2 * n_frames):
        src = (fld // 2) * height + (fld & 1) * hf
        asm[80 + fld * lpf:80 + fld * lpf + hf] = recs[src:src + hf]
    _, s, f = util.emu_deint(asm, 0, False, True, True, True, 128)
    s, f = s.reshape(-1, 2), f.reshape(-1, 2)
    assert len(s) == len(g["l"])
    assert np.array_equal(s[:, 0], g["l"]) and np.array_equal(s[:, 1], g["r"])
    assert np.array_equal(f[:, 0], g["flags_l"] & 7) and np.array_equal(f[:, 1], g["flags_r"] & 7)


@have_ref
def test_crc_against_reference():
    rng = np.random.RandomState(1)
    for _ in range(200):
        w = rng.randint(0, 1 << 14, 8)
        assert O.crc_stc007(w) == R.crc_stc007(w)
        w = rng.randint(0, 1 << 13, 6)
        assert O.crc_pcm1(w) == R.crc_pcm1(w)
        w = rng.randint(0, 1 << 16, 3)
        assert O.crc_pcm16x0(w) == R.crc_pcm16x0(w)


@have_ref
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_v2d_against_reference_live(mode):
    base = synth.make_stc007(2, seed=40 + mode, control_block=(mode == 1))["luma"]
    for luma in (base, synth.damage_stc007(base, seed=50 + mode)):
        ref = R.v2d_run(R.TYPE_STC007, mode, luma)
        ref = ref[(ref["service_type"] == 0) | (ref["service_type"] == 7)]
        got = O.v2d_stc007(mode, luma)
        for f in R.LINE_REC.names:
            if f != "pad":
                assert np.array_equal(ref[f], got[f]), (mode, f)


@have_ref
def test_binarize_lines_against_reference_live():
    luma = synth.damage_stc007(synth.make_stc007(1, seed=3)["luma"], seed=8)[0][:120]
    for presets in (dict(), dict(ref=100, black=30, white=190, start=14, stop=682), dict(ref=90)):
        ref = R.binarize_lines(R.TYPE_STC007, 2, luma, **presets)
        got = O.binarize_lines_stc007(2, luma, **presets)
        for f in R.LINE_REC.names:
            if f not in ("pad", "frame", "line"):
                assert np.array_equal(ref[f], got[f]), (presets, f)


@have_ref
def test_deint_against_reference_live():
    lines = _random_lines(2000, seed=99, p_bad=0.08)
    for res_mode in range(4):
        for pq in ((1, 1), (1, 0), (0, 0)):       # Q on implies P on in the reference
            a = (lines["words"][:, :8], (lines["flags"] & 3).astype(np.uint8), res_mode, False, True, pq[0], pq[1])
            ref, got = R.deint_stc007(*a), O.deint_stc007(*a)
            for f in ("words", "line_crc", "word_valid", "audio_state", "resolution", "samples"):
                assert np.array_equal(ref[f], got[f]), (res_mode, pq, f)
