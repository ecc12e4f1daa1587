"""PCM-16x0 (SI) deinterleave operator: oracle pinned against the reference (golden fixture + live differential runs),
the product's block logic (host-compiled and on the GPU) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oraclebind as O, refbind as R
from sdvpcmdecoder_b200 import capi, synth
from tests import util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pcm16x0_deint.npz")
SETTINGS = ((0, 1, 1), (0, 0, 1), (0, 1, 0), (0, 0, 0), (1, 0, 1), (1, 1, 1))      # (ignore_crc, force_check, p_corr)


def make_sublines(n_itl, seed, p_bad=0.05, p_pick=0.03, lie=0.002):
    """Interleave blocks with consistent parity (sub-lines i / i+35 / i+70 = A / A^C / C), CRC failures with corrupted
    words, valid-looking lines with wrong words (-> BROKEN), bit-picker marks."""
    rng = np.random.RandomState(seed)
    n = n_itl * 105
    w = np.zeros((n, 3), np.uint16)
    for m in range(n_itl):
        a = rng.randint(0, 1 << 16, size=(35, 3)).astype(np.uint16)
        c = rng.randint(0, 1 << 16, size=(35, 3)).astype(np.uint16)
        w[m * 105:m * 105 + 35], w[m * 105 + 35:m * 105 + 70], w[m * 105 + 70:m * 105 + 105] = a, a ^ c, c
    ok = rng.rand(n) >= p_bad
    liar = rng.rand(n) < lie
    w[liar, 1] ^= 0x0101
    bad = ~ok
    w[bad] ^= rng.randint(1, 1 << 16, size=(int(bad.sum()), 3)).astype(np.uint16)
    data = ok | (rng.rand(n) < 0.7)
    pr = rng.rand(n) < p_pick
    pl = np.where(rng.rand(n) < p_pick, rng.randint(1, 5, size=n), 0).astype(np.uint8)
    flags = (ok * 1 + data * 2 + pr * 8).astype(np.uint8)
    return w, flags, pl


def to_records(w, flags, pl):
    sub = np.zeros(len(w), capi.PCM16X0_SUBLINE)
    sub["words"], sub["flags"], sub["picked_left"] = w, flags, pl
    return sub


def emu(sub, ign, force, p, ei=False):
    n_itl = len(sub) // (1470 if ei else 105)
    nb = n_itl * (490 if ei else 35)
    s = np.zeros((nb, 6), np.int16); f = np.zeros((nb, 6), np.uint8); st = np.zeros((nb, 3), np.uint8)
    util.emu().emu_deint_pcm16x0(sub.ctypes.data_as(C.c_void_p), n_itl, ign, force, p, int(ei), s.ctypes.data_as(C.c_void_p),
                                 f.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
    return s, f, st


def test_oracle_against_golden():
    g = np.load(GOLD)
    for k, (ign, force, p) in enumerate(SETTINGS):
        s, f, st = O.deint_pcm16x0(g["words"], g["flags"], g["picked_left"], ign, force, p)
        assert np.array_equal(s, g[f"samples_{k}"]) and np.array_equal(f, g[f"sflags_{k}"]) and np.array_equal(st, g[f"states_{k}"]), k
    assert (g["states_0"] == 1).any() and (g["states_0"] == 2).any()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(4))
def test_oracle_against_reference_live(seed):
    w, fl, pl = make_sublines(30, seed, p_bad=[0.0, 0.03, 0.1, 0.2][seed], p_pick=[0, 0.05, 0.2, 0.5][seed])
    for (ign, force, p) in SETTINGS:
        a, b = R.deint_pcm16x0(w, fl, pl, ign, force, p), O.deint_pcm16x0(w, fl, pl, ign, force, p)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (ign, force, p)


@pytest.mark.parametrize("seed", range(4))
def test_device_logic_on_host_against_oracle(seed):
    w, fl, pl = make_sublines(40, 10 + seed, p_bad=[0.0, 0.03, 0.1, 0.2][seed], p_pick=[0, 0.05, 0.2, 0.5][seed])
    sub = to_records(w, fl, pl)
    for (ign, force, p) in SETTINGS:
        a, b = O.deint_pcm16x0(w, fl, pl, ign, force, p), emu(sub, ign, force, p)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (ign, force, p)


def test_generator_layout_decodes_to_source_pairs():
    """The generator's PCM-16x0 sub-line layout (synth.make_pcm16x0) deinterleaves back to the source pairs in order."""
    t = synth.make_pcm16x0(1)
    sw = t["sub_words"][:735]                        # first field: 7 interleave blocks of 105 sub-lines
    s, f, st = O.deint_pcm16x0(sw, np.full(735, 3, np.uint8), np.zeros(735, np.uint8))
    exp = t["pairs"][:735].astype(np.uint16).view(np.int16).reshape(245, 6)
    assert np.array_equal(s, exp) and (f == 7).all() and (st == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("setting", SETTINGS)
def test_gpu_deint_pcm16x0(setting):
    import torch
    from sdvpcmdecoder_b200 import operators
    ign, force, p = setting
    w, fl, pl = make_sublines(300, 77, p_bad=0.08, p_pick=0.1)
    sub = to_records(w, fl, pl)
    d = operators.PCM16X0Deinterleaver(capi.Handle(0))
    d.setIgnoreCRC(ign); d.setForcedErrorCheck(force); d.setPCorrection(p)
    s, f, st = d.processInterleaveBlocks(torch.from_numpy(sub.view(np.uint8).reshape(-1, 8)).cuda())
    torch.cuda.synchronize()
    es, ef, est = O.deint_pcm16x0(w, fl, pl, ign, force, p)
    assert np.array_equal(s.cpu().numpy(), es) and np.array_equal(f.cpu().numpy(), ef) and np.array_equal(st.cpu().numpy(), est)


# ---- EI format (PCM-1630): one unit = one frame of 1470 sub-lines, data block i from sub-lines i, i+490, i+980
def make_sublines_ei(n_frames, seed, p_bad=0.05, p_pick=0.03):
    rng = np.random.RandomState(seed)
    n = n_frames * 1470
    w = np.zeros((n, 3), np.uint16)
    for m in range(n_frames):
        a = rng.randint(0, 1 << 16, size=(490, 3)).astype(np.uint16)
        c = rng.randint(0, 1 << 16, size=(490, 3)).astype(np.uint16)
        w[m * 1470:m * 1470 + 490], w[m * 1470 + 490:m * 1470 + 980], w[m * 1470 + 980:m * 1470 + 1470] = a, a ^ c, c
    ok = rng.rand(n) >= p_bad
    w[~ok] ^= rng.randint(1, 1 << 16, size=(int((~ok).sum()), 3)).astype(np.uint16)
    liar = rng.rand(n) < 0.002
    w[liar, 1] ^= 0x0101
    data = ok | (rng.rand(n) < 0.7)
    pr = rng.rand(n) < p_pick
    pl = np.where(rng.rand(n) < p_pick, rng.randint(1, 5, size=n), 0).astype(np.uint8)
    return w, (ok * 1 + data * 2 + pr * 8).astype(np.uint8), pl


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(3))
def test_ei_oracle_and_device_logic_against_reference_live(seed):
    w, fl, pl = make_sublines_ei(3, 50 + seed, p_bad=[0.0, 0.05, 0.2][seed], p_pick=[0, 0.05, 0.3][seed])
    sub = to_records(w, fl, pl)
    for (ign, force, p) in SETTINGS:
        a = R.deint_pcm16x0(w, fl, pl, ign, force, p, ei=True)
        b = O.deint_pcm16x0(w, fl, pl, ign, force, p, ei=True)
        c = emu(sub, ign, force, p, ei=True)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (ign, force, p)
        assert all(np.array_equal(x, y) for x, y in zip(a, c)), (ign, force, p)
    assert len(a[0]) == 3 * 490


def test_ei_golden():
    g = np.load(os.path.join(os.path.dirname(GOLD), "pcm16x0_deint_ei.npz"))
    for k, (ign, force, p) in enumerate(SETTINGS):
        s, f, st = O.deint_pcm16x0(g["words"], g["flags"], g["picked_left"], ign, force, p, ei=True)
        assert np.array_equal(s, g[f"samples_{k}"]) and np.array_equal(f, g[f"sflags_{k}"]) and np.array_equal(st, g[f"states_{k}"]), k
        e = emu(to_records(g["words"], g["flags"], g["picked_left"]), ign, force, p, ei=True)
        assert np.array_equal(e[0], s) and np.array_equal(e[1], f) and np.array_equal(e[2], st), k


@pytest.mark.gpu
@pytest.mark.parametrize("setting", SETTINGS)
def test_gpu_deint_pcm16x0_ei(setting):
    import torch
    from sdvpcmdecoder_b200 import operators
    ign, force, p = setting
    w, fl, pl = make_sublines_ei(40, 78, p_bad=0.08, p_pick=0.1)
    sub = to_records(w, fl, pl)
    d = operators.PCM16X0Deinterleaver(capi.Handle(0))
    d.setIgnoreCRC(ign); d.setForcedErrorCheck(force); d.setPCorrection(p); d.setEIFormat()
    s, f, st = d.processInterleaveBlocks(torch.from_numpy(sub.view(np.uint8).reshape(-1, 8)).cuda())
    torch.cuda.synchronize()
    es, ef, est = O.deint_pcm16x0(w, fl, pl, ign, force, p, ei=True)
    assert np.array_equal(s.cpu().numpy(), es) and np.array_equal(f.cpu().numpy(), ef) and np.array_equal(st.cpu().numpy(), est)
