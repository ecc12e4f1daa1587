"""PCM-1 deinterleave operator: oracle pinned against the reference (golden fixture + live), product against the oracle."""
import os

import numpy as np
import pytest

from oracle import oraclebind as O, refbind as R
from sdvpcmdecoder_b200 import capi, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pcm1_deint.npz")


def make_sublines(n_fields, seed, p_bad=0.02):
    """Sub-lines of the synthetic PCM-1 tape (synth.make_pcm1 layout) with random CRC failures; fields padded to 735."""
    rng = np.random.RandomState(seed)
    n = n_fields * 735
    lr = rng.randint(0, 1 << 13, size=(n, 2)).astype(np.uint16)
    crc = rng.rand(n) >= p_bad
    crc[:735] = True                                 # one fully valid field (valid interleave blocks)
    bw = crc | (rng.rand(n) < 0.5)
    flags = (crc.astype(np.uint8) * capi.P1F_CRC_OK) | (bw.astype(np.uint8) * capi.P1F_BW_SET)
    return lr, flags


def test_oracle_against_golden():
    g = np.load(GOLD)
    for ign in (0, 1):
        s, f = O.deint_pcm1(g["lr"], g["flags"], ign)
        assert np.array_equal(s, g[f"samples_{ign}"]) and np.array_equal(f, g[f"sflags_{ign}"])
    assert (g["sflags_0"] & 1).any() and not (g["sflags_0"] & 1).all()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_oracle_against_reference_live():
    lr, flags = make_sublines(5, seed=9, p_bad=0.01)
    for ign in (0, 1):
        a, b = R.deint_pcm1(lr, flags, ign), O.deint_pcm1(lr, flags, ign)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_source_pairs_come_back_in_order():
    """Deinterleaving the generator's sub-line layout returns the source sample pairs in order (encode -> decode)."""
    t = synth.make_pcm1(1)
    words = t["line_words"].reshape(-1, 2)           # 2 fields x 245 lines x 3 sub-lines
    flags = np.full(len(words), 3, np.uint8)
    s, f = O.deint_pcm1(words, flags)
    exp = synth.pcm1_expand(t["pairs"][:2 * 735].reshape(-1))
    assert np.array_equal(s, exp) and (f == 3).all()


@pytest.mark.gpu
@pytest.mark.parametrize("ignore_crc", [False, True])
def test_gpu_deint_pcm1(ignore_crc):
    import torch
    from sdvpcmdecoder_b200 import operators
    lr, flags = make_sublines(40, seed=5, p_bad=0.003)
    sub = np.zeros(len(lr), capi.PCM1_SUBLINE)
    sub["left"], sub["right"], sub["flags"] = lr[:, 0], lr[:, 1], flags
    d = operators.PCM1Deinterleaver(capi.Handle(0))
    d.setIgnoreCRC(ignore_crc)
    s, f = d.processFields(torch.from_numpy(sub.view(np.uint8).reshape(-1, 8)).cuda())
    torch.cuda.synchronize()
    es, ef = O.deint_pcm1(lr, flags, ignore_crc)
    assert np.array_equal(s.cpu().numpy(), es) and np.array_equal(f.cpu().numpy(), ef)
    assert (ef & 1).any() and not (ef & 1).all()
