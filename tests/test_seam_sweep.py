"""Field-seam padding sweep (STC007DataStitcher::tryPadding): oracle pinned against the reference's private member
(golden fixture + live), product kernel against the oracle, and the end-to-end property: on a tape whose fields lost
k lines between them, padding k is the only one that is not BROKEN."""
import os

import numpy as np
import pytest

from oracle import oraclebind as O, refbind as R
from sdvpcmdecoder_b200 import capi, synth
from tests import util

GOLD = os.path.join(os.path.dirname(__file__), "golden", "stc007_try_padding.npz")
CASES = [(6, 0.0, 288, False), (6, 0.02, 288, False), (0, 0.01, 288, False), (11, 0.05, 240, False), (6, 0.0, 100, False),
         (6, 0.0, 288, True), (20, 0.2, 288, False), (3, 0.0, 60, False)]


def fields(seed, lost=6, p_bad=0.01, n_lines=288, silent=False):
    """Two consecutive fields of a continuous STC-007 stream; [lost] stream lines between them were not captured."""
    rng = np.random.RandomState(seed)
    lpf = n_lines + lost
    audio = rng.randint(0, 1 << 14, size=(2 * lpf + 200, 6)).astype(np.uint16)
    if silent:
        audio[:] = 0
    words = synth.stc007_line_words(audio, 2 * lpf + 200)
    f1, f2 = words[100:100 + n_lines].copy(), words[100 + lpf:100 + lpf + n_lines].copy()
    ok1 = (rng.rand(n_lines) >= p_bad).astype(np.uint8) * 3
    ok2 = (rng.rand(n_lines) >= p_bad).astype(np.uint8) * 3
    f1[ok1 == 0] ^= 0x1555
    f2[ok2 == 0] ^= 0x2AAA
    return f1, ok1, f2, ok2


def test_oracle_against_golden():
    g = np.load(GOLD)
    for i, (lost, p_bad, nl, sil) in enumerate(CASES):
        f1, ok1, f2, ok2 = fields(i, lost, p_bad, nl, sil)
        for j, pq in enumerate(((1, 1), (1, 0), (0, 0))):
            assert np.array_equal(O.try_padding(f1, ok1, f2, ok2, 32, 0, False, *pq), g[f"stats_{i}_{j}"]), (i, pq)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_oracle_against_reference_live():
    for i, (lost, p_bad, nl, sil) in enumerate(CASES):
        f1, ok1, f2, ok2 = fields(50 + i, lost, p_bad, nl, sil)
        for pq in ((1, 1), (1, 0), (0, 0)):
            assert np.array_equal(R.try_padding(f1, ok1, f2, ok2, 32, *pq), O.try_padding(f1, ok1, f2, ok2, 32, 0, False, *pq)), (i, pq)


def test_true_padding_is_the_only_unbroken_one():
    f1, ok1, f2, ok2 = fields(7, lost=6, p_bad=0.0)
    st = O.try_padding(f1, ok1, f2, ok2, 32)
    assert st[6, 5] == capi.DS_RET_OK and st[6, 4] == 0 and st[6, 1] > 100
    assert (np.delete(st[:, 5], 6) == capi.DS_RET_BROKE).all()


@pytest.mark.gpu
def test_gpu_try_padding():
    import torch
    from sdvpcmdecoder_b200 import operators
    h = capi.Handle(0)
    recs, seams, expect = [], [], {}
    pos = 0
    for i, (lost, p_bad, nl, sil) in enumerate(CASES):
        f1, ok1, f2, ok2 = fields(80 + i, lost, p_bad, nl, sil)
        for w, ok in ((f1, ok1), (f2, ok2)):
            r = np.zeros(len(w), capi.LINE_REC)
            r["words"][:, :8] = w
            r["flags"] = ok
            recs.append(r)
        seams.append((pos, len(f1), pos + len(f1), len(f2)))
        pos += len(f1) + len(f2)
        expect[i] = (f1, ok1, f2, ok2)
    recs = np.concatenate(recs)
    dev = torch.from_numpy(recs.view(np.uint8).reshape(-1, 32)).cuda()
    for pq in ((True, True), (True, False), (False, False)):
        st = operators.STC007DataStitcher(h)
        st.setPCorrection(pq[0]); st.setQCorrection(pq[1])
        got = st.tryPadding(dev, np.array(seams, dtype=capi.SEAM), 32)
        for i in range(len(CASES)):
            exp = O.try_padding(*expect[i], 32, 0, False, *pq)
            g = got[i]
            tab = np.stack([g["index"], g["valid"], g["silent"], g["unchecked"], g["broken"], g["result"].astype(np.uint16)], axis=1)
            assert np.array_equal(tab, exp), (i, pq)
