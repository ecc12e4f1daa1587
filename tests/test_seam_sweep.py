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


# ---- findPadding: the decision over the sweep
GOLD_FIND = os.path.join(os.path.dirname(__file__), "golden", "stc007_find_padding.npz")
FIND_SETTINGS = [(std, r16, pq) for std in (0, 1, 2) for r16 in (False, True) for pq in ((True, True), (True, False), (False, False))]


def find_cases(n=60, seed=5):
    """Random seams: lost lines 0..35, damage from none to heavy, full to very short fields (incl. fewer than 112 queued
    lines: NO_DATA), silent and near-silent fields."""
    rng = np.random.RandomState(seed)
    out = []
    for t in range(n):
        lost = int(rng.randint(0, 36)); p_bad = float(rng.choice([0, 0.01, 0.05, 0.1, 0.2, 0.3, 0.5]))
        nl = int(rng.choice([288, 240, 130, 100, 60, 56, 20])); sil = bool(rng.rand() < 0.1)
        f1, ok1, f2, ok2 = fields(1000 + t, lost, p_bad, nl, sil)
        if rng.rand() < 0.3:
            f1[:, :] &= 0x0003
        out.append((f1, ok1, f2, ok2))
    return out


def test_find_padding_oracle_against_golden():
    g = np.load(GOLD_FIND)
    for i, c in enumerate(find_cases()):
        for j, (std, r16, pq) in enumerate(FIND_SETTINGS):
            assert np.array_equal(np.array(O.find_padding(*c, std, r16, 0, False, *pq), dtype=np.uint16), g["res"][i, j]), (i, std, r16, pq)
    assert (g["res"][:, :, 1] == 4).any() and (g["res"][:, :, 1] == 3).any() and (g["res"][:, :, 1] == 1).any()


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_find_padding_oracle_against_reference_live():
    for i, c in enumerate(find_cases(40, seed=77)):
        for std, r16, pq in FIND_SETTINGS[::2]:
            assert R.find_padding(*c, std, r16, *pq) == O.find_padding(*c, std, r16, 0, False, *pq), (i, std, r16, pq)


def test_try_padding_short_queues():
    # fewer than 112 queued lines: NO_DATA and untouched (cleared) statistics; exactly 112: no block, NO_PAD
    f1, ok1, f2, ok2 = fields(1008, 5, 0.05, 20)
    assert np.array_equal(O.try_padding(f1, ok1, f2, ok2, 2), np.array([[0, 0, 255, 255, 255, 0]] * 2, dtype=np.uint16))
    f1, ok1, f2, ok2 = fields(1008, 5, 0.0, 56)
    assert np.array_equal(O.try_padding(f1, ok1, f2, ok2, 1), np.array([[0, 0, 0, 0, 0, 3]], dtype=np.uint16))


def test_find_padding_recovers_the_lost_lines():
    for lost in (0, 6, 11, 25):
        f1, ok1, f2, ok2 = fields(300 + lost, lost, 0.01, 288)
        assert O.find_padding(f1, ok1, f2, ok2, 1)[:2] == (lost, 4)


@pytest.mark.gpu
def test_gpu_find_padding():
    import torch
    from sdvpcmdecoder_b200 import operators
    h = capi.Handle(0)
    cases = find_cases()
    recs, seams = [], []
    pos = 0
    for f1, ok1, f2, ok2 in cases:
        for w, ok in ((f1, ok1), (f2, ok2)):
            r = np.zeros(len(w), capi.LINE_REC)
            r["words"][:, :8] = w
            r["flags"] = ok
            recs.append(r)
        seams.append((pos, len(f1), pos + len(f1), len(f2)))
        pos += len(f1) + len(f2)
    dev = torch.from_numpy(np.concatenate(recs).view(np.uint8).reshape(-1, 32)).cuda()
    seams = np.array(seams, dtype=capi.SEAM)
    for std, r16, pq in FIND_SETTINGS:
        st = operators.STC007DataStitcher(h)
        st.setPCorrection(pq[0]); st.setQCorrection(pq[1])
        got = st.findPadding(dev, seams, video_std=std, resolution_16bit=r16)
        for i, c in enumerate(cases):
            exp = O.find_padding(*c, std, r16, 0, False, *pq)
            assert (int(got[i]["padding"]), int(got[i]["result"]), int(got[i]["last_pad_counter"])) == exp, (i, std, r16, pq)


@pytest.mark.gpu
def test_gpu_try_padding_short_queues():
    import torch
    from sdvpcmdecoder_b200 import operators
    h = capi.Handle(0)
    for nl in (20, 56):
        f1, ok1, f2, ok2 = fields(1008, 5, 0.0, nl)
        r = np.zeros(2 * nl, capi.LINE_REC)
        r["words"][:, :8] = np.concatenate([f1, f2]); r["flags"] = np.concatenate([ok1, ok2])
        dev = torch.from_numpy(r.view(np.uint8).reshape(-1, 32)).cuda()
        g = operators.STC007DataStitcher(h).tryPadding(dev, np.array([(0, nl, nl, nl)], dtype=capi.SEAM), 3)[0]
        tab = np.stack([g["index"], g["valid"], g["silent"], g["unchecked"], g["broken"], g["result"].astype(np.uint16)], axis=1)
        assert np.array_equal(tab, O.try_padding(f1, ok1, f2, ok2, 3)), nl


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
