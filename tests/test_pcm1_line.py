"""PCM-1 line decode + chain + frame assembly: the device code built for the host (tests/hostemu, one-thread block) against
the compiled reference (oracle/_ref) and the golden fixtures.  The kernels themselves are checked by the -m gpu tests."""
import os

import numpy as np
import pytest

from oracle import refbind as R, oraclebind as O
from sdvpcmdecoder_b200 import synth
from sdvpcmdecoder_b200.capi import LINE_REC
from tests import util

have_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def pcm1_cases():
    """name -> luma; shared with tests/golden/make_golden.py and the GPU tests."""
    base = synth.make_pcm1(2)["luma"]
    return {
        "clean": base,
        "header": synth.make_pcm1(3, seed=7, header=True)["luma"],
        "damaged": synth.damage_stc007(base, seed=102),
        "noise": synth.damage_stc007(base, seed=202, jitter=False, blur=False, sigma=25., dropout_frac=0.05),
        "cutleft": synth.make_pcm1(2, seed=11, x0=-9, x1=705)["luma"],
        "cutright": synth.make_pcm1(2, seed=12, x0=10, x1=726)["luma"],
        "cutboth": synth.damage_stc007(synth.make_pcm1(2, seed=13, x0=-12, x1=728)["luma"], seed=5, jitter=False, blur=False,
                                       sigma=6., dropout_frac=0.02),
        "narrow": synth.make_pcm1(2, seed=14, x0=30, x1=690)["luma"],
    }


def ref_lines(luma, mode=2, dup=True):
    ref = R.v2d_run(R.TYPE_PCM1, mode, luma, line_dup=dup)
    return util.ref_lines_in_frame_order(ref)[:luma.shape[0] * luma.shape[1]]


def ref_samples(luma, mode=2, bff=False):
    cfg = R.StitchCfg()
    cfg.field_order = 2 if bff else 1
    cfg.auto_line_offset = 1
    pairs, _, _ = R.pipeline_run(R.TYPE_PCM1, mode, luma, cfg, taps=False)
    a = pairs[pairs["service_type"] == 0]
    smp = np.stack([a["l"], a["r"]], axis=1).reshape(-1)
    fl = (np.stack([a["flags_l"], a["flags_r"]], axis=1).reshape(-1) & 3).astype(np.uint8)
    return smp, fl


def emu_samples(rec, n_frames, height, bff=False):
    sub, info = util.emu_p1_assemble(rec, n_frames, height, bff)
    smp, fl = O.deint_pcm1(np.stack([sub["left"], sub["right"]], axis=1), sub["flags"])
    return smp.reshape(-1), (fl.reshape(-1) & 3).astype(np.uint8), info


@have_ref
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_lines_against_reference_live(mode):
    for name, luma in pcm1_cases().items():
        for dup in ((True, False) if name in ("clean", "header") else (True,)):
            rec, aux, _ = util.emu_p1_v2d(luma, mode, dup)
            bad = util.compare_line_records(ref_lines(luma, mode, dup), rec, aux, oracle_only_flags=1 << 11)
            assert not bad, (name, mode, dup, bad)


@have_ref
def test_samples_against_reference_live():
    for name, luma in pcm1_cases().items():
        rec, _, _ = util.emu_p1_v2d(luma, 2, True)
        for bff in (False, True):
            smp, fl, info = emu_samples(rec, luma.shape[0], luma.shape[1], bff)
            rs, rf = ref_samples(luma, 2, bff)
            assert np.array_equal(rs, smp) and np.array_equal(rf, fl), (name, bff)
        if name == "header":
            assert info["header_present"].tolist() == [0, 1, 1]         # the reference forgets the header of the file's first frame


def test_golden_lines_and_samples():
    g = np.load(os.path.join(GOLD, "pcm1_lines.npz"))
    cases = pcm1_cases()
    for name in ("clean", "header", "damaged", "cutboth"):
        luma = cases[name]
        rec, aux, _ = util.emu_p1_v2d(luma, 2, True)
        want = g[name + "_recs"].view(LINE_REC).reshape(-1)
        assert np.array_equal(want, rec), name
        smp, fl, _ = emu_samples(rec, luma.shape[0], luma.shape[1])
        assert np.array_equal(g[name + "_samples"], smp) and np.array_equal(g[name + "_sflags"], fl), name


def test_prescan_rows_and_reduce():
    # frame_buf of a 480-row frame: 240 odd lines, END_FIELD, 240 even lines, END_FIELD, END_FRAME (+ NEW_FILE on the first)
    rec, aux, ps = util.emu_p1_v2d(pcm1_cases()["clean"], 2, True)
    assert ps["valid"].all() and (ps["start"] == 8).all() and set(ps["stop"].tolist()) <= {712, 713}
    rec, aux, ps = util.emu_p1_v2d(pcm1_cases()["clean"], 0, True)     # MODE_DRAFT: no prescan
    assert not ps["valid"].any()


@have_ref
def test_manual_line_offsets_against_reference_live():
    """setAutoLineOffset(false) + setOdd/EvenLineOffset: skipped / padded top lines, bottom trimmed when the field overflows."""
    cases = pcm1_cases()
    for name, ofs in (("clean", (0, 0)), ("clean", (3, 2)), ("clean", (-8, 1)), ("clean", (-10, -10)), ("header", (-4, -5)), ("damaged", (-5, -5))):
        luma = cases[name]
        cfg = R.StitchCfg()
        cfg.field_order, cfg.auto_line_offset = 1, 0
        cfg.reserved[0], cfg.reserved[1] = ofs
        pairs, _, _ = R.pipeline_run(R.TYPE_PCM1, 2, luma, cfg, taps=False)
        a = pairs[pairs["service_type"] == 0]
        rec, _, _ = util.emu_p1_v2d(luma, 2, True)
        sub, _ = util.emu_p1_assemble(rec, luma.shape[0], luma.shape[1], False, True, offsets=ofs)
        smp, fl = O.deint_pcm1(np.stack([sub["left"], sub["right"]], axis=1), sub["flags"])
        assert np.array_equal(np.stack([a["l"], a["r"]], 1).reshape(-1), smp.reshape(-1)), (name, ofs)
        assert np.array_equal((np.stack([a["flags_l"], a["flags_r"]], 1).reshape(-1) & 3), fl.reshape(-1) & 3), (name, ofs)


@have_ref
def test_mode_insane_against_reference_live():
    """MODE_INSANE (reference level sweep over the coordinate search); tiny frames, the reference takes seconds per swept line."""
    base = synth.make_pcm1(1)["luma"]
    for luma in (base[:, :32], synth.damage_stc007(base, seed=102)[:, :64]):
        rec, aux, _ = util.emu_p1_v2d(luma, 3, True)
        bad = util.compare_line_records(ref_lines(luma, 3, True), rec, aux, oracle_only_flags=1 << 11)
        assert not bad, bad
