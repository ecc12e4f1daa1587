"""PCM-16x0 line decode + chain: the device code built for the host (tests/hostemu, one-thread block) against the compiled
reference (oracle/_ref) and the golden fixture.  The kernels themselves are checked by the -m gpu tests."""
import os

import numpy as np
import pytest

from oracle import refbind as R
from sdvpcmdecoder_b200 import synth
from sdvpcmdecoder_b200.capi import LINE_REC
from tests import util

have_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def pcm16x0_cases():
    base = synth.make_pcm16x0(2)["luma"]
    return {
        "clean": base,
        "damaged": synth.damage_stc007(base, seed=102),
        "noise": synth.damage_stc007(base, seed=202, jitter=False, blur=False, sigma=25., dropout_frac=0.05),
        "dropouts": synth.damage_stc007(base, seed=302, jitter=False, blur=False, sigma=3., dropout_frac=0.2),
        "cutleft": synth.make_pcm16x0(2, seed=11, x0=-5, x1=710)["luma"],
        "cutright": synth.make_pcm16x0(2, seed=12, x0=6, x1=723)["luma"],
        "cutboth": synth.damage_stc007(synth.make_pcm16x0(2, seed=13, x0=-7, x1=725)["luma"], seed=5, jitter=False, blur=False,
                                       sigma=6., dropout_frac=0.02),
        "drift": np.concatenate([synth.make_pcm16x0(2, seed=1)["luma"], synth.make_pcm16x0(2, seed=2, x0=9, x1=713)["luma"]]),
        "wide1440": synth.make_pcm16x0(1, seed=15, width=1440)["luma"],
    }


def ref_sublines(luma, mode=2, dup=True):
    ref = R.v2d_run(R.TYPE_PCM16X0, mode, luma, line_dup=dup)
    return ref[ref["service_type"] == 0][:luma.shape[0] * luma.shape[1] * 3]


def _compare(ref, rec, aux):
    bad = util.compare_line_records(util.x0_ref_to_product(ref), rec, aux, oracle_only_flags=0)
    if not np.array_equal(ref["line_part"], rec["reserved"]):
        bad.append("line_part")
    return bad


@have_ref
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_sublines_against_reference_live(mode):
    for name, luma in pcm16x0_cases().items():
        if mode != 2 and name in ("dropouts", "wide1440"):
            continue
        for dup in ((True, False) if name in ("clean", "drift") else (True,)):
            rec, aux, _ = util.emu_x0_v2d(luma, mode, dup)
            bad = _compare(ref_sublines(luma, mode, dup), rec, aux)
            assert not bad, (name, mode, dup, bad)


def test_golden_sublines():
    g = np.load(os.path.join(GOLD, "pcm16x0_lines.npz"))
    cases = pcm16x0_cases()
    for name in ("clean", "damaged", "cutboth", "drift"):
        rec, aux, _ = util.emu_x0_v2d(cases[name], 2, True)
        assert np.array_equal(g[name + "_recs"].view(LINE_REC).reshape(-1), rec), name


@have_ref
def test_mode_insane_against_reference_live():
    """MODE_INSANE incl. the carried-CRC quirk of the sweep's dummy line (a level whose stale CRC word happens to match is not
    searched, and VideoLine::scan_done keeps its previous value)."""
    base = synth.make_pcm16x0(1)["luma"]
    for luma in (base[:, :24], synth.damage_stc007(base, seed=102)[:, :40]):
        rec, aux, _ = util.emu_x0_v2d(luma, 3, True)
        bad = _compare(ref_sublines(luma, 3, True), rec, aux)
        assert not bad, bad
