"""Binarizer::setFineSettings(bin_preset_t): the numeric fine settings (AGC limits, reference-level limits, sweep acceptance,
marker search distance, bit-picker depth) through every line decode path, against the UNMODIFIED reference run with the same
bin_preset_t (oracle/_ref, VideoToDigital::setFineSettings).  CPU part: the host build of the device code; GPU part: the CUDA
path through sdv_bin_set_fine_settings."""
import numpy as np
import pytest

from oracle import refbind as R
from sdvpcmdecoder_b200 import synth, capi
from sdvpcmdecoder_b200.capi import LINE_REC
from tests import util
from tests.test_pcm1_line import pcm1_cases, ref_lines as p1_ref_lines
from tests.test_pcm16x0_line import pcm16x0_cases, ref_sublines, _compare as x0_compare

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")

PRESETS = {
    "tight_levels": dict(max_black_lvl=60, min_white_lvl=100, min_contrast=40, min_ref_lvl=40, max_ref_lvl=150),
    "loose_sweep": dict(min_valid_crcs=2, mark_max_dist=12, min_ref_lvl=3, max_ref_lvl=250),
    "strict_sweep": dict(min_valid_crcs=12, mark_max_dist=3),
    "no_bit_picker": dict(left_bit_pick=0, right_bit_pick=0),
    "short_bit_picker": dict(left_bit_pick=2, right_bit_pick=1, min_white_lvl=90),
    "no_coord_search": dict(en_coord_search=0),
    "no_first_line_dup": dict(en_first_line_dup=0),       # the first PCM line of a field is not forced bad (videotodigital.cpp:1199)
}


def stc007_tapes():
    t = synth.make_stc007(2, seed=471)["luma"]
    dark = np.clip((t.astype(np.float32) - 16) * 0.35 + 10, 0, 255).astype(np.uint8)       # white near 74: below some min_white_lvl settings
    return {"clean": t, "config4": synth.damage_stc007(t, seed=4567), "heavy": synth.damage_stc007(t, seed=472, sigma=20.0, dropout_frac=0.1, marker_kill_frac=0.05),
            "dark": synth.damage_stc007(dark, seed=473, sigma=3.0)}


@needs_ref
@pytest.mark.parametrize("preset", sorted(PRESETS))
def test_stc007_lines_with_fine_settings(preset):
    try:
        R.set_fine_settings(**PRESETS[preset]); util.emu_set_fine(**PRESETS[preset])
        changed = False
        for name, luma in stc007_tapes().items():
            for mode in ((2,) if name != "config4" else (1, 2)):
                ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, mode, luma), keep=(0, 7))
                rec, aux, _ = util.emu_v2d(luma, mode, True, hybrid=True)
                bad = util.compare_line_records(ref, rec, aux)
                assert not bad, (preset, name, mode, bad)
                R.set_fine_settings()
                ref0 = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, mode, luma), keep=(0, 7))
                R.set_fine_settings(**PRESETS[preset])
                changed = changed or not np.array_equal(ref0, ref)
        if preset not in ("no_bit_picker", "short_bit_picker", "no_coord_search"):
            assert changed, "the preset must change the reference's output on at least one tape"
    finally:
        R.set_fine_settings(); util.emu_set_fine()


@needs_ref
@pytest.mark.parametrize("preset", ["short_bit_picker", "tight_levels", "no_coord_search", "no_first_line_dup"])        # (the GPU test runs all presets)
def test_pcm1_pcm16x0_lines_with_fine_settings(preset):
    try:
        R.set_fine_settings(**PRESETS[preset]); util.emu_set_fine(**PRESETS[preset])
        extra = ("clean",) if preset == "no_first_line_dup" else ()
        for name in ("damaged", "cutboth")+extra:
            luma = pcm1_cases()[name]
            rec, aux, _ = util.emu_p1_v2d(luma, 2, True)
            bad = util.compare_line_records(p1_ref_lines(luma, 2, True), rec, aux, oracle_only_flags=1 << 11)
            assert not bad, (preset, "pcm1", name, bad)
        cases = pcm16x0_cases()
        for name in ("damaged", "cutboth")+extra:
            luma = cases[name]
            rec, aux, _ = util.emu_x0_v2d(luma, 2, True)
            bad = x0_compare(ref_sublines(luma, 2, True), rec, aux)
            assert not bad, (preset, "pcm16x0", name, bad)
    finally:
        R.set_fine_settings(); util.emu_set_fine()


@pytest.mark.gpu
@needs_ref
def test_gpu_fine_settings_all_formats():
    """sdv_bin_set_fine_settings on the CUDA path: STC-007 (bulk + chain + relay), PCM-1, PCM-16x0 against the reference with the
    same bin_preset_t; then back to the defaults on the same handle."""
    import torch
    from sdvpcmdecoder_b200 import operators as ops
    h = capi.Handle(0)
    v2d = ops.VideoToDigital(h)
    try:
        for preset in sorted(PRESETS):
            R.set_fine_settings(**PRESETS[preset])
            v2d.setFineSettings(v2d.getDefaultFineSettings(), **PRESETS[preset])
            cur = v2d.getCurrentFineSettings()
            assert all(getattr(cur, k) == v for k, v in PRESETS[preset].items())
            v2d.setPCMType(capi.TYPE_STC007)
            for name, luma in stc007_tapes().items():
                recs, aux = v2d.doBinarize(torch.from_numpy(luma).cuda(), want_aux=True)
                ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, 2, luma), keep=(0, 7))
                bad = util.compare_line_records(ref, ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, capi.LINE_AUX))
                assert not bad, (preset, name, bad)
            v2d.setPCMType(capi.TYPE_PCM1)
            for name in ("cutboth", "clean") if preset == "no_first_line_dup" else ("cutboth",):
                luma = pcm1_cases()[name]
                recs, aux = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma)).cuda(), want_aux=True)
                bad = util.compare_line_records(p1_ref_lines(luma, 2, True), ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, capi.LINE_AUX), oracle_only_flags=1 << 11)
                assert not bad, (preset, "pcm1", name, bad)
            v2d.setPCMType(capi.TYPE_PCM16X0)
            for name in ("damaged", "clean") if preset == "no_first_line_dup" else ("damaged",):
                luma = pcm16x0_cases()[name]
                recs, aux = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma)).cuda(), want_aux=True)
                bad = x0_compare(ref_sublines(luma, 2, True), ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, capi.LINE_AUX))
                assert not bad, (preset, "pcm16x0", name, bad)
        # the two switches still taken at their defaults only are refused, loudly
        with pytest.raises(capi.SdvError):
            v2d.setFineSettings(en_force_coords=1)
        with pytest.raises(capi.SdvError):
            v2d.setFineSettings(en_good_no_marker=0)
        R.set_fine_settings()
        v2d.setDefaultFineSettings()
        v2d.setPCMType(capi.TYPE_STC007)
        luma = stc007_tapes()["config4"]
        recs, aux = v2d.doBinarize(torch.from_numpy(luma).cuda(), want_aux=True)
        ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, 2, luma), keep=(0, 7))
        assert not util.compare_line_records(ref, ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, capi.LINE_AUX))
    finally:
        R.set_fine_settings()
