"""PCM-1 line decode on the GPU (through the C ABI) against the compiled reference (oracle/_ref) and the golden fixtures."""
import os

import numpy as np
import pytest

from oracle import refbind as R
from sdvpcmdecoder_b200 import synth, capi
from sdvpcmdecoder_b200.capi import LINE_REC, LINE_AUX
from tests import util

pytestmark = pytest.mark.gpu
have_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    import torch
    from sdvpcmdecoder_b200 import operators
    assert torch.cuda.is_available()
    return capi.Handle(0), operators, torch


def _decode(ctx, luma, mode=2, dup=True):
    h, ops, torch = ctx
    v2d = ops.VideoToDigital(h)
    v2d.setPCMType(capi.TYPE_PCM1)
    v2d.setBinarizationMode(mode)
    v2d.setCheckLineDup(dup)
    recs, aux = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma)).cuda(), want_aux=True)
    torch.cuda.synchronize()
    return ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, LINE_AUX), v2d.stats()


def _check(ctx, luma, mode=2, dup=True):
    ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_PCM1, mode, luma, line_dup=dup))[:luma.shape[0] * luma.shape[1]]
    rec, aux, st = _decode(ctx, luma, mode, dup)
    bad = util.compare_line_records(ref, rec, aux, oracle_only_flags=1 << 11)
    assert not bad, bad
    return ref, st


@have_ref
def test_clean_tape_all_fields_and_bulk_path(ctx):
    luma = synth.make_pcm1(12)["luma"]
    ref, st = _check(ctx, luma)
    assert st["frames_skipped"] == 12 and st["lines_chain"] == 0      # every frame taken from the bulk kernel
    assert (ref["flags"] & 1).mean() > 0.99


@have_ref
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_modes_dup_and_header(ctx, mode):
    base = synth.make_pcm1(3, seed=21 + mode)["luma"]
    _check(ctx, base, mode)
    _check(ctx, base, mode, dup=False)
    _check(ctx, synth.make_pcm1(3, seed=7, header=True)["luma"], mode)


@have_ref
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_damaged_and_cut_tapes(ctx, mode):
    base = synth.make_pcm1(2)["luma"]
    _check(ctx, synth.damage_stc007(base, seed=100 + mode), mode)
    _check(ctx, synth.damage_stc007(base, seed=200 + mode, jitter=False, blur=False, sigma=25., dropout_frac=0.05), mode)
    _check(ctx, synth.make_pcm1(2, seed=11, x0=-9, x1=705)["luma"], mode)
    _check(ctx, synth.make_pcm1(2, seed=12, x0=10, x1=726)["luma"], mode)
    _check(ctx, synth.damage_stc007(synth.make_pcm1(2, seed=13, x0=-12, x1=728)["luma"], seed=5, jitter=False, blur=False,
                                    sigma=6., dropout_frac=0.02), mode)
    _check(ctx, synth.make_pcm1(2, seed=14, x0=30, x1=690)["luma"], mode)


@have_ref
def test_sparse_damage_mixes_bulk_and_chain(ctx):
    luma = synth.make_pcm1(10, seed=5)["luma"].copy()
    luma[3, 100:104, 200:500] = 255          # a dropout in frame 3
    luma[7, 0, :] = 16                       # first line of frame 7 blank
    ref, st = _check(ctx, luma)
    assert 0 < st["frames_skipped"] < 10 and st["lines_chain"] > 0


@have_ref
def test_unaligned_width(ctx):
    _check(ctx, synth.make_pcm1(3, seed=9, width=722, x0=9, x1=713)["luma"])


@have_ref
def test_mode_insane_reference_level_sweep(ctx):
    """MODE_INSANE: a full coordinate search at every reference level for every line that fails the presets (and for the
    four prescan lines of every frame).  Small frames: the reference needs seconds per swept line."""
    base = synth.make_pcm1(1)["luma"]
    _check(ctx, base[:, :64], mode=3)
    ref, st = _check(ctx, synth.damage_stc007(base, seed=102)[:, :120], mode=3)
    assert ((ref["flags"] >> 5) & 1).sum() > 0          # lines decoded by the sweep


def _samples(ctx, luma, mode=2, bff=False, ignore_crc=False):
    h, ops, torch = ctx
    v2d = ops.VideoToDigital(h)
    v2d.setPCMType(capi.TYPE_PCM1)
    v2d.setBinarizationMode(mode)
    recs = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma)).cuda())
    st = ops.PCM1DataStitcher(h)
    st.setFieldOrder(st.ORDER_BFF if bff else st.ORDER_TFF)
    st.setIgnoreCRC(ignore_crc)
    smp, fl, info = st.doFrameReassemble(recs, luma.shape[0], luma.shape[1], want_info=True)
    torch.cuda.synchronize()
    return smp.cpu().numpy(), fl.cpu().numpy() & 3, ops.records_to_numpy(info, capi.PCM1_FRAME_INFO), ops.records_to_numpy(recs, LINE_REC)


@have_ref
def test_pipeline_samples_against_reference(ctx):
    from tests.test_pcm1_line import pcm1_cases, ref_samples
    for name, luma in pcm1_cases().items():
        for bff in (False, True):
            smp, fl, info, _ = _samples(ctx, luma, 2, bff)
            rs, rf = ref_samples(luma, 2, bff)
            assert np.array_equal(rs, smp) and np.array_equal(rf, fl), (name, bff)


def test_golden_lines_and_samples(ctx):
    from tests.test_pcm1_line import pcm1_cases
    g = np.load(os.path.join(GOLD, "pcm1_lines.npz"))
    cases = pcm1_cases()
    for name in ("clean", "header", "damaged", "cutboth"):
        smp, fl, info, rec = _samples(ctx, cases[name])
        assert np.array_equal(g[name + "_recs"].view(LINE_REC).reshape(-1), rec), name
        assert np.array_equal(g[name + "_samples"], smp) and np.array_equal(g[name + "_sflags"], fl), name


def test_config2_round_trip(ctx):
    """BASELINE config 2 at full size (1000 frames, tiled from 50): every source sample pair comes back, bit-exact."""
    t = synth.make_pcm1(50)
    luma = np.tile(t["luma"], (20, 1, 1))
    smp, fl, info, rec = _samples(ctx, luma)
    src = synth.pcm1_expand(t["pairs"]).reshape(50, 2, 735, 2)
    got = smp.reshape(1000, 2, 735, 2)
    valid = (fl.reshape(1000, 2, 735, 2) & 2) != 0
    want = np.tile(src, (20, 1, 1, 1))
    assert valid.mean() > 0.97                      # 5 uncaptured lines per field + the first line of every field
    assert np.array_equal(got[valid], want[valid])
    assert (rec["flags"] & 1).mean() > 0.99


@have_ref
def test_manual_line_offsets(ctx):
    from tests.test_pcm1_line import pcm1_cases
    h, ops, torch = ctx
    luma = pcm1_cases()["clean"]
    v2d = ops.VideoToDigital(h)
    v2d.setPCMType(capi.TYPE_PCM1)
    recs = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma)).cuda())
    for ofs in ((3, 2), (-8, 1), (-10, -10)):
        cfg = R.StitchCfg()
        cfg.field_order, cfg.auto_line_offset = 1, 0
        cfg.reserved[0], cfg.reserved[1] = ofs
        pairs, _, _ = R.pipeline_run(R.TYPE_PCM1, 2, luma, cfg, taps=False)
        a = pairs[pairs["service_type"] == 0]
        st = ops.PCM1DataStitcher(h)
        st.setAutoLineOffset(False); st.setOddLineOffset(ofs[0]); st.setEvenLineOffset(ofs[1])
        smp, fl = st.doFrameReassemble(recs, luma.shape[0], luma.shape[1])
        torch.cuda.synchronize()
        assert np.array_equal(np.stack([a["l"], a["r"]], 1).reshape(-1), smp.cpu().numpy()), ofs
        assert np.array_equal(np.stack([a["flags_l"], a["flags_r"]], 1).reshape(-1) & 3, fl.cpu().numpy() & 3), ofs


@have_ref
@pytest.mark.parametrize("width", [352, 1024, 1920])
def test_frame_widths(ctx, width):
    m = width / 720.0
    luma = synth.make_pcm1(3, seed=width, width=width, x0=int(8 * m), x1=width - int(8 * m))["luma"]
    _check(ctx, luma)
    _check(ctx, synth.damage_stc007(luma[:2], seed=width + 1, sigma=6.0, dropout_frac=0.03, jitter=False, blur=False))
