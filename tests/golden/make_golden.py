"""Generates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libsdvref.so, built from /root/reference by
oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

  stc007_pipeline_pal.npz : PCMSamplePair stream of VideoToDigital + STC007DataStitcher (PAL/TFF/14-bit preset, P+Q on,
                            CWD off) on synth.make_stc007(8, seed=1234)
  stc007_lines_clean.npz / stc007_lines_damaged.npz : every STC007Line the reference VideoToDigital emits for a clean
                            and a damaged (synth.damage_stc007) tape, MODE_NORMAL
  stc007_deint.npz        : STC007Deinterleaver::processBlock results on random erased lines, all resolution modes
  stc007_try_padding.npz  : STC007DataStitcher::tryPadding (private member) for paddings 0..31 on eight field seams
  stc007_find_padding.npz : STC007DataStitcher::findPadding (private member) on 60 random seams x 18 settings
  pcm16x0_frame_info.npz  : sample rate / emphasis per frame of the reference's PCM-16x0 PCMSamplePair stream (control-bit votes + history)
  pcm16x0_deint.npz       : PCM16X0Deinterleaver::processBlock (SI) over 24 interleave blocks, six settings
  pcm1_deint.npz          : PCM1Deinterleaver::processBlock over 6 fields of random sub-lines, CRC checked / ignored
  pcm16x0_lines.npz       : every PCM16X0SubLine of VideoToDigital (MODE_NORMAL) for four tapes of
                            tests.test_pcm16x0_line.pcm16x0_cases()
  pcm16x0_ei_stitch.npz   : PCM-16x0 EI format: sub-line records of VideoToDigital and the PCMSamplePair stream of PCM16X0DataStitcher
                            (setFormat(FORMAT_EI); TFF / BFF / no P correction) for two tapes of tests.test_pcm16x0_stitch.ei_cases()
  stc007_stitch.npz       : line records of VideoToDigital and the PCMSamplePair stream of STC007DataStitcher (PAL/TFF/14-bit preset,
                            trim, paddings and masking its own) for two heavily damaged tapes of tests.test_stc007_stitch.stitch_cases()
                            (only this file: python tests/golden/make_golden.py stitch)
  pcm1_lines.npz          : every PCM1Line of VideoToDigital (MODE_NORMAL) and the PCMSamplePair stream of PCM1DataStitcher
                            (TFF, automatic line offset) for four tapes of tests.test_pcm1_line.pcm1_cases()
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402
from sdvpcmdecoder_b200 import synth  # noqa: E402
from tests.test_hostemu import _random_lines  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def stitch_golden():
    from tests import util
    from tests.test_stc007_stitch import stitch_cases, reference_stream
    out = {}
    for name in ("heavy", "vertical_jitter_heavy"):
        luma, std, order, res, p, q = stitch_cases()[name]
        pairs, _ = reference_stream(luma, std, order, res, p, q)
        recs = util.lines_from_oracle(util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, R.MODE_NORMAL, luma), keep=(0, 7)))
        out[name + "_recs"] = recs.view(np.uint8).reshape(len(recs), -1)
        out[name + "_shape"] = np.array(luma.shape)
        out[name + "_l"], out[name + "_r"], out[name + "_fl"], out[name + "_fr"] = pairs["l"], pairs["r"], pairs["flags_l"], pairs["flags_r"]
    np.savez_compressed(os.path.join(HERE, "stc007_stitch.npz"), **out)


def ei_stitch_golden():
    """PCM-16x0 EI format: the reference's sub-line records of two tapes + its PCMSamplePair stream for three settings."""
    from tests import util
    from tests.test_pcm16x0_line import ref_sublines
    from tests.test_pcm16x0_stitch import ei_cases, ref_pairs
    out = {}
    for name in ("variantB", "shift-60"):
        luma = ei_cases()[name]
        ref = ref_sublines(luma)
        rec = util.lines_from_oracle(util.x0_ref_to_product(ref))
        rec["flags"], rec["reserved"] = ref["flags"], ref["line_part"]       # the control bit travels in the flags (bit 11)
        out[name + "_rec"] = rec.view(np.uint8).reshape(len(rec), -1)
        out[name + "_frames"] = np.array(luma.shape[0])
        for i, (bff, p_corr) in enumerate(((False, True), (True, True), (False, False))):
            smp, fl = ref_pairs(luma, bff, p_corr, ei=True)
            got = util.emu_x0_stitch_auto(rec, luma.shape[0], luma.shape[1], bff, p_corr=p_corr, ei=True)
            assert np.array_equal(smp, got[0]) and np.array_equal(fl, got[1]), (name, bff, p_corr)
            out[f"{name}_smp{i}"], out[f"{name}_fl{i}"] = smp, fl
    np.savez_compressed(os.path.join(HERE, "pcm16x0_ei_stitch.npz"), **out)


def main():
    assert R.available(), "build oracle/_ref first (make -C oracle ref)"
    if sys.argv[1:] == ["stitch"]:
        return stitch_golden()
    if sys.argv[1:] == ["ei"]:
        return ei_stitch_golden()
    stitch_golden()
    ei_stitch_golden()
    # ---- pipeline
    n_frames, seed = 8, 1234
    tape = synth.make_stc007(n_frames, seed=seed)
    cfg = R.StitchCfg(video_std=1, field_order=1, resolution=1, p_corr=1, q_corr=1, cwd=0)
    pairs, _, _ = R.pipeline_run(R.TYPE_STC007, R.MODE_NORMAL, tape["luma"], cfg, taps=False)
    p = pairs[pairs["service_type"] == 0]
    np.savez_compressed(os.path.join(HERE, "stc007_pipeline_pal.npz"), n_frames=n_frames, seed=seed, first_pair=0,
                        l=p["l"], r=p["r"], flags_l=p["flags_l"], flags_r=p["flags_r"])
    # ---- line records
    for name, luma, mode in (("clean", synth.make_stc007(3, seed=11)["luma"], R.MODE_NORMAL),
                             ("damaged", synth.damage_stc007(synth.make_stc007(2, seed=12)["luma"], seed=4567), R.MODE_NORMAL)):
        r = R.v2d_run(R.TYPE_STC007, mode, luma)
        r = r[(r["service_type"] == 0) | (r["service_type"] == 7)]
        np.savez_compressed(os.path.join(HERE, f"stc007_lines_{name}.npz"), recs=r.view(np.uint8).reshape(len(r), -1), mode=mode)
    # ---- deinterleaver
    out = {}
    lines = _random_lines(2500, seed=321, p_bad=0.06, burst=True)
    out["words"] = lines["words"][:, :8]
    out["crc_ok"] = (lines["flags"] & 3).astype(np.uint8)
    for res_mode in range(4):
        b = R.deint_stc007(out["words"], out["crc_ok"], res_mode, False, True, True, True)
        out[f"blocks_{res_mode}"] = b.view(np.uint8).reshape(len(b), -1)
    np.savez_compressed(os.path.join(HERE, "stc007_deint.npz"), **out)
    # ---- PCM-1 deinterleaver
    from tests.test_pcm1 import make_sublines
    lr, flags = make_sublines(6, seed=123, p_bad=0.004)
    out = {"lr": lr, "flags": flags}
    for ign in (0, 1):
        smp, sfl = R.deint_pcm1(lr, flags, ign)
        out[f"samples_{ign}"], out[f"sflags_{ign}"] = smp, sfl
    np.savez_compressed(os.path.join(HERE, "pcm1_deint.npz"), **out)
    # ---- PCM-16x0 deinterleaver (SI)
    from tests.test_pcm16x0 import make_sublines as make16, SETTINGS
    w, fl, pl = make16(24, seed=456, p_bad=0.06, p_pick=0.08)
    out = {"words": w, "flags": fl, "picked_left": pl}
    for k, (ign, force, pc) in enumerate(SETTINGS):
        smp, sfl, st = R.deint_pcm16x0(w, fl, pl, ign, force, pc)
        out[f"samples_{k}"], out[f"sflags_{k}"], out[f"states_{k}"] = smp, sfl, st
    np.savez_compressed(os.path.join(HERE, "pcm16x0_deint.npz"), **out)
    # ---- PCM-16x0 deinterleaver, EI format
    from tests.test_pcm16x0 import make_sublines_ei, SETTINGS as X0_SETTINGS
    w, fl, pl = make_sublines_ei(2, 60, p_bad=0.08, p_pick=0.08)
    out = dict(words=w, flags=fl, picked_left=pl)
    for k, (ign, force, p) in enumerate(X0_SETTINGS):
        out[f"samples_{k}"], out[f"sflags_{k}"], out[f"states_{k}"] = R.deint_pcm16x0(w, fl, pl, ign, force, p, ei=True)
    np.savez_compressed(os.path.join(HERE, "pcm16x0_deint_ei.npz"), **out)
    # ---- STC-007 seam padding sweep (tryPadding)
    from tests.test_seam_sweep import fields, CASES
    out = {}
    for i, (lost, p_bad, nl, sil) in enumerate(CASES):
        f1, ok1, f2, ok2 = fields(i, lost, p_bad, nl, sil)
        for j, pq in enumerate(((1, 1), (1, 0), (0, 0))):
            out[f"stats_{i}_{j}"] = R.try_padding(f1, ok1, f2, ok2, 32, *pq)
    np.savez_compressed(os.path.join(HERE, "stc007_try_padding.npz"), **out)
    # ---- STC-007 seam padding decision (findPadding)
    from tests.test_seam_sweep import find_cases, FIND_SETTINGS
    res = np.zeros((len(find_cases()), len(FIND_SETTINGS), 3), dtype=np.uint16)
    for i, c in enumerate(find_cases()):
        for j, (std, r16, pq) in enumerate(FIND_SETTINGS):
            res[i, j] = R.find_padding(*c, std, r16, *pq)
    np.savez_compressed(os.path.join(HERE, "stc007_find_padding.npz"), res=res)
    # ---- PCM-16x0 control-bit decisions per frame
    from tests.test_pcm16x0_stitch import info_cases, ref_frame_info
    np.savez_compressed(os.path.join(HERE, "pcm16x0_frame_info.npz"), **{k: ref_frame_info(v) for k, v in info_cases().items()})
    # ---- PCM-1 line decode + stitcher
    from tests.test_pcm1_line import pcm1_cases, ref_lines, ref_samples
    from tests.util import lines_from_oracle
    import tests.util as U
    cases = pcm1_cases()
    out = {}
    for name in ("clean", "header", "damaged", "cutboth"):
        luma = cases[name]
        ref = ref_lines(luma, 2, True)
        r = lines_from_oracle(ref)
        r["flags"] = ref["flags"] & ~np.uint16(1 << 11)
        out[name + "_recs"] = r.view(np.uint8).reshape(len(r), -1)
        out[name + "_samples"], out[name + "_sflags"] = ref_samples(luma, 2, False)
    np.savez_compressed(os.path.join(HERE, "pcm1_lines.npz"), **out)
    # ---- PCM-16x0 line decode (three sub-line records per video line)
    from tests.test_pcm16x0_line import pcm16x0_cases, ref_sublines
    cases = pcm16x0_cases()
    out = {}
    for name in ("clean", "damaged", "cutboth", "drift"):
        ref = ref_sublines(cases[name], 2, True)
        r = lines_from_oracle(U.x0_ref_to_product(ref))
        r["flags"] = ref["flags"]
        r["reserved"] = ref["line_part"]
        out[name + "_recs"] = r.view(np.uint8).reshape(len(r), -1)
    np.savez_compressed(os.path.join(HERE, "pcm16x0_lines.npz"), **out)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
