"""Fuzz of the PCM-16x0 stitcher (SI and EI formats) on the host build of the device code against the reference pipeline
(oracle/_ref): random tapes (control-bit layouts, blanked spans / bands / frames, noise, heavy damage, vertical shifts), three
settings each (TFF, BFF, no P correction); every frame of the PCMSamplePair stream must be equal.
    python tools/parity_fuzz_x0_stitch.py [seed] [cases] [si|ei|both]"""
import sys

import numpy as np

from oracle import refbind as R
from sdvpcmdecoder_b200 import synth
from tests import util
from tests.test_pcm16x0_stitch import variant_b, shift_rows, ref_pairs, edge_damage


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    fmts = {"si": (False,), "ei": (True,), "both": (False, True)}[sys.argv[3] if len(sys.argv) > 3 else "both"]
    assert R.available(), "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.RandomState(seed)
    n_bad = 0
    for it in range(cases):
        ctrl = [(1, 2), (2,), (), (0, 1, 2, 3), (1,)][rng.randint(5)]
        ei = fmts[rng.randint(len(fmts))]
        base = synth.make_pcm16x0(4, seed=rng.randint(1 << 20), ei=ei, ctrl_lines=ctrl)["luma"]
        kind = rng.randint(7)
        luma = base
        if kind == 1:
            luma = variant_b(base, seed=rng.randint(1000), frac=rng.choice([0.05, 0.2, 0.5]))
        if kind == 2:
            luma = synth.damage_stc007(base, seed=rng.randint(1000), jitter=False, blur=False, sigma=float(rng.choice([15, 25, 40])),
                                       dropout_frac=float(rng.choice([0.02, 0.1, 0.3])))
        if kind == 3:
            luma = synth.damage_stc007(base, seed=rng.randint(1000))
        if kind == 4:
            luma = base.copy()
            a = rng.randint(0, 400)
            luma[rng.randint(4), a:a + rng.randint(20, 300)] = 16
        if kind == 5:
            luma = base.copy()
            luma[rng.randint(4)] = 16
        if kind == 6:
            luma = edge_damage(base, seed=rng.randint(1000), per_field=int(rng.randint(25, 50)))
        sh = int(rng.choice([0, 0, 0, 3, -7, 12, -20, 50, -60, 100]))
        luma = shift_rows(luma, sh)
        rec, _, _ = util.emu_x0_v2d(luma, 2, True)
        n = luma.shape[0]
        for bff, p_corr in ((False, True), (True, True), (False, False)):
            ref = ref_pairs(luma, bff, p_corr, ei=ei)
            smp, fl, al = util.emu_x0_stitch_auto(rec, n, luma.shape[1], bff, p_corr=p_corr, ei=ei)
            ok = ref[0].shape == smp.shape and np.array_equal(ref[0], smp) and np.array_equal(ref[1], fl)
            tag = ("EI" if ei else "SI", it, ctrl, kind, sh, bff, p_corr)
            if not ok:
                n_bad += 1
                bad = [f for f in range(n) if ref[0].shape == smp.shape and not (
                    np.array_equal(ref[0][f * 490:(f + 1) * 490], smp[f * 490:(f + 1) * 490]) and
                    np.array_equal(ref[1][f * 490:(f + 1) * 490], fl[f * 490:(f + 1) * 490]))]
                print("MISMATCH", *tag, ref[0].shape, smp.shape, bad, al.tolist(), flush=True)
            else:
                print("ok", *tag, [(a["result"].tolist(), int(a["mask_seams"])) for a in al], flush=True)
    print("ALL OK" if n_bad == 0 else f"{n_bad} MISMATCHES")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
