"""Fuzz of the PCM-1 path to samples (line decode + chain + PCM1DataStitcher + deinterleave) on the host build of the device code
against the reference pipeline (oracle/_ref): random tapes (header line shown or not, blanked spans / bands / frames, broken CRCs
at the field edges, noise, heavy damage, vertical shifts), TFF and BFF; the whole PCMSamplePair stream must be equal.
    python tools/parity_fuzz_p1_stitch.py [seed] [cases]"""
import sys

import numpy as np

from oracle import refbind as R
from sdvpcmdecoder_b200 import synth
from tests import util
from tests.test_pcm1_line import ref_samples, emu_samples
from tests.test_pcm16x0_stitch import variant_b, shift_rows, edge_damage


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    assert R.available(), "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.RandomState(seed)
    n_bad = 0
    for it in range(cases):
        header = bool(rng.randint(2))
        base = synth.make_pcm1(4, seed=rng.randint(1 << 20), header=header)["luma"]
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
            luma = edge_damage(base, seed=rng.randint(1000), per_field=int(rng.randint(25, 60)))
        sh = int(rng.choice([0, 0, 0, 1, 3, -2, -7, 12, -20, 50, -60]))
        luma = shift_rows(luma, sh)
        rec, _, _ = util.emu_p1_v2d(luma, 2, True)
        n = luma.shape[0]
        for bff in (False, True):
            ref = ref_samples(luma, 2, bff)
            smp, fl, info = emu_samples(rec, n, luma.shape[1], bff)
            ok = ref[0].shape == smp.shape and np.array_equal(ref[0], smp) and np.array_equal(ref[1], fl)
            tag = (it, "header" if header else "plain", kind, sh, bff)
            if not ok:
                n_bad += 1
                bad = []
                if ref[0].shape == smp.shape:
                    per = len(smp) // n
                    bad = [f for f in range(n) if not (np.array_equal(ref[0][f * per:(f + 1) * per], smp[f * per:(f + 1) * per]) and
                                                       np.array_equal(ref[1][f * per:(f + 1) * per], fl[f * per:(f + 1) * per]))]
                print("MISMATCH", *tag, ref[0].shape, smp.shape, bad, info.tolist(), flush=True)
            else:
                print("ok", *tag, flush=True)
    print("ALL OK" if n_bad == 0 else f"{n_bad} MISMATCHES")
    return 1 if n_bad else 0


if __name__ == "__main__":
    sys.exit(main())
