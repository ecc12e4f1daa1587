"""Randomised parity sweep for the PCM-1 / PCM-16x0 line decode + chain (device code built for the host) against the
compiled reference: random data coordinates (incl. off-screen), levels, noise, blur, dropouts, modes, duplicate check."""
import sys
import numpy as np
from sdvpcmdecoder_b200 import synth
from oracle import refbind as R
from tests import util

def one(seed):
    rng = np.random.RandomState(seed)
    fmt = "pcm1" if rng.rand() < 0.5 else "pcm16x0"
    mode = int(rng.randint(0, 3))
    dup = bool(rng.rand() < 0.7)
    x0 = int(rng.randint(-14, 40)); x1 = int(720 - rng.randint(-14, 40))
    black = int(rng.randint(5, 60)); white = int(rng.randint(120, 250))
    n = int(rng.randint(1, 4))
    if fmt == "pcm1":
        luma = synth.make_pcm1(n, seed=seed, x0=x0, x1=x1, black=black, white=white, header=bool(rng.rand() < 0.3))["luma"]
    else:
        luma = synth.make_pcm16x0(n, seed=seed, x0=x0, x1=x1, black=black, white=white)["luma"]
    if rng.rand() < 0.8:
        luma = synth.damage_stc007(luma, seed=seed + 1, sigma=float(rng.choice([0., 4., 10., 20.])), jitter=bool(rng.rand() < 0.4),
                                   blur=bool(rng.rand() < 0.4), dropout_frac=float(rng.choice([0., 0.02, 0.1])),
                                   marker_kill_frac=float(rng.choice([0., 0.02])), src_black=black, src_white=white)
    if fmt == "pcm1":
        ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_PCM1, mode, luma, line_dup=dup))[:luma.shape[0] * 480]
        rec, aux, _ = util.emu_p1_v2d(luma, mode, dup)
        bad = util.compare_line_records(ref, rec, aux, oracle_only_flags=1 << 11)
    else:
        ref = R.v2d_run(R.TYPE_PCM16X0, mode, luma, line_dup=dup)
        ref = ref[ref["service_type"] == 0][:luma.shape[0] * 480 * 3]
        rec, aux, _ = util.emu_x0_v2d(luma, mode, dup)
        bad = util.compare_line_records(util.x0_ref_to_product(ref), rec, aux, oracle_only_flags=0)
    print(seed, fmt, "mode", mode, "dup", dup, "x", x0, x1, "bw", black, white, "valid %.3f" % (ref["flags"] & 1).mean(), "OK" if not bad else bad, flush=True)
    return not bad

if __name__ == "__main__":
    a, b = int(sys.argv[1]), int(sys.argv[2])
    ok = all([one(s) for s in range(a, b)])
    print("ALL OK" if ok else "FAILURES")
