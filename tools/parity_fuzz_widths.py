"""Randomised parity sweep over frame WIDTHS (and data coordinates scaled with them) for all three line decoders + chains
(device code built for the host) against the compiled reference.  The reference's ingest doubles narrow frames, so 1440
and other widths are real inputs; every pixel-per-bit dependent constant of the search is exercised here."""
import sys
import numpy as np
from sdvpcmdecoder_b200 import synth
from oracle import refbind as R
from tests import util

WIDTHS = [320, 352, 480, 640, 704, 720, 768, 960, 1024, 1280, 1440, 1920]

def one(seed):
    rng = np.random.RandomState(seed)
    fmt = ["stc007", "pcm1", "pcm16x0"][seed % 3]
    W = int(WIDTHS[rng.randint(0, len(WIDTHS))])
    mode = int(rng.randint(0, 3))
    dup = bool(rng.rand() < 0.7)
    m = W / 720.0
    x0 = int(round(rng.randint(-6, 30) * m)); x1 = int(W - round(rng.randint(-6, 30) * m))
    black = int(rng.randint(5, 60)); white = int(rng.randint(120, 250))
    n = int(rng.randint(1, 3))
    if fmt == "stc007":
        x0 = max(x0, 2); x1 = min(x1, W - 3)
        luma = synth.make_stc007(n, seed=seed, pal=bool(rng.rand() < 0.5), width=W, x0=x0, x1=x1, black=black, white=white)["luma"]
    elif fmt == "pcm1":
        luma = synth.make_pcm1(n, seed=seed, width=W, x0=x0, x1=x1, black=black, white=white, header=bool(rng.rand() < 0.3))["luma"]
    else:
        luma = synth.make_pcm16x0(n, seed=seed, width=W, x0=x0, x1=x1, black=black, white=white)["luma"]
    if rng.rand() < 0.7:
        luma = synth.damage_stc007(luma, seed=seed + 1, sigma=float(rng.choice([0., 4., 10.])), jitter=bool(rng.rand() < 0.4),
                                   blur=bool(rng.rand() < 0.4), dropout_frac=float(rng.choice([0., 0.02, 0.1])),
                                   marker_kill_frac=float(rng.choice([0., 0.02])), src_black=black, src_white=white)
    H = luma.shape[1]
    if fmt == "stc007":
        ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, mode, luma, line_dup=dup))[:luma.shape[0] * H]
        rec, aux, _ = util.emu_v2d(luma, mode, dup)
        bad = util.compare_line_records(ref, rec, aux)
    elif fmt == "pcm1":
        ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_PCM1, mode, luma, line_dup=dup))[:luma.shape[0] * H]
        rec, aux, _ = util.emu_p1_v2d(luma, mode, dup)
        bad = util.compare_line_records(ref, rec, aux, oracle_only_flags=1 << 11)
    else:
        ref = R.v2d_run(R.TYPE_PCM16X0, mode, luma, line_dup=dup)
        ref = ref[ref["service_type"] == 0][:luma.shape[0] * H * 3]
        rec, aux, _ = util.emu_x0_v2d(luma, mode, dup)
        bad = util.compare_line_records(util.x0_ref_to_product(ref), rec, aux, oracle_only_flags=0)
    print(seed, fmt, "W", W, "mode", mode, "dup", dup, "x", x0, x1, "bw", black, white, "valid %.3f" % (ref["flags"] & 1).mean(), "OK" if not bad else bad, flush=True)
    return not bad

if __name__ == "__main__":
    a, b = int(sys.argv[1]), int(sys.argv[2])
    ok = all([one(s) for s in range(a, b)])
    print("ALL OK" if ok else "FAILURES")
