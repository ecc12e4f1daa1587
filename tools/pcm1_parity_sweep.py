import numpy as np, time, sys
from sdvpcmdecoder_b200 import synth
from oracle import refbind as R
from tests import util
util.ORACLE_ONLY_FLAGS = np.uint16(1<<11)
def check(name, luma, mode=2, dup=True):
    ref = util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_PCM1, mode, luma, line_dup=dup))[:luma.shape[0]*luma.shape[1]]
    rec, aux, ps = util.emu_p1_v2d(luma, mode, dup)
    bad = util.compare_line_records(ref, rec, aux)
    print(name, "mode", mode, "dup", dup, "valid %.3f"%(ref["flags"]&1).mean(), "hdr", int((ref["service_type"]==6).sum()), "picked", int((ref["mark_st_stage"]>0).sum()), "OK" if not bad else bad)
    return not bad
ok = True
base = synth.make_pcm1(2)["luma"]
for mode in (0,1,2):
    ok &= check("clean", base, mode)
    ok &= check("clean", base, mode, dup=False)
    ok &= check("header", synth.make_pcm1(2, seed=7, header=True)["luma"], mode)
    ok &= check("damaged", synth.damage_stc007(base, seed=100+mode), mode)
    ok &= check("noise", synth.damage_stc007(base, seed=200+mode, jitter=False, blur=False, sigma=25., dropout_frac=0.05), mode)
    ok &= check("cutleft", synth.make_pcm1(2, seed=11, x0=-9, x1=705)["luma"], mode)
    ok &= check("cutright", synth.make_pcm1(2, seed=12, x0=10, x1=726)["luma"], mode)
    ok &= check("cutboth", synth.damage_stc007(synth.make_pcm1(2, seed=13, x0=-12, x1=728)["luma"], seed=5, jitter=False, blur=False, sigma=6., dropout_frac=0.02), mode)
    ok &= check("narrow", synth.make_pcm1(2, seed=14, x0=30, x1=690)["luma"], mode)
print("ALL OK" if ok else "FAILURES")
