import numpy as np, time, sys
from sdvpcmdecoder_b200 import synth
from oracle import refbind as R
from tests import util
def check(name, luma, mode=2, dup=True):
    t = time.time()
    ref = R.v2d_run(R.TYPE_PCM16X0, mode, luma, line_dup=dup)
    ref = ref[ref["service_type"] == 0][:luma.shape[0]*luma.shape[1]*3]
    t1 = time.time()-t
    rec, aux, ps = util.emu_x0_v2d(luma, mode, dup)
    bad = util.compare_line_records(util.x0_ref_to_product(ref), rec, aux, oracle_only_flags=0)
    if not np.array_equal(ref["line_part"], rec["reserved"]): bad.append("line_part")
    print(name, "mode", mode, "dup", dup, "valid %.3f"%(ref["flags"]&1).mean(), "picked", int((ref["mark_st_stage"]>0).sum()), int((ref["mark_ed_stage"]>0).sum()),
          "ctrl0", int(((ref["flags"]>>11)&1==0).sum()), "ref %.2fs"%t1, "OK" if not bad else bad)
    return not bad
ok = True
base = synth.make_pcm16x0(2)["luma"]
for mode in (2, 0, 1):
    ok &= check("clean", base, mode)
    ok &= check("clean", base, mode, dup=False)
    ok &= check("damaged", synth.damage_stc007(base, seed=100+mode), mode)
    ok &= check("noise", synth.damage_stc007(base, seed=200+mode, jitter=False, blur=False, sigma=25., dropout_frac=0.05), mode)
    ok &= check("dropouts", synth.damage_stc007(base, seed=300+mode, jitter=False, blur=False, sigma=3., dropout_frac=0.2), mode)
    ok &= check("cutleft", synth.make_pcm16x0(2, seed=11, x0=-5, x1=710)["luma"], mode)
    ok &= check("cutright", synth.make_pcm16x0(2, seed=12, x0=6, x1=723)["luma"], mode)
    ok &= check("cutboth", synth.damage_stc007(synth.make_pcm16x0(2, seed=13, x0=-7, x1=725)["luma"], seed=5, jitter=False, blur=False, sigma=6., dropout_frac=0.02), mode)
    ok &= check("narrow", synth.make_pcm16x0(2, seed=14, x0=30, x1=690)["luma"], mode)
    ok &= check("wide1440", synth.make_pcm16x0(1, seed=15, width=1440)["luma"], mode)
print("ALL OK" if ok else "FAILURES")
