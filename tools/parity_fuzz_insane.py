"""Parity sweep in MODE_INSANE (reference-level sweep on every line that needs it) for STC-007, M2, PCM-1 and PCM-16x0 on
slices of damaged frames (device code built for the host against the compiled reference).  usage: parity_fuzz_insane.py <first seed> <last seed + 1>"""
import sys
import numpy as np
from sdvpcmdecoder_b200 import synth
from oracle import refbind as R
from tests import util
ok=True
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng=np.random.RandomState(seed)
    fmt=["stc007","stc007m2","pcm1","pcm16x0"][seed%4]
    W=int(rng.choice([640,720,960])); m=W/720.0
    x0=int(round(rng.randint(2,24)*m)); x1=int(W-round(rng.randint(2,24)*m))
    black=int(rng.randint(5,60)); white=int(rng.randint(120,250)); dup=bool(rng.rand()<0.7)
    if fmt.startswith("stc007"):
        luma=synth.make_stc007(1,seed=seed,pal=bool(rng.rand()<0.5),width=W,x0=x0,x1=x1,black=black,white=white,quiet_frac=0.2 if fmt.endswith("m2") else 0.0)["luma"]
    elif fmt=="pcm1":
        luma=synth.make_pcm1(1,seed=seed,width=W,x0=x0,x1=x1,black=black,white=white)["luma"]
    else:
        luma=synth.make_pcm16x0(1,seed=seed,width=W,x0=x0,x1=x1,black=black,white=white)["luma"]
    luma=synth.damage_stc007(luma,seed=seed+1,sigma=float(rng.choice([4.,10.])),jitter=True,blur=bool(rng.rand()<0.5),dropout_frac=0.05,marker_kill_frac=0.02,src_black=black,src_white=white)
    luma=luma[:, :120].copy()        # INSANE is slow on the host: a slice of the frame
    H=luma.shape[1]
    if fmt=="stc007":
        ref=util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007,3,luma,line_dup=dup))[:H]; rec,aux,_=util.emu_v2d(luma,3,dup); bad=util.compare_line_records(ref,rec,aux)
    elif fmt=="stc007m2":
        ref=util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_M2,3,luma,line_dup=dup))[:H]; rec,aux,_=util.emu_v2d(luma,3,dup,m2=True); bad=util.compare_line_records(ref,rec,aux)
    elif fmt=="pcm1":
        ref=util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_PCM1,3,luma,line_dup=dup))[:H]; rec,aux,_=util.emu_p1_v2d(luma,3,dup); bad=util.compare_line_records(ref,rec,aux,oracle_only_flags=1<<11)
    else:
        ref=R.v2d_run(R.TYPE_PCM16X0,3,luma,line_dup=dup); ref=ref[ref["service_type"]==0][:H*3]; rec,aux,_=util.emu_x0_v2d(luma,3,dup); bad=util.compare_line_records(util.x0_ref_to_product(ref),rec,aux,oracle_only_flags=0)
    print(seed,fmt,W,dup,"valid %.3f"%(ref["flags"]&1).mean(),"OK" if not bad else bad,flush=True)
    ok&=not bad
print("ALL OK" if ok else "FAILURES")
