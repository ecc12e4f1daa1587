#!/usr/bin/env python
"""Config 4 (damaged STC-007 tape: gain/offset jitter, noise, blur, dropouts, killed markers): GPU chain path vs the
reference's VideoToDigital on the host, with a full-field parity check against the oracle."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sdvpcmdecoder_b200 import capi, operators, synth
from oracle import oraclebind as O, refbind as R
from tests import util

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 30
luma = synth.damage_stc007(synth.make_stc007(frames, seed=4)["luma"], seed=4567)
h = capi.Handle(0)
v2d = operators.VideoToDigital(h)
dev = torch.from_numpy(luma).cuda()
recs, aux = v2d.doBinarize(dev, want_aux=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    recs, aux = v2d.doBinarize(dev, want_aux=True)
torch.cuda.synchronize()
gpu_s = (time.perf_counter() - t0) / 3
st = v2d.stats()
t0 = time.perf_counter()
o = O.v2d_stc007(2, luma, True)
cpu_oracle_s = time.perf_counter() - t0
ref_s = None
if R.available():
    t0 = time.perf_counter()
    R.v2d_run(R.TYPE_STC007, R.MODE_NORMAL, luma)
    ref_s = time.perf_counter() - t0
bad = util.compare_line_records(o, operators.records_to_numpy(recs, capi.LINE_REC), operators.records_to_numpy(aux, capi.LINE_AUX))
n = frames * 576
print(json.dumps({"frames": frames, "lines": n, "gpu_lines_per_s": n / gpu_s, "gpu_ms": gpu_s * 1e3,
                  "oracle_1thread_lines_per_s": n / cpu_oracle_s, "reference_v2d_1thread_lines_per_s": (n / ref_s) if ref_s else None,
                  "lines_swept": st["reserved"], "lines_chain": st["lines_chain"], "valid_frac": float((o["flags"] & 1).mean()),
                  "mismatches": bad}))
