#!/usr/bin/env python
"""Config 4 (damaged STC-007 tape) timing split: line decode (relay mode) and the stitcher with its own alignment."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from sdvpcmdecoder_b200 import capi, operators, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
luma = torch.from_numpy(synth.damage_stc007(synth.make_stc007(frames, seed=4)["luma"], seed=4567)).cuda()
h = capi.Handle(0)
v2d = operators.VideoToDigital(h)
st = operators.STC007DataStitcher(h)


def t(fn, reps=3):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out.append(time.perf_counter() - t0)
    return float(np.median(out)) * 1e3, r


ms_dec, recs = t(lambda: v2d.doBinarize(luma))
stats = v2d.stats()
ms_auto, _ = t(lambda: st.doFrameReassembleAuto(recs, frames, 576, video_std=1))
ms_preset, _ = t(lambda: st.doFrameReassemble(recs, frames, 576))
st.setCWDCorrection(True)
ms_cwd, _ = t(lambda: st.doFrameReassembleAuto(recs, frames, 576, video_std=1))
cwd_stats = h.last_stats()
st.setCWDCorrection(False)
v2d.relay = False
ms_seq = None
if frames <= 64:
    ms_seq, _ = t(lambda: v2d.doBinarize(luma), 1)
print(json.dumps({"frames": frames, "relay_len": os.environ.get("SDV_RELAY_LEN"), "decode_ms": ms_dec, "decode_lines_per_s": frames * 576 / ms_dec * 1e3,
                  "stitch_auto_ms": ms_auto, "stitch_preset_ms": ms_preset, "stitch_auto_cwd_ms": ms_cwd, "cwd_chains": cwd_stats["reserved"], "cwd_redo_rounds": cwd_stats["frames_skipped"], "relay_pieces": stats["reserved"] >> 16, "relay_redone": stats["reserved"] & 0xFFFF,
                  "sequential_ms": ms_seq}))
