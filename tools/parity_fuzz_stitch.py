#!/usr/bin/env python
"""Random STC-007 tapes through the stitcher (own alignment, audio-resolution detection, Cross-Word Decoding) -- host build of the
device code + the library's decision chains -- against the unmodified reference pipeline.  usage: parity_fuzz_stitch.py first_seed count"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import refbind as R
from sdvpcmdecoder_b200 import synth
from tests import util
from tests.test_stc007_stitch import stream_mismatch, block_mismatch, shift_rows

first, count = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(first, first + count):
    rng = np.random.RandomState(seed)
    pal = rng.rand() < 0.7
    f16 = rng.rand() < 0.4
    frames = int(rng.randint(3, 6))
    t = synth.make_stc007(frames, seed=seed, pal=pal, f1_16bit=f16)
    luma = t["luma"]
    if rng.rand() < 0.3:
        luma = shift_rows(luma, int(rng.randint(-3, 4)))
    kind = rng.randint(0, 6)
    if kind == 0:
        luma = synth.damage_stc007(luma, seed=seed + 1)
    elif kind == 1:
        luma = synth.damage_stc007(luma, seed=seed + 1, sigma=float(rng.uniform(10, 24)), dropout_frac=float(rng.uniform(0.02, 0.15)), marker_kill_frac=float(rng.uniform(0, 0.06)))
    elif kind == 2:
        luma = synth.damage_stc007(luma, seed=seed + 1, sigma=float(rng.uniform(2, 6)), dropout_frac=float(rng.uniform(0.1, 0.35)), marker_kill_frac=0.0)
    elif kind == 3:
        luma = luma.copy()
        f = int(rng.randint(0, frames))
        luma[f] = synth.damage_stc007(luma[f:f + 1], seed=seed + 1, sigma=4.0, dropout_frac=0.3)[0]
    elif kind == 4:
        # CRCs broken (levels and markers kept) in the first / last line of every field and in random lines: the number of good
        # lines per field lands around the trim's MIN_GOOD_LINES_PF threshold
        from tests.test_pcm16x0_stitch import edge_damage
        luma = edge_damage(luma, seed=seed + 1, per_field=int(rng.randint(3, 90)))
    else:
        luma = shift_rows(luma, int(rng.choice([-40, -12, -6, 5, 9, 20, 60])))
    if rng.rand() < 0.15:
        luma = luma.copy(); luma[int(rng.randint(0, frames))] = 16
    std = int(rng.choice([0, 1 if pal else 2]))
    order = int(rng.choice([0, 1]))
    res = int(rng.choice([0, 2 if f16 else 1]))
    p, q = (1, 1) if rng.rand() < 0.8 else (1, 0)
    cwd = int(rng.rand() < 0.7)
    cfg = R.StitchCfg()
    cfg.video_std, cfg.field_order, cfg.resolution, cfg.p_corr, cfg.q_corr, cfg.cwd = std, order, res, p, q, cwd
    pairs, _, ref_blocks = R.pipeline_run(R.TYPE_STC007, R.MODE_NORMAL, luma, cfg)
    pairs = pairs[pairs["service_type"] == 0]
    recs = util.lines_from_oracle(util.ref_lines_in_frame_order(R.v2d_run(R.TYPE_STC007, R.MODE_NORMAL, luma), keep=(0, 7)))
    blocks, samples, flags, info = util.emu_stc007_stitch(recs, luma.shape[0], luma.shape[1], video_std=std, field_order=order,
                                                          res16={0: None, 1: False, 2: True}[res], p_corr=bool(p), q_corr=bool(q), cwd=bool(cwd))
    m1, m2 = stream_mismatch(pairs, samples, flags), block_mismatch(ref_blocks, blocks)
    tag = f"seed {seed} pal={pal} f16={f16} frames={frames} kind={kind} std={std} order={order} res={res} pq={p}{q} cwd={cwd}"
    if m1 or m2:
        bad += 1
        print("MISMATCH", tag, m1, m2, flush=True)
    else:
        print("ok", tag, flush=True)
print("mismatches:", bad)
