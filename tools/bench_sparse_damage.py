"""PCM-1 / PCM-16x0 tapes with one damaged line in every frame (what real captures look like): every frame goes through
the chain kernel, which takes the bulk pass's records as hints for the lines that did decode with the current presets."""
import json
import sys
import time

import numpy as np
import torch

from sdvpcmdecoder_b200 import synth, capi, operators as ops

for fmt in (sys.argv[1:] or ["pcm1", "pcm16x0"]):
    h = capi.Handle(0)
    base = (synth.make_pcm1(50) if fmt == "pcm1" else synth.make_pcm16x0(50))["luma"]
    luma = np.tile(base, (4, 1, 1)).copy()
    rng = np.random.RandomState(7)
    for f in range(luma.shape[0]):
        r = rng.randint(4, 476)
        luma[f, r, 200:420] = 235
    t = torch.from_numpy(luma).cuda()
    v2d = ops.VideoToDigital(h)
    v2d.setPCMType(capi.TYPE_PCM1 if fmt == "pcm1" else capi.TYPE_PCM16X0)
    best = 1e9
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        recs = v2d.doBinarize(t)
        torch.cuda.synchronize()
        best = min(best, time.time() - t0)
    st = v2d.stats()
    print(json.dumps({"format": fmt, "frames": luma.shape[0], "ms": best * 1e3, "lines_per_s": luma.shape[0] * 480 / best,
                      "frames_from_bulk": st["frames_skipped"], "lines_chain": st["lines_chain"], "chain_lines_hinted": st["reserved"]}))
