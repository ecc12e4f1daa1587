"""PCM-16x0 'damaged' golden case (2 frames, most sub-lines fail the preset decode) through the CUDA path, a few times: the
workload of the pcm16x0_chain_kernel entry in the smoke() launch list.  For ncu captures."""
import sys
import time

import numpy as np
import torch

from sdvpcmdecoder_b200 import capi, operators as ops
from tests.test_pcm16x0_line import pcm16x0_cases

luma = torch.from_numpy(np.ascontiguousarray(pcm16x0_cases()[sys.argv[1] if len(sys.argv) > 1 else "damaged"])).cuda()
v2d = ops.VideoToDigital(capi.Handle(0))
v2d.setPCMType(capi.TYPE_PCM16X0)
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    v2d.doBinarize(luma)
    torch.cuda.synchronize()
    print("ms", (time.time() - t0) * 1e3, v2d.stats())
