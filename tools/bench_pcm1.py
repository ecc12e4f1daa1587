import numpy as np, torch, time
from sdvpcmdecoder_b200 import synth, capi, operators as ops
h = capi.Handle(0)
v2d = ops.VideoToDigital(h); v2d.setPCMType(capi.TYPE_PCM1); v2d.setBinarizationMode(2)
base = synth.make_pcm1(50)["luma"]
luma = torch.from_numpy(np.ascontiguousarray(np.tile(base, (20,1,1)))).cuda()   # 1000 frames
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    recs = v2d.doBinarize(luma)
    torch.cuda.synchronize(); dt=time.time()-t
    print("1000 frames: %.2f ms, %.3g lines/s"%(dt*1e3, luma.shape[0]*480/dt), v2d.stats())
