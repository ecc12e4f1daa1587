import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from sdvpcmdecoder_b200 import capi, operators, synth
frames=int(sys.argv[1]); S=int(sys.argv[2])
base=synth.damage_stc007(synth.make_stc007(30, seed=4)["luma"], seed=4567)
luma=np.concatenate([base]*((frames+29)//30))[:frames]
h=capi.Handle(0); v2d=operators.VideoToDigital(h); v2d.chain_segments=S
dev=torch.from_numpy(luma).cuda()
recs=v2d.doBinarize(dev); torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(2): recs=v2d.doBinarize(dev)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/2
r=operators.records_to_numpy(recs, capi.LINE_REC)
print(json.dumps({"frames":frames,"segments":S,"lines_per_s":frames*576/dt,"ms":dt*1e3,"valid_frac":float((r['flags']&1).mean())}))
