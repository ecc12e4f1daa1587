"""Per-phase device time of the config-5 step on N GPUs (torchrun): line decode / halo exchange / to-samples."""
import os, json, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sdvpcmdecoder_b200 import capi, operators, sharding

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
F, H, W, period = 90000, 576, 720, 60
a, b = sharding.frame_range(F, rank, world); n = b - a
seg = bench.make_segment(period)
luma = torch.from_numpy(seg["luma"]).to(dev)[(torch.arange(a, b, device=dev) % period)]
h = capi.Handle(local); v2d = operators.VideoToDigital(h); st = operators.STC007DataStitcher(h); st.lead_in = sharding.shard_lead_in(rank)
nb = st.block_count(n)
recs = torch.empty((n * H, 32), dtype=torch.uint8, device=dev); samples = torch.empty((nb, 6), dtype=torch.int16, device=dev); flags = torch.empty((nb, 6), dtype=torch.uint8, device=dev)
halo = torch.zeros((112, 32), dtype=torch.uint8, device=dev) if rank < world - 1 else None
mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
gath = torch.zeros((world, 112, 32), dtype=torch.uint8, device=dev)
def exchange():
    if world == 1: return None
    if mode == "p2p": return sharding.exchange_halo(recs, halo, rank, world)
    dist.all_gather_into_tensor(gath, recs[:112].contiguous())
    return gath[rank + 1] if rank < world - 1 else None
K = 20
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
for it in range(K + 3):
    e = ev[max(it - 3, 0)]
    e[0].record(); v2d.doBinarize(luma, out=recs); e[1].record(); hh = exchange(); e[2].record(); st.doFrameReassemble(recs, n, H, samples=samples, flags=flags, halo=hh); e[3].record()
torch.cuda.synchronize()
if world > 1: dist.barrier()
d = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in ev]); tot = np.array([ev[i][0].elapsed_time(ev[i + 1][0]) for i in range(K - 1)])
print(json.dumps({"rank": rank, "mode": mode, "decode_ms": d[:, 0].mean(), "exchange_ms": d[:, 1].mean(), "samples_ms": d[:, 2].mean(), "step_ms": tot.mean()}), flush=True)
if world > 1: dist.destroy_process_group()
