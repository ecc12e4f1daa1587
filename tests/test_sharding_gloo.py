"""Frame-range sharding + halo exchange with two CPU processes (gloo): the sharded decode must reproduce the unsharded
sample stream exactly.  Line decode and deinterleave run through the host-compiled device logic (tests/hostemu);
the exchange and the index arithmetic are the product's own (sdvpcmdecoder_b200/sharding.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FRAMES, LPF, H = 6, 294, 576


def _assemble(recs, n_frames, lead_in, halo):
    from sdvpcmdecoder_b200.capi import LINE_REC
    hf = H // 2
    nb = lead_in + n_frames * 2 * LPF
    asm = np.zeros(nb + 112, LINE_REC)
    for fld in range(2 * n_frames):
        src = (fld // 2) * H + (fld & 1) * hf
        asm[lead_in + fld * LPF:lead_in + fld * LPF + hf] = recs[src:src + hf]
    if halo is not None:
        asm[nb:nb + 112] = halo
    return asm, nb


def _tamper(recs, first_frame, n_frames):
    """Give a few lines the words of another valid line (CRC flags stay valid): every block they feed fails its parity check with
    no CRC error to blame -> BROKEN, which opens the 128-block countdown.  Lines chosen so that windows straddle the boundary
    between two shards (frame 3 of 6) and sit elsewhere on the tape; applied by global line index, so that the sharded and the
    unsharded run see the same tape."""
    hf = H // 2
    for (frame, field, j) in ((2, 1, 285), (1, 0, 100), (4, 1, 20)):
        if first_frame <= frame < first_frame + n_frames:
            i = (frame - first_frame) * H + field * hf + j
            recs["words"][i] = recs["words"][i - 7]
    return recs


def _worker(rank, world, port, out_dir, tamper=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdvpcmdecoder_b200 import synth, sharding
    from sdvpcmdecoder_b200.capi import LINE_REC
    from tests import util
    luma = synth.make_stc007(N_FRAMES, seed=77)["luma"]
    a, b = sharding.frame_range(N_FRAMES, rank, world)
    recs, _, _ = util.emu_v2d(luma[a:b], 2, True, hybrid=True)           # each shard starts its chain empty
    if tamper:
        recs = _tamper(recs, a, b - a)
    rt = torch.from_numpy(recs.view(np.uint8).reshape(-1, 32).copy())
    halo_t = torch.zeros((sharding.HALO_LINES, 32), dtype=torch.uint8) if rank < world - 1 else None
    # the product posts the exchange from the first-frame hook and collects it after the decode: same two halves here
    reqs = sharding.exchange_halo_start(rt, halo_t, rank, world)
    got = sharding.exchange_halo_finish(reqs, halo_t, rank, world)
    halo = got.numpy().reshape(-1).view(LINE_REC) if got is not None else None
    asm, nb = _assemble(recs, b - a, sharding.shard_lead_in(rank), halo)
    assert nb == sharding.block_count(N_FRAMES, rank, world, LPF)
    # every shard first with no countdown carried in, then the hand-off from shard to shard (sharding.carry_countdowns)
    res = {}

    def run(countdown_in):
        _, res["s"], res["f"], out = util.emu_deint_carry(asm, countdown_in)
        return {"countdown_in": countdown_in, "countdown_out": out, "depends_on_in": True}
    state = sharding.carry_countdowns(run(0), run, rank, world)
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"), s=res["s"], f=res["f"], first=sharding.first_block(N_FRAMES, rank, world, LPF),
             countdown_in=state["countdown_in"], countdown_out=state["countdown_out"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("tamper", [False, True])
def test_two_shards_equal_unsharded(tmp_path, tamper):
    """tamper: BROKEN blocks shortly before the shard boundary -- the countdown they open has to reach into the next shard."""
    sys.path.insert(0, ROOT)
    from sdvpcmdecoder_b200 import synth, sharding
    from tests import util
    port = 29500 + (os.getpid() % 400) + (400 if tamper else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), tamper), nprocs=2, join=True)
    luma = synth.make_stc007(N_FRAMES, seed=77)["luma"]
    recs, _, _ = util.emu_v2d(luma, 2, True, hybrid=True)
    if tamper:
        recs = _tamper(recs, 0, N_FRAMES)
    asm, nb = _assemble(recs, N_FRAMES, sharding.LEAD_IN_LINES, None)
    _, s, f = util.emu_deint(asm, 0, False, True, True, True, 128)
    pos = 0
    for r in range(2):
        g = np.load(os.path.join(str(tmp_path), f"shard{r}.npz"))
        assert int(g["first"]) == pos
        n = len(g["s"])
        assert np.array_equal(g["s"], s[pos:pos + n]) and np.array_equal(g["f"], f[pos:pos + n]), f"shard {r}"
        pos += n
        if tamper and r == 1:
            assert int(g["countdown_in"]) > 0, "the test tape was meant to carry a countdown across the boundary"
    assert pos == nb


def test_frame_ranges_cover_the_tape():
    from sdvpcmdecoder_b200 import sharding
    for n in (1, 7, 90000):
        for world in (1, 2, 4, 8):
            edges = [sharding.frame_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            assert sum(sharding.block_count(n, r, world, 294) for r in range(world)) == 80 + n * 588


# ---- PCM-1 and PCM-16x0 (SI): frames are deinterleaved field by field, so the shards need no halo at all; only the first
# shard opens the file (PCM1DataStitcher forgets the header alignment of the file's first frame).
def _decode_fmt(fmt, luma, file_start):
    from oracle import oraclebind as O
    from tests import util
    n, h = luma.shape[0], luma.shape[1]
    if fmt == "pcm1":
        rec, _, _ = util.emu_p1_v2d(luma, 2, True)
        sub, _ = util.emu_p1_assemble(rec, n, h, False, file_start)
        s, f = O.deint_pcm1(np.stack([sub["left"], sub["right"]], axis=1), sub["flags"])
        return s.reshape(-1), f.reshape(-1)
    rec, _, _ = util.emu_x0_v2d(luma, 2, True)
    s, f = util.emu_x0_stitch(rec, n, h)
    return s.reshape(-1), f.reshape(-1)


def _fmt_tape(fmt):
    from sdvpcmdecoder_b200 import synth
    return synth.make_pcm1(4, seed=31, header=True)["luma"] if fmt == "pcm1" else synth.make_pcm16x0(4, seed=32)["luma"]


def _worker_fmt(rank, world, port, out_dir, fmt):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdvpcmdecoder_b200 import sharding
    luma = _fmt_tape(fmt)
    a, b = sharding.frame_range(luma.shape[0], rank, world)
    s, f = _decode_fmt(fmt, luma[a:b], file_start=(rank == 0))
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(s)], dtype=torch.int64))          # bookkeeping only: nothing on the data path
    assert sum(int(c) for c in counts) == luma.shape[0] * 2 * 1470
    np.savez(os.path.join(out_dir, f"{fmt}{rank}.npz"), s=s, f=f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("fmt", ["pcm1", "pcm16x0"])
def test_two_shards_equal_unsharded_pcm1_pcm16x0(tmp_path, fmt):
    sys.path.insert(0, ROOT)
    port = 29900 + (os.getpid() % 90) + (0 if fmt == "pcm1" else 1)
    mp.spawn(_worker_fmt, args=(2, port, str(tmp_path), fmt), nprocs=2, join=True)
    s, f = _decode_fmt(fmt, _fmt_tape(fmt), True)
    gs = np.concatenate([np.load(os.path.join(str(tmp_path), f"{fmt}{r}.npz"))["s"] for r in range(2)])
    gf = np.concatenate([np.load(os.path.join(str(tmp_path), f"{fmt}{r}.npz"))["f"] for r in range(2)])
    assert np.array_equal(gs, s) and np.array_equal(gf, f)
