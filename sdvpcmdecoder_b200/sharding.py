"""Frame-range sharding of a tape over the ranks of one node (one process per GPU, torch.distributed).

The decode path shards by contiguous frame ranges.  Line decode needs nothing from the neighbours (each shard starts
its VideoToDigital chain empty, exactly like the reference at file start); the STC-007 deinterleaver reads 112 lines
ahead (7 x 16 lines, stc007datablock.h:44-58), so shard g needs the first 112 line records of shard g+1: one small
point-to-point message per boundary (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests).  The one thing that
flows the other way is the stitcher's broken-block countdown: a BROKEN block near the end of shard g masks the first blocks
of shard g+1 (carry_countdowns: one 8-byte all_gather per decode, a second round only for the shards it affects).
Samples stay sharded.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

HALO_LINES = 112            # STC007DataBlock::MIN_DEINT_DATA
LEAD_IN_LINES = 80          # lines queued ahead of the first frame of a file (stc007datastitcher.cpp:4733-4737)


def frame_range(n_frames: int, rank: int, world: int):
    """Frames [a, b) of the tape decoded by this rank."""
    return rank * n_frames // world, (rank + 1) * n_frames // world


def shard_lead_in(rank: int) -> int:
    """Only the first shard carries the file's lead-in lines."""
    return LEAD_IN_LINES if rank == 0 else 0


def first_block(n_frames: int, rank: int, world: int, lines_per_field: int) -> int:
    """Index, in the unsharded block stream, of this rank's first data block."""
    a, _ = frame_range(n_frames, rank, world)
    return 0 if rank == 0 else LEAD_IN_LINES + a * 2 * lines_per_field


def block_count(n_frames: int, rank: int, world: int, lines_per_field: int) -> int:
    a, b = frame_range(n_frames, rank, world)
    return shard_lead_in(rank) + (b - a) * 2 * lines_per_field


def exchange_halo(recs: torch.Tensor, halo: torch.Tensor | None, rank: int, world: int):
    """Send this shard's first 112 line records to the previous rank, receive the next rank's into [halo].

    recs: [n_lines, 32] uint8 record buffer of this shard; halo: [112, 32] uint8 (None on the last rank).
    Returns [halo] (None on the last rank)."""
    return exchange_halo_finish(exchange_halo_start(recs, halo, rank, world), halo, rank, world)


def exchange_halo_start(recs: torch.Tensor, halo: torch.Tensor | None, rank: int, world: int):
    """First half of exchange_halo(): post the send / receive and return the requests.  Only the shard's first 112 line
    records need to be final -- VideoToDigital.doBinarize(on_first_frame=...) calls back at that point, while the rest of
    the shard is still being decoded."""
    if world == 1:
        return []
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, recs[:HALO_LINES], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.irecv, halo, rank + 1))
    return dist.batch_isend_irecv(ops) if ops else []


def exchange_halo_finish(reqs, halo: torch.Tensor | None, rank: int, world: int):
    for r in reqs:
        r.wait()
    return halo if (world > 1 and rank < world - 1) else None


def carry_countdowns(state: dict, redo, rank: int, world: int, device=None) -> dict:
    """Hand the broken-block countdown from shard to shard.

    STC007DataStitcher::broken_countdown is a member (stc007datastitcher.cpp:79, 6785-6863): a BROKEN block within the last
    broken_mask_dur blocks of shard g masks the first blocks of shard g+1.  Every rank runs its deinterleave pass first with
    countdown_in = 0 (all ranks at once); [state] is what the library reports for that pass (STC007DataStitcher.countdown():
    countdown_in / countdown_out / depends_on_in).  Then, the same on every rank: gather all countdown_out; the true input
    of rank g is the output of rank g-1; ranks whose input was something else call redo(countdown_in) -> new state (the
    library only redoes the blocks inside the windows) and the gather repeats.  Rank 0's output is final from the start,
    rank 1's after one round, ...: at most [world] rounds, one round (a single 8-byte all_gather) on a tape without
    BROKEN blocks near the shard boundaries.  Returns the final state of this rank."""
    if world == 1:
        return state
    used = [0] * world
    for _ in range(world + 1):
        mine = torch.tensor([int(state["countdown_out"]), int(bool(state["depends_on_in"]))], dtype=torch.int32, device=device)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        outs = [int(v[0]) for v in allv]
        want = [0] + outs[:-1]
        changed = [g for g in range(world) if want[g] != used[g]]
        if not changed:
            return state
        if rank in changed:
            state = redo(want[rank])
        used = want
    raise RuntimeError("countdown hand-off did not settle")


def carry_countdowns_device(state_dev: torch.Tensor, gathered: torch.Tensor, redo, rank: int, world: int) -> int:
    """carry_countdowns() for the GPU path: [state_dev] is the int32[4] device tensor STC007DataStitcher.countdown_to() filled
    (countdown_in, countdown_out, windows, depends), [gathered] an int32 [world, 4] device tensor.  One all_gather over NCCL
    straight from device memory and one read-back per round; redo(countdown_in) runs the deinterleave call again and refills
    [state_dev].  Returns the number of rounds in which this rank had to redo its windows (0 on a tape without BROKEN blocks at
    the shard boundaries)."""
    if world == 1:
        return 0
    used = [0] * world
    redone = 0
    for _ in range(world + 1):
        dist.all_gather_into_tensor(gathered.view(-1), state_dev.view(-1)[:4].contiguous())
        outs = [int(v) for v in gathered.view(world, 4)[:, 1].cpu()]
        want = [0] + outs[:-1]
        changed = [g for g in range(world) if want[g] != used[g]]
        if not changed:
            return redone
        if rank in changed:
            redo(want[rank])
            redone += 1
        used = want
    raise RuntimeError("countdown hand-off did not settle")


class CountdownHandoff:
    """carry_countdowns_device() without a host round trip in the step: post() gathers the countdown states over NCCL and starts
    their copy into pinned host memory (nothing blocks); settle() -- called when the host synchronises with the device anyway,
    e.g. at the start of the next step, and before the results are used -- looks at the copy and runs the redo rounds if a
    shard's predecessor left a window open (never on a tape without BROKEN blocks at the shard boundaries)."""

    def __init__(self, rank: int, world: int, device):
        self.rank, self.world = rank, world
        self.gathered = torch.zeros((world, 4), dtype=torch.int32, device=device)
        self.host = torch.zeros((world, 4), dtype=torch.int32).pin_memory()
        self.event = torch.cuda.Event()
        self.pending = None
        self.redone = 0

    def post(self, state_dev: torch.Tensor, redo):
        if self.world == 1:
            return
        dist.all_gather_into_tensor(self.gathered.view(-1), state_dev.view(-1)[:4].contiguous())
        self.host.copy_(self.gathered, non_blocking=True)
        self.event.record()
        self.pending = (state_dev, redo)

    def settle(self):
        if self.pending is None:
            return
        state_dev, redo = self.pending
        self.pending = None
        self.event.synchronize()
        outs = [int(v) for v in self.host[:, 1]]
        used = [0] * self.world
        for _ in range(self.world + 1):
            want = [0] + outs[:-1]
            changed = [g for g in range(self.world) if want[g] != used[g]]
            if not changed:
                return
            if self.rank in changed:
                redo(want[self.rank])
                self.redone += 1
            used = want
            dist.all_gather_into_tensor(self.gathered.view(-1), state_dev.view(-1)[:4].contiguous())
            outs = [int(v) for v in self.gathered[:, 1].cpu()]
        raise RuntimeError("countdown hand-off did not settle")


def bind_to_gpu_numa_node(device_index: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that host buffers allocated afterwards (the
    pinned luma / sample buffers of the host entry points) are local to the GPU's PCIe root: with one process per GPU
    every H2D stream then reads its own socket's memory.  Returns the node, or None when the platform does not say
    (no sysfs entry, a single node, or a VM that hides the topology) -- nothing is changed in that case."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
