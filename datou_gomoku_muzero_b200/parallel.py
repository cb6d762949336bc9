"""Multi-GPU plumbing: one process per GPU, games sharded across ranks, no collective on the
per-simulation path.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is used
only for the two exchanges the reference does through multiprocessing queues:

  * trainer -> workers weight publication (`model_update_queue`, workers.py:587-593, 331-335)
    -> `broadcast_weights`: one flattened-buffer broadcast from the trainer rank;
  * workers -> data loader trajectories (`data_queue`, workers.py:230, 399)
    -> `gather_finished_games`: counts all-gathered, records sent to the collecting rank.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_games(total_games: int, world_size: int, rank: int):
    """Contiguous shard [start, stop) of the global game indices owned by `rank` (sizes differ by <= 1)."""
    base, extra = divmod(int(total_games), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def rank_noise_seed(base_seed: int, rank: int) -> int:
    """Distinct Gumbel-noise stream per rank (games on different GPUs must not share noise)."""
    return (int(base_seed) * 0x9E3779B1 + 0x85EBCA6B * (rank + 1)) & 0x7FFFFFFFFFFFFFFF


def broadcast_weights(module: torch.nn.Module, src: int = 0, group=None):
    """Publish `module`'s parameters and buffers from `src` to every rank with ONE broadcast of a flat
    buffer per dtype (88 MB fp32 at 15x15: launch latency matters, link count does not)."""
    tensors = [t for t in list(module.parameters()) + list(module.buffers())]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            n = t.numel()
            with torch.no_grad():
                t.copy_(flat[off:off + n].reshape(t.shape))
            off += n
    return module


class FlatWeights:
    """A module's parameters and buffers re-homed as views into ONE flat tensor per dtype, so that publishing
    the weights (`model_update_queue`, workers.py:587-593) is a single in-place broadcast per dtype with no
    packing or unpacking kernels -- the link, not ~300 small copies, sets the time.  The module keeps working
    as before (its tensors now alias the flat buffers); every rank must wrap an identically shaped module."""

    def __init__(self, module: torch.nn.Module):
        self.module = module
        groups = {}
        for t in list(module.parameters()) + list(module.buffers()):
            groups.setdefault((t.dtype, t.device), []).append(t)
        self.flats = []
        for (dtype, device), ts in groups.items():
            flat = torch.empty(sum(t.numel() for t in ts), dtype=dtype, device=device)
            off = 0
            for t in ts:
                n = t.numel()
                view = flat[off:off + n].view(t.shape)
                with torch.no_grad():
                    view.copy_(t)
                t.data = view                     # parameter / buffer now aliases the flat storage
                off += n
            self.flats.append(flat)

    @property
    def nbytes(self):
        return sum(f.numel() * f.element_size() for f in self.flats)

    def broadcast(self, src: int = 0, group=None):
        for f in self.flats:
            dist.broadcast(f, src=src, group=group)
        return self.module


def max_over_ranks(value: float, device=None, group=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device=None, group=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def gather_finished_games(records, dst: int = 0, group=None):
    """Harvested games (list of dicts from TrajectoryStore.harvest) of every rank -> `dst`.
    Fixed-stride packing: one int64 header row + float64 payload per game, counts all-gathered first."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [None] * world
    dist.all_gather_object(counts, len(records), group=group)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(records, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    out = []
    for r, recs in enumerate(gathered):
        assert len(recs) == counts[r]
        for rec in recs:
            rec = dict(rec)
            rec["rank"] = r
            out.append(rec)
    return out
