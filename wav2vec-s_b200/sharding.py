"""Batch sharding across the GPUs of one box.

Utterances are independent, so the encoder forward shards with no data-path collective: every rank
encodes a contiguous slice of the global batch with replicated weights.  Collectives (NCCL over
NVLink on GPUs, gloo in the CPU tests) are used only to gather the variable-length outputs and the
per-rank timings.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of n_items over `world` ranks; the first n_items % world ranks get one more."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def balanced_order(lengths: List[int], world: int) -> List[List[int]]:
    """Length-aware assignment: attention cost grows with T^2, so greedily give the next-longest
    utterance to the least-loaded rank (cost model T + T^2/4096).  Returns per-rank index lists."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += lengths[i] + lengths[i] ** 2 / 4096.0
    return [sorted(x) for x in out]


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device time of the timed region)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_outputs(x: torch.Tensor, n_frames: torch.Tensor, dst: Optional[int] = None):
    """Gather encoder outputs of all ranks.

    x: [B_r, T_r, D] outputs of this rank, n_frames: [B_r] valid frames per utterance.  Ranks may have
    different B_r / T_r: shapes are exchanged first, tensors are padded to the global maximum and
    all-gathered, and the padding is stripped again.  Returns (list of [T_i, D] tensors in global batch
    order, [B] frame counts) on every rank (or only on `dst` when given, None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [x[i, : int(n_frames[i])] for i in range(x.size(0))], n_frames.clone()
    world, rank = dist.get_world_size(), dist.get_rank()
    shape = torch.tensor([x.size(0), x.size(1)], dtype=torch.int64, device=x.device)
    shapes = shape.new_empty((world * 2,))
    dist.all_gather_into_tensor(shapes, shape)
    shapes = shapes.view(world, 2).tolist()                 # one device -> host read for all ranks' shapes
    Bm = max(s_[0] for s_ in shapes)
    Tm = max(s_[1] for s_ in shapes)
    if x.size(0) == Bm and x.size(1) == Tm and x.is_contiguous():
        pad = x                                       # equal shards (the usual case): no staging copy
    else:
        pad = x.new_zeros((Bm, Tm, x.size(2)))
        pad[: x.size(0), : x.size(1)] = x
    nf = torch.zeros(Bm, dtype=torch.int64, device=x.device)
    nf[: x.size(0)] = n_frames.to(torch.int64)
    # one collective per tensor into a single [world, ...] buffer (NCCL all-gather over NVLink on GPUs)
    xs = x.new_empty((world * Bm,) + tuple(pad.shape[1:]))
    nfs = nf.new_empty((world * Bm,))
    dist.all_gather_into_tensor(xs, pad)
    dist.all_gather_into_tensor(nfs, nf)
    xs, nfs = xs.view((world, Bm) + tuple(pad.shape[1:])), nfs.view(world, Bm)
    if dst is not None and rank != dst:
        return None
    nfs_host = nfs.tolist()                                 # one read for every utterance's frame count
    outs, counts = [], []
    for r in range(world):
        for i in range(shapes[r][0]):
            n = nfs_host[r][i]
            outs.append(xs[r, i, :n])                       # views into the gathered buffer: no copies
            counts.append(n)
    return outs, torch.tensor(counts, dtype=torch.int64, device=x.device)
