"""Frame-parallel driver: one process per GPU, frames sharded contiguously, the per-clip pre-pass
sharded over ranks with NCCL all-gathers of mergeable per-channel partial statistics.

The reference has no distributed code at all; this is the multi-GPU design of SURVEY
8(e).  The per-frame loop needs no communication: every frame is a pure function of the frame, the
weights and ~60 KB of per-clip constants.  The only exchange is in ``Decoder.compute``
(test/style_network_global.py:425-439): at each of its 11 statistic points (and 3 filter
predictions) every rank reduces its local samples to ``double[5][C] = {count, sum, M2, min, max}``,
all ranks all-gather those (<= 20 KB per rank) and merge them in rank order with Chan's parallel
variance formula (csrc/stats.cu: stats_merge_kernel), so all ranks hold bit-identical tables.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n items for ``rank``; the first n % world ranks get one more."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sample_indices(n_frames: int, interval: int = 8):
    """Frames the reference samples for the pre-pass (generate_real_video.py:133-143): every
    ``interval``-th frame for s in range((n-1)//interval), then the last frame."""
    return [s * interval for s in range((n_frames - 1) // interval)] + [n_frames - 1]


def allgather_parts(part: torch.Tensor, group=None) -> torch.Tensor:
    """[5, C] float64 partial statistics -> [world, 5, C] in rank order (NCCL on GPU, gloo on CPU)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(out, part.contiguous(), group=group)      # concatenated along dim 0 (gloo and nccl)
    return out.view((world,) + tuple(part.shape))


def sharded_prepass(fw, sample_frames, rank: int, world: int, group=None):
    """``clean(); add(...)*; compute()`` of framework.Stylization with the sampled frames sharded
    over ranks.  ``sample_frames``: the full list of uint8 BGR frames (every rank passes the same
    list; each encodes only its shard, plus the clip's first sample for quirk Q1)."""
    if len(sample_frames) < world:
        raise ValueError(f"{len(sample_frames)} sampled frames cannot be sharded over {world} ranks")
    eng = fw.model._eng()
    fw.clean()
    lo, hi = shard_range(len(sample_frames), rank, world)
    eng.stats_allgather = (lambda part: allgather_parts(part, group)) if world > 1 else None
    if lo != 0:
        eng.add_q1(fw._upload(sample_frames[0]), kind=1)
    for f in sample_frames[lo:hi]:
        fw.add(f)
    try:
        fw.compute()
    finally:
        eng.stats_allgather = None
