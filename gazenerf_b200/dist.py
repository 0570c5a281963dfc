"""Multi-GPU plumbing for the render path: one process per GPU, faces (batch items) sharded across ranks, weights
replicated, and ONE all-gather of the rendered outputs when the caller needs the whole batch.

The reference is single-GPU (trainer/base.py:31-35); every face and every ray is independent through the whole path, so the
only exchange step is gathering results.  No data-path collective exists inside the render itself.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of the batch owned by `rank` (first `rem` ranks get one extra face)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(kwargs: Dict[str, Optional[torch.Tensor]], rank: int, world: int) -> Dict[str, Optional[torch.Tensor]]:
    """Slice every batched forward() argument of GazeNeRFNet to this rank's faces."""
    gb = kwargs["batch_xy"].shape[0]
    lo, hi = shard_range(gb, rank, world)
    out = {}
    for k, v in kwargs.items():
        out[k] = v[lo:hi] if (torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == gb) else v
    return out


def all_gather_images(local: Dict[str, torch.Tensor], global_batch: int, group=None) -> Dict[str, torch.Tensor]:
    """Gather {merge_img_face, merge_img_eyes, merge_img} [B_local,3,P,P] from all ranks into [global_batch,3,P,P] with a
    single collective (the three images are stacked so it is one all-gather); bg_img is rank-independent."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    keys = ["merge_img_face", "merge_img_eyes", "merge_img"]
    stacked = torch.stack([local[k] for k in keys], dim=1).contiguous()  # [B_local,3,3,P,P]
    sizes = [shard_range(global_batch, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    if len(set(counts)) == 1:
        full = torch.empty((global_batch,) + tuple(stacked.shape[1:]), device=stacked.device, dtype=stacked.dtype)
        dist.all_gather_into_tensor(full, stacked, group=group)
    else:  # ragged shards: pad to the largest shard, gather, then trim
        m = max(counts)
        pad = torch.zeros((m,) + tuple(stacked.shape[1:]), device=stacked.device, dtype=stacked.dtype)
        pad[: counts[rank]] = stacked
        buf = torch.empty((world * m,) + tuple(stacked.shape[1:]), device=stacked.device, dtype=stacked.dtype)
        dist.all_gather_into_tensor(buf, pad, group=group)
        full = torch.cat([buf[r * m: r * m + counts[r]] for r in range(world)], 0)
    out = {k: full[:, i] for i, k in enumerate(keys)}
    out["bg_img"] = local["bg_img"]
    return out


class BatchShardedRenderer(object):
    """net("test", **global_kwargs) over a process group: each rank renders its slice, one all-gather returns the batch."""

    def __init__(self, net, group=None):
        self.net = net
        self.group = group

    @torch.no_grad()
    def __call__(self, mode: str, **kwargs) -> Dict[str, Dict[str, torch.Tensor]]:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.net(mode, **kwargs)
        rank = dist.get_rank(self.group)
        gb = kwargs["batch_xy"].shape[0]
        local = self.net(mode, **shard_inputs(kwargs, rank, world))
        return {"coarse_dict": all_gather_images(local["coarse_dict"], gb, self.group)}
