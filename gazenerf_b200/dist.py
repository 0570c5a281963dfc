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


def allreduce_gradients(params, group=None) -> None:
    """Data-parallel training: average the gradients of `params` over the ranks with ONE flat all-reduce (the network has 5.0 M
    parameters = 20 MB).  Parameters whose .grad is None on this rank contribute zeros, so every rank issues the same
    collective.  In place."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    flat /= world
    for p, g, synced in zip(params, grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        if p.grad is None:
            p.grad = synced.clone()
        else:
            g.copy_(synced)


class PeerAllGather(object):
    """All-gather of the rendered images FUSED into the kernel that writes them (include/gnrf.h, gnrf_neural_render_tc_fwd_gather).

    Every rank owns two symmetric buffers [3 keys][global_batch][3][P][P] (torch symmetric memory: each rank's buffer is mapped into
    every process; an NVSwitch multicast address covers all of them).  The last neural-render kernel stores each RGB value once to
    the multicast address (``multimem.st``, NVLS) -- or once per peer over NVLink when multicast is unavailable -- so the whole batch
    materialises on every GPU without a separate collective.  ``finish()`` runs one device-side barrier (signal pads) on the current
    stream and returns the gathered images.  ``n_buf`` buffers rotate per step: a rank can only be one barrier ahead of its slowest
    peer, so a peer's reads of step i (stream-ordered before its step i+1) are complete before anyone rewrites that buffer in step
    i + n_buf; with the default of three, a caller that reads the aliased buffer on ANOTHER stream (bench.py's D2H copy stream) has two
    whole steps to finish before it must fence.
    """

    def __init__(self, b_local: int, img_size: int, device, group=None, n_buf: int = 3):
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= 8:
            raise RuntimeError("PeerAllGather supports 2..8 ranks of one NVLink / NVSwitch domain")
        self.b_local, self.gb, self.P = b_local, b_local * self.world, img_size
        shape = (3, self.gb, 3, img_size, img_size)
        self.n_buf = max(2, int(n_buf))
        self.bufs = [symm.empty(shape, dtype=torch.float32, device=device) for _ in range(self.n_buf)]
        self.hdls = [symm.rendezvous(b, group) for b in self.bufs]
        self.use_multicast = all(int(h.multicast_ptr) != 0 for h in self.hdls)
        self.step = 0
        self._forced: Optional[int] = None   # set while a CUDA graph is captured for one particular buffer

    def current(self, index: Optional[int] = None):
        """(peer pointers, multicast pointer or 0) of the buffer the CURRENT step writes (``index``: of that buffer; graph capture)."""
        if index is None:
            index = self._forced if self._forced is not None else self.step % self.n_buf
        h = self.hdls[index]
        return [int(p) for p in h.buffer_ptrs], (int(h.multicast_ptr) if self.use_multicast else 0)

    def finish(self, alias: bool = False) -> Dict[str, torch.Tensor]:
        """Device barrier across the ranks (on the current stream), then this step's gathered images.

        ``alias=False`` (default) returns fresh tensors (one D2D copy), like the NCCL path.  ``alias=True`` returns zero-copy VIEWS of
        the symmetric buffer: peers overwrite it remotely two steps later, so the caller must have finished reading it -- on this
        stream, or behind an event it waits on before entering the next-but-one step -- by then (bench.py does this for its D2H)."""
        i = self.step % self.n_buf
        with torch.cuda.device(self.bufs[i].device):
            self.hdls[i].barrier(channel=0)
        self.step += 1
        buf = self.bufs[i]
        if not alias:
            buf = buf.clone()
        return {"merge_img_face": buf[0], "merge_img_eyes": buf[1], "merge_img": buf[2]}


class BatchShardedRenderer(object):
    """net("test", **global_kwargs) over a process group: each rank renders its slice, one all-gather returns the batch."""

    def __init__(self, net, group=None, fused_gather: bool = False, alias_outputs: bool = False):
        self.net = net
        self.group = group
        self.fused_gather = fused_gather   # gather inside the last neural-render kernel (PeerAllGather) instead of NCCL
        self.alias_outputs = alias_outputs  # fused path: return views of the symmetric buffer (see PeerAllGather.finish)
        self._peer: Optional[PeerAllGather] = None

    @torch.no_grad()
    def __call__(self, mode: str, **kwargs) -> Dict[str, Dict[str, torch.Tensor]]:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.net(mode, **kwargs)
        rank = dist.get_rank(self.group)
        gb = kwargs["batch_xy"].shape[0]
        if gb < world:   # checked on every rank BEFORE any collective: a rank with no face would otherwise leave its peers blocked
            raise ValueError("global batch %d < world size %d: every rank needs at least one face" % (gb, world))
        mine = shard_inputs(kwargs, rank, world)
        b_local = mine["batch_xy"].shape[0]
        # the fused gather lives in the coarse, all-three-images neural-render call: hierarchical nets (two image sets) and
        # only_merge callers take the NCCL path, as do ragged shards
        fused_ok = (self.fused_gather and gb == b_local * world and not self.net.hier_sampling and not kwargs.get("only_merge", False))
        if fused_ok:
            if self._peer is None or self._peer.b_local != b_local:
                self._peer = PeerAllGather(b_local, self.net.pred_img_size, mine["batch_xy"].device, self.group)
            self.net.gather_ctx = self._peer
            self.net.gather_used = False
            try:
                local = self.net(mode, **mine)
            finally:
                self.net.gather_ctx = None
            if not self.net.gather_used:
                raise RuntimeError("fused gather requested but the forward did not run the gathering neural-render kernel")
            out = self._peer.finish(alias=self.alias_outputs)
            out["bg_img"] = local["coarse_dict"]["bg_img"]
            return {"coarse_dict": out}
        local = self.net(mode, **mine)
        res = {"coarse_dict": all_gather_images(local["coarse_dict"], gb, self.group)}
        if "fine_dict" in local:
            res["fine_dict"] = all_gather_images(local["fine_dict"], gb, self.group)
        return res
