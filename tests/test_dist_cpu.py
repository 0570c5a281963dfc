"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: batch sharding and the single all-gather of rendered images."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gazenerf_b200.dist import all_gather_images, allreduce_gradients, shard_inputs, shard_range


def test_shard_range_partitions_batch():
    for gb in (1, 2, 7, 8, 45):
        for world in (1, 2, 4, 8):
            spans = [shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, gb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = 8
        full = {k: torch.arange(gb * 3 * P * P, dtype=torch.float32).view(gb, 3, P, P) + 1000 * i
                for i, k in enumerate(["merge_img_face", "merge_img_eyes", "merge_img"])}
        kw = {"batch_xy": torch.zeros(gb, 2, 4), "shape_code": torch.arange(gb).float().view(gb, 1).expand(gb, 179).contiguous(), "batch_uv": None}
        mine = shard_inputs(kw, rank, world)
        lo, hi = shard_range(gb, rank, world)
        assert mine["shape_code"].shape[0] == hi - lo and mine["batch_uv"] is None
        assert float(mine["shape_code"][0, 0]) == lo if hi > lo else True
        local = {k: v[lo:hi].clone() for k, v in full.items()}
        local["bg_img"] = torch.ones(1, 3, P, P)
        out = all_gather_images(local, gb)
        ok = all(torch.equal(out[k], full[k]) for k in full) and out["bg_img"].shape == (1, 3, P, P)
        # data-parallel gradient averaging: one flat all-reduce; a parameter without a gradient on one rank still takes part
        params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
        params[0].grad = torch.full((3, 4), float(rank + 1))
        params[1].grad = torch.arange(5.0) * (rank + 1)
        if rank == 0:
            params[2].grad = torch.ones(2)
        allreduce_gradients(params)
        ok = ok and torch.allclose(params[0].grad, torch.full((3, 4), 1.5)) and torch.allclose(params[1].grad, torch.arange(5.0) * 1.5)
        ok = ok and torch.allclose(params[2].grad, torch.full((2,), 0.5))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gb", [2, 4, 5])
def test_all_gather_images_world2_gloo(gb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + gb
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]


class _StubNet(object):
    """Stands in for GazeNeRFNet in the host-logic test of BatchShardedRenderer (CPU, gloo): images are a function of the codes."""

    def __init__(self, hier):
        self.hier_sampling, self.pred_img_size, self.gather_ctx, self.gather_used = hier, 4, None, False

    def __call__(self, mode, **kw):
        b = kw["batch_xy"].shape[0]
        img = lambda s: (kw["shape_code"][:, :1].view(b, 1, 1, 1) + s).expand(b, 3, 4, 4).contiguous()
        d = {"merge_img_face": img(0.0), "merge_img_eyes": img(1.0), "merge_img": img(2.0), "bg_img": torch.ones(1, 3, 4, 4)}
        out = {"coarse_dict": d}
        if self.hier_sampling:
            out["fine_dict"] = {k: v + 10.0 for k, v in d.items()}
        return out


def _renderer_worker(rank, world, port, q):
    from gazenerf_b200.dist import BatchShardedRenderer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        gb = 5   # ragged: 3 + 2
        kw = {"batch_xy": torch.zeros(gb, 2, 4), "shape_code": torch.arange(gb).float().view(gb, 1).expand(gb, 179).contiguous(), "batch_uv": None}
        # hierarchical net with fused_gather requested: must take the collective path and return BOTH dicts for the whole batch
        r = BatchShardedRenderer(_StubNet(hier=True), fused_gather=True)
        out = r("test", **kw)
        want = torch.arange(gb).float().view(gb, 1, 1, 1).expand(gb, 3, 4, 4)
        ok = ok and torch.equal(out["coarse_dict"]["merge_img"], want + 2.0) and torch.equal(out["fine_dict"]["merge_img_eyes"], want + 11.0)
        # only_merge callers never enter the fused path either
        out = BatchShardedRenderer(_StubNet(hier=False), fused_gather=True)("test", only_merge=True, **kw)
        ok = ok and torch.equal(out["coarse_dict"]["merge_img_face"], want)
        # global batch smaller than the world: every rank raises BEFORE any collective (no rank is left blocked)
        small = {"batch_xy": torch.zeros(1, 2, 4), "shape_code": torch.zeros(1, 179), "batch_uv": None}
        try:
            BatchShardedRenderer(_StubNet(hier=False))("test", **small)
            ok = False
        except ValueError:
            pass
        dist.barrier()   # still in step with the peer
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_batch_sharded_renderer_host_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 17
    procs = [ctx.Process(target=_renderer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=10) for _ in range(2)) == [(0, True), (1, True)]
