"""Per-rank body of tests/test_full_size.py::test_two_gpu_fused_gather_matches_nccl (launched by torch.distributed.run, one process
per GPU): the all-gather fused into the last neural-render kernel (gazenerf_b200.dist.PeerAllGather: multimem.st through the NVSwitch
multicast address, or peer stores over NVLink) must be bit-identical to the NCCL all-gather of the same local images, over several
steps with changing inputs (the two symmetric buffers alternate), through BatchShardedRenderer as a caller would use it."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import gazenerf_b200 as G  # noqa: E402
from gazenerf_b200.dist import BatchShardedRenderer, all_gather_images, shard_inputs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 128})
    opt.num_sample_coarse = 16
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).eval()
    ru = G.RenderUtils(45, dev, opt)
    F = 2                      # faces per rank
    gb = F * world
    fused = BatchShardedRenderer(net, fused_gather=True)
    ok = True
    mode = None
    for step in range(5):
        g = torch.Generator().manual_seed(100 + step)   # identical global batch on every rank
        shape = (torch.randn(gb, 179, generator=g) * 0.3).to(dev)
        appea = (torch.randn(gb, 127, generator=g) * 0.3).to(dev)
        gaze = (torch.rand(gb, 2, generator=g) - 0.5).to(dev)
        cams = [ru.cam_info_list[(7 * step + i) % 45] for i in range(gb)]
        kw = dict(batch_xy=ru.ray_xy.expand(gb, -1, -1), batch_uv=None, bg_code=None, shape_code=shape, appea_code=appea, gaze_code=gaze,
                  **{k: torch.cat([c[k] for c in cams], 0) for k in cams[0]})
        out = fused("test", **kw)["coarse_dict"]
        out = {k: v.clone() for k, v in out.items()}
        mode = "multicast" if fused._peer.use_multicast else "peer stores"
        with torch.no_grad():
            local_out = net("test", **shard_inputs(kw, rank, world))["coarse_dict"]
        ref = all_gather_images(local_out, gb)
        for k in ("merge_img_face", "merge_img_eyes", "merge_img"):
            assert out[k].shape == ref[k].shape == (gb, 3, 128, 128)
            ok = ok and torch.equal(out[k], ref[k])
        ok = ok and torch.equal(out["bg_img"], ref["bg_img"])
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("GATHER_CHECK %s (%s, world %d)" % ("bit-identical" if int(flag.item()) == 1 else "FAILED", mode, world))
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
