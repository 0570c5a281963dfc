"""Debug helper (not a test): per-layer timeline of CTA 0 of the fused MLP kernel, from in-kernel clock64 stamps.
usage on the GPU box:  python tests/tc_timeline.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gazenerf_b200 as G
sys.argv = [sys.argv[0]]
import bench

dev = torch.device("cuda:0")
prof = torch.zeros(4 * 10 * 16, dtype=torch.int64, device=dev)
opt = G.BaseOptions()
torch.manual_seed(45)
net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).eval()
net.tc_debug = (None, prof, 2)   # -> gnrf_mlp_tc_fwd_debug(..., timeline = prof, cluster_size = 2)
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.synthetic_inputs(torch, G, opt, 1, 0).items()}
for _ in range(3):
    net("test", **kw)
torch.cuda.synchronize()
p = prof.cpu().view(4, 10, 16)
for item in (1, 2):
    t0 = int(p[item, 0, 0])
    print("item", item, "(cycles relative to the tile's first MMA-warp stamp)")
    print(" layer | mma_start   mma_end  dur   w_stall a_stall blk0_dur | acc_full  kb01_rel  all_rel ")
    for l in range(9):
        r = p[item, l]
        ws = int(r[4] - r[1]); as_ = int(r[5] - r[2])
        print("  %d    | %8d %8d %6d  %6d %6d %6d | %8d %8d %8d" % (l, int(r[0]) - t0, int(r[3]) - t0, int(r[3] - r[0]), ws, as_, int(r[10] - r[0]) if l > 0 else 0,
              int(r[7]) - t0, int(r[8]) - t0 if l < 8 else 0, int(r[9]) - t0))
    l = 2
    r = p[item, l]; base = int(p[item, l - 1, 7])  # acc_full of layer 1 = start of the drain that feeds layer 2
    print("  layer 2 detail (cycles after acc_full of layer 1): drain released kb1 %d kb3 %d kb5 %d | MMA blk0 issued kb1..5: %s | blk0 end %d, layer end %d" % (
        int(p[item, l - 1, 8]) - base, int(p[item, l - 1, 6]) - base, int(p[item, l - 1, 9]) - base,
        [int(r[k]) - base for k in range(11, 16)], int(r[10]) - base, int(r[3]) - base))
    print("  prologue done:", int(p[item, 0, 6]) - t0, " next tile start:", int(p[item + 1, 0, 0]) - t0)
