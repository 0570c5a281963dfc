"""Full-size training-step probe (BASELINE config 5 shape: B=2, 64x64 rays x 64 samples -> 512x512): phase timings + peak memory."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gazenerf_b200 as G
from bench import synthetic_inputs

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
opt = G.BaseOptions()
torch.manual_seed(45)
net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).train()
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic_inputs(torch, G, opt, B, 0).items()}
optim = torch.optim.Adam(net.parameters(), lr=1e-4)
gt = torch.rand(B, 3, 512, 512, device=dev)
L = G.lib()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); n0 = L.gnrf_launch_count()
    out = net("train", **kw)["coarse_dict"]
    torch.cuda.synchronize(); t1 = time.perf_counter()
    loss = sum((out[k] - gt).abs().mean() for k in ("merge_img_face", "merge_img_eyes", "merge_img")) + ((out["bg_img"] - 1.0) ** 2).mean()
    optim.zero_grad(set_to_none=True)
    loss.backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    optim.step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("iter %d: fwd %.1f ms  loss+bwd %.1f ms  adam %.1f ms  total %.1f ms  loss %.5f  launches %d  peak mem %.1f GB" % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0), float(loss), L.gnrf_launch_count() - n0,
        torch.cuda.max_memory_allocated() / 1e9), flush=True)
