"""Profiling driver: a few eager training steps at BASELINE config[4] size (B = 2) for ncu (launch list / --set full captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gazenerf_b200 as G
from bench import synthetic_inputs

dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
opt = G.BaseOptions()
torch.manual_seed(45)
net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).train()
net.train_precision = prec
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic_inputs(torch, G, opt, 2, 0).items()}
gt = torch.rand(2, 3, 512, 512, device=dev)
for it in range(steps):
    out = net("train", **kw)["coarse_dict"]
    loss = sum((out[k] - gt).abs().mean() for k in ("merge_img_face", "merge_img_eyes", "merge_img"))
    for p in net.parameters():
        p.grad = None
    loss.backward()
torch.cuda.synchronize()
print("done", float(loss))
