"""Profiling helper (not a test): N steady-state forwards of the default bench workload, for `ncu` launch lists / captures.
usage:  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tests/prof_step.py [steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gazenerf_b200 as G
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sys.argv = [sys.argv[0]]
import bench

dev = torch.device("cuda:0")
opt = G.BaseOptions()
torch.manual_seed(45)
net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).eval()
kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.synthetic_inputs(torch, G, opt, 1, 0).items()}
with torch.no_grad():
    for _ in range(steps):
        net("test", **kw)
torch.cuda.synchronize()
print("done")
