"""Does the row stride (HW) of the channel-major layout limit conv_tc / wgrad_tc?  Same total work, different HW x n_img splits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gazenerf_b200.train import _Ops

dev = torch.device("cuda:0")
o = _Ops(dev)
N = K = 384
total = 524288
W = (torch.randn(N, K) / K ** 0.5).to(dev)
b = torch.zeros(N, device=dev)
pk = o.pack(W, b, N, K)
for hw in (262144, 32768, 4096, 512, 128):
    n_img = total // hw
    X = torch.randn(n_img, K, hw, device=dev)
    out = torch.empty(n_img, N, hw, device=dev)
    for name, fn in (("conv ", lambda: o.conv(pk, N, K, X.data_ptr(), 0, out.data_ptr(), 0, n_img, hw, act=1)),
                     ("wgrad", lambda: o.wgrad(out.data_ptr(), 0, X.data_ptr(), 0, N, K, n_img, hw, "sum"))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("%s HW=%7d n_img=%5d  %.3f ms" % (name, hw, n_img, e0.elapsed_time(e1) / 5), flush=True)
    del X, out
