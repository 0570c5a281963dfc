"""Gradient golden vectors from the UNMODIFIED reference (imported from /root/reference): forward in the reference's own
autograd graph, ``loss = sum_k <img_k, Wt_k>`` with seeded random weights, ``loss.backward()``.

    python oracle/gen_golden_grad.py        # writes tests/golden/tiny_grad.npz, std_dense_train_grad.npz

tiny_grad.npz            every parameter / input gradient in full (weights are those of tiny.npz).
std_dense_train_grad.npz real layer widths, train mode (jitter from std_dense_train.npz): input gradients (codes, gaze, R, T) in
                         full; per-parameter gradients as [sum, abs-sum, l2, <g, r>] with r ~ N(0,1) from a per-tensor seed
                         (the full set is 20 MB).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as GG  # noqa: E402

IMG_KEYS = ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img")


def loss_weights(shapes, seed=99):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(shapes[k], generator=g) for k in IMG_KEYS}


def proj_vec(shape, idx):
    g = torch.Generator().manual_seed(1000 + idx)
    return torch.randn(shape, generator=g)


def run(net, mode, xy, uv, shape, appea, gaze, cam, seed_rng=None):
    leaves = {"shape": shape.clone().requires_grad_(True), "appea": appea.clone().requires_grad_(True),
              "gaze": gaze.clone().requires_grad_(True), "R": cam["batch_Rmats"].clone().requires_grad_(True),
              "T": cam["batch_Tvecs"].clone().requires_grad_(True)}
    for p in net.parameters():
        p.grad = None
    if seed_rng is not None:
        torch.manual_seed(seed_rng)
    out = net(mode, xy, uv, None, leaves["shape"], leaves["appea"], leaves["gaze"], leaves["R"], leaves["T"], cam["batch_inv_inmats"])
    imgs = out["coarse_dict"]
    wt = loss_weights({k: imgs[k].shape for k in IMG_KEYS})
    loss = sum((imgs[k] * wt[k]).sum() for k in IMG_KEYS)
    loss.backward()
    return float(loss), leaves


def main():
    GG.install_kornia_shim()
    sys.path.insert(0, GG.REF)
    os.chdir(GG.REF)
    from configs.gazenerf_options import BaseOptions
    from models.gaze_nerf import GazeNeRFNet
    from utils.render_utils import RenderUtils

    def inputs(opt, b):
        ru = RenderUtils(45, "cpu", opt)
        g = torch.Generator().manual_seed(0)
        shape = torch.randn(b, 179, generator=g) * 0.3
        appea = torch.randn(b, 127, generator=g) * 0.3
        gaze = torch.rand(b, 2, generator=g) - 0.5
        cams = [ru.base_cam_info, ru.cam_info_list[7]]
        cam = {k: torch.cat([cams[i % 2][k] for i in range(b)], 0) for k in cams[0]}
        return ru.ray_xy.expand(b, -1, -1), ru.ray_uv.expand(b, -1, -1), shape, appea, gaze, cam

    # ---- tiny: weights from tiny.npz
    tiny = np.load(os.path.join(GG.OUT, "tiny.npz"))
    opt = BaseOptions({"featmap_size": 8, "featmap_nc": 48, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    opt.mlp_hidden_nchannels = 32
    net = GazeNeRFNet(opt, include_vd=False, hier_sampling=False).eval()
    net.load_state_dict({k[3:]: torch.from_numpy(tiny[k]) for k in tiny.files if k.startswith("sd/")}, strict=True)
    xy, uv, shape, appea, gaze, cam = inputs(opt, 2)
    loss, leaves = run(net, "test", xy, uv, shape, appea, gaze, cam)
    out = {"loss": np.array([loss])}
    for k, v in leaves.items():
        out["gin/" + k] = GG.np32(v.grad)
    for k, p in net.named_parameters():
        out["gp/" + k] = GG.np32(p.grad if p.grad is not None else torch.zeros_like(p))
    np.savez_compressed(os.path.join(GG.OUT, "tiny_grad.npz"), **out)
    print("tiny_grad.npz loss", loss, sum(v.nbytes for v in out.values()) // 1024, "KiB raw")

    # ---- std dense train
    std = np.load(os.path.join(GG.OUT, "std_dense_train.npz"))
    opt = BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    torch.manual_seed(45)
    net = GazeNeRFNet(opt, include_vd=False, hier_sampling=False).eval()
    xy, uv, shape, appea, gaze, cam = inputs(opt, 2)
    GG.make_dense(net, xy, shape, appea, gaze, cam, opt.num_sample_coarse)
    loss, leaves = run(net, "train", xy, uv, shape, appea, gaze, cam, seed_rng=123)
    out = {"loss": np.array([loss])}
    for k, v in leaves.items():
        out["gin/" + k] = GG.np32(v.grad)
    for i, (k, p) in enumerate(net.named_parameters()):
        g = (p.grad if p.grad is not None else torch.zeros_like(p)).double()
        out["gs/" + k] = np.array([float(g.sum()), float(g.abs().sum()), float(g.norm()), float((g * proj_vec(g.shape, i).double()).sum())])
    # sanity: the forward of this run equals the committed forward golden
    assert np.allclose(std["in_shape"], GG.np32(shape))
    np.savez_compressed(os.path.join(GG.OUT, "std_dense_train_grad.npz"), **out)
    print("std_dense_train_grad.npz loss", loss, sum(v.nbytes for v in out.values()) // 1024, "KiB raw")




def loss_goldens():
    """GazeNeRFLoss data terms + image gradients from the reference's own loss module (use_vgg_loss=False)."""
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import gazenerf_oracle as O
    from losses.gazenerf_loss import GazeNeRFLoss  # reference

    out = {}
    for use_l1 in (True, False):
        pred, gt, face, full_eye, left, right = O.synthetic_loss_inputs(2, 64, seed=5)
        pred = {k: v.clone().requires_grad_(True) for k, v in pred.items()}
        lf = GazeNeRFLoss(eye_loss_importance=1.0, vgg_importance=1.0, use_vgg_loss=False, use_l1_loss=use_l1)
        opt_code = {"iden": torch.zeros(2, 100), "expr": torch.zeros(2, 79), "appea": torch.zeros(2, 127), "bg": None}
        ld = lf.calc_total_loss(None, opt_code, {"coarse_dict": pred}, gt, face, full_eye, left, right, None, None, 0, 0)
        w = {"bg_loss": 1.0, "eyes_loss": 2.0, "face_loss": 3.0, "nonhead_loss": 4.0, "head_loss": 5.0}
        sum(w[k] * ld[k] for k in w).backward()
        tag = "l1" if use_l1 else "mse"
        for k in w:
            out["%s/%s" % (tag, k)] = np.array([float(ld[k])])
        out["%s/total_loss" % tag] = np.array([float(ld["total_loss"])])
        for k, v in pred.items():
            out["%s/g_%s" % (tag, k)] = GG.np32(v.grad)
    np.savez_compressed(os.path.join(GG.OUT, "loss.npz"), **out)
    print("loss.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    if os.environ.get("GNRF_LOSS_ONLY"):   # only tests/golden/loss.npz
        GG.install_kornia_shim()
        sys.path.insert(0, GG.REF)
        os.chdir(GG.REF)
    else:
        main()
    loss_goldens()
