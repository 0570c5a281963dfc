"""Callers either side of the path: batched view sweeps and the trainer's code / camera assembly."""
import numpy as np
import pytest
import torch

import gazenerf_b200 as G
from gazenerf_b200.trainer_utils import build_code_and_cam, eulurangle2Rmat


def _ref_euler(angles):
    """trainer/base.py:92-124 restated literally (identity matrices + slice writes)."""
    b = angles.shape[0]
    sx, sy, sz = torch.sin(angles[:, 0]), torch.sin(angles[:, 1]), torch.sin(angles[:, 2])
    cx, cy, cz = torch.cos(angles[:, 0]), torch.cos(angles[:, 1]), torch.cos(angles[:, 2])
    rx = torch.eye(3).view(1, 3, 3).repeat(b, 1, 1)
    ry, rz = rx.clone(), rx.clone()
    rx[:, 1, 1], rx[:, 1, 2], rx[:, 2, 1], rx[:, 2, 2] = cx, -sx, sx, cx
    ry[:, 0, 0], ry[:, 0, 2], ry[:, 2, 0], ry[:, 2, 2] = cy, sy, -sy, cy
    rz[:, 0, 0], rz[:, 0, 1], rz[:, 1, 0], rz[:, 1, 1] = cz, -sz, sz, cz
    return rz.bmm(ry.bmm(rx))


def test_euler_and_code_cam_assembly():
    g = torch.Generator().manual_seed(2)
    ang = (torch.rand(5, 3, generator=g) - 0.5).requires_grad_(True)
    a, b = eulurangle2Rmat(ang), _ref_euler(ang)
    assert torch.equal(a, b)
    ga, = torch.autograd.grad((a * torch.arange(9.).view(1, 3, 3)).sum(), ang)
    gb, = torch.autograd.grad((b * torch.arange(9.).view(1, 3, 3)).sum(), ang)
    assert torch.allclose(ga, gb, atol=1e-6)
    B = 2
    base = {"iden": torch.randn(B, 100, generator=g), "expr": torch.randn(B, 79, generator=g), "text": torch.randn(B, 100, generator=g),
            "illu": torch.randn(B, 27, generator=g), "gaze": torch.randn(B, 2, generator=g)}
    off = {"iden": torch.randn(6, 100, generator=g), "expr": torch.randn(6, 79, generator=g), "appea": torch.randn(6, 127, generator=g)}
    cam = {"batch_Rmats": torch.randn(B, 3, 3, generator=g), "batch_Tvecs": torch.randn(B, 3, 1, generator=g), "batch_inv_inmats": torch.eye(3).expand(B, 3, 3)}
    de, dt = torch.randn(6, 3, generator=g) * 0.1, torch.randn(6, 3, 1, generator=g) * 0.1
    code, optc, cam2, dcam = build_code_and_cam(base, off, cam, pos=2, batch_size=B, delta_eulur=de, delta_tvecs=dt)
    assert code["shape_code"].shape == (B, 179) and code["appea_code"].shape == (B, 127) and code["bg_code"] is None
    assert torch.equal(code["shape_code"][:, :100], base["iden"] + off["iden"][2:4])
    dr = _ref_euler(de[2:4])
    assert torch.allclose(cam2["batch_Rmats"], dr.bmm(cam["batch_Rmats"])) and torch.allclose(cam2["batch_Tvecs"], dr.bmm(cam["batch_Tvecs"]) + dt[2:4])
    assert torch.equal(dcam["delta_eulur"], de[2:4]) and torch.equal(optc["appea"], off["appea"][2:4])
    code, optc, cam3, dcam = build_code_and_cam(base, off, cam, pos=0, batch_size=B)
    assert cam3 is cam and dcam is None


@pytest.mark.gpu
def test_batched_novel_views_match_sequential_forwards():
    """render_novel_views submits the sweep as one batch; every image must equal the reference-style batch-1 forward."""
    dev = torch.device("cuda:0")
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 16
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    g = torch.Generator().manual_seed(0)
    shape, appea = torch.randn(1, 179, generator=g) * 0.3, torch.randn(1, 127, generator=g) * 0.3
    ru_c = G.RenderUtils(9, "cpu", opt)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    oo = O.OracleOptions(featmap_size=16, featmap_nc=258, pred_img_size=64, num_sample_coarse=16)
    bias = O.calibrate_dense_bias(sd, oo, ru_c.ray_xy, shape, appea, torch.zeros(1, 2), ru_c.base_cam_info["batch_Rmats"],
                                  ru_c.base_cam_info["batch_Tvecs"], ru_c.base_cam_info["batch_inv_inmats"])   # non-vacuous density
    net.load_state_dict(O.densify(sd, *bias))
    net = net.to(dev).eval()
    ru = G.RenderUtils(9, dev, opt)
    code = {"bg_code": None, "shape_code": shape.to(dev), "appea_code": appea.to(dev), "gaze_code": torch.zeros(1, 2, device=dev)}
    imgs = ru.render_novel_views(net, code, move_gaze=True)
    assert len(imgs) == 9 and imgs[0].shape == (64, 64, 3) and imgs[0].dtype == np.uint8
    assert torch.allclose(code["gaze_code"].cpu(), torch.tensor([[ru._SWEEP_H[8], ru._SWEEP_V[8]]]))
    for i in range(9):
        c = dict(code)
        c["gaze_code"] = torch.tensor([[ru._SWEEP_H[i], ru._SWEEP_V[i]]], device=dev)
        with torch.no_grad():
            ref = net("test", ru.ray_xy, ru.ray_uv, **c, **ru.cam_info_list[i])["coarse_dict"]["merge_img"]
        ref = (ref[0].cpu().permute(1, 2, 0).numpy() * 255).astype(np.uint8)
        assert np.abs(ref.astype(np.int32) - imgs[i].astype(np.int32)).max() <= 1, i
    assert len({im.tobytes() for im in imgs}) == 9   # the views differ
    gz = ru.render_novel_views_gaze(net, code, ru.base_cam_info)
    assert len(gz) == 11 + 11 + 10 + 10
    mo = ru.render_morphing_res(net, code, {**code, "shape_code": code["shape_code"] * 0.5}, 4)
    assert len(mo) == 4


# ------------------------------------------------------------------------------------------------ GazeNeRFLoss
from conftest import load_golden, rel_l2  # noqa: E402
from oracle import gazenerf_oracle as O  # noqa: E402

LOSS_W = {"bg_loss": 1.0, "eyes_loss": 2.0, "face_loss": 3.0, "nonhead_loss": 4.0, "head_loss": 5.0}


@pytest.mark.parametrize("use_l1", [True, False])
def test_oracle_data_loss_matches_reference(use_l1):
    gold, tag = load_golden("loss"), "l1" if use_l1 else "mse"
    pred, gt, face, full_eye, left, right = O.synthetic_loss_inputs(2, 64, seed=5)
    pred = {k: v.clone().requires_grad_(True) for k, v in pred.items()}
    terms = O.data_loss_terms(pred, gt, face, full_eye, left, right, use_l1=use_l1)
    for k in LOSS_W:
        assert abs(float(terms[k]) - float(gold["%s/%s" % (tag, k)][0])) < 1e-6, k
    sum(LOSS_W[k] * terms[k] for k in LOSS_W).backward()
    for k, v in pred.items():
        assert rel_l2(v.grad, gold["%s/g_%s" % (tag, k)]) < 1e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize("use_l1", [True, False])
def test_fused_data_loss_matches_reference(use_l1):
    dev = torch.device("cuda:0")
    gold, tag = load_golden("loss"), "l1" if use_l1 else "mse"
    pred, gt, face, full_eye, left, right = O.synthetic_loss_inputs(2, 64, seed=5)
    pred = {k: v.to(dev).requires_grad_(True) for k, v in pred.items()}
    lf = G.GazeNeRFLoss(eye_loss_importance=1.0, vgg_importance=1.0, use_vgg_loss=False, use_l1_loss=use_l1)
    opt_code = {"iden": torch.zeros(2, 100, device=dev), "expr": torch.zeros(2, 79, device=dev), "appea": torch.zeros(2, 127, device=dev), "bg": None}
    ld = lf.calc_total_loss(None, opt_code, {"coarse_dict": pred}, gt.to(dev), face.to(dev), full_eye.to(dev), left.to(dev), right.to(dev), None, None, 0, 0)
    for k in LOSS_W:
        assert abs(float(ld[k]) - float(gold["%s/%s" % (tag, k)][0])) < 2e-6 * max(1.0, abs(float(gold["%s/%s" % (tag, k)][0]))), k
    assert abs(float(ld["total_loss"]) - float(gold["%s/total_loss" % tag][0])) < 1e-5
    sum(LOSS_W[k] * ld[k] for k in LOSS_W).backward()
    for k, v in pred.items():
        assert rel_l2(v.grad.cpu(), gold["%s/g_%s" % (tag, k)]) < 1e-5, k
    with pytest.raises(NotImplementedError):
        G.GazeNeRFLoss(1.0, 1.0)   # the reference default (VGG perceptual term) needs downloaded weights
