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


@pytest.mark.gpu
def test_graphed_train_step_matches_eager_steps():
    """build_code_and_cam -> net("train") -> GazeNeRFLoss -> backward -> Adam captured into ONE CUDA graph.
    With the same (fed) jitter, N replays must leave the network, code offsets and camera deltas where N eager steps leave them, and
    the graph must contain the libgnrf launches (weight packing included: the weights change on every replay)."""
    import copy
    from gazenerf_b200.trainer_utils import GraphedTrainStep
    from bench import synthetic_inputs, synthetic_targets
    dev = torch.device("cuda:0")
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 16
    F = 2

    def make():
        torch.manual_seed(45)
        net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).train()
        with torch.no_grad():
            for m in (net.fg_CD_predictor_face, net.fg_CD_predictor_eyes):
                m.density_module.weight.mul_(30.0)
        off = {"iden": torch.zeros(F, 100, device=dev, requires_grad=True), "expr": torch.zeros(F, 79, device=dev, requires_grad=True),
               "appea": torch.zeros(F, 127, device=dev, requires_grad=True)}
        d_eul = torch.zeros(F, 3, device=dev, requires_grad=True)
        d_tv = torch.zeros(F, 3, 1, device=dev, requires_grad=True)
        optim = torch.optim.Adam([{"params": list(net.parameters()), "lr": 1e-3}, {"params": list(off.values()), "lr": 1.5e-3},
                                  {"params": [d_eul, d_tv], "lr": 1e-4}], capturable=True)
        return net, off, d_eul, d_tv, optim

    kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic_inputs(torch, G, opt, F, seed=0).items()}
    tg = {k: v.to(dev) for k, v in synthetic_targets(torch, F, opt.pred_img_size, seed=0).items()}
    base = {"iden": kw["shape_code"][:, :100].contiguous(), "expr": kw["shape_code"][:, 100:].contiguous(), "text": kw["appea_code"][:, :100].contiguous(),
            "illu": kw["appea_code"][:, 100:].contiguous(), "gaze": kw["gaze_code"]}
    cam = {k: kw[k] for k in ("batch_Rmats", "batch_Tvecs", "batch_inv_inmats")}
    loss_fn = G.GazeNeRFLoss(eye_loss_importance=1.0, vgg_importance=1.0, use_vgg_loss=False, use_l1_loss=True)
    ju = [torch.rand(F, 256, 17, device=dev, generator=torch.Generator(device=dev).manual_seed(i)) for i in range(3)]
    n_steps = 3

    # eager reference run
    net_e, off_e, de_e, dt_e, optim_e = make()
    for i in range(n_steps):
        code_info, opt_code, cam_info, delta_cam = build_code_and_cam(base, off_e, cam, 0, F, de_e, dt_e)
        pred = net_e("train", kw["batch_xy"], None, **code_info, **cam_info, jitter_u=ju[i])
        loss = loss_fn.calc_total_loss(delta_cam, opt_code, pred, tg["gt"], tg["head"], tg["full_eye"], tg["left_eye"], tg["right_eye"], None, None, 0, 0)
        optim_e.zero_grad(set_to_none=True)
        loss["total_loss"].backward()
        optim_e.step()
    loss_e = float(loss["total_loss"])

    # graphed run: warm-up steps are taken on COPIES so that the captured run starts from the same state as the eager one
    net_g, off_g, de_g, dt_g, optim_g = make()
    snap = (copy.deepcopy(net_g.state_dict()), {k: v.detach().clone() for k, v in off_g.items()}, de_g.detach().clone(), dt_g.detach().clone(),
            copy.deepcopy(optim_g.state_dict()))
    gs = GraphedTrainStep(net_g, loss_fn, optim_g, kw["batch_xy"], base, off_g, cam, tg, de_g, dt_g, jitter_u=ju[0])
    assert gs.launches_per_replay > 100
    with torch.no_grad():   # rewind to the initial state (in place: the graph holds these tensors)
        for k, v in net_g.state_dict().items():
            v.copy_(snap[0][k])
        for k in off_g:
            off_g[k].copy_(snap[1][k])
        de_g.copy_(snap[2]); dt_g.copy_(snap[3])
        for st in optim_g.state.values():
            for k, v in st.items():
                if torch.is_tensor(v):
                    v.zero_()
    net_g.invalidate_caches()
    for i in range(n_steps):
        out = gs.step(jitter_u=ju[i])
    torch.cuda.synchronize()
    assert abs(float(out["total_loss"]) - loss_e) < 1e-4 * max(1.0, abs(loss_e))
    worst = max(float((a - b).abs().max()) for a, b in zip(net_g.state_dict().values(), net_e.state_dict().values()))
    assert worst < 1e-5, worst
    for a, b in ((off_g["iden"], off_e["iden"]), (off_g["appea"], off_e["appea"]), (de_g, de_e), (dt_g, dt_e)):
        assert float((a - b).abs().max()) < 1e-5
    assert float((net_g.fg_CD_predictor_face.FeaExt_module_3.weight - snap[0]["fg_CD_predictor_face.FeaExt_module_3.weight"]).abs().max()) > 1e-4   # it trained


@pytest.mark.gpu
def test_graphed_train_step_draws_jitter_on_device():
    """Without a fed jitter tensor the stratified draws come from the graph-safe device generator: two replays see different samples."""
    from gazenerf_b200.trainer_utils import GraphedTrainStep
    from bench import synthetic_inputs, synthetic_targets
    dev = torch.device("cuda:0")
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 16
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False).to(dev).train()
    with torch.no_grad():
        for m in (net.fg_CD_predictor_face, net.fg_CD_predictor_eyes):
            m.density_module.weight.mul_(30.0)
    off = {"iden": torch.zeros(1, 100, device=dev, requires_grad=True), "expr": torch.zeros(1, 79, device=dev, requires_grad=True),
           "appea": torch.zeros(1, 127, device=dev, requires_grad=True)}
    optim = torch.optim.Adam(list(net.parameters()) + list(off.values()), lr=0.0, capturable=True)   # lr = 0: only the jitter differs
    kw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic_inputs(torch, G, opt, 1, seed=0).items()}
    tg = {k: v.to(dev) for k, v in synthetic_targets(torch, 1, opt.pred_img_size, seed=0).items()}
    base = {"iden": kw["shape_code"][:, :100].contiguous(), "expr": kw["shape_code"][:, 100:].contiguous(), "text": kw["appea_code"][:, :100].contiguous(),
            "illu": kw["appea_code"][:, 100:].contiguous(), "gaze": kw["gaze_code"]}
    cam = {k: kw[k] for k in ("batch_Rmats", "batch_Tvecs", "batch_inv_inmats")}
    loss_fn = G.GazeNeRFLoss(eye_loss_importance=1.0, vgg_importance=1.0, use_vgg_loss=False, use_l1_loss=True)
    gs = GraphedTrainStep(net, loss_fn, optim, kw["batch_xy"], base, off, cam, tg)
    a = float(gs.step()["total_loss"])
    b = float(gs.step()["total_loss"])
    assert a != b and abs(a - b) < 0.1 * abs(a)
