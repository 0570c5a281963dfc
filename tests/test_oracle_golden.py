"""Pin the oracle (oracle/gazenerf_oracle.py) against vectors produced by the reference itself (oracle/gen_golden.py).

CPU only.  tiny.npz stores all weights, so this does not depend on RNG replication; std_*.npz rebuild the weights with
the drop-in module's reference-identical init (checked separately in test_module_api.py)."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden, max_rel, rel_l2
from oracle import gazenerf_oracle as O

TOL = 2e-6  # fp32 re-association noise between two CPU evaluations of the same graph


def _opt(meta, n_fine=8):
    return O.OracleOptions(featmap_size=int(meta[0]), featmap_nc=int(meta[1]), pred_img_size=int(meta[2]),
                           num_sample_coarse=int(meta[3]), mlp_hidden_nchannels=int(meta[4]), num_sample_fine=n_fine)


def _inputs(g):
    t = lambda k: torch.from_numpy(g[k])
    return t("in_xy"), t("in_shape"), t("in_appea"), t("in_gaze"), t("in_R"), t("in_T"), t("in_Kinv")


def test_host_fixtures_match_reference(tiny_golden):
    g = tiny_golden
    xy, uv = O.pixel_grid(8)
    assert np.array_equal(xy.numpy(), g["in_xy"][:1])
    assert np.array_equal(uv.numpy(), g["ru_uv"])
    assert np.array_equal(O.scaled_inv_intrinsics(8).numpy(), g["ru_inv_inmat"])
    r, t = O.base_camera()
    assert np.array_equal(r.numpy(), g["in_R"][:1]) and np.array_equal(t.numpy(), g["in_T"][:1])
    cams = O.orbit_cameras(45)
    assert np.allclose(cams[7][0].numpy(), g["ru_orbit7_R"], atol=1e-7)
    assert np.allclose(cams[7][1].numpy(), g["ru_orbit7_T"], atol=1e-6)


def test_sampling_and_posenc(tiny_golden):
    g = tiny_golden
    opt = _opt(g["meta"])
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    s = O.sample_points(xy, R, T, K, opt.num_sample_coarse, opt.world_z1, opt.world_z2)
    assert np.array_equal(s["pts"].numpy(), g["smp_pts"])  # same torch ops -> bit-equal
    assert np.array_equal(s["zvals"].numpy(), g["smp_zvals"])
    assert np.array_equal(s["z_dists"].numpy(), g["smp_z_dists"])
    assert np.array_equal(s["ray_d"].unsqueeze(-1).numpy(), g["smp_batch_ray_d"])
    assert np.array_equal(s["ray_l"].unsqueeze(-1).numpy(), g["smp_batch_ray_l"])
    pe = O.posenc(s["pts"], 10, True)
    assert pe.shape[1] == 63
    assert np.array_equal(pe.numpy(), g["pe"])


def test_mlp_composite_tiny(tiny_golden):
    g = tiny_golden
    opt = _opt(g["meta"])
    sd = golden_state_dict(g)
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    s = O.sample_points(xy, R, T, K, opt.num_sample_coarse, opt.world_z1, opt.world_z2)
    br = O.render_branches(sd, opt, s["pts"], s["z_dists"], s["zvals"], shape, appea, gaze)
    for name in ("face", "eyes"):
        assert rel_l2(br[name][0], g["feat_" + name]) < TOL
        assert max_rel(br[name][1], g["bg_alpha_" + name]) < 2e-5  # 1 - sum(w): cancellation
        assert rel_l2(br[name][2], g["w_" + name]) < 1e-5
    # the dense variant must not be vacuous (SURVEY §8c caveat 1)
    assert g["bg_alpha_face"].min() < 0.5


def test_mlp_points_match_reference_per_point(tiny_golden):
    g = tiny_golden
    opt = _opt(g["meta"])
    sd = golden_state_dict(g)
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    s = O.sample_points(xy, R, T, K, opt.num_sample_coarse, opt.world_z1, opt.world_z2)
    pe = O.posenc(s["pts"])
    feat, sigma = O.mlp_branch(sd, "fg_CD_predictor_face", pe, torch.cat([shape, gaze], 1), appea)
    n_r = pe.shape[2]
    assert rel_l2(feat[:, :, :: max(1, n_r // 16), :], g["mlp_feat_face"]) < TOL
    assert rel_l2(sigma, g["mlp_sigma_face"]) < TOL


@pytest.mark.parametrize("name", ["tiny"])
def test_full_forward_tiny(name):
    g = load_golden(name)
    opt = _opt(g["meta"])
    sd = golden_state_dict(g)
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    out = O.forward(sd, opt, "test", xy, shape, appea, gaze, R, T, K, return_stages=True)
    for k in ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img"):
        ref = g["img_" + k]
        assert out["coarse_dict"][k].shape == ref.shape
        assert float((out["coarse_dict"][k] - torch.from_numpy(ref)).abs().max()) < 2e-6, k


def test_neural_render_pieces():
    # bilinear_up2 / blur3x3 / pixel_shuffle2 restatements against the torch ops the reference calls
    import torch.nn.functional as F
    torch.manual_seed(1)
    x = torch.randn(2, 5, 6, 8)
    up = torch.nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)(x)
    assert float((O.bilinear_up2(x) - up).abs().max()) < 1e-6
    assert torch.equal(O.pixel_shuffle2(torch.arange(2 * 8 * 3 * 3).float().view(2, 8, 3, 3)),
                       F.pixel_shuffle(torch.arange(2 * 8 * 3 * 3).float().view(2, 8, 3, 3), 2))
    k = torch.tensor([1.0, 2.0, 1.0])
    k2 = (k[None, :] * k[:, None] / 16.0).view(1, 1, 3, 3).repeat(5, 1, 1, 1)
    ref = F.conv2d(F.pad(x, [1, 1, 1, 1], mode="reflect"), k2, groups=5)
    assert float((O.blur3x3(x) - ref).abs().max()) < 1e-6


def test_fine_sample_tiny(tiny_golden):
    g = tiny_golden
    opt = _opt(g["meta"])
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    s = O.sample_points(xy, R, T, K, opt.num_sample_coarse, opt.world_z1, opt.world_z2)
    fs = O.fine_sample(torch.from_numpy(g["w_face"]), s["zvals"], s["ray_o"], s["ray_d"], s["ray_l"], n_fine=8)
    assert np.array_equal(fs["inds"].numpy(), g["fine_inds"])  # integer work: bit-exact
    assert np.array_equal(fs["zvals"].numpy(), g["fine_zvals"])
    assert np.array_equal(fs["z_dists"].numpy(), g["fine_z_dists"])
    assert np.array_equal(fs["pts"].numpy(), g["fine_pts"])
    assert fs["zvals"].shape[-1] == 8 + 8  # N_c + N_f sorted samples


def test_jitter_matches_reference():
    g = load_golden("std_dense_train")
    opt = _opt(g["meta"])
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    s = O.sample_points(xy, R, T, K, opt.num_sample_coarse, opt.world_z1, opt.world_z2, jitter_u=torch.from_numpy(g["jitter_u"]))
    assert np.array_equal(s["zvals"].numpy(), g["smp_zvals"])
    assert np.array_equal(s["pts"].numpy(), g["smp_pts"])


def _mid_state_dict(g):
    """Weights of the mid fixture: the drop-in module's reference-identical seeded init (checksums stored in the fixture) +
    the dense-density variant with the fixture's scale / biases."""
    import gazenerf_b200 as G
    meta = g["meta"]
    opt = G.BaseOptions({"featmap_size": int(meta[0]), "featmap_nc": int(meta[1]), "pred_img_size": int(meta[2])})
    opt.num_sample_coarse = int(meta[3])
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        chk = g["chk/" + k]
        assert abs(float(v.double().sum()) - chk[0]) <= 1e-9 * max(1.0, abs(chk[1])), k
    return O.densify(sd, *g["dense_bias"], scale=float(g["dense_scale"][0]))


def test_mid_size_reference_golden_16x16x64():
    """16x16 rays x 64 samples (two rays per 128-point tile of the fused kernel), real layer widths, B = 2: the oracle against the
    reference's own outputs -- features, bg_alpha, weights, the four 128x128 images and FineSample(64) int64 indices."""
    g = load_golden("mid_dense_test")
    opt = _opt(g["meta"], n_fine=64)
    sd = _mid_state_dict(g)
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    out = O.forward(sd, opt, "test", xy, shape, appea, gaze, R, T, K, return_stages=True)
    st = out["stages"]
    b = xy.shape[0]
    for name in ("face", "eyes"):
        assert rel_l2(st["feat_" + name].reshape(b, 258, -1), g["feat_" + name]) < TOL, name
        assert max_rel(st["bg_alpha_" + name].reshape(b, 1, -1), g["bg_alpha_" + name]) < 2e-5, name
        assert rel_l2(st["w_" + name], g["w_" + name]) < 1e-5, name
    for k in ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img"):
        assert float((out["coarse_dict"][k] - torch.from_numpy(g["img_" + k])).abs().max()) < 5e-6, k
    fs = O.fine_sample(torch.from_numpy(g["w_face"]), st["zvals"], st["ray_o"], st["ray_d"], st["ray_l"], n_fine=64)
    assert np.array_equal(fs["inds"].numpy(), g["fine_inds"])   # integer work: bit-exact
    assert np.array_equal(fs["zvals"].numpy(), g["fine_zvals"])
    assert fs["zvals"].shape[-1] == 64 + 64
    assert 0.1 < float(g["bg_alpha_face"].mean()) < 0.9   # non-vacuous


def _vd_state_dict(g):
    """Weights of the include_vd fixture: the drop-in module's seeded init (checksums + shapes stored) + the fixture's dense variant."""
    import gazenerf_b200 as G
    opt = G.BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=True, hier_sampling=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["shape/" + k]), k          # state_dict shapes of the reference's include_vd=True network
        chk = g["chk/" + k]
        assert abs(float(v.double().sum()) - chk[0]) <= 1e-9 * max(1.0, abs(chk[1])), k
    sd = O.densify(sd, *g["dense_bias"])
    for br in ("face", "eyes"):
        k = "fg_CD_predictor_%s.RGB_layer_1.weight" % br
        w = sd[k].clone()
        w[:, 384:384 + 27] *= float(g["vd_col_scale"][0])
        sd[k] = w
    return net, sd


def test_include_vd_reference_golden():
    """include_vd=True (models/gaze_nerf.py:70-80,140-141): the oracle with the 27-channel view-direction encoding against the
    reference's own outputs; the drop-in module reproduces the reference's parameter shapes (RGB_layer_1: 192 x 538) and seeded init."""
    g = load_golden("std_dense_vd_test")
    opt = _opt(g["meta"])
    net, sd = _vd_state_dict(g)
    assert sd["fg_CD_predictor_face.RGB_layer_1.weight"].shape == (192, 384 + 27 + 127, 1, 1)
    xy, shape, appea, gaze, R, T, K = _inputs(g)
    out = O.forward(sd, opt, "test", xy, shape, appea, gaze, R, T, K, return_stages=True, include_vd=True)
    st = out["stages"]
    vd = O.posenc(st["ray_d"].unsqueeze(-1), 4, True)[..., 0]
    assert np.array_equal(vd.numpy(), g["vd_pe"])
    for name in ("face", "eyes"):
        assert rel_l2(st["feat_" + name].reshape(2, 258, -1), g["feat_" + name]) < TOL, name
        assert max_rel(st["bg_alpha_" + name].reshape(2, 1, -1), g["bg_alpha_" + name]) < 2e-5, name
    for k in ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img"):
        assert float((out["coarse_dict"][k] - torch.from_numpy(g["img_" + k])).abs().max()) < 5e-6, k
    # the view direction matters in this fixture: dropping it changes the features visibly
    no_vd = {k: (v[:, list(range(384)) + list(range(411, 538))] if k.endswith("RGB_layer_1.weight") else v) for k, v in sd.items()}
    out0 = O.forward(no_vd, opt, "test", xy, shape, appea, gaze, R, T, K, return_stages=True)
    assert rel_l2(out0["stages"]["feat_face"].reshape(2, 258, -1), g["feat_face"]) > 1e-3
