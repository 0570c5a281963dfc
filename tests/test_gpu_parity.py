"""GPU parity tests: every libgnrf stage and the full drop-in forward against the CPU oracle / the reference goldens.

All calls go through the C ABI (ctypes) -- either directly or via gazenerf_b200.GazeNeRFNet.  Tolerances:
  * integer / index work (fine-sample searchsorted indices): bit-exact;
  * geometry (rays, depths, sample positions): <= 2e-6 relative (1-ulp re-association between CPU BLAS and the kernel);
  * rendered features / RGB: <= 1e-3 relative (north_star), measured both as relative L2 and as max|err|/max|ref|.
    The fp32 CUDA-core path and the bf16x3 tensor-core path are additionally held to 2e-4, far inside the bar.
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden, max_rel, rel_l2
from oracle import gazenerf_oracle as O

import gazenerf_b200 as G
from gazenerf_b200 import _lib

pytestmark = pytest.mark.gpu

TOL_FEAT = 1e-3     # north_star bar
TOL_TIGHT = 2e-4    # what the kernels are actually held to
TOL_GEOM = 2e-6


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    L = _lib.lib()
    _lib.check(L.gnrf_device_check(), "gnrf_device_check")
    return torch.device("cuda:0")


def S():
    return torch.cuda.current_stream().cuda_stream


def _net_from_golden(g, dev, mlp_impl, dense=True, hier=False):
    meta = g["meta"]
    opt = G.BaseOptions({"featmap_size": int(meta[0]), "featmap_nc": int(meta[1]), "pred_img_size": int(meta[2])})
    opt.num_sample_coarse = int(meta[3])
    opt.mlp_hidden_nchannels = int(meta[4])
    opt.num_sample_fine = 8
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=hier, mlp_impl=mlp_impl)
    if any(k.startswith("sd/") for k in g):
        sd = golden_state_dict(g)
        if hier:
            sd.update({k: v for k, v in net.state_dict().items() if k.startswith("fine_fg_CD_predictor")})
        net.load_state_dict(sd, strict=True)  # a reference checkpoint loads strictly
    elif int(meta[5]):
        sd = O.densify({k: v.clone() for k, v in net.state_dict().items()}, *g["dense_bias"])
        net.load_state_dict(sd, strict=True)
    return opt, net.to(dev).eval()


def _inputs(g, dev):
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    return dict(batch_xy=t("in_xy"), batch_uv=None, bg_code=None, shape_code=t("in_shape"), appea_code=t("in_appea"),
                gaze_code=t("in_gaze"), batch_Rmats=t("in_R"), batch_Tvecs=t("in_T"), batch_inv_inmats=t("in_Kinv"))


# ----------------------------------------------------------------------------------------------- geometry
def test_ray_setup_and_depths(dev, tiny_golden):
    g = tiny_golden
    L = _lib.lib()
    xy, R, T, K = (torch.from_numpy(g[k]).to(dev) for k in ("in_xy", "in_R", "in_T", "in_Kinv"))
    B, _, n_r = xy.shape
    n_s = int(g["meta"][3])
    ray_dl = torch.empty(B, n_r, 4, device=dev)
    _lib.check(L.gnrf_ray_setup(xy.data_ptr(), R.data_ptr(), K.data_ptr(), B, n_r, ray_dl.data_ptr(), S()))
    d_ref = torch.from_numpy(g["smp_batch_ray_d"])[..., 0].permute(0, 2, 1)
    l_ref = torch.from_numpy(g["smp_batch_ray_l"])[:, 0, :, 0]
    assert max_rel(ray_dl[..., :3].cpu(), d_ref) < TOL_GEOM
    assert max_rel(ray_dl[..., 3].cpu(), l_ref) < TOL_GEOM
    tv = torch.linspace(0, 1, n_s + 1).to(dev)
    z = torch.empty(B, n_r, n_s + 1, device=dev)
    _lib.check(L.gnrf_coarse_depths(T.reshape(B, 3).contiguous().data_ptr(), tv.data_ptr(), None, B, n_r, n_s, 2.5, -3.5, z.data_ptr(), S()))
    assert np.array_equal(z[..., :-1].cpu().numpy(), g["smp_zvals"][:, 0])  # individually rounded ops: bit-equal
    zd = (z[..., 1:] - z[..., :-1]) * ray_dl[..., 3:4]
    assert max_rel(zd.cpu(), g["smp_z_dists"][:, 0]) < TOL_GEOM


def test_jittered_depths_bit_equal(dev):
    g = load_golden("std_dense_train")
    L = _lib.lib()
    T = torch.from_numpy(g["in_T"]).to(dev).reshape(-1, 3).contiguous()
    u = torch.from_numpy(g["jitter_u"]).to(dev)
    B, n_r, n_e = u.shape
    tv = torch.linspace(0, 1, n_e).to(dev)
    z = torch.empty(B, n_r, n_e, device=dev)
    _lib.check(L.gnrf_coarse_depths(T.data_ptr(), tv.data_ptr(), u.data_ptr(), B, n_r, n_e - 1, 2.5, -3.5, z.data_ptr(), S()))
    assert np.array_equal(z[..., :-1].cpu().numpy(), g["smp_zvals"][:, 0])


def _fine_case(dev, w, z_edges, n_f1, u=None):
    L = _lib.lib()
    B, n_r, n_c = w.shape
    per_ray = u is not None
    if u is None:
        u = torch.linspace(0, 1, n_f1)
    inds = torch.empty(B * n_r, n_f1, dtype=torch.int64, device=dev)
    zf = torch.empty(B, n_r, n_c + n_f1, device=dev)
    w_d, z_d, u_d = w.to(dev).contiguous(), z_edges.to(dev).contiguous(), u.to(dev).contiguous()  # keep alive across the call
    _lib.check(L.gnrf_fine_depths(w_d.data_ptr(), z_d.data_ptr(), u_d.data_ptr(), 1 if per_ray else 0, B, n_r, n_c, n_f1,
                                  inds.data_ptr(), zf.data_ptr(), S()))
    torch.cuda.synchronize()
    return inds.cpu(), zf.cpu()


def test_fine_sample_indices_bit_exact_vs_reference(dev, tiny_golden):
    g = tiny_golden
    w = torch.from_numpy(g["w_face"])[:, 0]
    zc = torch.from_numpy(g["smp_zvals"])[:, 0]
    z_edges = torch.cat([zc, zc[..., -1:] + 1.0], -1)  # the last edge is not used by FineSample
    inds, zf = _fine_case(dev, w, z_edges, 9)
    assert np.array_equal(inds.numpy(), g["fine_inds"])
    # golden fine_zvals = sorted[:-1]
    assert max_rel(zf[..., :-1], g["fine_zvals"][:, 0]) < TOL_GEOM


@pytest.mark.parametrize("n_c,n_f,per_ray", [(64, 64, False), (64, 128, False), (16, 32, True), (3, 5, False)])
def test_fine_sample_vs_oracle_random(dev, n_c, n_f, per_ray):
    gen = torch.Generator().manual_seed(n_c * 1000 + n_f)
    B, n_r = 2, 96
    w = torch.rand(B, n_r, n_c, generator=gen) ** 3
    w[0, :5] = 0.0          # all-zero weights (empty ray): denom < 1e-5 branch
    w[1, 7, 10:] = 0.0      # ragged support
    zc = torch.sort(torch.rand(B, n_r, n_c + 1, generator=gen) * 6 + 9.5, -1)[0]
    u = torch.rand(B * n_r, n_f + 1, generator=gen) if per_ray else None
    inds, zf = _fine_case(dev, w, zc, n_f + 1, u)
    o = torch.zeros(B, 3, n_r)
    d = torch.ones(B, 3, n_r)
    l = torch.ones(B, 1, n_r)
    fs = O.fine_sample(w.unsqueeze(1), zc[..., :-1].unsqueeze(1), o, d, l, n_f, u)
    mism = (inds != fs["inds"]).float().mean().item()
    assert mism == 0.0, "searchsorted indices differ on %.4f%% of draws" % (100 * mism)
    # depths: the inverse-CDF lerp divides by (cdf[above]-cdf[below]) >= 1e-5, so a 1-ulp difference in the cdf (torch's
    # vectorised float sum vs the kernel's correctly-rounded sum) moves a fine depth by up to 6e-8/1e-5 * bin_width ~ 6e-4
    # absolute (4e-5 of the depth range) in near-empty bins; everywhere else agreement is ~1e-7.
    assert max_rel(zf, fs["z_sorted"]) < 1e-4
    assert float((zf - fs["z_sorted"]).abs().median()) < 2e-6
    assert bool((zf[..., 1:] >= zf[..., :-1]).all())  # sortedness property


# ----------------------------------------------------------------------------------------------- composite / compose
def test_composite_vs_oracle(dev):
    L = _lib.lib()
    gen = torch.Generator().manual_seed(3)
    B, n_r, n_s, C = 2, 37, 24, 258
    feat = torch.randn(B, n_r, n_s, C, generator=gen)
    sigma = torch.relu(torch.randn(B, n_r, n_s, generator=gen)) * 8
    z = torch.sort(torch.rand(B, n_r, n_s + 1, generator=gen) * 6 + 9.5, -1)[0]
    ray_dl = torch.randn(B, n_r, 4, generator=gen)
    ray_dl[..., 3] = 1.0 + torch.rand(B, n_r, generator=gen) * 0.1
    out_f = torch.empty(B, C, n_r, device=dev); out_a = torch.empty(B, n_r, device=dev)
    out_d = torch.empty(B, n_r, device=dev); out_w = torch.empty(B, n_r, n_s, device=dev)
    feat_d, sigma_d, z_d, ray_d = feat.to(dev), sigma.to(dev), z.to(dev), ray_dl.to(dev)  # keep alive across the call
    _lib.check(L.gnrf_composite_fwd(feat_d.data_ptr(), sigma_d.data_ptr(), z_d.data_ptr(), ray_d.data_ptr(),
                                    B, n_r, n_s, C, out_f.data_ptr(), out_a.data_ptr(), out_d.data_ptr(), out_w.data_ptr(), S()))
    torch.cuda.synchronize()
    zd = ((z[..., 1:] - z[..., :-1]) * ray_dl[..., 3:4]).unsqueeze(1)
    fr, ba, dep, w = O.composite(feat.permute(0, 3, 1, 2), sigma.unsqueeze(1), zd, z[..., :-1].unsqueeze(1))
    assert rel_l2(out_f.cpu(), fr) < 2e-6
    assert max_rel(out_a.cpu(), ba[:, 0]) < 2e-5  # 1 - sum(w): cancellation
    assert rel_l2(out_d.cpu(), dep[:, 0]) < 2e-6
    assert rel_l2(out_w.cpu(), w[:, 0]) < 2e-6


def test_compose_vs_oracle(dev):
    L = _lib.lib()
    gen = torch.Generator().manual_seed(5)
    B, C, Ssz = 3, 258, 8
    P = Ssz * Ssz
    ff, fe = torch.randn(B, C, P, generator=gen), torch.randn(B, C, P, generator=gen)
    af, ae = torch.rand(B, P, generator=gen), torch.rand(B, P, generator=gen)
    bg = torch.randn(C, P, generator=gen)
    gaze = torch.rand(B, 2, generator=gen) - 0.5
    out = torch.empty(3, B, C, P, device=dev)
    d_ = [t.to(dev) for t in (ff, af, fe, ae, bg, gaze)]  # keep alive across the call
    _lib.check(L.gnrf_compose_fwd(*[t.data_ptr() for t in d_], B, C, P, out.data_ptr(), S()))
    torch.cuda.synchronize()
    v = lambda t: t.view(B, -1, Ssz, Ssz)
    mf, ep, mg = O.compose_featmaps(v(ff), v(af), v(fe), v(ae), bg.view(1, C, Ssz, Ssz), gaze)
    o = out.cpu().view(3, B, C, Ssz, Ssz)
    assert max_rel(o[0], mf) < 2e-6 and max_rel(o[1], ep) < 2e-6 and max_rel(o[2], mg) < 2e-6


# ----------------------------------------------------------------------------------------------- neural renderer
@pytest.mark.parametrize("impl", ["simt", "tc", "tc_layerwise"])
@pytest.mark.parametrize("C,Ssz,nb", [(48, 8, 3), (258, 8, 3), (258, 16, 2), (258, 32, 2), (96, 16, 3)])
def test_neural_render_vs_oracle(dev, C, Ssz, nb, impl):
    torch.manual_seed(11)
    nr = G.NeuralRendererParams(feat_nc=C, featmap_size=Ssz, img_size=Ssz << nb)
    sd = {"neural_render." + k: v for k, v in nr.state_dict().items()}
    x = torch.randn(3, C, Ssz, Ssz)
    ref = O.neural_render(sd, x, nb)
    nr = nr.to(dev)
    nr.impl = impl  # fp32 CUDA-core convs | tcgen05 bf16x3: fused level kernels where supported (S*S % 128 == 0) | layer-wise convs
    img = nr(x.to(dev)).cpu()
    assert img.shape == ref.shape
    assert float((img - ref).abs().max()) < 2e-5
    # edge rows/cols are where reflect/clamp borders live
    assert float((img[..., 0, :] - ref[..., 0, :]).abs().max()) < 2e-5 and float((img[..., :, -1] - ref[..., :, -1]).abs().max()) < 2e-5


# ----------------------------------------------------------------------------------------------- full forward
def _check_forward(net, g, dev, tol):
    net.keep_stages = True
    kw = _inputs(g, dev)
    train = bool(g["meta"][6])
    extra = {"jitter_u": torch.from_numpy(g["jitter_u"]).to(dev)} if train else {}
    with torch.no_grad():
        out = net("train" if train else "test", **kw, **extra)
    st = net.last_stages
    for br in ("face", "eyes"):
        assert rel_l2(st["feat_" + br].cpu(), g["feat_" + br]) < tol, br
        assert max_rel(st["feat_" + br].cpu(), g["feat_" + br]) < tol, br
        assert max_rel(st["bg_alpha_" + br].cpu(), g["bg_alpha_" + br][:, 0]) < tol, br
        assert rel_l2(st["w_" + br].cpu(), g["w_" + br][:, 0]) < tol, br
    for k in ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img"):
        ref = torch.from_numpy(g["img_" + k])
        got = out["coarse_dict"][k].cpu()
        assert got.shape == ref.shape, k
        assert float((got - ref).abs().max()) < tol, k  # images live in (0,1): absolute == relative-to-range
    return out


def test_forward_simt_tiny_golden(dev, tiny_golden):
    opt, net = _net_from_golden(tiny_golden, dev, "simt")
    _check_forward(net, tiny_golden, dev, TOL_TIGHT)


@pytest.mark.parametrize("name", ["std_refinit_test", "std_dense_test", "std_dense_train"])
def test_forward_simt_std_golden(dev, name):
    g = load_golden(name)
    opt, net = _net_from_golden(g, dev, "simt")
    _check_forward(net, g, dev, TOL_TIGHT)


@pytest.mark.parametrize("name", ["std_refinit_test", "std_dense_test", "std_dense_train"])
def test_forward_tc_std_golden(dev, name):
    g = load_golden(name)
    opt, net = _net_from_golden(g, dev, "tc")
    assert net._tc_supported(opt.num_sample_coarse)
    _check_forward(net, g, dev, TOL_TIGHT)


def test_tc_matches_simt_and_oracle_midsize(dev):
    """16x16 rays x 64 samples (2 rays per 128-row tile), orbit cameras, both kernels vs the oracle."""
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 64
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    ru = G.RenderUtils(45, "cpu", opt)
    B = 2
    shape, appea, gaze = O.synthetic_codes(B)
    cams = [ru.cam_info_list[3], ru.cam_info_list[20]]
    cam = {k: torch.cat([c[k] for c in cams], 0) for k in cams[0]}
    xy = ru.ray_xy.expand(B, -1, -1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    oo = O.OracleOptions(featmap_size=16, featmap_nc=258, pred_img_size=64, num_sample_coarse=64)
    bf, be = O.calibrate_dense_bias(sd, oo, xy, shape, appea, gaze, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"],
                                    scale=4.0)
    sd = O.densify(sd, bf, be, scale=4.0)
    net.load_state_dict(sd)
    ref = O.forward(sd, oo, "test", xy, shape, appea, gaze, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"], return_stages=True)
    assert 0.05 < float(ref["stages"]["bg_alpha_face"].mean()) < 0.95  # non-vacuous
    net = net.to(dev).eval()
    net.keep_stages = True
    kw = dict(batch_xy=xy.to(dev), batch_uv=None, bg_code=None, shape_code=shape.to(dev), appea_code=appea.to(dev), gaze_code=gaze.to(dev),
              **{k: v.to(dev) for k, v in cam.items()})
    res = {}
    for impl in ("simt", "tc"):
        net.mlp_impl = impl
        out = net("test", **kw)
        st = net.last_stages
        for br in ("face", "eyes"):
            assert rel_l2(st["feat_" + br].cpu().view(B, 258, 16, 16), ref["stages"]["feat_" + br]) < TOL_TIGHT, (impl, br)
            assert max_rel(st["bg_alpha_" + br].cpu().view(B, 1, 16, 16), ref["stages"]["bg_alpha_" + br]) < TOL_TIGHT, (impl, br)
        for k, v in ref["coarse_dict"].items():
            assert float((out["coarse_dict"][k].cpu() - v).abs().max()) < TOL_TIGHT, (impl, k)
        res[impl] = out
    for k in res["tc"]["coarse_dict"]:
        assert float((res["tc"]["coarse_dict"][k] - res["simt"]["coarse_dict"][k]).abs().max()) < TOL_TIGHT


def test_hier_forward_vs_oracle(dev):
    """BASELINE config 3 building blocks at small size: coarse 16 -> fine 16+... sorted samples, indices bit-exact."""
    g = load_golden("std_dense_test")
    for impl in ("simt", "tc"):
        opt, net = _net_from_golden(g, dev, impl, hier=True)
        net.num_sample_fine = 8
        net.keep_stages = True
        out = net("test", **_inputs(g, dev))
        st = net.last_stages
        assert "fine_dict" in out and out["fine_dict"]["merge_img"].shape == out["coarse_dict"]["merge_img"].shape
        # feed the ORACLE's weights to the kernel for the bit-exact index claim (SURVEY §8c caveat 2) -> done in
        # test_fine_sample_indices_bit_exact_vs_reference; here the pipeline's own weights must reproduce the oracle pipeline
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        oo = O.OracleOptions(featmap_size=8, featmap_nc=258, pred_img_size=64, num_sample_coarse=8)
        t = lambda k: torch.from_numpy(g[k])
        ref = O.forward_hier(sd, oo, t("in_xy"), t("in_shape"), t("in_appea"), t("in_gaze"), t("in_R"), t("in_T"), t("in_Kinv"), n_fine=8)
        assert (st["fine_inds"].cpu() != ref["fine_sample"]["inds"]).float().mean().item() < 0.01
        assert max_rel(st["z_fine"].cpu(), ref["fine_sample"]["z_sorted"]) < 1e-4
        for i, br in enumerate(("face", "eyes")):
            assert rel_l2(st["fine_feat_" + br].cpu(), ref["fine"][br][0]) < 5e-3, (impl, br)


def test_errors_raise_not_abort(dev):
    L = _lib.lib()
    assert L.gnrf_ray_setup(None, None, None, 1, 1, None, S()) == 1
    assert b"argument check failed" in L.gnrf_last_error()
    with pytest.raises(RuntimeError, match="libgnrf"):
        _lib.check(L.gnrf_compose_fwd(1, 1, 1, 1, 1, 1, 1, 4, 4, 1, S()), "gnrf_compose_fwd")


# ----------------------------------------------------------------------------------------------- full-size / edge configurations
def _default_net(dev, hier=False, n_s=64, n_fine=64):
    opt = G.BaseOptions()
    opt.num_sample_coarse = n_s
    opt.num_sample_fine = n_fine
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=hier)
    return opt, net


def _full_inputs(opt, B, cams_idx):
    ru = G.RenderUtils(45, "cpu", opt)
    shape, appea, gaze = O.synthetic_codes(B)
    cams = [ru.cam_info_list[i] for i in cams_idx]
    cam = {k: torch.cat([c[k] for c in cams], 0) for k in cams[0]}
    return ru, dict(batch_xy=ru.ray_xy.expand(B, -1, -1), batch_uv=None, bg_code=None, shape_code=shape, appea_code=appea, gaze_code=gaze, **cam)


def test_full_size_config2_tc_vs_simt_and_oracle_rays(dev):
    """BASELINE config[1] size (64x64 rays x 64 samples -> 512x512), B=2 with two different orbit cameras:
    (a) the folded bf16x3 tcgen05 kernel and the literal fp32 kernel agree on every ray / every output pixel;
    (b) rays are independent, so the CPU oracle on every 32nd ray must match those rays of the full GPU render."""
    opt, net = _default_net(dev)
    B = 2
    ru, kw = _full_inputs(opt, B, [5, 31])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    oo = O.OracleOptions()
    bf, be = O.calibrate_dense_bias(sd, oo, kw["batch_xy"], kw["shape_code"], kw["appea_code"], kw["gaze_code"], kw["batch_Rmats"],
                                    kw["batch_Tvecs"], kw["batch_inv_inmats"], scale=4.0)
    sd = O.densify(sd, bf, be, scale=4.0)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    net.keep_stages = True
    dkw = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()}
    outs, stages = {}, {}
    for impl in ("tc", "simt"):
        net.mlp_impl = impl
        net.neural_render.impl = impl
        outs[impl] = {k: v.clone() for k, v in net("test", **dkw)["coarse_dict"].items()}
        stages[impl] = {k: net.last_stages[k].clone() for k in ("feat_face", "feat_eyes", "bg_alpha_face", "bg_alpha_eyes", "w_face")}
    for k in stages["tc"]:
        assert rel_l2(stages["tc"][k].cpu(), stages["simt"][k].cpu()) < TOL_TIGHT, k
    for k in outs["tc"]:
        assert outs["tc"][k].shape[-1] == 512
        assert float((outs["tc"][k] - outs["simt"][k]).abs().max()) < TOL_TIGHT, k
    a = stages["tc"]["bg_alpha_face"]
    assert 0.05 < float(a.mean()) < 0.95 and float(a.min()) < 0.2 and float(a.max()) > 0.5  # non-vacuous: mixed opacity
    # (b) oracle on a ray subset
    step = 32
    sub = O.sample_points(kw["batch_xy"][:, :, ::step].contiguous(), kw["batch_Rmats"], kw["batch_Tvecs"], kw["batch_inv_inmats"], 64, 2.5, -3.5)
    br = O.render_branches(sd, oo, sub["pts"], sub["z_dists"], sub["zvals"], kw["shape_code"], kw["appea_code"], kw["gaze_code"])
    for name in ("face", "eyes"):
        got = stages["tc"]["feat_" + name].cpu()[:, :, ::step]
        assert rel_l2(got, br[name][0]) < TOL_TIGHT, name
        assert max_rel(stages["tc"]["bg_alpha_" + name].cpu()[:, ::step], br[name][1][:, 0]) < TOL_TIGHT, name


def test_unsupported_sample_count_uses_fp32_cuda_path(dev):
    """N_s = 24 does not divide the 128-row tile of the tcgen05 kernel: the module must run the literal fp32 CUDA kernels (still
    libgnrf, never a CPU path) and match the oracle."""
    opt = G.BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 24
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    assert not net._tc_supported(24)
    ru, kw = _full_inputs(opt, 1, [9])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    oo = O.OracleOptions(featmap_size=8, featmap_nc=258, pred_img_size=64, num_sample_coarse=24)
    ref = O.forward(sd, oo, "test", kw["batch_xy"], kw["shape_code"], kw["appea_code"], kw["gaze_code"], kw["batch_Rmats"], kw["batch_Tvecs"],
                    kw["batch_inv_inmats"])
    net = net.to(dev).eval()
    n0 = _lib.lib().gnrf_launch_count()
    out = net("test", **{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()})
    assert _lib.lib().gnrf_launch_count() > n0
    for k, v in ref["coarse_dict"].items():
        assert float((out["coarse_dict"][k].cpu() - v).abs().max()) < TOL_TIGHT, k


def test_hier_config3_sizes_one_ray_per_tile(dev):
    """coarse 64 + fine 64 -> 128 sorted samples per ray (1 ray per 128-row tile), 16x16 rays: fused kernel vs the oracle pipeline."""
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 64
    opt.num_sample_fine = 64
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=True)
    ru, kw = _full_inputs(opt, 1, [14])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    oo = O.OracleOptions(featmap_size=16, featmap_nc=258, pred_img_size=64, num_sample_coarse=64, num_sample_fine=64)
    bf, be = O.calibrate_dense_bias(sd, oo, kw["batch_xy"], kw["shape_code"], kw["appea_code"], kw["gaze_code"], kw["batch_Rmats"],
                                    kw["batch_Tvecs"], kw["batch_inv_inmats"], scale=4.0)
    sd = O.densify(sd, bf, be, scale=4.0)
    net.load_state_dict(sd)
    ref = O.forward_hier(sd, oo, kw["batch_xy"], kw["shape_code"], kw["appea_code"], kw["gaze_code"], kw["batch_Rmats"], kw["batch_Tvecs"],
                         kw["batch_inv_inmats"], n_fine=64)
    net = net.to(dev).eval()
    net.keep_stages = True
    assert net._tc_supported(128)
    out = net("test", **{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()})
    st = net.last_stages
    assert st["z_fine"].shape[-1] == 64 + 65 and "fine_dict" in out
    assert (st["fine_inds"].cpu() != ref["fine_sample"]["inds"]).float().mean().item() < 0.01
    assert bool((st["z_fine"][..., 1:] >= st["z_fine"][..., :-1]).all())
    for br in ("face", "eyes"):
        assert rel_l2(st["fine_feat_" + br].cpu(), ref["fine"][br][0]) < 5e-3, br


def test_repeat_calls_and_weight_update_refresh_packed_cache(dev):
    """Packed weights are a cache keyed on parameter versions: an in-place optimizer-style update must be picked up."""
    g = load_golden("std_dense_test")
    opt, net = _net_from_golden(g, dev, "tc")
    kw = _inputs(g, dev)
    a = net("test", **kw)["coarse_dict"]["merge_img"].clone()
    b = net("test", **kw)["coarse_dict"]["merge_img"].clone()
    assert torch.equal(a, b)  # deterministic
    with torch.no_grad():
        net.fg_CD_predictor_face.RGB_layer_2.bias.add_(0.5)
        net.neural_render.feat_layers[0].bias.add_(0.1)
    c = net("test", **kw)["coarse_dict"]["merge_img"]
    assert float((c - a).abs().max()) > 1e-3
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    oo = O.OracleOptions(featmap_size=8, featmap_nc=258, pred_img_size=64, num_sample_coarse=8)
    t = lambda k: torch.from_numpy(g[k])
    ref = O.forward(sd, oo, "test", t("in_xy"), t("in_shape"), t("in_appea"), t("in_gaze"), t("in_R"), t("in_T"), t("in_Kinv"))
    assert float((c.cpu() - ref["coarse_dict"]["merge_img"]).abs().max()) < TOL_TIGHT


def test_graphed_forward_matches_eager(dev):
    """net.graphed(): the forward captured into a CUDA graph and replayed with NEW inputs equals the eager forward bit for bit."""
    g = load_golden("std_dense_test")
    opt, net = _net_from_golden(g, dev, "tc")
    kw = _inputs(g, dev)
    gf = net.graphed("test", **kw)
    assert gf.launches_per_replay > 10
    kw2 = dict(kw)
    kw2["gaze_code"] = kw["gaze_code"] + 0.1
    kw2["shape_code"] = kw["shape_code"] * 0.5
    with torch.no_grad():
        ref = {k: v.clone() for k, v in net("test", **kw2)["coarse_dict"].items()}
    out = gf(**kw2)["coarse_dict"]
    torch.cuda.synchronize()
    for k in ref:
        assert torch.equal(out[k], ref[k]), k
    with torch.no_grad():
        net.neural_render.bg_featmap.mul_(0.9)
    with pytest.raises(RuntimeError):
        gf(**kw2)


def test_include_vd_forward_vs_reference_golden_and_oracle(dev):
    """GazeNeRFNet(include_vd=True): the 27-channel view-direction encoding (models/gaze_nerf.py:70-80,140-141) as a per-ray bias of the
    fused kernel's last stage, against the reference's own include_vd=True outputs (8x8 rays) and against the oracle at 16x16x64."""
    from test_oracle_golden import _vd_state_dict
    g = load_golden("std_dense_vd_test")
    net, sd = _vd_state_dict(g)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    net.keep_stages = True
    with torch.no_grad():
        out = net("test", **_inputs(g, dev))
    st = net.last_stages
    for br in ("face", "eyes"):
        assert rel_l2(st["feat_" + br].cpu(), g["feat_" + br]) < TOL_TIGHT, br
        assert max_rel(st["bg_alpha_" + br].cpu(), g["bg_alpha_" + br][:, 0]) < TOL_TIGHT, br
    for k in ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img"):
        assert float((out["coarse_dict"][k].cpu() - torch.from_numpy(g["img_" + k])).abs().max()) < TOL_TIGHT, k
    with pytest.raises(NotImplementedError):
        net("train", **{k: (v.requires_grad_(True) if k == "shape_code" else v) for k, v in _inputs(g, dev).items()})
    # mid size (two rays per tile, orbit cameras) vs the oracle
    opt = G.BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 64
    torch.manual_seed(45)
    net2 = G.GazeNeRFNet(opt, include_vd=True, hier_sampling=False)
    ru = G.RenderUtils(45, "cpu", opt)
    shape, appea, gaze = O.synthetic_codes(2)
    cams = [ru.cam_info_list[3], ru.cam_info_list[20]]
    cam = {k: torch.cat([c[k] for c in cams], 0) for k in cams[0]}
    xy = ru.ray_xy.expand(2, -1, -1)
    sd2 = {k: v.clone() for k, v in net2.state_dict().items()}
    for br in ("face", "eyes"):
        k = "fg_CD_predictor_%s." % br
        sd2[k + "density_module.weight"] = sd2[k + "density_module.weight"] * 4.0
        w = sd2[k + "RGB_layer_1.weight"].clone()
        w[:, 384:411] *= 8.0
        sd2[k + "RGB_layer_1.weight"] = w
    net2.load_state_dict(sd2)
    oo = O.OracleOptions(featmap_size=16, featmap_nc=258, pred_img_size=64, num_sample_coarse=64)
    with torch.no_grad():
        ref = O.forward(sd2, oo, "test", xy, shape, appea, gaze, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"], return_stages=True,
                        include_vd=True)
    net2 = net2.to(dev).eval()
    net2.keep_stages = True
    with torch.no_grad():
        out2 = net2("test", batch_xy=xy.to(dev), batch_uv=None, bg_code=None, shape_code=shape.to(dev), appea_code=appea.to(dev), gaze_code=gaze.to(dev),
                    **{k: v.to(dev) for k, v in cam.items()})
    for br in ("face", "eyes"):
        assert rel_l2(net2.last_stages["feat_" + br].cpu().view(2, 258, 16, 16), ref["stages"]["feat_" + br]) < TOL_TIGHT, br
    for k, v in ref["coarse_dict"].items():
        assert float((out2["coarse_dict"][k].cpu() - v).abs().max()) < TOL_TIGHT, k


def test_abi_reentrant_across_streams_and_devices(dev):
    """include/gnrf.h: no device-global tables, no process-wide `static` device state.  Packing + forward issued on two user streams at
    once, and (with >= 2 GPUs) the FIRST use of the library on a second device after the first device has been initialised, must
    reproduce the default-stream result bit for bit."""
    g = load_golden("std_dense_test")
    opt, net = _net_from_golden(g, dev, "tc")
    kw = _inputs(g, dev)
    with torch.no_grad():
        ref = {k: v.clone() for k, v in net("test", **kw)["coarse_dict"].items()}
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for s in streams:                       # both streams re-pack (fresh caches) and run concurrently
        net.invalidate_caches()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            outs.append({k: v.clone() for k, v in net("test", **kw)["coarse_dict"].items()})
    torch.cuda.synchronize()
    for o in outs:
        for k in ref:
            assert torch.equal(o[k], ref[k]), k
    if torch.cuda.device_count() >= 2:
        dev1 = torch.device("cuda:1")
        opt1, net1 = _net_from_golden(g, dev1, "tc")
        with torch.cuda.device(dev1), torch.no_grad():
            out1 = net1("test", **_inputs(g, dev1))["coarse_dict"]
            tr = net1.neural_render   # training forward kernels (conv_tc / wgrad opt-ins) on the second device, too
        for k in ref:
            assert torch.equal(out1[k].cpu(), ref[k].cpu()), k
