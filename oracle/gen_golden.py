"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference needs ``kornia.filters.filter2d`` (kornia==0.6.4, not installed, not vendored); a shim with the
documented semantics is installed into ``sys.modules`` before the import.  Nothing from the
reference is copied: we import it, call its modules on seeded inputs and store inputs/outputs as fixtures.

Fixtures
  tiny.npz      hidden=32, featmap_nc=48, 8x8 rays, 8 samples, 64x64 image, B=2; ALL weights stored, so the
                oracle pin does not depend on RNG replication.
  std_*.npz     hidden=384, featmap_nc=258 (the real layer shapes), 8x8 rays, 8 samples, 64x64 image, B=2,
                reference init under torch.manual_seed(45) (train.py:53); weights are NOT stored (20 MB) --
                only per-parameter checksums; tests rebuild them with the drop-in module's identical init.
                variants: ref-init "test", dense-density "test", dense-density "train" (jitter).
  fine_*        FineSample on the coarse weights of the above (config-3 building block), int64 indices included.
  mid_dense_test.npz  hidden=384, featmap_nc=258, 16x16 rays x 64 samples (2 rays per 128-point tile of the fused kernel), 128x128
                image, B=2, dense density, FineSample with num_sample_fine=64 (the config-3 sample counts); slim (no per-point arrays).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = os.environ.get("GNRF_REFERENCE", "/root/reference")


def install_kornia_shim():
    k = types.ModuleType("kornia")
    kf = types.ModuleType("kornia.filters")

    def filter2d(x, kernel, border_type="reflect", normalized=False, padding="same"):
        b, c, h, w = x.shape
        tmp = kernel.unsqueeze(1).to(x)
        if normalized:
            tmp = tmp / tmp.abs().sum(dim=(-2, -1), keepdim=True)
        kh, kw = kernel.shape[-2:]
        tmp = tmp.expand(-1, c, -1, -1).reshape(-1, 1, kh, kw)
        xp = F.pad(x, [kw // 2, kw // 2, kh // 2, kh // 2], mode=border_type)
        return F.conv2d(xp, tmp, groups=c)

    kf.filter2d = filter2d
    k.filters = kf
    sys.modules["kornia"] = k
    sys.modules["kornia.filters"] = kf


def np32(t):
    return t.detach().cpu().numpy()



def make_dense(net, xy, shape, appea, gaze, cam, n_s, scale=30.0):
    """Non-vacuous density (SURVEY §8c caveat 1): scale the density head and centre it on the median raw density of
    these inputs, so about half of the sample points are opaque in BOTH branches.  Returns the two biases."""
    raw = {}
    hooks = []
    for name, br in (("face", net.fg_CD_predictor_face), ("eyes", net.fg_CD_predictor_eyes)):
        hooks.append(br.density_module.register_forward_hook(lambda m, i, o, name=name: raw.__setitem__(name, o.detach().clone())))
    with torch.no_grad():
        smp = net.sample_func(xy, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"], False)
        pe = net.vp_encoder_face(smp["pts"])
        n_r = xy.shape[2]
        ext = torch.cat([shape, gaze], 1)[:, :, None, None].expand(-1, -1, n_r, n_s)
        app = appea[:, :, None, None].expand(-1, -1, n_r, n_s)
        vp = torch.cat([pe, ext], 1)
        net.fg_CD_predictor_face(vp, app)
        net.fg_CD_predictor_eyes(vp, app)
    for h in hooks:
        h.remove()
    biases = []
    with torch.no_grad():
        for name, br in (("face", net.fg_CD_predictor_face), ("eyes", net.fg_CD_predictor_eyes)):
            b0 = float(br.density_module.bias[0])
            med = float(raw[name].median()) - b0
            br.density_module.weight *= scale
            br.density_module.bias.fill_(-scale * med)
            biases.append(-scale * med)
    return np.array(biases, dtype=np.float64)


def run_case(net, opt, mode, xy, uv, shape, appea, gaze, cam, dense, seed_rng=None, n_fine=8, slim=False):
    from utils.model_utils import FineSample  # reference

    b = xy.shape[0]
    n_r = xy.shape[2]
    n_s = opt.num_sample_coarse
    out = {}
    with torch.no_grad():
        if seed_rng is not None:
            torch.manual_seed(seed_rng)
        full = net(mode, xy, uv, None, shape, appea, gaze, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"])
        for k, v in full["coarse_dict"].items():
            out["img_" + k] = np32(v)
        # stage by stage through the reference's own modules (same RNG draw for train mode)
        if seed_rng is not None:
            torch.manual_seed(seed_rng)
            # rand_like(zvals[B,N_r,N_s+1]) is the first draw inside sample_func (utils/model_utils.py:306)
            out["jitter_u"] = np32(torch.rand(b, n_r, n_s + 1))
            torch.manual_seed(seed_rng)
        smp = net.sample_func(xy, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"], mode == "train")
        for k in ("pts", "zvals", "z_dists", "batch_ray_o", "batch_ray_d", "batch_ray_l"):
            out["smp_" + k] = np32(smp[k])
        pe = net.vp_encoder_face(smp["pts"])
        out["pe"] = np32(pe)
        shape_ext = torch.cat([shape, gaze], 1)
        ext = shape_ext[:, :, None, None].expand(-1, -1, n_r, n_s)
        app = appea[:, :, None, None].expand(-1, -1, n_r, n_s)
        vp = torch.cat([pe, ext], 1)
        w_face = None
        for br, mlp in (("face", net.fg_CD_predictor_face), ("eyes", net.fg_CD_predictor_eyes)):
            rgb, sig = mlp(vp, app)
            fr, ba, dep, w = net.calc_color_func(smp["pts"], rgb, sig, smp["z_dists"], smp["zvals"])
            out["mlp_feat_" + br] = np32(rgb[:, :, :: max(1, n_r // 16), :])  # subsample rays
            out["mlp_sigma_" + br] = np32(sig)
            out["feat_" + br] = np32(fr)
            out["bg_alpha_" + br] = np32(ba)
            out["depth_" + br] = np32(dep)
            out["w_" + br] = np32(w)
            if br == "face":
                w_face = w
        # FineSample (the only working part of the hier path), deterministic u
        fopt = types.SimpleNamespace(num_sample_fine=n_fine)
        fs = FineSample(fopt)(w_face, smp, False)
        for k in ("pts", "zvals", "z_dists"):
            out["fine_" + k] = np32(fs[k])
        # int64 indices, recomputed with the reference's exact expression sequence (utils/model_utils.py:417-445)
        tw = w_face[:, :, :, 1:-1].reshape(-1, n_s - 2)
        pdf = tw / torch.sum(tw + 1e-5, dim=-1, keepdim=True)
        cdf = F.pad(torch.cumsum(pdf, dim=-1), pad=[1, 0, 0, 0], mode="constant", value=0.0).contiguous()
        u = torch.linspace(0.0, 1.0, steps=n_fine + 1).view(1, n_fine + 1).expand(cdf.size(0), n_fine + 1).contiguous()
        out["fine_inds"] = torch.searchsorted(cdf, u, right=True).numpy().astype(np.int64)
    if slim:  # mid-size fixture: keep what the parity tests read, drop the bulky per-point arrays
        for k in ("smp_pts", "pe", "mlp_feat_face", "mlp_feat_eyes", "mlp_sigma_face", "mlp_sigma_eyes", "fine_pts", "fine_z_dists",
                  "depth_face", "depth_eyes"):
            out.pop(k, None)
    out["in_xy"] = np32(xy)
    out["in_shape"] = np32(shape)
    out["in_appea"] = np32(appea)
    out["in_gaze"] = np32(gaze)
    out["in_R"] = np32(cam["batch_Rmats"])
    out["in_T"] = np32(cam["batch_Tvecs"])
    out["in_Kinv"] = np32(cam["batch_inv_inmats"])
    out["meta"] = np.array([opt.featmap_size, opt.featmap_nc, opt.pred_img_size, opt.num_sample_coarse,
                            opt.mlp_hidden_nchannels, int(dense), int(mode == "train")], dtype=np.int64)
    return out


def main():
    only = set(sys.argv[1:])   # e.g. `python oracle/gen_golden.py mid` regenerates just that fixture family (tiny | std | mid)
    want = lambda name: not only or name in only
    install_kornia_shim()
    sys.path.insert(0, REF)
    os.chdir(REF)  # RenderUtils opens configs/... relative to cwd (utils/render_utils.py:36)
    from configs.gazenerf_options import BaseOptions
    from models.gaze_nerf import GazeNeRFNet
    from utils.render_utils import RenderUtils

    os.makedirs(OUT, exist_ok=True)

    def inputs(opt, b):
        ru = RenderUtils(45, "cpu", opt)
        g = torch.Generator().manual_seed(0)
        shape = torch.randn(b, 179, generator=g) * 0.3
        appea = torch.randn(b, 127, generator=g) * 0.3
        gaze = torch.rand(b, 2, generator=g) - 0.5
        # item 0: base camera, item 1: orbit camera #7 (non-trivial rotation)
        cams = [ru.base_cam_info, ru.cam_info_list[7]]
        cam = {k: torch.cat([cams[i % 2][k] for i in range(b)], 0) for k in cams[0]}
        return ru, ru.ray_xy.expand(b, -1, -1), ru.ray_uv.expand(b, -1, -1), shape, appea, gaze, cam

    # ---------------- tiny: all weights stored ----------------
    if want("tiny"):
        _gen_tiny(BaseOptions, GazeNeRFNet, inputs)
    if want("std"):
        _gen_std(BaseOptions, GazeNeRFNet, inputs)
    if want("mid"):
        _gen_mid(BaseOptions, GazeNeRFNet, inputs)
    if want("vd"):
        _gen_vd(BaseOptions, GazeNeRFNet, inputs)


def _gen_tiny(BaseOptions, GazeNeRFNet, inputs):
    opt = BaseOptions({"featmap_size": 8, "featmap_nc": 48, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    opt.mlp_hidden_nchannels = 32
    torch.manual_seed(7)
    net = GazeNeRFNet(opt, include_vd=False, hier_sampling=False).eval()
    with torch.no_grad():
        net.neural_render.bg_featmap.mul_(0.5).add_(0.1 * torch.randn_like(net.neural_render.bg_featmap))
    ru, xy, uv, shape, appea, gaze, cam = inputs(opt, 2)
    dense_bias = make_dense(net, xy, shape, appea, gaze, cam, opt.num_sample_coarse)
    out = run_case(net, opt, "test", xy, uv, shape, appea, gaze, cam, dense=True)
    out["dense_bias"] = dense_bias
    for k, v in net.state_dict().items():
        out["sd/" + k] = np32(v)
    out["ru_inv_inmat"] = np32(ru.inv_inmat)
    out["ru_orbit7_R"] = np32(ru.cam_info_list[7]["batch_Rmats"])
    out["ru_orbit7_T"] = np32(ru.cam_info_list[7]["batch_Tvecs"])
    out["ru_uv"] = np32(ru.ray_uv)
    np.savez_compressed(os.path.join(OUT, "tiny.npz"), **out)
    print("tiny.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")



def _gen_std(BaseOptions, GazeNeRFNet, inputs):
    # ---------------- std: real layer widths, weights by seed ----------------
    opt = BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    for name, dense, mode, seed_rng in (("std_refinit_test", False, "test", None),
                                        ("std_dense_test", True, "test", None),
                                        ("std_dense_train", True, "train", 123)):
        torch.manual_seed(45)
        net = GazeNeRFNet(opt, include_vd=False, hier_sampling=False).eval()
        chk = {"chk/" + k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in net.state_dict().items()}
        ru, xy, uv, shape, appea, gaze, cam = inputs(opt, 2)
        dense_bias = make_dense(net, xy, shape, appea, gaze, cam, opt.num_sample_coarse) if dense else np.zeros(2)
        out = run_case(net, opt, mode, xy, uv, shape, appea, gaze, cam, dense=dense, seed_rng=seed_rng)
        out.pop("pe")  # large and already pinned by tiny.npz
        out["dense_bias"] = dense_bias
        out.update(chk)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, sum(v.nbytes for v in out.values()) // 1024, "KiB raw")



def _gen_mid(BaseOptions, GazeNeRFNet, inputs):
    # ---------------- mid: 16x16 rays x 64 samples, reference-anchored multi-ray-per-tile case ----------------
    opt = BaseOptions({"featmap_size": 16, "featmap_nc": 258, "pred_img_size": 128})
    opt.num_sample_coarse = 64
    torch.manual_seed(45)
    net = GazeNeRFNet(opt, include_vd=False, hier_sampling=False).eval()
    chk = {"chk/" + k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in net.state_dict().items()}
    ru, xy, uv, shape, appea, gaze, cam = inputs(opt, 2)
    dense_bias = make_dense(net, xy, shape, appea, gaze, cam, opt.num_sample_coarse, scale=4.0)
    out = run_case(net, opt, "test", xy, uv, shape, appea, gaze, cam, dense=True, n_fine=64, slim=True)
    out["dense_bias"] = dense_bias
    out["dense_scale"] = np.array([4.0])
    out.update(chk)
    np.savez_compressed(os.path.join(OUT, "mid_dense_test.npz"), **out)
    print("mid_dense_test", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


def _gen_vd(BaseOptions, GazeNeRFNet, inputs):
    # ---------------- include_vd=True (view-direction input, models/gaze_nerf.py:70-80): real layer widths, 8x8 rays x 8 samples ----------------
    opt = BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    torch.manual_seed(45)
    net = GazeNeRFNet(opt, include_vd=True, hier_sampling=False).eval()
    chk = {"chk/" + k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in net.state_dict().items()}
    shapes = {"shape/" + k: np.array(v.shape, dtype=np.int64) for k, v in net.state_dict().items()}
    ru, xy, uv, shape, appea, gaze, cam = inputs(opt, 2)
    # dense density (same recipe as make_dense, with the view-direction input wired in)
    raw, hooks = {}, []
    for name, br in (("face", net.fg_CD_predictor_face), ("eyes", net.fg_CD_predictor_eyes)):
        hooks.append(br.density_module.register_forward_hook(lambda m, i, o, name=name: raw.__setitem__(name, o.detach().clone())))
    with torch.no_grad():
        net("test", xy, uv, None, shape, appea, gaze, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"])
    for h in hooks:
        h.remove()
    biases = []
    with torch.no_grad():
        for name, br in (("face", net.fg_CD_predictor_face), ("eyes", net.fg_CD_predictor_eyes)):
            med = float(raw[name].median()) - float(br.density_module.bias[0])
            br.density_module.weight *= 30.0
            br.density_module.bias.fill_(-30.0 * med)
            biases.append(-30.0 * med)
        # the default init of RGB_layer_1 gives the 27 view-direction columns little weight: scale them so that the test is sensitive to them
        for br in (net.fg_CD_predictor_face, net.fg_CD_predictor_eyes):
            br.RGB_layer_1.weight[:, 384:384 + 27] *= 8.0
    out = {}
    with torch.no_grad():
        full = net("test", xy, uv, None, shape, appea, gaze, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"])
        for k, v in full["coarse_dict"].items():
            out["img_" + k] = np32(v)
        smp = net.sample_func(xy, cam["batch_Rmats"], cam["batch_Tvecs"], cam["batch_inv_inmats"], False)
        pe = net.vp_encoder_face(smp["pts"])
        vd = net.vd_encoder(smp["dirs"])
        out["vd_pe"] = np32(vd[:, :, :, 0])
        n_r, n_s = xy.shape[2], opt.num_sample_coarse
        ext = torch.cat([shape, gaze], 1)[:, :, None, None].expand(-1, -1, n_r, n_s)
        app = torch.cat([vd, appea[:, :, None, None].expand(-1, -1, n_r, n_s)], 1)
        vp = torch.cat([pe, ext], 1)
        for br, mlp in (("face", net.fg_CD_predictor_face), ("eyes", net.fg_CD_predictor_eyes)):
            rgb, sig = mlp(vp, app)
            fr, ba, dep, w = net.calc_color_func(smp["pts"], rgb, sig, smp["z_dists"], smp["zvals"])
            out["feat_" + br], out["bg_alpha_" + br], out["w_" + br] = np32(fr), np32(ba), np32(w)
    out["in_xy"], out["in_shape"], out["in_appea"], out["in_gaze"] = np32(xy), np32(shape), np32(appea), np32(gaze)
    out["in_R"], out["in_T"], out["in_Kinv"] = np32(cam["batch_Rmats"]), np32(cam["batch_Tvecs"]), np32(cam["batch_inv_inmats"])
    out["meta"] = np.array([8, 258, 64, 8, opt.mlp_hidden_nchannels, 1, 0], dtype=np.int64)
    out["dense_bias"] = np.array(biases, dtype=np.float64)
    out["vd_col_scale"] = np.array([8.0])
    out.update(chk)
    out.update(shapes)
    np.savez_compressed(os.path.join(OUT, "std_dense_vd_test.npz"), **out)
    print("std_dense_vd_test", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    main()
