"""CPU-side checks of the drop-in boundary: parameter names/shapes/order, reference-identical seeded init, oracle on the
real layer widths (std_*.npz), and that libgnrf.so loads and exports every symbol include/gnrf.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, max_rel, rel_l2
from oracle import gazenerf_oracle as O

import gazenerf_b200 as G


def _std_net(dense, g=None):
    opt = G.BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    torch.manual_seed(45)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return opt, net, (O.densify(sd, *g["dense_bias"]) if dense else sd)


def test_state_dict_matches_reference_names_and_init():
    g = load_golden("std_refinit_test")
    opt, net, sd = _std_net(False)
    ref_keys = [k[4:] for k in g if k.startswith("chk/")]
    assert list(net.state_dict().keys()) == ref_keys  # names AND order (optimizer state / strict load)
    for k in ref_keys:
        v = sd[k].double()
        chk = g["chk/" + k]
        assert abs(float(v.sum()) - chk[0]) <= 1e-9 * max(1.0, abs(chk[0])), k
        assert abs(float(v.abs().sum()) - chk[1]) <= 1e-9 * max(1.0, abs(chk[1])), k
    # parameter order handed to Adam (trainer/gazenerf_trainer.py:464)
    names = [n for n, _ in net.named_parameters()]
    assert names[0] == "fg_CD_predictor_eyes.FeaExt_module_0.weight" and names[48] == "neural_render.bg_featmap"
    assert sum(p.numel() for p in net.parameters()) == 2 * 1518979 + (1977756 - 258 * 64 * 64 + 258 * 8 * 8)


def test_default_sizes_param_count():
    torch.manual_seed(0)
    net = G.GazeNeRFNet(G.BaseOptions(), include_vd=False, hier_sampling=False)
    assert sum(p.numel() for p in net.parameters()) == 5015714
    assert tuple(net.fg_CD_predictor_face.FeaExt_module_5.weight.shape) == (384, 628, 1, 1)
    assert tuple(net.neural_render.bg_featmap.shape) == (1, 258, 64, 64)


@pytest.mark.parametrize("name,dense", [("std_refinit_test", False), ("std_dense_test", True), ("std_dense_train", True)])
def test_oracle_on_real_layer_widths(name, dense):
    g = load_golden(name)
    opt, net, sd = _std_net(dense, g)
    oo = O.OracleOptions(featmap_size=8, featmap_nc=258, pred_img_size=64, num_sample_coarse=8)
    t = lambda k: torch.from_numpy(g[k])
    train = bool(g["meta"][6])
    out = O.forward(sd, oo, "train" if train else "test", t("in_xy"), t("in_shape"), t("in_appea"), t("in_gaze"), t("in_R"), t("in_T"),
                    t("in_Kinv"), jitter_u=t("jitter_u") if train else None, return_stages=True)
    st = out["stages"]
    assert rel_l2(st["feat_face"].flatten(2), g["feat_face"]) < 5e-6
    assert rel_l2(st["feat_eyes"].flatten(2), g["feat_eyes"]) < 5e-6
    assert max_rel(st["bg_alpha_face"].flatten(2), g["bg_alpha_face"]) < 2e-5
    for k in ("merge_img_face", "merge_img_eyes", "merge_img", "bg_img"):
        assert float((out["coarse_dict"][k] - t("img_" + k)).abs().max()) < 5e-6, k
    if dense:
        assert g["bg_alpha_face"].min() < 0.2 and g["bg_alpha_eyes"].min() < 0.2  # non-vacuous in both branches
        assert (g["mlp_sigma_face"] == 0).mean() > 0.2 and (g["mlp_sigma_face"] > 0).mean() > 0.2


def test_render_utils_mirror(tiny_golden):
    g = tiny_golden
    opt = G.BaseOptions({"featmap_size": 8, "featmap_nc": 48, "pred_img_size": 64})
    ru = G.RenderUtils(45, "cpu", opt)
    assert np.array_equal(ru.ray_xy.numpy(), g["in_xy"][:1])
    assert np.array_equal(ru.ray_uv.numpy(), g["ru_uv"])
    assert np.array_equal(ru.inv_inmat.numpy(), g["ru_inv_inmat"])
    assert np.array_equal(ru.cam_info_list[7]["batch_Rmats"].numpy(), g["ru_orbit7_R"])
    assert np.array_equal(ru.cam_info_list[7]["batch_Tvecs"].numpy(), g["ru_orbit7_T"])
    assert np.array_equal(ru.base_cam_info["batch_Rmats"].numpy(), g["in_R"][:1])
    assert len(ru.cam_info_list) == 45


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gnrf.h")).read()
    declared = sorted(set(re.findall(r"\b(gnrf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 17
    path = G.build()
    L = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(L, name), "libgnrf.so does not export %s" % name
    L.gnrf_abi_version.restype = ctypes.c_int
    assert L.gnrf_abi_version() == 1
    assert sorted(G._lib.SYMBOLS) == declared


def test_product_never_imports_oracle():
    """The product package must not import, link or execute anything under oracle/ (comments may mention it)."""
    pat_py = re.compile(r"^\s*(import|from)\s+oracle\b|oracle[/.]gazenerf_oracle|importlib.*oracle", re.M)
    pat_c = re.compile(r"#include\s*[<\"].*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gazenerf_b200")):
        for f in files:
            src = open(os.path.join(dirpath, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert not pat_py.search(src), f
            elif src:
                assert not pat_c.search(src), f


def test_cpu_tensors_fail_loudly():
    opt = G.BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    opt.num_sample_coarse = 8
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    ru = G.RenderUtils(45, "cpu", opt)
    s, a, gz = O.synthetic_codes(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        net("test", ru.ray_xy, ru.ray_uv, None, s, a, gz, **ru.base_cam_info)
