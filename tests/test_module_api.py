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


def test_abi_argument_checks_fail_cleanly_without_a_gpu():
    """Entry points validate their arguments before touching the device: status code + gnrf_last_error(), never an abort."""
    L = G.lib()
    assert L.gnrf_conv_tc(None, 0, 0, None, 0, None, None, 0, 0, None, 0, 0, 0.0, None, 0, 0, 1, 1, None) == 1   # GNRF_ERR_ARG
    assert b"gnrf_conv_tc" in L.gnrf_last_error()
    assert L.gnrf_wgrad_tc(None, 0, None, 0, 4, 4, 1, 64, None, None, 0, 0, None, 0, None) == 1
    assert L.gnrf_compose_bwd(*([None] * 7), 1, 3, 4, *([None] * 6), None) == 1
    assert L.gnrf_neural_render_tc_fwd_gather(None, None, None, 4, 258, 8, 3, 32, None, None, 0, None, None, 2, 0, 1, 2, None) == 1
    assert L.gnrf_data_loss_fwd(*([None] * 9), 1, 64, 1, 1.0, None, None, None, None) == 1
    assert L.gnrf_conv_tc_packed_bytes(384, 384) > 2 * 384 * 384 * 2 and L.gnrf_conv_tc_packed_bytes(0, 4) == 0
    assert L.gnrf_wgrad_tc_workspace_bytes(384, 384, 2, 262144) >= 128 * 208 * 4
    assert L.gnrf_nr_train_saved_bytes(7, 258, 64, 3, 32) > 0 and L.gnrf_nr_train_bwd_workspace_bytes(7, 258, 64, 3, 32) > 0
    # plane kernels of the training path (csrc/lin_hl.cu)
    assert L.gnrf_lin_hl(None, 384, 384, 2, None, 0, 0, None, 1, 0, None, 0, 0, 384, None, 0, None, 0, 0, None, 0, 0, 2, 512, None) == 1
    assert b"gnrf_lin_hl" in L.gnrf_last_error()
    assert L.gnrf_lin_hl_pack(None, None, 384, 384, 0, 2, None, None) == 1
    assert L.gnrf_wgrad_hl(None, 0, 0, None, 0, 0, 2, 384, 384, 2, 512, None, None, 1, None, 0, None) == 1
    assert L.gnrf_pe_fwd_hl(None, None, None, 1, 64, 8, None, 0, None, 0, 0, 2, None) == 1
    assert L.gnrf_composite_cm_bwd_hl(None, None, None, 0, None, 0, None, None, None, 1, 64, 8, 192, None, 0, 0, None, 0, 0, 2, None, None, None) == 1
    assert L.gnrf_lin_hl_packed_bytes(384, 384, 2) >= 2 * 384 * 384 * 2 + 384 * 4 and L.gnrf_lin_hl_packed_bytes(384, 384, 1) < L.gnrf_lin_hl_packed_bytes(384, 384, 2)
    assert L.gnrf_lin_hl_packed_bytes(384, 384, 3) == 0 and L.gnrf_wgrad_hl_workspace_bytes(384, 384, 2, 262144) >= 128 * 400 * 4
    assert L.gnrf_compose_bwd_groups(258) == 8 and L.gnrf_compose_bwd_blocks(4096, 258) == 32 * 8


def test_training_path_effective_tensors_reproduce_the_reference_layers():
    """gazenerf_b200.train.branch_tensors (host logic, differentiable torch ops): the folded per-face biases and the re-arranged
    skip-layer / head matrices must give the same pre-activations as the reference layers on their concatenated inputs."""
    from gazenerf_b200.train import branch_tensors, PE
    opt = G.BaseOptions({"featmap_size": 8, "featmap_nc": 258, "pred_img_size": 64})
    torch.manual_seed(3)
    net = G.GazeNeRFNet(opt, include_vd=False, hier_sampling=False)
    mlp = net.fg_CD_predictor_face
    B, n = 2, 5
    g = torch.Generator().manual_seed(1)
    shape_ext, appea = torch.randn(B, 181, generator=g), torch.randn(B, 127, generator=g)
    pe, h4, h7 = torch.randn(B, n, PE, generator=g), torch.randn(B, n, 384, generator=g), torch.randn(B, n, 384, generator=g)
    T = branch_tensors(mlp, shape_ext, appea)
    W = lambda name: mlp._modules[name].weight.flatten(1)
    b = lambda name: mlp._modules[name].bias
    code = shape_ext[:, None, :].expand(B, n, 181)
    # layer 0 on cat([pe, codes])  (models/gaze_nerf.py:250-262)
    ref0 = torch.cat([pe, code], -1) @ W("FeaExt_module_0").t() + b("FeaExt_module_0")
    assert torch.allclose(pe @ T[0].t() + T[1][:, None, :], ref0, atol=1e-5)
    # skip layer on cat([pe, codes, h4])  (models/mlp_nerf.py:106-107); operand order of the GEMM is [h4 | pe]
    ref5 = torch.cat([pe, code, h4], -1) @ W("FeaExt_module_5").t() + b("FeaExt_module_5")
    assert torch.allclose(torch.cat([h4, pe], -1) @ T[10].t() + T[11][:, None, :], ref5, atol=1e-5)
    # RGB_layer_0 + density share the input; RGB_layer_1 on cat([rgb0, appea]); RGB_layer_2 with its bias as an extra column
    r0 = h7 @ T[16].t() + T[17]
    assert torch.allclose(r0[..., :384], h7 @ W("RGB_layer_0").t() + b("RGB_layer_0"), atol=1e-5)
    assert torch.allclose(r0[..., 384], (h7 @ W("density_module").t() + b("density_module"))[..., 0], atol=1e-5)
    ref1 = torch.cat([r0[..., :384], appea[:, None, :].expand(B, n, 127)], -1) @ W("RGB_layer_1").t() + b("RGB_layer_1")
    assert torch.allclose(r0[..., :384] @ T[18].t() + T[19][:, None, :], ref1, atol=1e-5)
    hc = torch.relu(ref1)
    ref2 = hc @ W("RGB_layer_2").t() + b("RGB_layer_2")
    assert torch.allclose(torch.cat([hc, torch.ones(B, n, 1)], -1) @ T[20].t(), ref2, atol=1e-5)
    # autograd maps a per-face bias gradient back to codes and weight columns
    (T[1].sum() + T[19].sum()).backward()
    assert mlp.FeaExt_module_0.weight.grad[:, PE:].abs().sum() > 0 and mlp.RGB_layer_1.weight.grad[:, 384:].abs().sum() > 0


def test_train_precision_is_validated_before_any_device_work():
    """net.train_precision (gazenerf_b200/train.py) selects the activation storage of the differentiable path; an unknown mode must
    fail loudly in Python, not fall through to some default."""
    import pytest
    import torch
    from gazenerf_b200 import train

    class _Net:
        train_precision = "fp8"

    with pytest.raises(ValueError, match="train_precision"):
        train.render_featmaps(_Net(), *([None] * 9))
    assert set(train.TRAIN_PRECISIONS) == {"bf16x3", "mixed", "bf16", "f32"}
