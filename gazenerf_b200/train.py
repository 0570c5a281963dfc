"""Training path of the drop-in ``GazeNeRFNet``: forward that keeps activations + hand-written backward, as ONE
``torch.autograd.Function`` over libgnrf's C ABI (include/gnrf.h, "Training path").

The reference trains through torch autograd over its eager graph (trainer/gazenerf_trainer.py:479-528:
``pred = net("train", ...)``, ``loss.backward()``, Adam).  Here the same graph -- rays -> positional encoding -> both
radiance MLPs (models/mlp_nerf.py:95-119) -> alpha composite (utils/model_utils.py:493-534) -> compose
(models/gaze_nerf.py:175-203) -> neural renderer (models/neural_renderer.py:98-113) -- is evaluated layer by layer on
channel-major activations kept in HBM (B200: 180 GB; ~4 GB per face and branch at 64x64x64), every dense layer and both of
its gradients on tcgen05 tensor cores, everything else in streaming CUDA kernels (train_ops.cu, nr_train.cu).  The tape holds
two Functions: FeatureMapFn (rays -> feature maps) and NeuralRenderFn (feature maps -> images).

Per-point activations of the radiance MLPs are stored PRE-SPLIT as bf16 planes (hi | lo, csrc/lin_hl.cu) so that the GEMM kernels
feed them to the tensor cores straight from TMA: ``net.train_precision`` = "bf16x3" (default: hi + lo planes, the bf16x3 scheme of
the inference kernel), "mixed" (forward as "bf16x3" -- images and loss identical to it -- but every GRADIENT tensor stored as one
bf16 plane and both backward GEMMs single-pass on the hi planes), "bf16" (hi plane only everywhere: single-pass bf16, what BASELINE
config[4] names; half the bytes, a third of the MMAs; tolerances of both in tests/test_train_grad.py) or "f32" (r1 path: fp32 activations re-split inside conv_tc.cu /
wgrad_tc.cu; also taken automatically when a shape is outside the plane kernels' tiling: points % 256, hidden width < 256).

Exact rewrites used (identities of the reference graph, as in the inference kernel):
  * the per-face code columns of FeaExt_module_0 / FeaExt_module_5 / RGB_layer_1 are folded into per-face bias vectors.  The
    fold itself (``bias + codes @ W_code.T``, a [B,181]x[181,384] product) is written with differentiable torch ops in
    ``branch_tensors`` so that autograd maps the per-face bias gradients back to the weight columns, the biases and the codes;
  * density_module is evaluated as row 384 of the RGB_layer_0 GEMM; RGB_layer_2 (linear) is applied after compositing.
PyTorch is used for device buffers, the autograd tape around the Function and those O(B x 384) per-face vector products.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

PE = 63
N_BRANCH_T = 21  # effective tensors per branch, see branch_tensors()
# net.train_precision -> (planes of the saved activations, planes of the gradient tensors); 0 = fp32 tensors (r1 kernels)
TRAIN_PRECISIONS = {"bf16x3": (2, 2), "mixed": (2, 1), "bf16": (1, 1), "f32": (0, 0)}


def branch_tensors(mlp, shape_ext: torch.Tensor, appea: torch.Tensor) -> List[torch.Tensor]:
    """Effective (folded / concatenated) fp32 tensors of one radiance MLP, built with differentiable torch ops."""
    H = mlp.h_channel
    W = lambda n: mlp._modules[n].weight.flatten(1)
    b = lambda n: mlp._modules[n].bias
    n_code = shape_ext.shape[1]
    W0, W5, W1r = W("FeaExt_module_0"), W("FeaExt_module_5"), W("RGB_layer_1")
    t = [W0[:, :PE].contiguous(), b("FeaExt_module_0") + shape_ext @ W0[:, PE:].t()]
    for i in (1, 2, 3, 4):
        t += [W("FeaExt_module_%d" % i).contiguous(), b("FeaExt_module_%d" % i)]
    # skip layer input = cat([PE 63 | codes 181 | hidden H]) (models/mlp_nerf.py:106-107); GEMM operand = [hidden H | PE 63]
    t += [torch.cat([W5[:, PE + n_code:], W5[:, :PE]], 1).contiguous(), b("FeaExt_module_5") + shape_ext @ W5[:, PE:PE + n_code].t()]
    for i in (6, 7):
        t += [W("FeaExt_module_%d" % i).contiguous(), b("FeaExt_module_%d" % i)]
    t += [torch.cat([W("RGB_layer_0"), W("density_module")], 0).contiguous(), torch.cat([b("RGB_layer_0"), b("density_module")], 0)]
    t += [W1r[:, :H].contiguous(), b("RGB_layer_1") + appea @ W1r[:, H:].t()]
    t += [torch.cat([W("RGB_layer_2"), b("RGB_layer_2")[:, None]], 1).contiguous()]
    assert len(t) == N_BRANCH_T
    return t


class _Ops:
    """Thin ctypes call helpers (pointer arithmetic on device buffers; every call is enqueued on the current stream)."""

    def __init__(self, device):
        self.L = _lib.lib()
        self.dev = device
        self.st = torch.cuda.current_stream().cuda_stream
        self._ws: Optional[torch.Tensor] = None

    def empty(self, *shape):
        return torch.empty(shape, device=self.dev, dtype=torch.float32)

    def zeros(self, *shape):
        return torch.zeros(shape, device=self.dev, dtype=torch.float32)

    def pack(self, W: torch.Tensor, bias: Optional[torch.Tensor], N: int, K: int, transposed: bool = False) -> torch.Tensor:
        """N, K = GEMM output / input channels.  transposed: W is the forward weight [K][N]."""
        assert W.is_contiguous() and W.dtype == torch.float32 and W.shape == ((K, N) if transposed else (N, K)), (W.shape, N, K)
        pk = torch.empty((self.L.gnrf_conv_tc_packed_bytes(N, K),), device=self.dev, dtype=torch.uint8)
        _lib.check(self.L.gnrf_conv_tc_pack(W.data_ptr(), bias.data_ptr() if bias is not None else None, N, K, 1 if transposed else 0,
                                            pk.data_ptr(), self.st), "gnrf_conv_tc_pack")
        return pk

    def conv(self, pk, N, K, x_ptr, xs, out_ptr, os_, n_img, HW, act=0, bias_img=None, mask_ptr=None, ms=0, mask_rows=0, slope=0.0,
             add_ptr=None, as_=0, add_rows=0):
        _lib.check(self.L.gnrf_conv_tc(pk.data_ptr(), N, K, x_ptr, xs, bias_img.data_ptr() if bias_img is not None else None, out_ptr, os_,
                                       act, mask_ptr, ms, mask_rows, slope, add_ptr, as_, add_rows, n_img, HW, self.st), "gnrf_conv_tc")

    def wgrad(self, dy_ptr, dys, x_ptr, xs, N, K, n_img, HW, db_mode: str):
        """-> dW [N][K], db ([n_img][N] for 'img', [N] for 'sum', None for 'none')."""
        need = self.L.gnrf_wgrad_tc_workspace_bytes(N, K, n_img, HW)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), device=self.dev, dtype=torch.uint8)
        dW = self.empty(N, K)
        db = self.empty(n_img, N) if db_mode == "img" else (self.empty(N) if db_mode == "sum" else None)
        _lib.check(self.L.gnrf_wgrad_tc(dy_ptr, dys, x_ptr, xs, N, K, n_img, HW, dW.data_ptr(), db.data_ptr() if db is not None else None,
                                        1 if db_mode == "sum" else 0, 0, self._ws.data_ptr(), self._ws.numel(), self.st), "gnrf_wgrad_tc")
        return dW, db


RELU, NONE = 1, 0


def _branch_forward(o: _Ops, T: Sequence[torch.Tensor], B, n_r, n_s, H, C, ray_dl, tvecs, z_edges) -> Dict[str, torch.Tensor]:
    """One radiance MLP + composite on channel-major activations; returns everything the backward needs."""
    L, P = o.L, n_r * n_s
    H2, S0 = H // 2, (H + 64) * P
    f4 = 4
    buf0 = o.empty(B, H + 64, P)           # rows [0,H): layer-4 output; rows [H, H+63): positional encoding (shared by layers 0 and 5)
    pe_ptr = buf0.data_ptr() + H * P * f4
    _lib.check(L.gnrf_pe_fwd(ray_dl.data_ptr(), tvecs.data_ptr(), z_edges.data_ptr(), B, n_r, n_s, pe_ptr, S0, o.st), "gnrf_pe_fwd")
    h: List[Optional[torch.Tensor]] = [None] * 8
    h[0] = o.empty(B, H, P)
    o.conv(o.pack(T[0], None, H, PE), H, PE, pe_ptr, S0, h[0].data_ptr(), 0, B, P, act=RELU, bias_img=T[1])
    for i in (1, 2, 3):
        h[i] = o.empty(B, H, P)
        o.conv(o.pack(T[2 * i], T[2 * i + 1], H, H), H, H, h[i - 1].data_ptr(), 0, h[i].data_ptr(), 0, B, P, act=RELU)
    o.conv(o.pack(T[8], T[9], H, H), H, H, h[3].data_ptr(), 0, buf0.data_ptr(), S0, B, P, act=RELU)          # layer 4 -> buf0 rows [0,H)
    h[5] = o.empty(B, H, P)
    o.conv(o.pack(T[10], None, H, H + PE), H, H + PE, buf0.data_ptr(), S0, h[5].data_ptr(), 0, B, P, act=RELU, bias_img=T[11])
    for i, k in ((6, 12), (7, 14)):
        h[i] = o.empty(B, H, P)
        o.conv(o.pack(T[k], T[k + 1], H, H), H, H, h[i - 1].data_ptr(), 0, h[i].data_ptr(), 0, B, P, act=RELU)
    r0 = o.empty(B, H + 1, P)              # rows [0,H): RGB_layer_0 output; row H: raw density
    o.conv(o.pack(T[16], T[17], H + 1, H), H + 1, H, h[7].data_ptr(), 0, r0.data_ptr(), 0, B, P, act=NONE)
    hc = o.empty(B, H2, P)
    o.conv(o.pack(T[18], None, H2, H), H2, H, r0.data_ptr(), (H + 1) * P, hc.data_ptr(), 0, B, P, act=RELU, bias_img=T[19])
    Hc, bg_alpha, w = o.empty(B, H2 + 1, n_r), o.empty(B, n_r), o.empty(B, n_r, n_s)
    sig_ptr = r0.data_ptr() + H * P * f4
    _lib.check(L.gnrf_composite_cm_fwd(hc.data_ptr(), 0 + H2 * P, sig_ptr, (H + 1) * P, z_edges.data_ptr(), ray_dl.data_ptr(), B, n_r, n_s, H2,
                                       Hc.data_ptr(), bg_alpha.data_ptr(), w.data_ptr(), o.st), "gnrf_composite_cm_fwd")
    feat = o.empty(B, C, n_r)
    o.conv(o.pack(T[20], None, C, H2 + 1), C, H2 + 1, Hc.data_ptr(), 0, feat.data_ptr(), 0, B, n_r, act=NONE)
    return {"buf0": buf0, "h": h, "r0": r0, "hc": hc, "Hc": Hc, "bg_alpha": bg_alpha, "w": w, "feat": feat}


def _branch_backward(o: _Ops, T: Sequence[torch.Tensor], sv: Dict[str, torch.Tensor], B, n_r, n_s, H, C, ray_dl, z_edges, g_feat, g_alpha,
                     g_m, g_o, g_z, g_l) -> List[torch.Tensor]:
    """Gradients of the 21 effective tensors; accumulates the per-ray geometry gradients into g_m / g_o / g_z / g_l."""
    L, P = o.L, n_r * n_s
    H2, S0, f4 = H // 2, (H + 64) * P, 4
    buf0, h, r0, hc, Hc, w = sv["buf0"], sv["h"], sv["r0"], sv["hc"], sv["Hc"], sv["w"]
    pe_ptr = buf0.data_ptr() + H * P * f4
    g: List[Optional[torch.Tensor]] = [None] * N_BRANCH_T
    # RGB_layer_2 (after the composite): feat = [W2 | b2] [Hc ; sum w]
    g[20], _ = o.wgrad(g_feat.data_ptr(), 0, Hc.data_ptr(), 0, C, H2 + 1, B, n_r, "none")
    g_Hc = o.empty(B, H2 + 1, n_r)
    o.conv(o.pack(T[20], None, H2 + 1, C, transposed=True), H2 + 1, C, g_feat.data_ptr(), 0, g_Hc.data_ptr(), 0, B, n_r)
    g_r0, g_hc = o.empty(B, H + 1, P), o.empty(B, H2, P)
    _lib.check(L.gnrf_composite_cm_bwd(g_Hc.data_ptr(), g_alpha.data_ptr(), hc.data_ptr(), H2 * P, r0.data_ptr() + H * P * f4, (H + 1) * P,
                                       w.data_ptr(), z_edges.data_ptr(), ray_dl.data_ptr(), B, n_r, n_s, H2, g_hc.data_ptr(), H2 * P,
                                       g_r0.data_ptr() + H * P * f4, (H + 1) * P, g_z.data_ptr(), g_l.data_ptr(), o.st), "gnrf_composite_cm_bwd")
    # RGB_layer_1 (hidden part; the appearance columns live in the per-face bias)
    g[18], g[19] = o.wgrad(g_hc.data_ptr(), 0, r0.data_ptr(), (H + 1) * P, H2, H, B, P, "img")
    o.conv(o.pack(T[18], None, H, H2, transposed=True), H, H2, g_hc.data_ptr(), 0, g_r0.data_ptr(), (H + 1) * P, B, P)
    del g_hc
    # RGB_layer_0 + density_module on h7
    g[16], g[17] = o.wgrad(g_r0.data_ptr(), 0, h[7].data_ptr(), 0, H + 1, H, B, P, "sum")
    ga = o.empty(B, H, P)
    o.conv(o.pack(T[16], None, H, H + 1, transposed=True), H, H + 1, g_r0.data_ptr(), 0, ga.data_ptr(), 0, B, P, mask_ptr=h[7].data_ptr())
    del g_r0
    for i, k in ((7, 14), (6, 12)):
        g[k], g[k + 1] = o.wgrad(ga.data_ptr(), 0, h[i - 1].data_ptr(), 0, H, H, B, P, "sum")
        gb = o.empty(B, H, P)
        o.conv(o.pack(T[k], None, H, H, transposed=True), H, H, ga.data_ptr(), 0, gb.data_ptr(), 0, B, P, mask_ptr=h[i - 1].data_ptr())
        ga = gb
    # skip layer 5: operand [h4 | PE] in buf0
    g[10], g[11] = o.wgrad(ga.data_ptr(), 0, buf0.data_ptr(), S0, H, H + PE, B, P, "img")
    g_buf0 = o.empty(B, H + 64, P)
    o.conv(o.pack(T[10], None, H + PE, H, transposed=True), H + PE, H, ga.data_ptr(), 0, g_buf0.data_ptr(), S0, B, P,
           mask_ptr=buf0.data_ptr(), ms=S0, mask_rows=H)
    # layer 4 (its output gradient = rows [0,H) of g_buf0)
    g[8], g[9] = o.wgrad(g_buf0.data_ptr(), S0, h[3].data_ptr(), 0, H, H, B, P, "sum")
    ga = o.empty(B, H, P)
    o.conv(o.pack(T[8], None, H, H, transposed=True), H, H, g_buf0.data_ptr(), S0, ga.data_ptr(), 0, B, P, mask_ptr=h[3].data_ptr())
    for i in (3, 2, 1):
        g[2 * i], g[2 * i + 1] = o.wgrad(ga.data_ptr(), 0, h[i - 1].data_ptr(), 0, H, H, B, P, "sum")
        gb = o.empty(B, H, P)
        o.conv(o.pack(T[2 * i], None, H, H, transposed=True), H, H, ga.data_ptr(), 0, gb.data_ptr(), 0, B, P, mask_ptr=h[i - 1].data_ptr())
        ga = gb
    # layer 0 on the positional encoding
    g[0], g[1] = o.wgrad(ga.data_ptr(), 0, pe_ptr, S0, H, PE, B, P, "img")
    g_pe = o.empty(B, PE, P)
    o.conv(o.pack(T[0], None, PE, H, transposed=True), PE, H, ga.data_ptr(), 0, g_pe.data_ptr(), 0, B, P)
    _lib.check(L.gnrf_pe_bwd(g_pe.data_ptr(), PE * P, g_buf0.data_ptr() + H * P * f4, S0, pe_ptr, S0, ray_dl.data_ptr(), z_edges.data_ptr(),
                             B, n_r, n_s, g_m.data_ptr(), g_o.data_ptr(), g_z.data_ptr(), o.st), "gnrf_pe_bwd")
    return g  # type: ignore[return-value]


class _HL:
    """Call helpers of the plane kernels (csrc/lin_hl.cu).  A plane tensor is a bf16 torch tensor [PL][B][rows][P]."""

    def __init__(self, o: _Ops, planes: int, B: int, P: int):
        self.o, self.L, self.PL, self.B, self.P = o, o.L, planes, B, P
        self._ws: Optional[torch.Tensor] = None

    def empty(self, rows: int) -> torch.Tensor:
        return torch.empty((self.PL, self.B, rows, self.P), device=self.o.dev, dtype=torch.bfloat16)

    def pack(self, W: torch.Tensor, bias: Optional[torch.Tensor], N: int, K: int, transposed: bool = False) -> torch.Tensor:
        assert W.is_contiguous() and W.dtype == torch.float32 and W.shape == ((K, N) if transposed else (N, K)), (W.shape, N, K)
        pk = torch.empty((self.L.gnrf_lin_hl_packed_bytes(N, K, self.PL),), device=self.o.dev, dtype=torch.uint8)
        _lib.check(self.L.gnrf_lin_hl_pack(W.data_ptr(), bias.data_ptr() if bias is not None else None, N, K, 1 if transposed else 0,
                                           self.PL, pk.data_ptr(), self.o.st), "gnrf_lin_hl_pack")
        return pk

    def bits(self, rows: int) -> torch.Tensor:
        """ReLU sign bits of a [rows][P] activation: one bit per element (written by lin(sign_out=...), read by lin(mask=...))."""
        return torch.empty((self.B, rows, self.P // 32), device=self.o.dev, dtype=torch.int32)

    def lin(self, pk, N, K, x: torch.Tensor, x_row0: int = 0, out: Optional[torch.Tensor] = None, out_row0: int = 0, hl_rows: Optional[int] = None,
            out_f32: Optional[torch.Tensor] = None, act=0, bias_img=None, mask: Optional[torch.Tensor] = None, mask_rows: int = 0,
            sign_out: Optional[torch.Tensor] = None, act_rows: int = 0):
        """x / out: plane tensors, addressed from row x_row0 / out_row0 (row windows of a wider tensor keep its strides);
        mask / sign_out: bit tensors from bits()."""
        P = self.P
        hl_rows = N if hl_rows is None else hl_rows
        _lib.check(self.L.gnrf_lin_hl(
            pk.data_ptr(), N, K, self.PL, x.data_ptr() + x_row0 * P * 2, x.shape[2] * P, x.stride(0), bias_img.data_ptr() if bias_img is not None else None,
            act, act_rows, (out.data_ptr() + out_row0 * P * 2) if out is not None else None, out.shape[2] * P if out is not None else 0,
            out.stride(0) if out is not None else 0, hl_rows, out_f32.data_ptr() if out_f32 is not None else None,
            out_f32.shape[1] * P if out_f32 is not None else 0, mask.data_ptr() if mask is not None else None,
            mask.shape[1] * (P // 32) if mask is not None else 0, mask_rows,
            sign_out.data_ptr() if sign_out is not None else None, sign_out.shape[1] * (P // 32) if sign_out is not None else 0,
            sign_out.shape[1] if sign_out is not None else 0, self.B, P, self.o.st), "gnrf_lin_hl")

    def wgrad(self, dy: torch.Tensor, N: int, x: torch.Tensor, K: int, db_mode: str, x_row0: int = 0):
        """-> dW [N][K], db ([B][N] for 'img', [N] for 'sum')."""
        P, o = self.P, self.o
        need = self.L.gnrf_wgrad_hl_workspace_bytes(N, K, self.B, P)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((need,), device=o.dev, dtype=torch.uint8)
        dW = o.empty(N, K)
        db = o.empty(self.B, N) if db_mode == "img" else o.empty(N)
        _lib.check(self.L.gnrf_wgrad_hl(dy.data_ptr(), dy.shape[2] * P, dy.stride(0), x.data_ptr() + x_row0 * P * 2, x.shape[2] * P, x.stride(0),
                                        self.PL, N, K, self.B, P, dW.data_ptr(), db.data_ptr(), 1 if db_mode == "sum" else 0,
                                        self._ws.data_ptr(), self._ws.numel(), o.st), "gnrf_wgrad_hl")
        return dW, db


def hl_supported(H: int, P: int) -> bool:
    return P % 256 == 0 and H >= 256 and H % 2 == 0


def _branch_forward_hl(o: _Ops, T: Sequence[torch.Tensor], B, n_r, n_s, H, C, ray_dl, tvecs, z_edges, planes: int) -> Dict[str, torch.Tensor]:
    """_branch_forward on bf16 plane tensors (no conversion pass in any GEMM); fp32 only where a non-GEMM kernel consumes it."""
    L, P = o.L, n_r * n_s
    H2 = H // 2
    q = _HL(o, planes, B, P)
    buf0 = q.empty(H + 64)                 # rows [0,H): layer-4 output; rows [H, H+63): positional encoding (layers 0 and 5)
    pe32 = o.empty(B, PE, P)               # fp32 copy for the encoding's backward
    _lib.check(L.gnrf_pe_fwd_hl(ray_dl.data_ptr(), tvecs.data_ptr(), z_edges.data_ptr(), B, n_r, n_s, pe32.data_ptr(), PE * P,
                                buf0.data_ptr() + H * P * 2, (H + 64) * P, buf0.stride(0), planes, o.st), "gnrf_pe_fwd_hl")
    h: List[Optional[torch.Tensor]] = [None] * 8
    sg = [q.bits(H) for _ in range(8)]     # ReLU sign bits of the eight hidden activations (1 bit per element) for the backward
    h[0] = q.empty(H)
    q.lin(q.pack(T[0], None, H, PE), H, PE, buf0, x_row0=H, out=h[0], act=RELU, bias_img=T[1], sign_out=sg[0])
    for i in (1, 2, 3):
        h[i] = q.empty(H)
        q.lin(q.pack(T[2 * i], T[2 * i + 1], H, H), H, H, h[i - 1], out=h[i], act=RELU, sign_out=sg[i])
    q.lin(q.pack(T[8], T[9], H, H), H, H, h[3], out=buf0, act=RELU, sign_out=sg[4])                   # layer 4 -> buf0 rows [0,H)
    h[5] = q.empty(H)
    q.lin(q.pack(T[10], None, H, H + PE), H, H + PE, buf0, out=h[5], act=RELU, bias_img=T[11], sign_out=sg[5])
    for i, k in ((6, 12), (7, 14)):
        h[i] = q.empty(H)
        q.lin(q.pack(T[k], T[k + 1], H, H), H, H, h[i - 1], out=h[i], act=RELU, sign_out=sg[i])
    # RGB_layer_0 has no activation (models/mlp_nerf.py:110-111), so RGB_layer_1(RGB_layer_0(h7)) is ONE linear map of h7 -- the same
    # exact fold the inference kernel uses; the density row rides along:
    #   [hc_pre ; sigma_raw] = [W1h W0 ; wd] h7 + [b1_face + W1h b0 ; bd]      (193 rows; ReLU on the first 192 only)
    # One 2-M-tile GEMM replaces a 385-row and a 192-row one in the forward and in both backward products; the 384-row RGB_layer_0
    # output is never materialised.  The gradients of the four original tensors follow from the product rule in _branch_backward_hl.
    W0, b0, wd, bd, W1h = T[16][:H], T[17][:H], T[16][H:], T[17][H:], T[18]
    W_eff = torch.cat([(W1h.double() @ W0.double()).float(), wd], 0).contiguous()                     # [H2 + 1][H]
    b_eff = torch.cat([T[19] + (W1h @ b0)[None], bd[None].expand(B, 1)], 1).contiguous()              # [B][H2 + 1] per-face bias
    hs = o.empty(B, H2 + 1, P)             # rows [0,H2): RGB_layer_1 output (post-ReLU); row H2: raw density -- fp32 for the composite
    q.lin(q.pack(W_eff, None, H2 + 1, H), H2 + 1, H, h[7], hl_rows=0, out_f32=hs, act=RELU, act_rows=H2, bias_img=b_eff)
    Hc, bg_alpha, w = o.empty(B, H2 + 1, n_r), o.empty(B, n_r), o.empty(B, n_r, n_s)
    _lib.check(L.gnrf_composite_cm_fwd(hs.data_ptr(), (H2 + 1) * P, hs.data_ptr() + H2 * P * 4, (H2 + 1) * P, z_edges.data_ptr(),
                                       ray_dl.data_ptr(), B, n_r, n_s, H2, Hc.data_ptr(), bg_alpha.data_ptr(), w.data_ptr(), o.st),
               "gnrf_composite_cm_fwd")
    feat = o.empty(B, C, n_r)
    o.conv(o.pack(T[20], None, C, H2 + 1), C, H2 + 1, Hc.data_ptr(), 0, feat.data_ptr(), 0, B, n_r, act=NONE)
    return {"buf0": buf0, "pe32": pe32, "h": h, "sg": sg, "hs": hs, "W_eff": W_eff, "Hc": Hc, "bg_alpha": bg_alpha, "w": w,
            "feat": feat, "planes": planes, "planes_bwd": planes}


def _branch_backward_hl(o: _Ops, T: Sequence[torch.Tensor], sv: Dict[str, torch.Tensor], B, n_r, n_s, H, C, ray_dl, z_edges, g_feat, g_alpha,
                        g_m, g_o, g_z, g_l) -> List[torch.Tensor]:
    """_branch_backward on plane tensors: every output / input gradient that feeds a GEMM is written as planes by its producer.
    sv["planes_bwd"] = 1 with 2-plane saved activations is the "mixed" mode: gradients are stored as one bf16 plane and both
    backward GEMMs run single-pass on the hi planes (the saved activations' hi plane is plane 0 of the same tensor)."""
    L, P = o.L, n_r * n_s
    H2 = H // 2
    planes = sv["planes_bwd"]
    q = _HL(o, planes, B, P)
    buf0, h, hs, Hc, w, sg, W_eff = sv["buf0"], sv["h"], sv["hs"], sv["Hc"], sv["w"], sv["sg"], sv["W_eff"]
    g: List[Optional[torch.Tensor]] = [None] * N_BRANCH_T
    g[20], _ = o.wgrad(g_feat.data_ptr(), 0, Hc.data_ptr(), 0, C, H2 + 1, B, n_r, "none")
    g_Hc = o.empty(B, H2 + 1, n_r)
    o.conv(o.pack(T[20], None, H2 + 1, C, transposed=True), H2 + 1, C, g_feat.data_ptr(), 0, g_Hc.data_ptr(), 0, B, n_r)
    g_hs = q.empty(H2 + 1)                         # rows [0,H2): gradient of the RGB_layer_1 pre-activation; row H2: of the raw density
    _lib.check(L.gnrf_composite_cm_bwd_hl(g_Hc.data_ptr(), g_alpha.data_ptr(), hs.data_ptr(), (H2 + 1) * P, hs.data_ptr() + H2 * P * 4,
                                          (H2 + 1) * P, w.data_ptr(), z_edges.data_ptr(), ray_dl.data_ptr(), B, n_r, n_s, H2,
                                          g_hs.data_ptr(), (H2 + 1) * P, g_hs.stride(0), g_hs.data_ptr() + H2 * P * 2, (H2 + 1) * P,
                                          g_hs.stride(0), planes, g_z.data_ptr(), g_l.data_ptr(), o.st), "gnrf_composite_cm_bwd_hl")
    # folded RGB head on h7: dW_eff [H2+1][H], per-face bias gradient [B][H2+1]; product rule back to the four original tensors
    gW_eff, gb_eff = q.wgrad(g_hs, H2 + 1, h[7], H, "img")
    W0, b0, W1h = T[16][:H], T[17][:H], T[18]
    gWp, gbp = gW_eff[:H2], gb_eff[:, :H2]                                                            # the RGB_layer_1 part
    g[16] = torch.cat([W1h.t() @ gWp, gW_eff[H2:]], 0)                                              # dW0 = W1h^T dW' ; density row
    g[17] = torch.cat([W1h.t() @ gbp.sum(0), gb_eff[:, H2].sum(0, keepdim=True)], 0)                # db0 = W1h^T sum_b db' ; density bias
    g[18] = gWp @ W0.t() + torch.outer(gbp.sum(0), b0)                                              # dW1h = dW' W0^T + db' b0^T
    g[19] = gbp.contiguous()                                                                        # per-face bias of RGB_layer_1
    ga = q.empty(H)
    q.lin(q.pack(W_eff, None, H, H2 + 1, transposed=True), H, H2 + 1, g_hs, out=ga, mask=sg[7])
    del g_hs
    for i, k in ((7, 14), (6, 12)):
        g[k], g[k + 1] = q.wgrad(ga, H, h[i - 1], H, "sum")
        gb = q.empty(H)
        q.lin(q.pack(T[k], None, H, H, transposed=True), H, H, ga, out=gb, mask=sg[i - 1])
        ga = gb
    # skip layer 5: operand [h4 | PE] = rows [0, H+63) of buf0; its input gradient: rows < H (h4, masked) as planes, PE rows fp32
    g[10], g[11] = q.wgrad(ga, H, buf0, H + PE, "img")
    g4, g_pe_b = q.empty(H), o.empty(B, PE, P)
    q.lin(q.pack(T[10], None, H + PE, H, transposed=True), H + PE, H, ga, out=g4, hl_rows=H, out_f32=g_pe_b, mask=sg[4], mask_rows=H)
    # layer 4
    g[8], g[9] = q.wgrad(g4, H, h[3], H, "sum")
    ga = q.empty(H)
    q.lin(q.pack(T[8], None, H, H, transposed=True), H, H, g4, out=ga, mask=sg[3])
    del g4
    for i in (3, 2, 1):
        g[2 * i], g[2 * i + 1] = q.wgrad(ga, H, h[i - 1], H, "sum")
        gb = q.empty(H)
        q.lin(q.pack(T[2 * i], None, H, H, transposed=True), H, H, ga, out=gb, mask=sg[i - 1])
        ga = gb
    # layer 0 on the positional encoding
    g[0], g[1] = q.wgrad(ga, H, buf0, PE, "img", x_row0=H)
    g_pe_a = o.empty(B, PE, P)
    q.lin(q.pack(T[0], None, PE, H, transposed=True), PE, H, ga, hl_rows=0, out_f32=g_pe_a)
    _lib.check(L.gnrf_pe_bwd(g_pe_a.data_ptr(), PE * P, g_pe_b.data_ptr(), PE * P, sv["pe32"].data_ptr(), PE * P, ray_dl.data_ptr(),
                             z_edges.data_ptr(), B, n_r, n_s, g_m.data_ptr(), g_o.data_ptr(), g_z.data_ptr(), o.st), "gnrf_pe_bwd")
    return g  # type: ignore[return-value]


class FeatureMapFn(torch.autograd.Function):
    """rays -> positional encoding -> both radiance MLPs -> composite -> compose:  feature maps [3B+1, C, S, S] =
    (merge_face | eyes_planes | merge | bg_featmap), with gradients to every argument tensor."""

    @staticmethod
    def forward(ctx, cfg, xy, rmats, tvecs, kinv, gaze, z_edges, bg_featmap, *tensors):
        H, C, S, n_s = cfg["H"], cfg["C"], cfg["S"], cfg["n_s"]
        B, n_r = xy.shape[0], xy.shape[2]
        o = _Ops(xy.device)
        L = o.L
        T = [t.detach().contiguous() for t in tensors]
        Tf, Te = T[:N_BRANCH_T], T[N_BRANCH_T:2 * N_BRANCH_T]
        ray_dl = o.empty(B, n_r, 4)
        _lib.check(L.gnrf_ray_setup(xy.data_ptr(), rmats.data_ptr(), kinv.data_ptr(), B, n_r, ray_dl.data_ptr(), o.st), "gnrf_ray_setup")
        planes, planes_bwd = TRAIN_PRECISIONS[cfg.get("precision", "bf16x3")]
        if planes and not hl_supported(H, n_r * n_s):
            planes = 0
        if planes:
            sv_f = _branch_forward_hl(o, Tf, B, n_r, n_s, H, C, ray_dl, tvecs, z_edges, planes)
            sv_e = _branch_forward_hl(o, Te, B, n_r, n_s, H, C, ray_dl, tvecs, z_edges, planes)
            sv_f["planes_bwd"] = sv_e["planes_bwd"] = planes_bwd
        else:
            sv_f = _branch_forward(o, Tf, B, n_r, n_s, H, C, ray_dl, tvecs, z_edges)
            sv_e = _branch_forward(o, Te, B, n_r, n_s, H, C, ray_dl, tvecs, z_edges)
        n_img = 3 * B + 1
        fm = o.empty(n_img, C, S, S)
        bg = bg_featmap.detach().contiguous()
        _lib.check(L.gnrf_compose_fwd(sv_f["feat"].data_ptr(), sv_f["bg_alpha"].data_ptr(), sv_e["feat"].data_ptr(), sv_e["bg_alpha"].data_ptr(),
                                      bg.data_ptr(), gaze.data_ptr(), B, C, S * S, fm.data_ptr(), o.st), "gnrf_compose_fwd")
        fm[3 * B].copy_(bg[0])
        ctx.cfg = cfg
        ctx.keep = dict(xy=xy, rmats=rmats, kinv=kinv, gaze=gaze, z_edges=z_edges, bg=bg, ray_dl=ray_dl, T=T, sv_f=sv_f, sv_e=sv_e)
        if cfg.get("stages") is not None:
            cfg["stages"].update({"feat_face": sv_f["feat"], "feat_eyes": sv_e["feat"], "bg_alpha_face": sv_f["bg_alpha"],
                                  "bg_alpha_eyes": sv_e["bg_alpha"], "w_face": sv_f["w"], "w_eyes": sv_e["w"], "featmaps": fm})
        return fm

    @staticmethod
    def backward(ctx, g_fm):
        cfg, k = ctx.cfg, ctx.keep
        H, C, S, n_s = cfg["H"], cfg["C"], cfg["S"], cfg["n_s"]
        xy, ray_dl, z_edges, T = k["xy"], k["ray_dl"], k["z_edges"], k["T"]
        B, n_r = xy.shape[0], xy.shape[2]
        o = _Ops(xy.device)
        L = o.L
        Tf, Te = T[:N_BRANCH_T], T[N_BRANCH_T:2 * N_BRANCH_T]
        g_fm = g_fm.contiguous().float()
        # ---- compose
        P2 = S * S
        sv_f, sv_e = k["sv_f"], k["sv_e"]
        g_ff, g_fe = o.empty(B, C, P2), o.empty(B, C, P2)
        n_grp = L.gnrf_compose_bwd_groups(C)               # channel groups: partial sums, added up below in a fixed order
        g_af, g_ae = o.empty(n_grp, B, P2), o.empty(n_grp, B, P2)
        g_bg = o.empty(1, C, S, S)
        nblk = L.gnrf_compose_bwd_blocks(P2, C)
        g_gz = o.empty(B, nblk, 2)
        _lib.check(L.gnrf_compose_bwd(g_fm.data_ptr(), sv_f["feat"].data_ptr(), sv_f["bg_alpha"].data_ptr(), sv_e["feat"].data_ptr(),
                                      sv_e["bg_alpha"].data_ptr(), k["bg"].data_ptr(), k["gaze"].data_ptr(), B, C, P2, g_ff.data_ptr(),
                                      g_af.data_ptr(), g_fe.data_ptr(), g_ae.data_ptr(), g_bg.data_ptr(), g_gz.data_ptr(), o.st),
                   "gnrf_compose_bwd")
        g_bg = g_bg + g_fm[3 * B:3 * B + 1]
        g_gaze = g_gz.sum(1)
        g_af, g_ae = g_af.sum(0), g_ae.sum(0)
        # ---- radiance MLPs + composite + positional encoding
        g_m, g_o, g_l = o.zeros(B, n_r, 3), o.zeros(B, n_r, 3), o.zeros(B, n_r)
        g_z = o.zeros(B, n_r, n_s + 1)
        bwd = _branch_backward_hl if "planes" in sv_f else _branch_backward
        g_f = bwd(o, Tf, sv_f, B, n_r, n_s, H, C, ray_dl, z_edges, g_ff, g_af, g_m, g_o, g_z, g_l)
        g_e = bwd(o, Te, sv_e, B, n_r, n_s, H, C, ray_dl, z_edges, g_fe, g_ae, g_m, g_o, g_z, g_l)
        # ---- geometry
        contrib = o.empty(B, n_r, 12)
        _lib.check(L.gnrf_geom_bwd(xy.data_ptr(), k["rmats"].data_ptr(), k["kinv"].data_ptr(), g_m.data_ptr(), g_o.data_ptr(), g_l.data_ptr(),
                                   g_z.data_ptr(), B, n_r, n_s, contrib.data_ptr(), o.st), "gnrf_geom_bwd")
        cs = contrib.sum(1)
        g_R, g_T = cs[:, :9].reshape(B, 3, 3), cs[:, 9:].reshape(B, 3)
        ctx.keep = None
        grads = [None, None, g_R, g_T, None, g_gaze, None, g_bg] + g_f + g_e
        need = ctx.needs_input_grad
        return tuple(gr if (i < len(need) and need[i]) else None for i, gr in enumerate(grads))


class NeuralRenderFn(torch.autograd.Function):
    """feature maps [N, C, S, S] -> images [N, 3, P, P] (models/neural_renderer.py:98-113) with gradients to the maps and the
    renderer's parameters."""

    @staticmethod
    def forward(ctx, cfg, fm, *tensors):
        net, C, S = cfg["net"], cfg["C"], cfg["S"]
        nr = net.neural_render
        o = _Ops(fm.device)
        L = o.L
        Tn = [t.detach().contiguous() for t in tensors]
        fm = fm.detach().contiguous()
        n_img = fm.shape[0]
        Pimg = S << nr.n_blocks
        imgs = o.empty(n_img, 3, Pimg, Pimg)
        saved_bytes = L.gnrf_nr_train_saved_bytes(n_img, C, S, nr.n_blocks, nr.min_feat)
        saved = torch.empty((saved_bytes,), device=fm.device, dtype=torch.uint8)
        packed = nr.packed_tc()
        _lib.check(L.gnrf_nr_train_fwd(_lib.ptr_array([t.data_ptr() for t in Tn]), packed.data_ptr(), fm.data_ptr(), n_img, C, S, nr.n_blocks,
                                       nr.min_feat, imgs.data_ptr(), saved.data_ptr(), saved_bytes, o.st), "gnrf_nr_train_fwd")
        ctx.cfg = cfg
        ctx.keep = dict(Tn=Tn, fm=fm, imgs=imgs, saved=saved)
        return imgs

    @staticmethod
    def backward(ctx, g_imgs):
        cfg, k = ctx.cfg, ctx.keep
        net, C, S = cfg["net"], cfg["C"], cfg["S"]
        nr = net.neural_render
        Tn, fm = k["Tn"], k["fm"]
        o = _Ops(fm.device)
        L = o.L
        n_img = fm.shape[0]
        g_imgs = g_imgs.contiguous().float()
        g_fm = o.empty(n_img, C, S, S)
        g_n = [torch.empty_like(t) for t in Tn]
        ws_bytes = L.gnrf_nr_train_bwd_workspace_bytes(n_img, C, S, nr.n_blocks, nr.min_feat)
        ws = torch.empty((ws_bytes,), device=fm.device, dtype=torch.uint8)
        _lib.check(L.gnrf_nr_train_bwd(_lib.ptr_array([t.data_ptr() for t in Tn]), fm.data_ptr(), k["saved"].data_ptr(),
                                       k["imgs"].data_ptr(), g_imgs.data_ptr(), n_img, C, S, nr.n_blocks, nr.min_feat, g_fm.data_ptr(),
                                       _lib.ptr_array([t.data_ptr() for t in g_n]), ws.data_ptr(), ws_bytes, o.st), "gnrf_nr_train_bwd")
        ctx.keep = None
        grads = [None, g_fm] + g_n
        need = ctx.needs_input_grad
        return tuple(gr if (i < len(need) and need[i]) else None for i, gr in enumerate(grads))


def render_featmaps(net, xy, rmats, tvecs, kinv, gaze, shape_ext, appea, z_edges, stages=None) -> torch.Tensor:
    """Differentiable rays -> feature maps [3B+1, C, S, S] (merge_face | eyes_planes | merge | bg_featmap)."""
    precision = getattr(net, "train_precision", "bf16x3")
    if precision not in TRAIN_PRECISIONS:
        raise ValueError("train_precision must be one of %s, got %r" % (sorted(TRAIN_PRECISIONS), precision))
    tf = branch_tensors(net.fg_CD_predictor_face, shape_ext, appea)
    te = branch_tensors(net.fg_CD_predictor_eyes, shape_ext, appea)
    cfg = {"net": net, "H": net.mlp_h_channel, "C": net.featmap_nc, "S": net.featmap_size, "n_s": net.num_sample_coarse, "stages": stages,
           "precision": precision}
    return FeatureMapFn.apply(cfg, xy, rmats, tvecs, kinv, gaze, z_edges, net.neural_render.bg_featmap, *tf, *te)


def neural_render_train(net, fm: torch.Tensor) -> torch.Tensor:
    """Differentiable feature maps -> images through the neural renderer."""
    tn = [p if p.dim() == 1 else p.flatten(1) for p in net.neural_render.param_list()]
    cfg = {"net": net, "C": net.featmap_nc, "S": net.featmap_size}
    return NeuralRenderFn.apply(cfg, fm, *tn)


def forward_train(net, xy, rmats, tvecs, kinv, gaze, shape_ext, appea, z_edges, stages=None) -> torch.Tensor:
    """Differentiable render: returns images [3B+1,3,P,P].  ``shape_ext`` / ``appea`` / ``gaze`` / ``rmats`` / ``tvecs`` may require
    grad (the reference optimises code offsets and camera deltas, trainer/gazenerf_trainer.py:338-405).  Two chained
    ``autograd.Function``s: rays -> feature maps, feature maps -> images."""
    fm = render_featmaps(net, xy, rmats, tvecs, kinv, gaze, shape_ext, appea, z_edges, stages=stages)
    return neural_render_train(net, fm)
