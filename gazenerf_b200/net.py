"""Drop-in ``GazeNeRFNet``: the reference's module API on top of libgnrf's sm_100a kernels.

Mirrors models/gaze_nerf.py:13-351 of the reference: same constructor ``GazeNeRFNet(opt, include_vd, hier_sampling)``,
same ``forward(mode, batch_xy, batch_uv, bg_code, shape_code, appea_code, gaze_code, batch_Rmats, batch_Tvecs,
batch_inv_inmats, dist_expr=False, **kwargs)``, same output dict ``{"coarse_dict": {merge_img_face, merge_img_eyes,
merge_img, bg_img}}``, and the same ``state_dict()`` keys / shapes / registration order (checkpoints are loaded with
strict ``load_state_dict``, trainer/gazenerf_trainer.py:116).  Sub-modules hold parameters only -- all arithmetic runs in
libgnrf (include/gnrf.h); there is no eager-PyTorch or CPU implementation of the path in this package.

Parameters are created in the reference's construction order with the reference's initialisers
(models/gaze_nerf.py:49-119, models/mlp_nerf.py:29-93, models/neural_renderer.py:35-96), so the same
``torch.manual_seed`` yields bit-identical initial weights.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .options import BaseOptions

PE_DIMS = 63
TC_TILE = 128


def _conv(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0)


class RadianceMLP(nn.Module):
    """Parameter container with the layout of MLPforNeRF (models/mlp_nerf.py:13-93)."""

    LAYER_NAMES = ["FeaExt_module_%d" % i for i in range(8)] + ["density_module", "RGB_layer_0", "RGB_layer_1", "RGB_layer_2"]

    def __init__(self, vp_channels: int, vd_channels: int, n_layers: int = 8, h_channel: int = 256, res_nfeat: int = 3):
        super().__init__()
        assert n_layers == 8, "libgnrf is specialised for the reference's 8-layer trunk"
        self.vp_channels, self.vd_channels, self.n_layers = vp_channels, vd_channels, n_layers
        self.h_channel, self.res_nfeat = h_channel, res_nfeat
        self.skips = [n_layers // 2]
        self.add_module("FeaExt_module_0", _conv(vp_channels, h_channel))  # default init (mlp_nerf.py:30-35)
        for i in range(n_layers - 1):
            cin = h_channel + vp_channels if i in self.skips else h_channel
            layer = _conv(cin, h_channel)
            self.add_module("FeaExt_module_%d" % (i + 1), layer)
            nn.init.xavier_uniform_(layer.weight.data)  # mlp_nerf.py:61
        self.add_module("density_module", _conv(h_channel, 1))
        nn.init.xavier_uniform_(self.density_module.weight.data)
        self.density_module.bias.data[:] = 0.0  # mlp_nerf.py:66-67
        self.add_module("RGB_layer_0", _conv(h_channel, h_channel))
        nn.init.xavier_uniform_(self.RGB_layer_0.weight.data)
        self.add_module("RGB_layer_1", _conv(h_channel + vd_channels, h_channel // 2))
        self.add_module("RGB_layer_2", _conv(h_channel // 2, res_nfeat))

    def param_list(self) -> List[torch.Tensor]:
        out = []
        for n in self.LAYER_NAMES:
            m = self._modules[n]
            out += [m.weight, m.bias]
        return out

    def forward(self, *a, **k):  # pragma: no cover - parameters only
        raise RuntimeError("RadianceMLP holds parameters only; evaluation happens inside libgnrf (GazeNeRFNet.forward)")


class _Blur(nn.Module):
    """Holds the [1,2,1] buffer so state_dict keys match (pixel_shuffle_upsample.py:7-11); the filter is baked in the kernel."""

    def __init__(self):
        super().__init__()
        self.register_buffer("f", torch.Tensor([1, 2, 1]))


class _PixelShuffleUpsampleParams(nn.Module):
    def __init__(self, in_feature: int):
        super().__init__()
        self.in_feature = in_feature
        self.layer_1 = _conv(in_feature, in_feature * 2)
        self.layer_2 = _conv(in_feature * 2, in_feature * 4)
        self.blur_layer = _Blur()


class NeuralRendererParams(nn.Module):
    """Parameter container with the layout of NeuralRenderer (models/neural_renderer.py:12-96)."""

    def __init__(self, bg_type="white", feat_nc=256, out_dim=3, final_actvn=True, min_feat=32, featmap_size=32, img_size=256):
        super().__init__()
        assert out_dim == 3 and final_actvn
        self.bg_type, self.featmap_size, self.n_feat, self.min_feat = bg_type, featmap_size, feat_nc, min_feat
        self.n_blocks = int(math.log2(img_size) - math.log2(featmap_size))
        # "tc": tcgen05 bf16x3, fused level kernels where the shape allows (default) | "tc_layerwise": one tcgen05 launch per 1x1 conv |
        # "simt": fp32 CUDA-core convs
        self.impl = "tc"
        w = lambda i: max(feat_nc // (2 ** i), min_feat)
        self.feat_upsample_list = nn.ModuleList([_PixelShuffleUpsampleParams(w(i)) for i in range(self.n_blocks)])
        self.rgb_upsample = nn.Sequential(nn.Identity(), _Blur())  # index 1 carries the "f" buffer
        self.feat_2_rgb_list = nn.ModuleList([_conv(feat_nc, out_dim)] + [_conv(w(i + 1), out_dim) for i in range(self.n_blocks)])
        self.feat_layers = nn.ModuleList([_conv(w(i), w(i + 1)) for i in range(self.n_blocks)])
        if bg_type == "white":
            bg = torch.ones((1, feat_nc, featmap_size, featmap_size), dtype=torch.float32)
        elif bg_type == "black":
            bg = torch.zeros((1, feat_nc, featmap_size, featmap_size), dtype=torch.float32)
        else:
            raise ValueError("Error bg_type")
        self.register_parameter("bg_featmap", nn.Parameter(bg))

    def get_bg_featmap(self):
        return self.bg_featmap

    def param_list(self) -> List[torch.Tensor]:
        out = []
        for m in self.feat_upsample_list:
            out += [m.layer_1.weight, m.layer_1.bias, m.layer_2.weight, m.layer_2.bias]
        for m in self.feat_2_rgb_list:
            out += [m.weight, m.bias]
        for m in self.feat_layers:
            out += [m.weight, m.bias]
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N,C,S,S] -> [N,3,P,P] through libgnrf (same call shape as the reference's NeuralRenderer.forward)."""
        return neural_render(self, x)

    def packed_tc(self) -> torch.Tensor:
        """bf16 hi/lo weight streams for the tcgen05 convs; a derived cache keyed on parameter versions, never saved."""
        L = _lib.lib()
        params = [p.detach() for p in self.param_list()]
        key = tuple((p.data_ptr(), p._version) for p in params)
        hit = getattr(self, "_tc_pack", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        nbytes = L.gnrf_nr_tc_packed_bytes(self.n_feat, self.n_blocks, self.min_feat)
        packed = torch.empty((nbytes,), device=params[0].device, dtype=torch.uint8)
        _lib.check(L.gnrf_nr_tc_pack(_ptrs(params), self.n_feat, self.n_blocks, self.min_feat, packed.data_ptr(), _stream()), "gnrf_nr_tc_pack")
        object.__setattr__(self, "_tc_pack", (key, packed))
        return packed


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (libgnrf has no CPU path); got device %s" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptrs(tensors: List[torch.Tensor]):
    for t in tensors:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    return _lib.ptr_array([t.data_ptr() for t in tensors])


def neural_render(nr: NeuralRendererParams, x: torch.Tensor, gather=None) -> torch.Tensor:
    """``gather``: a dist.PeerAllGather -- the first 3 * b_local images are also stored into every rank's gathered buffer by the
    kernel that produces them (fused all-gather over NVLink / NVSwitch multicast)."""
    L = _lib.lib()
    x = _dev_f32(x, "featmap")
    n, c, s, _ = x.shape
    P = s << nr.n_blocks
    img = torch.empty((n, 3, P, P), device=x.device, dtype=torch.float32)
    ws_bytes = L.gnrf_nr_workspace_bytes(n, c, s, nr.n_blocks, nr.min_feat)
    ws = torch.empty((ws_bytes,), device=x.device, dtype=torch.uint8)
    params = [p.detach() for p in nr.param_list()]
    if gather is not None:
        if nr.impl != "tc":
            raise RuntimeError("the fused all-gather is part of the tcgen05 neural-render path (impl='tc')")
        peers, mc = gather.current()
        assert n >= 3 * gather.b_local and P == gather.P
        packed = nr.packed_tc()
        _lib.check(L.gnrf_neural_render_tc_fwd_gather(_ptrs(params), packed.data_ptr(), x.data_ptr(), n, c, s, nr.n_blocks, nr.min_feat,
                                                      img.data_ptr(), ws.data_ptr(), ws_bytes, _lib.ptr_array(peers), mc or None, gather.world,
                                                      gather.rank, gather.b_local, gather.gb, _stream()), "gnrf_neural_render_tc_fwd_gather")
    elif nr.impl in ("tc", "tc_layerwise"):
        packed = nr.packed_tc()
        fn = L.gnrf_neural_render_tc_fwd if nr.impl == "tc" else L.gnrf_neural_render_tc_layerwise_fwd
        _lib.check(fn(_ptrs(params), packed.data_ptr(), x.data_ptr(), n, c, s, nr.n_blocks, nr.min_feat,
                      img.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "gnrf_neural_render_tc_fwd")
    else:
        _lib.check(L.gnrf_neural_render_fwd(_ptrs(params), x.data_ptr(), n, c, s, nr.n_blocks, nr.min_feat, img.data_ptr(),
                                            ws.data_ptr(), ws_bytes, _stream()), "gnrf_neural_render_fwd")
    return img


class GraphedForward(object):
    """``net("test", ...)`` for fixed shapes captured ONCE into a CUDA graph and replayed (inference serving: the kernel launches,
    ctypes calls and tensor allocations of a forward cost more host time than the small kernels at the head of the step take on the
    GPU).  Inputs are copied into static device buffers, the returned dict holds static output tensors that the next call overwrites.
    The packed weight buffers are baked into the graph: re-capture after changing parameters.

    With ``gather`` (a dist.PeerAllGather) the fused all-gather is part of the capture: one graph per rotating symmetric buffer (its
    peer / multicast addresses are launch parameters), and a call replays the graph of the buffer the current step writes; the caller
    then runs ``gather.finish()`` as usual."""

    def __init__(self, net: "GazeNeRFNet", mode: str, kwargs: Dict[str, Optional[torch.Tensor]], gather=None):
        if mode != "test":
            raise ValueError("only the deterministic 'test' forward is captured (train mode draws fresh jitter every call)")
        if net.keep_stages or (gather is None and net.gather_ctx is not None):
            raise RuntimeError("graph capture: no stage dumps; pass the PeerAllGather as `gather=` instead of setting net.gather_ctx")
        if gather is not None and (net.hier_sampling or kwargs.get("only_merge")):
            raise RuntimeError("the fused gather lives in the coarse, all-three-images neural-render call")
        self.net, self.mode, self.gather = net, mode, gather
        # every float32 input lives in ONE flat static device buffer (views below), so a caller that keeps its host inputs in the same
        # layout (pack_inputs) feeds a whole step with a single H2D copy instead of one small copy per tensor
        self.input_layout, off = {}, 0
        for k, v in kwargs.items():
            if torch.is_tensor(v):
                self.input_layout[k] = (off, tuple(v.shape))
                off += (v.numel() + 3) // 4 * 4          # 16-byte aligned slots
        dev = next(v.device for v in kwargs.values() if torch.is_tensor(v))
        self.flat_input = torch.zeros((off,), device=dev, dtype=torch.float32)
        self.static_in = {}
        for k, v in kwargs.items():
            if torch.is_tensor(v):
                o, shp = self.input_layout[k]
                view = self.flat_input[o:o + v.numel()].view(shp)
                view.copy_(v.detach().float())
                self.static_in[k] = view
            else:
                self.static_in[k] = v
        L = _lib.lib()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):   # lazy initialisation (function attributes, packed weights, bg_img cache, t-value tables) outside the capture
                net(mode, **self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self._versions = tuple((p.data_ptr(), p._version) for p in net.parameters())
        self.graphs, self.static_outs = [], []
        n_graphs = gather.n_buf if gather is not None else 1
        for b in range(n_graphs):
            g = torch.cuda.CUDAGraph()
            n0 = L.gnrf_launch_count()
            if gather is not None:
                gather._forced = b
                net.gather_ctx = gather
            try:
                with torch.no_grad(), torch.cuda.graph(g):
                    out = net(mode, **self.static_in)
            finally:
                if gather is not None:
                    gather._forced = None
                    net.gather_ctx = None
            self.launches_per_replay = int(L.gnrf_launch_count() - n0)
            self.graphs.append(g)
            self.static_outs.append(out)
        self.graph, self.static_out = self.graphs[0], self.static_outs[0]

    def __call__(self, **kwargs) -> Dict[str, Dict[str, torch.Tensor]]:
        if tuple((p.data_ptr(), p._version) for p in self.net.parameters()) != self._versions:
            raise RuntimeError("parameters changed since the graph was captured: call net.graphed(...) again")
        for k, v in kwargs.items():
            if torch.is_tensor(v):
                self.static_in[k].copy_(v, non_blocking=True)
        return self.replay()

    def replay(self) -> Dict[str, Dict[str, torch.Tensor]]:
        """Replay on whatever the static input buffers hold (see ``flat_input`` / ``pack_inputs`` / ``static_in``)."""
        b = (self.gather.step % self.gather.n_buf) if self.gather is not None else 0
        self.graphs[b].replay()
        return self.static_outs[b]

    def pack_inputs(self, kwargs: Dict[str, Optional[torch.Tensor]], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host tensors -> one pinned flat float32 buffer in the layout of ``flat_input``; then per step
        ``gf.flat_input.copy_(packed, non_blocking=True); gf.replay()`` is ONE host-to-device copy + one graph launch."""
        if out is None:
            out = torch.zeros((self.flat_input.numel(),), dtype=torch.float32).pin_memory()
        for k, (o, shp) in self.input_layout.items():
            v = kwargs[k]
            out[o:o + v.numel()].view(shp).copy_(v.detach().float())
        return out


class GazeNeRFNet(nn.Module):
    def __init__(self, opt: BaseOptions, include_vd, hier_sampling, mlp_impl: Optional[str] = None) -> None:
        super().__init__()
        self.hier_sampling = hier_sampling
        self.include_vd = include_vd
        self.opt = opt
        self.num_sample_coarse = opt.num_sample_coarse
        self.num_sample_fine = opt.num_sample_fine
        self.vp_n_freqs = 10
        self.mlp_h_channel = opt.mlp_hidden_nchannels
        self.base_shape_code_dims = opt.iden_code_dims + opt.expr_code_dims
        self.base_appea_code_dims = opt.text_code_dims + opt.illu_code_dims
        self.base_gaze_dims = opt.eye_code_dims
        self.featmap_size, self.featmap_nc, self.pred_img_size = opt.featmap_size, opt.featmap_nc, opt.pred_img_size
        assert self.base_shape_code_dims + self.base_gaze_dims == 181 and self.base_appea_code_dims == 127, \
            "libgnrf is specialised for the reference's code dims (179 + 2, 127)"
        vp_channels = self.base_shape_code_dims + self.base_gaze_dims + self.vp_n_freqs * 6 + 3
        # include_vd (models/gaze_nerf.py:70-80): the normalised ray direction, encoded with 4 frequencies + the input (27 channels),
        # is concatenated in front of the appearance code at RGB_layer_1's input.  Supported by the fused tcgen05 inference path
        # (a per-ray bias term of the last stage); the differentiable path and the literal fp32 path raise.
        self.vd_n_freqs = 4
        self.vd_pe_dims = (self.vd_n_freqs * 6 + 3) if include_vd else 0
        vd_channels = self.base_appea_code_dims + self.vd_pe_dims
        # construction order == models/gaze_nerf.py:87-119 (eyes first) so seeded init matches the reference
        self.fg_CD_predictor_eyes = RadianceMLP(vp_channels, vd_channels, h_channel=self.mlp_h_channel, res_nfeat=self.featmap_nc)
        self.fg_CD_predictor_face = RadianceMLP(vp_channels, vd_channels, h_channel=self.mlp_h_channel, res_nfeat=self.featmap_nc)
        if self.hier_sampling:
            # present in the reference's state_dict (models/gaze_nerf.py:102-108); never evaluated there (dead path)
            self.fine_fg_CD_predictor = RadianceMLP(vp_channels, vd_channels, h_channel=self.mlp_h_channel, res_nfeat=self.featmap_nc)
        self.neural_render = NeuralRendererParams(bg_type=opt.bg_type, feat_nc=self.featmap_nc, out_dim=3, final_actvn=True,
                                                  min_feat=32, featmap_size=self.featmap_size, img_size=self.pred_img_size)
        # "tc" = fused tcgen05 kernel (default); "simt" = literal fp32 CUDA-core kernels
        self.mlp_impl = mlp_impl
        self.train_precision = "bf16x3"   # per-point activation storage of the training path: "bf16x3" | "bf16" | "f32" (train.py)
        self._tc_cache: Dict[str, Tuple[tuple, torch.Tensor]] = {}
        self._tvals_cache: Dict[Tuple[int, str], torch.Tensor] = {}
        self.last_stages: Optional[Dict[str, torch.Tensor]] = None
        self.keep_stages = False
        # bench hook: when a list, (start, end) CUDA events bracketing the fused radiance-MLP launch (gnrf_mlp_tc_fwd: mlp_tc_kernel +
        # rgb_head_kernel, issued back to back by one C call, so no host gap is inside) are appended per call
        self.mlp_events: Optional[list] = None
        self.tc_debug = None   # developer hook: (dump tensor | None, timeline tensor | None, cluster size) -> gnrf_mlp_tc_fwd_debug
        # bg_img = NeuralRenderer(bg_featmap) depends on parameters only (models/gaze_nerf.py:175-176); in no-grad inference it
        # is cached per parameter version instead of being re-rendered on every call (set False to re-render every call).
        self.cache_bg_img = True
        self._bg_cache: Optional[Tuple[tuple, torch.Tensor]] = None
        # multi-GPU: when set (dist.BatchShardedRenderer / bench.py), the coarse images are all-gathered by the kernel that writes them
        self.gather_ctx = None
        self.gather_used = False   # set by _images when the gathering neural-render kernel actually ran

    # ------------------------------------------------------------------ helpers
    def invalidate_caches(self) -> None:
        """Drop the derived caches (packed tcgen05 weight streams, cached bg_img).  They are keyed on (data_ptr, Parameter._version):
        optimizer steps, ``load_state_dict`` and in-place ops on the Parameter are seen automatically, writes through ``p.data``
        (which do not bump ``_version``) are NOT -- call this after such a write."""
        self._tc_cache.clear()
        self._bg_cache = None
        if hasattr(self.neural_render, "_tc_pack"):
            object.__setattr__(self.neural_render, "_tc_pack", None)

    def _tc_supported(self, n_s: int) -> bool:
        return self.mlp_h_channel == 384 and self.featmap_nc == 258 and n_s >= 1 and TC_TILE % n_s == 0

    def _t_vals(self, n: int, device) -> torch.Tensor:
        key = (n, str(device))
        if key not in self._tvals_cache:
            # host linspace -> device, so t has torch-CPU rounding (the oracle's), SURVEY §7 hard part 4
            self._tvals_cache[key] = torch.linspace(0.0, 1.0, n + 1, dtype=torch.float32).to(device)
        return self._tvals_cache[key]

    def _packed_tc(self, name: str, mlp: RadianceMLP) -> torch.Tensor:
        L = _lib.lib()
        params = [p.detach() for p in mlp.param_list()]
        key = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._tc_cache.get(name)
        if hit is not None and hit[0] == key:
            return hit[1]
        packed = torch.empty((L.gnrf_mlp_tc_packed_bytes(),), device=params[0].device, dtype=torch.uint8)
        _lib.check(L.gnrf_mlp_tc_pack_vd(_ptrs(params), int(self.vd_pe_dims), packed.data_ptr(), _stream()), "gnrf_mlp_tc_pack_vd")
        self._tc_cache[name] = (key, packed)
        return packed

    def _render_branches(self, ray_dl, tvecs, z_edges, shape_ext, appea, n_s: int, impl: str, want_weights: bool):
        """-> feat_ray [2][B,C,N_r], bg_alpha [2][B,N_r], weights [2][B,N_r,N_s] or None   (index 0 = face, 1 = eyes)"""
        L = _lib.lib()
        B, n_r = ray_dl.shape[0], ray_dl.shape[1]
        dev = ray_dl.device
        C = self.featmap_nc
        branches = [("face", self.fg_CD_predictor_face), ("eyes", self.fg_CD_predictor_eyes)]
        feat = [torch.empty((B, C, n_r), device=dev, dtype=torch.float32) for _ in range(2)]
        alpha = [torch.empty((B, n_r), device=dev, dtype=torch.float32) for _ in range(2)]
        wts = [torch.empty((B, n_r, n_s), device=dev, dtype=torch.float32) for _ in range(2)] if want_weights else None
        if impl == "tc":
            packed = [self._packed_tc(n, m) for n, m in branches]
            nb = L.gnrf_mlp_tc_bias_floats()
            bias = [torch.empty((B, nb), device=dev, dtype=torch.float32) for _ in range(2)]
            for i in range(2):
                _lib.check(L.gnrf_mlp_tc_fold(packed[i].data_ptr(), shape_ext.data_ptr(), appea.data_ptr(), B, bias[i].data_ptr(),
                                              _stream()), "gnrf_mlp_tc_fold")
            ws_bytes = L.gnrf_mlp_tc_workspace_bytes(2, B, n_r)
            ws = torch.empty((max(ws_bytes, 1),), device=dev, dtype=torch.uint8)
            wp = _lib.ptr_array([w.data_ptr() for w in wts]) if wts is not None else None
            args = (2, _lib.ptr_array([p.data_ptr() for p in packed]), _lib.ptr_array([b.data_ptr() for b in bias]),
                    ray_dl.data_ptr(), tvecs.data_ptr(), z_edges.data_ptr(), B, n_r, n_s,
                    _lib.ptr_array([f.data_ptr() for f in feat]), _lib.ptr_array([a.data_ptr() for a in alpha]), wp, ws.data_ptr(), ws_bytes)
            vdb = None
            if self.include_vd:   # per-ray bias of the last stage: W1[:, 384:411] . PE4(ray direction)  (models/gaze_nerf.py:70-80,140-141)
                vdb = [torch.empty((B, n_r, 192), device=dev, dtype=torch.float32) for _ in range(2)]
                for i in range(2):
                    _lib.check(L.gnrf_mlp_tc_vd_bias(packed[i].data_ptr(), ray_dl.data_ptr(), B, n_r, vdb[i].data_ptr(), _stream()),
                               "gnrf_mlp_tc_vd_bias")
            if self.mlp_events is not None:   # bench hook: events around the ONE C call that launches mlp_tc_kernel + rgb_head_kernel
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
            if vdb is not None:
                _lib.check(L.gnrf_mlp_tc_fwd_vd(args[0], args[1], args[2], _lib.ptr_array([v.data_ptr() for v in vdb]), *args[3:], _stream()),
                           "gnrf_mlp_tc_fwd_vd")
            elif self.tc_debug is None:
                _lib.check(L.gnrf_mlp_tc_fwd(*args, _stream()), "gnrf_mlp_tc_fwd")
            else:   # developer instrumentation (tests/tc_timeline.py): (dump tensor or None, timeline tensor or None, cluster size)
                dbg, prof, csize = self.tc_debug
                _lib.check(L.gnrf_mlp_tc_fwd_debug(*args, dbg.data_ptr() if dbg is not None else None,
                                                   prof.data_ptr() if prof is not None else None, int(csize), _stream()), "gnrf_mlp_tc_fwd_debug")
            if self.mlp_events is not None:
                ev1.record()
                self.mlp_events.append((ev0, ev1))
        elif impl == "simt":
            if self.include_vd:
                raise NotImplementedError("include_vd=True runs on the fused tcgen05 path only (num_sample_coarse must divide 128; "
                                          "the literal fp32 kernels have no view-direction operand)")
            feat_pts = torch.empty((B, n_r, n_s, C), device=dev, dtype=torch.float32)
            sigma_pts = torch.empty((B, n_r, n_s), device=dev, dtype=torch.float32)
            for i, (_, mlp) in enumerate(branches):
                params = [p.detach() for p in mlp.param_list()]
                _lib.check(L.gnrf_mlp_simt_fwd(_ptrs(params), ray_dl.data_ptr(), tvecs.data_ptr(), z_edges.data_ptr(),
                                               shape_ext.data_ptr(), appea.data_ptr(), B, n_r, n_s, self.mlp_h_channel, C,
                                               feat_pts.data_ptr(), sigma_pts.data_ptr(), _stream()), "gnrf_mlp_simt_fwd")
                _lib.check(L.gnrf_composite_fwd(feat_pts.data_ptr(), sigma_pts.data_ptr(), z_edges.data_ptr(), ray_dl.data_ptr(),
                                                B, n_r, n_s, C, feat[i].data_ptr(), alpha[i].data_ptr(), None,
                                                wts[i].data_ptr() if wts is not None else None, _stream()), "gnrf_composite_fwd")
        else:
            raise ValueError("mlp_impl must be 'tc' or 'simt', got %r" % (impl,))
        return feat, alpha, wts

    def _images(self, feat, alpha, gaze, B: int, only_merge: bool = False) -> Dict[str, torch.Tensor]:
        """compose (bg blend, rotate, max-merge) + ONE batched neural-render call over [face|eyes|merge|bg]
        (``only_merge``: just [merge|bg] -- what the view-sweep callers consume, utils/render_utils.py:214-219)."""
        L = _lib.lib()
        C, S = self.featmap_nc, self.featmap_size
        P = S * S
        dev = feat[0].device
        bg = _dev_f32(self.neural_render.bg_featmap.detach(), "bg_featmap")
        key = tuple((p.data_ptr(), p._version) for p in self.neural_render.parameters())
        bg_img = self._bg_cache[1] if (self.cache_bg_img and self._bg_cache is not None and self._bg_cache[0] == key) else None
        n_img = 3 * B + (0 if bg_img is not None else 1)
        fm = torch.empty((n_img, C, S, S), device=dev, dtype=torch.float32)
        _lib.check(L.gnrf_compose_fwd(feat[0].data_ptr(), alpha[0].data_ptr(), feat[1].data_ptr(), alpha[1].data_ptr(), bg.data_ptr(),
                                      gaze.data_ptr(), B, C, P, fm.data_ptr(), _stream()), "gnrf_compose_fwd")
        if bg_img is None:
            fm[3 * B].copy_(bg[0])
        if self.keep_stages:
            self.last_stages.update({"merge_face": fm[:B], "eyes_planes": fm[B:2 * B], "merge": fm[2 * B:3 * B]})
        if only_merge:
            imgs = neural_render(self.neural_render, fm[2 * B:])
            if bg_img is None:
                bg_img = imgs[B:]
                if self.cache_bg_img:
                    self._bg_cache = (key, bg_img.clone())
            return {"merge_img": imgs[:B], "bg_img": bg_img}
        gather = self.gather_ctx if (self.gather_ctx is not None and not self.hier_sampling) else None
        imgs = neural_render(self.neural_render, fm, gather=gather)  # ONE batched call over [face | eyes | merge (| bg)]
        self.gather_used = gather is not None
        if bg_img is None:
            bg_img = imgs[3 * B:]
            if self.cache_bg_img:
                self._bg_cache = (key, bg_img.clone())
        return {"merge_img_face": imgs[:B], "merge_img_eyes": imgs[B:2 * B], "merge_img": imgs[2 * B:3 * B], "bg_img": bg_img}

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def _forward(self, for_train, batch_xy, batch_uv, bg_code, shape_code, appea_code, gaze_dir, batch_Rmats, batch_Tvecs,
                 batch_inv_inmats, dist_expr, jitter_u=None, only_merge=False):
        L = _lib.lib()
        batch_size, tv, n_r = batch_xy.size()
        assert tv == 2
        assert bg_code is None  # models/gaze_nerf.py:229
        if n_r != self.featmap_size * self.featmap_size:
            raise RuntimeError("batch_xy carries %d rays, expected featmap_size^2 = %d" % (n_r, self.featmap_size ** 2))
        xy = _dev_f32(batch_xy, "batch_xy")
        dev = xy.device
        with torch.cuda.device(dev):
            _lib.check(L.gnrf_device_check(), "gnrf_device_check")
            rm = _dev_f32(batch_Rmats, "batch_Rmats").reshape(batch_size, 3, 3)
            tvecs = _dev_f32(batch_Tvecs, "batch_Tvecs").reshape(batch_size, 3)
            kinv = _dev_f32(batch_inv_inmats, "batch_inv_inmats").reshape(batch_size, 3, 3)
            gaze = _dev_f32(gaze_dir, "gaze_code").reshape(batch_size, 2)
            shape_ext = torch.cat([_dev_f32(shape_code, "shape_code"), gaze], dim=1).contiguous()  # models/gaze_nerf.py:248
            appea = _dev_f32(appea_code, "appea_code")
            assert shape_ext.shape == (batch_size, 181) and appea.shape == (batch_size, 127)
            n_s = self.num_sample_coarse
            self.last_stages = {} if self.keep_stages else None

            ray_dl = torch.empty((batch_size, n_r, 4), device=dev, dtype=torch.float32)
            _lib.check(L.gnrf_ray_setup(xy.data_ptr(), rm.data_ptr(), kinv.data_ptr(), batch_size, n_r, ray_dl.data_ptr(), _stream()),
                       "gnrf_ray_setup")
            if for_train and jitter_u is None:
                # same draw shape as torch.rand_like(zvals[B,N_r,N_s+1]) on the global generator (utils/model_utils.py:306)
                jitter_u = torch.rand((batch_size, n_r, n_s + 1), device=dev, dtype=torch.float32)
            ju = _dev_f32(jitter_u, "jitter_u") if for_train else None
            z_edges = torch.empty((batch_size, n_r, n_s + 1), device=dev, dtype=torch.float32)
            _lib.check(L.gnrf_coarse_depths(tvecs.data_ptr(), self._t_vals(n_s, dev).data_ptr(), ju.data_ptr() if ju is not None else None,
                                            batch_size, n_r, n_s, float(self.opt.world_z1), float(self.opt.world_z2), z_edges.data_ptr(),
                                            _stream()), "gnrf_coarse_depths")

            impl = self.mlp_impl or "tc"
            impl_c = impl if (impl != "tc" or self._tc_supported(n_s)) else "simt"
            feat, alpha, wts = self._render_branches(ray_dl, tvecs, z_edges, shape_ext, appea, n_s, impl_c,
                                                     want_weights=self.hier_sampling or self.keep_stages)
            if self.keep_stages:
                self.last_stages.update({"ray_dl": ray_dl, "z_edges": z_edges, "feat_face": feat[0], "feat_eyes": feat[1],
                                         "bg_alpha_face": alpha[0], "bg_alpha_eyes": alpha[1], "w_face": wts[0], "w_eyes": wts[1]})
            res_dict = {"coarse_dict": self._images(feat, alpha, gaze, batch_size, only_merge=only_merge and not self.hier_sampling)}

            if self.hier_sampling:
                # BASELINE config 3.  The reference's own hier branch is dead code; this composes its working
                # modules: FineSample on the face-branch weights (models/gaze_nerf.py:209,284), then both branch MLPs on the
                # N_c + N_f sorted samples, composite, compose, neural render.
                n_f1 = self.num_sample_fine + 1
                if for_train:
                    u = torch.rand((batch_size * n_r, n_f1), device=dev, dtype=torch.float32)  # utils/model_utils.py:425-427
                else:
                    u = self._u_lin(n_f1, dev)
                z_fine = torch.empty((batch_size, n_r, n_s + n_f1), device=dev, dtype=torch.float32)
                inds = torch.empty((batch_size * n_r, n_f1), device=dev, dtype=torch.int64) if self.keep_stages else None
                _lib.check(L.gnrf_fine_depths(wts[0].data_ptr(), z_edges.data_ptr(), u.data_ptr(), 1 if for_train else 0, batch_size, n_r,
                                              n_s, n_f1, inds.data_ptr() if inds is not None else None, z_fine.data_ptr(), _stream()),
                           "gnrf_fine_depths")
                n_sf = n_s + n_f1 - 1
                impl_f = impl if (impl != "tc" or self._tc_supported(n_sf)) else "simt"
                feat_f, alpha_f, _ = self._render_branches(ray_dl, tvecs, z_fine, shape_ext, appea, n_sf, impl_f, want_weights=False)
                if self.keep_stages:
                    self.last_stages.update({"fine_inds": inds, "z_fine": z_fine, "fine_feat_face": feat_f[0], "fine_feat_eyes": feat_f[1],
                                             "fine_bg_alpha_face": alpha_f[0], "fine_bg_alpha_eyes": alpha_f[1]})
                    coarse_stage = {k: self.last_stages[k] for k in ("merge_face", "eyes_planes", "merge")}
                res_dict["fine_dict"] = self._images(feat_f, alpha_f, gaze, batch_size)
                if self.keep_stages:
                    self.last_stages.update(coarse_stage)
        return res_dict

    def graphed(self, mode, batch_xy, batch_uv, bg_code, shape_code, appea_code, gaze_code, batch_Rmats, batch_Tvecs, batch_inv_inmats,
                gather=None, **kwargs) -> GraphedForward:
        """Capture ``forward(mode, ...)`` for these shapes into a CUDA graph; the result is called with the same keyword tensors.
        ``gather``: a dist.PeerAllGather whose fused all-gather becomes part of the captured forward (one graph per symmetric buffer)."""
        kw = dict(batch_xy=batch_xy, batch_uv=batch_uv, bg_code=bg_code, shape_code=shape_code, appea_code=appea_code, gaze_code=gaze_code,
                  batch_Rmats=batch_Rmats, batch_Tvecs=batch_Tvecs, batch_inv_inmats=batch_inv_inmats, **kwargs)
        return GraphedForward(self, mode, kw, gather=gather)

    def _wants_grad(self, train_mode: bool, *inputs) -> bool:
        """"train" with trainable parameters, or any mode with an input that requires grad (codes / gaze / camera)."""
        return (train_mode and any(p.requires_grad for p in self.parameters())) or any(torch.is_tensor(t) and t.requires_grad for t in inputs)

    def _forward_train(self, jitter, batch_xy, bg_code, shape_code, appea_code, gaze_dir, batch_Rmats, batch_Tvecs, batch_inv_inmats,
                       jitter_u=None):
        """Same outputs as ``_forward`` with gradients to parameters, codes, gaze and camera (R, T); see gazenerf_b200/train.py."""
        from .train import forward_train

        if self.include_vd:
            raise NotImplementedError("include_vd=True is inference-only in libgnrf (no reference entry point trains with the view-direction "
                                      "input, train.py:43); use torch.no_grad()")
        if self.hier_sampling:
            raise NotImplementedError("the differentiable path covers the coarse render only (the reference's hier branch is dead code, "
                                      "SURVEY §0); use torch.no_grad() for hier_sampling=True")
        if not (self.mlp_h_channel % 2 == 0 and self.featmap_size % 4 == 0 and self.num_sample_coarse % 4 == 0):
            raise RuntimeError("training path needs even mlp_hidden_nchannels, featmap_size % 4 == 0 and num_sample_coarse % 4 == 0")
        L = _lib.lib()
        batch_size, tv, n_r = batch_xy.size()
        assert tv == 2
        assert bg_code is None  # models/gaze_nerf.py:229
        if n_r != self.featmap_size * self.featmap_size:
            raise RuntimeError("batch_xy carries %d rays, expected featmap_size^2 = %d" % (n_r, self.featmap_size ** 2))
        xy = _dev_f32(batch_xy, "batch_xy")
        dev = xy.device
        with torch.cuda.device(dev):
            _lib.check(L.gnrf_device_check(), "gnrf_device_check")
            rm = _dev_f32(batch_Rmats, "batch_Rmats").reshape(batch_size, 3, 3)
            tvecs = _dev_f32(batch_Tvecs, "batch_Tvecs").reshape(batch_size, 3)
            kinv = _dev_f32(batch_inv_inmats, "batch_inv_inmats").reshape(batch_size, 3, 3)
            gaze = _dev_f32(gaze_dir, "gaze_code").reshape(batch_size, 2)
            shape_ext = torch.cat([_dev_f32(shape_code, "shape_code"), gaze], dim=1)  # models/gaze_nerf.py:248
            appea = _dev_f32(appea_code, "appea_code")
            assert shape_ext.shape == (batch_size, 181) and appea.shape == (batch_size, 127)
            n_s = self.num_sample_coarse
            if jitter and jitter_u is None:
                jitter_u = torch.rand((batch_size, n_r, n_s + 1), device=dev, dtype=torch.float32)  # utils/model_utils.py:306
            ju = _dev_f32(jitter_u, "jitter_u") if jitter else None
            z_edges = torch.empty((batch_size, n_r, n_s + 1), device=dev, dtype=torch.float32)
            _lib.check(L.gnrf_coarse_depths(tvecs.detach().data_ptr(), self._t_vals(n_s, dev).data_ptr(), ju.data_ptr() if ju is not None else None,
                                            batch_size, n_r, n_s, float(self.opt.world_z1), float(self.opt.world_z2), z_edges.data_ptr(),
                                            _stream()), "gnrf_coarse_depths")
            self.last_stages = {} if self.keep_stages else None
            imgs = forward_train(self, xy, rm, tvecs, kinv, gaze, shape_ext, appea, z_edges, stages=self.last_stages)
            B = batch_size
            return {"coarse_dict": {"merge_img_face": imgs[:B], "merge_img_eyes": imgs[B:2 * B], "merge_img": imgs[2 * B:3 * B],
                                    "bg_img": imgs[3 * B:]}}

    def _u_lin(self, n_f1: int, device) -> torch.Tensor:
        key = (-n_f1, str(device))
        if key not in self._tvals_cache:
            self._tvals_cache[key] = torch.linspace(0.0, 1.0, n_f1, dtype=torch.float32).to(device)
        return self._tvals_cache[key]

    def forward(self, mode, batch_xy, batch_uv, bg_code, shape_code, appea_code, gaze_code, batch_Rmats, batch_Tvecs,
                batch_inv_inmats, dist_expr=False, **kwargs):
        assert mode in ["train", "test"]
        inputs = (shape_code, appea_code, gaze_code, batch_Rmats, batch_Tvecs)
        if torch.is_grad_enabled() and self._wants_grad(mode == "train", *inputs):
            # differentiable path (train.py / trainer/gazenerf_trainer.py:479-528): layer-wise forward that keeps activations
            return self._forward_train(mode == "train", batch_xy, bg_code, shape_code, appea_code, gaze_code, batch_Rmats, batch_Tvecs,
                                       batch_inv_inmats, jitter_u=kwargs.get("jitter_u"))
        return self._forward(mode == "train", batch_xy, batch_uv, bg_code, shape_code, appea_code, gaze_code, batch_Rmats,
                             batch_Tvecs, batch_inv_inmats, dist_expr, jitter_u=kwargs.get("jitter_u"),
                             only_merge=bool(kwargs.get("only_merge", False)))
