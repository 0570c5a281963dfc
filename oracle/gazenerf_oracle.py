"""CPU oracle for the GazeNeRF render hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch CPU fp32, functional, operating on a
flat ``state_dict``) of the algorithm implemented by the reference's
``GazeNeRFNet.forward``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``gazenerf_b200`` never does (it fails loudly when the CUDA library is missing).

Parity pin: the reference ships no tests / golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, generated in the build
container by ``oracle/gen_golden.py`` (which imports ``/root/reference``) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every stage against them.
The one third-party op, ``kornia.filters.filter2d`` (kornia==0.6.4, requirements.txt:9), is not
vendored by the reference; ``blur3x3`` restates its documented semantics (normalised kernel,
reflect border, depthwise cross-correlation) -- that single op is "parity unpinned" against
kornia itself and pinned only against the shim used to import the reference.

All ``file:line`` citations are relative to the reference repository root.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration (configs/gazenerf_options.py:1-35)
# --------------------------------------------------------------------------------------
@dataclass
class OracleOptions:
    featmap_size: int = 64
    featmap_nc: int = 258
    pred_img_size: int = 512
    num_sample_coarse: int = 64
    num_sample_fine: int = 128
    world_z1: float = 2.5
    world_z2: float = -3.5
    mlp_hidden_nchannels: int = 384
    iden_code_dims: int = 100
    expr_code_dims: int = 79
    text_code_dims: int = 100
    illu_code_dims: int = 27
    eye_code_dims: int = 2
    bg_type: str = "white"
    vp_n_freqs: int = 10  # models/gaze_nerf.py:27
    min_feat: int = 32  # models/gaze_nerf.py:117

    @property
    def n_blocks(self) -> int:  # models/neural_renderer.py:30
        return int(math.log2(self.pred_img_size) - math.log2(self.featmap_size))


# camera intrinsics fixture, configs/config_files/cam_inmat_info_32x32.json ("inv_inmat")
INV_INMAT_32 = (
    (0.007790804840624332, 0.0, -0.12553827464580536),
    (0.0, 0.007790804840624332, -0.12832458317279816),
    (0.0, 0.0, 1.0),
)


# --------------------------------------------------------------------------------------
# host-side input fixtures (utils/render_utils.py:20-99)
# --------------------------------------------------------------------------------------
def pixel_grid(featmap_size: int) -> Tuple[Tensor, Tensor]:
    """ray_xy [1,2,S*S] (row 0 = x, row 1 = y) and ray_uv [1,S*S,2]; utils/render_utils.py:20-34."""
    s = featmap_size
    idx = torch.arange(s * s)
    x = (idx % s).float()
    y = torch.div(idx, s, rounding_mode="floor").float()
    xy = torch.stack([x, y], 0).unsqueeze(0)
    uv = torch.stack([x / float(s), y / float(s)], -1).unsqueeze(0)
    return xy, uv


def scaled_inv_intrinsics(featmap_size: int) -> Tensor:
    """[1,3,3]; the 32x32 inverse intrinsics with [:2,:2] divided by S/32; utils/render_utils.py:36-40."""
    k = torch.tensor(INV_INMAT_32, dtype=torch.float32)
    k[:2, :2] /= featmap_size / 32.0
    return k.view(1, 3, 3)


def base_camera() -> Tuple[Tensor, Tensor]:
    """R = diag(1,-1,-1), T = (0,0,12); utils/render_utils.py:86-91."""
    r = torch.eye(3)
    r[1:, :] *= -1
    t = torch.zeros(3, 1)
    t[2, 0] = 0.5 + 11.5
    return r.view(1, 3, 3), t.view(1, 3, 1)


def orbit_cameras(view_num: int):
    """List of (R [1,3,3], T [1,3,1]) look-at cameras on a circle; utils/render_utils.py:42-84."""
    tv_z, tv_x = 0.5 + 11.5, 5.3
    center = np.zeros(3)
    radius = math.sqrt(np.sum((np.array([tv_x, 0.0, tv_z]) - center) ** 2) - np.sum((np.array([0.0, 0.0, tv_z]) - center) ** 2))
    up = np.array([0.0, -1.0, 0.0])
    cams = []
    for angle in np.linspace(0, 360.0, view_num):
        th = angle / 180.0 * 3.1415926535
        vp = np.array([math.cos(th) * radius, math.sin(th) * radius, tv_z])
        d1 = center - vp
        d2 = np.cross(up, d1)
        d3 = np.cross(d1, d2)
        d1, d2, d3 = (v / np.linalg.norm(v) for v in (d1, d2, d3))
        r = np.zeros((3, 3), dtype=np.float32)
        r[:, 0], r[:, 1], r[:, 2] = d2, d3, d1
        cams.append((torch.from_numpy(r).view(1, 3, 3), torch.from_numpy(vp).view(1, 3, 1).float()))
    return cams


# --------------------------------------------------------------------------------------
# ray generation + depth sampling (utils/model_utils.py:283-375)
# --------------------------------------------------------------------------------------
def gen_rays(xy: Tensor, rmat: Tensor, tvec: Tensor, inv_inmat: Tensor):
    """d = normalize(R K^-1 [x,y,1]); l = -1/d_z; o = T.  utils/model_utils.py:364-372.

    xy [B,2,N_r]; returns o,d [B,3,N_r], l [B,1,N_r].
    """
    b, _, n_r = xy.shape
    hom = torch.cat([xy, torch.ones(b, 1, n_r, dtype=xy.dtype)], 1)
    d = torch.bmm(rmat, torch.bmm(inv_inmat, hom))
    d = d / torch.norm(d, dim=1, keepdim=True)
    l = -1.0 / d[:, 2:3, :]
    o = tvec.expand(b, 3, n_r)
    return o, d, l


def coarse_depths(o: Tensor, n_samples: int, z1: float, z2: float) -> Tensor:
    """z_k = (o_z - z1)(1 - t_k) + (o_z - z2) t_k, t = linspace(0,1,N+1) -> [B,N_r,N+1]; utils/model_utils.py:332-357."""
    rel1 = (o[:, 2, :] - z1).unsqueeze(-1)
    rel2 = (o[:, 2, :] - z2).unsqueeze(-1)
    t = torch.linspace(0.0, 1.0, n_samples + 1, dtype=o.dtype).view(1, 1, -1)
    return rel1 * (1.0 - t) + rel2 * t


def jitter_depths(z: Tensor, u: Tensor) -> Tensor:
    """Stratified jitter, z' = lower + (upper - lower) * u; utils/model_utils.py:302-307 (u = rand_like(z))."""
    mid = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mid, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mid], -1)
    return lower + (upper - lower) * u


def points_from_depths(z: Tensor, o: Tensor, d: Tensor, l: Tensor):
    """delta = diff(z) * l ; z = z[:-1] ; pts = o + d*l*z.  utils/model_utils.py:309-315 (and :393-401).

    z [B,N_r,N+1]; o,d [B,3,N_r]; l [B,1,N_r] -> pts [B,3,N_r,N], zvals, z_dists [B,1,N_r,N].
    """
    o4, d4, l4 = o.unsqueeze(-1), d.unsqueeze(-1), l.unsqueeze(-1)
    z_dists = (z[..., 1:] - z[..., :-1]).unsqueeze(1) * l4
    zvals = z[..., :-1].unsqueeze(1)
    pts = o4 + d4 * l4 * zvals
    return pts, zvals, z_dists


def sample_points(xy, rmat, tvec, inv_inmat, n_samples, z1, z2, jitter_u: Optional[Tensor] = None):
    """GenSamplePoints.forward; utils/model_utils.py:364-375. ``jitter_u`` = the rand_like draw in train mode."""
    o, d, l = gen_rays(xy, rmat, tvec, inv_inmat)
    z = coarse_depths(o, n_samples, z1, z2)
    if jitter_u is not None:
        z = jitter_depths(z, jitter_u)
    pts, zvals, z_dists = points_from_depths(z, o, d, l)
    return {"pts": pts, "zvals": zvals, "z_dists": z_dists, "ray_o": o, "ray_d": d, "ray_l": l}


# --------------------------------------------------------------------------------------
# positional encoding (utils/model_utils.py:240-280)
# --------------------------------------------------------------------------------------
def posenc(x: Tensor, n_freqs: int = 10, include_input: bool = True) -> Tensor:
    """[B,3,...] -> [B,3+6F,...]: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(F-1) x), cos(2^(F-1) x)]."""
    freqs = 2.0 ** torch.linspace(0.0, n_freqs - 1, n_freqs)
    out = [x] if include_input else []
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, 1)


# --------------------------------------------------------------------------------------
# radiance MLP (models/mlp_nerf.py:95-119) on point-major matrices
# --------------------------------------------------------------------------------------
def _lin(sd: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    w = sd[name + ".weight"]
    w = w.reshape(w.shape[0], -1)
    return x @ w.t() + sd[name + ".bias"]


def mlp_points(sd: Dict[str, Tensor], prefix: str, vp: Tensor, vd: Tensor, n_layers: int = 8):
    """vp [P,C1], vd [P,C2] -> feat [P,res_nfeat], sigma [P].  models/mlp_nerf.py:95-119.

    8 x (1x1 conv + ReLU) with cat([vp, x]) after layer n_layers//2 (:24,:106-107); density =
    ReLU(density_module(x)) (:109,:115); feat = RGB_2(ReLU(RGB_1(cat[RGB_0(x), vd]))) (:110-113);
    sigmoid only when res_nfeat == 3 (:116-117).
    """
    p = prefix + "."
    x = vp
    for i in range(n_layers):
        x = torch.relu(_lin(sd, p + "FeaExt_module_%d" % i, x))
        if i == n_layers // 2:
            x = torch.cat([vp, x], 1)
    sigma = torch.relu(_lin(sd, p + "density_module", x))[:, 0]
    h = _lin(sd, p + "RGB_layer_0", x)
    h = torch.relu(_lin(sd, p + "RGB_layer_1", torch.cat([h, vd], 1)))
    feat = _lin(sd, p + "RGB_layer_2", h)
    if feat.shape[1] == 3:
        feat = torch.sigmoid(feat)
    return feat, sigma


def mlp_branch(sd, prefix, pe: Tensor, shape_ext: Tensor, appea: Tensor, chunk: int = 1 << 16, vd_pe: Optional[Tensor] = None):
    """pe [B,63,N_r,N_s]; shape_ext [B,181]; appea [B,127] -> feat [B,C,N_r,N_s], sigma [B,1,N_r,N_s].

    Builds the broadcast input of models/gaze_nerf.py:248-262,136-143 per point (chunked to bound memory).
    ``vd_pe`` [B,27,N_r,N_s] (include_vd=True): the view-direction encoding, concatenated IN FRONT of the appearance code
    (``torch.cat([FGvd_embedder, appea_code], dim=1)``, models/gaze_nerf.py:140-141).
    """
    b, c, n_r, n_s = pe.shape
    pts = pe.permute(0, 2, 3, 1).reshape(b, n_r * n_s, c)
    vds = vd_pe.permute(0, 2, 3, 1).reshape(b, n_r * n_s, -1) if vd_pe is not None else None
    feats, sigmas = [], []
    for i in range(b):
        fo, so = [], []
        for s in range(0, n_r * n_s, chunk):
            x = pts[i, s : s + chunk]
            vp = torch.cat([x, shape_ext[i].expand(x.shape[0], -1)], 1)
            vd = appea[i].expand(x.shape[0], -1)
            if vds is not None:
                vd = torch.cat([vds[i, s : s + chunk], vd], 1)
            f, sg = mlp_points(sd, prefix, vp, vd)
            fo.append(f)
            so.append(sg)
        feats.append(torch.cat(fo, 0))
        sigmas.append(torch.cat(so, 0))
    feat = torch.stack(feats, 0).reshape(b, n_r, n_s, -1).permute(0, 3, 1, 2).contiguous()
    sigma = torch.stack(sigmas, 0).reshape(b, 1, n_r, n_s)
    return feat, sigma


# --------------------------------------------------------------------------------------
# alpha compositing (utils/model_utils.py:493-534)
# --------------------------------------------------------------------------------------
def composite(feat: Tensor, sigma: Tensor, z_dists: Tensor, zvals: Tensor):
    """alpha = 1 - exp(-sigma*delta); T = cumprod([1, 1-alpha+1e-10]); w = alpha*T[:-1].

    Returns feat_ray [B,C,N_r], bg_alpha [B,1,N_r], depth [B,1,N_r], w [B,1,N_r,N_s].
    """
    alpha = 1.0 - torch.exp(-sigma * z_dists)
    x = 1.0 - alpha + 1e-10
    x = torch.cat([torch.ones_like(x[..., :1]), x], -1)
    trans = torch.cumprod(x, -1)
    w = alpha * trans[..., :-1]
    feat_ray = torch.sum(w * feat, -1)
    depth = torch.sum(w * zvals, -1)
    bg_alpha = 1.0 - torch.sum(w, -1)
    return feat_ray, bg_alpha, depth, w


# --------------------------------------------------------------------------------------
# gaze rotation of feature triplets + merge (utils/model_utils.py:11-46, models/gaze_nerf.py:175-203)
# --------------------------------------------------------------------------------------
def gaze_rotation(gaze: Tensor) -> Tensor:
    """[B,2] -> [B,3,3], R = Ry(g1) @ Rx(g0); utils/model_utils.py:11-29."""
    c, s = torch.cos(gaze), torch.sin(gaze)
    one, zero = torch.ones_like(c[:, 0]), torch.zeros_like(c[:, 0])
    rx = torch.stack([one, zero, zero, zero, c[:, 0], -s[:, 0], zero, s[:, 0], c[:, 0]], 1).view(-1, 3, 3)
    ry = torch.stack([c[:, 1], zero, s[:, 1], zero, one, zero, -s[:, 1], zero, c[:, 1]], 1).view(-1, 3, 3)
    return ry @ rx


def rotate_triplets(fmap: Tensor, gaze: Tensor) -> Tensor:
    """out[b,3k+j] = sum_i fmap[b,3k+i] * R_b[i,j]; models/gaze_nerf.py:181-197 + utils/model_utils.py:32-46."""
    b, c, h, w = fmap.shape
    rot = gaze_rotation(gaze.reshape(-1, 2))
    v = fmap.reshape(b, c // 3, 3, h, w)
    out = torch.einsum("bkihw,bij->bkjhw", v, rot)
    return out.reshape(b, c, h, w)


def compose_featmaps(feat_face, a_face, feat_eyes, a_eyes, bg_featmap, gaze):
    """merge_face, eyes_planes, merge; models/gaze_nerf.py:178-203. feat_* [B,C,S,S], a_* [B,1,S,S]."""
    merge_face = feat_face + a_face * bg_featmap
    merge_eyes = feat_eyes + a_eyes * bg_featmap
    eyes_planes = rotate_triplets(merge_eyes, gaze)
    merged = torch.maximum(merge_face, eyes_planes)
    return merge_face, eyes_planes, merged


# --------------------------------------------------------------------------------------
# 2-D neural renderer (models/neural_renderer.py:98-113, models/pixel_shuffle_upsample.py:7-42)
# --------------------------------------------------------------------------------------
def _reflect_idx(i: int, n: int) -> int:
    if i < 0:
        return -i
    if i >= n:
        return 2 * n - 2 - i
    return i


def blur3x3(x: Tensor) -> Tensor:
    """Depthwise [1,2,1]x[1,2,1]/16, reflect (no edge repeat) border; pixel_shuffle_upsample.py:7-16 ->
    kornia.filters.filter2d(normalized=True) (kornia 0.6.4; restated, see module docstring)."""
    h, w = x.shape[-2:]
    ih = [[_reflect_idx(i + d, h) for i in range(h)] for d in (-1, 0, 1)]
    iw = [[_reflect_idx(i + d, w) for i in range(w)] for d in (-1, 0, 1)]
    k = (1.0, 2.0, 1.0)
    out = torch.zeros_like(x)
    for a in range(3):
        rows = x[..., ih[a], :]
        for b in range(3):
            out = out + (k[a] * k[b] / 16.0) * rows[..., iw[b]]
    return out


def bilinear_up2(x: Tensor) -> Tensor:
    """nn.Upsample(scale_factor=2, bilinear, align_corners=False); neural_renderer.py:65-67.
    out[2i] = .25 in[i-1] + .75 in[i]; out[2i+1] = .75 in[i] + .25 in[i+1]; edges clamped."""

    def up_last(t):
        n = t.shape[-1]
        prev = t[..., [max(i - 1, 0) for i in range(n)]]
        nxt = t[..., [min(i + 1, n - 1) for i in range(n)]]
        even = 0.25 * prev + 0.75 * t
        odd = 0.75 * t + 0.25 * nxt
        return torch.stack([even, odd], -1).reshape(*t.shape[:-1], 2 * n)

    x = up_last(x)
    x = up_last(x.transpose(-1, -2)).transpose(-1, -2)
    return x


def conv1x1(sd, name: str, x: Tensor) -> Tensor:
    w = sd[name + ".weight"]
    w = w.reshape(w.shape[0], -1)
    return torch.einsum("oc,bchw->bohw", w, x) + sd[name + ".bias"].view(1, -1, 1, 1)


def leaky(x: Tensor) -> Tensor:
    return torch.where(x >= 0, x, 0.2 * x)


def pixel_shuffle2(x: Tensor) -> Tensor:
    """out[c, 2h+i, 2w+j] = in[4c + 2i + j, h, w] (F.pixel_shuffle(.,2))."""
    b, c4, h, w = x.shape
    c = c4 // 4
    x = x.reshape(b, c, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3)
    return x.reshape(b, c, 2 * h, 2 * w)


def psu(sd, prefix: str, x: Tensor) -> Tensor:
    """PixelShuffleUpsample.forward; pixel_shuffle_upsample.py:33-42."""
    y = x.repeat(1, 4, 1, 1)
    out = leaky(conv1x1(sd, prefix + ".layer_1", x))
    out = leaky(conv1x1(sd, prefix + ".layer_2", out))
    out = out + y
    return blur3x3(pixel_shuffle2(out))


def neural_render(sd, x: Tensor, n_blocks: int, prefix: str = "neural_render") -> Tensor:
    """NeuralRenderer.forward; models/neural_renderer.py:98-113."""
    p = prefix + "."
    rgb = blur3x3(bilinear_up2(conv1x1(sd, p + "feat_2_rgb_list.0", x)))
    net = x
    for i in range(n_blocks):
        net = leaky(conv1x1(sd, p + "feat_layers.%d" % i, psu(sd, p + "feat_upsample_list.%d" % i, net)))
        rgb = rgb + conv1x1(sd, p + "feat_2_rgb_list.%d" % (i + 1), net)
        if i < n_blocks - 1:
            rgb = blur3x3(bilinear_up2(rgb))
    return torch.sigmoid(rgb)


# --------------------------------------------------------------------------------------
# hierarchical fine sampling (utils/model_utils.py:378-490)
# --------------------------------------------------------------------------------------
def fine_sample(w: Tensor, zvals: Tensor, o: Tensor, d: Tensor, l: Tensor, n_fine: int, u: Optional[Tensor] = None):
    """FineSample.forward; utils/model_utils.py:404-490.

    w [B,1,N_r,N_c] coarse weights, zvals [B,1,N_r,N_c]; n_fine = opt.num_sample_fine (the module draws
    n_fine+1 samples, :381); u = optional [B*N_r, n_fine+1] uniform draw (train), else linspace(0,1).
    Returns dict with inds (int64 [B*N_r, n_fine+1]), z_sorted [B,N_r,N_c+n_fine+1], pts, zvals, z_dists.
    """
    nf = n_fine + 1
    b, _, n_r, n_c = w.shape
    tw = w[:, :, :, 1:-1].reshape(-1, n_c - 2)
    m = n_c - 2
    pdf = tw / torch.sum(tw + 1e-5, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1).contiguous()  # [N_t, m+1]
    n_t = cdf.shape[0]
    if u is None:
        u = torch.linspace(0.0, 1.0, nf, dtype=w.dtype).view(1, nf).expand(n_t, nf)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=m)
    zc = zvals.reshape(n_t, n_c)
    bins = 0.5 * (zc[:, 1:] + zc[:, :-1])  # [N_t, m+1]
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    zf = bin_b + t * (bin_a - bin_b)
    z_sorted, _ = torch.sort(torch.cat([zc, zf], -1), -1)
    z_sorted = z_sorted.view(b, n_r, nf + n_c)
    pts, zv, zd = points_from_depths(z_sorted, o, d, l)
    return {"inds": inds, "z_sorted": z_sorted, "pts": pts, "zvals": zv, "z_dists": zd}


# --------------------------------------------------------------------------------------
# full forward (models/gaze_nerf.py:121-351)
# --------------------------------------------------------------------------------------
def render_branches(sd, opt: OracleOptions, pts, z_dists, zvals, shape_code, appea_code, gaze_code, ray_d: Optional[Tensor] = None):
    """PE -> both branch MLPs -> composite. Returns per-branch (feat_ray, bg_alpha, w).
    ``ray_d`` [B,3,N_r] (include_vd=True, models/gaze_nerf.py:70-80,240-243): the normalised ray directions, expanded over the samples
    (utils/model_utils.py:317-318) and encoded with 4 frequencies + the input (27 channels)."""
    pe = posenc(pts, opt.vp_n_freqs, True)
    shape_ext = torch.cat([shape_code, gaze_code], 1)  # models/gaze_nerf.py:248
    vd_pe = None
    if ray_d is not None:
        vd_pe = posenc(ray_d.unsqueeze(-1).expand(-1, -1, -1, pts.shape[-1]), 4, True)
    out = {}
    for name in ("face", "eyes"):
        feat, sigma = mlp_branch(sd, "fg_CD_predictor_" + name, pe, shape_ext, appea_code, vd_pe=vd_pe)
        fr, ba, _, w = composite(feat, sigma, z_dists, zvals)
        out[name] = (fr, ba, w)
    return out


def forward(sd, opt: OracleOptions, mode: str, batch_xy, shape_code, appea_code, gaze_code, rmats, tvecs,
            inv_inmats, jitter_u: Optional[Tensor] = None, return_stages: bool = False, include_vd: bool = False):
    """GazeNeRFNet.forward (hier_sampling=False); models/gaze_nerf.py:211-351.  ``include_vd``: the view-direction input (:70-80).

    In "train" mode the caller passes ``jitter_u`` (what ``torch.rand_like(zvals)`` would draw).
    """
    assert mode in ("train", "test")
    b = batch_xy.shape[0]
    s = opt.featmap_size
    smp = sample_points(batch_xy, rmats, tvecs, inv_inmats, opt.num_sample_coarse, opt.world_z1, opt.world_z2,
                        jitter_u if mode == "train" else None)
    br = render_branches(sd, opt, smp["pts"], smp["z_dists"], smp["zvals"], shape_code, appea_code, gaze_code,
                         ray_d=smp["ray_d"] if include_vd else None)
    c = opt.featmap_nc
    feat_face = br["face"][0].view(b, c, s, s)
    feat_eyes = br["eyes"][0].view(b, c, s, s)
    a_face = br["face"][1].view(b, 1, s, s)
    a_eyes = br["eyes"][1].view(b, 1, s, s)
    bg = sd["neural_render.bg_featmap"]
    merge_face, eyes_planes, merged = compose_featmaps(feat_face, a_face, feat_eyes, a_eyes, bg, gaze_code)
    nb = opt.n_blocks
    res = {
        "merge_img_face": neural_render(sd, merge_face, nb),
        "merge_img_eyes": neural_render(sd, eyes_planes, nb),
        "merge_img": neural_render(sd, merged, nb),
        "bg_img": neural_render(sd, bg, nb),
    }
    out = {"coarse_dict": res}
    if return_stages:
        out["stages"] = {
            "pts": smp["pts"], "zvals": smp["zvals"], "z_dists": smp["z_dists"],
            "feat_face": feat_face, "feat_eyes": feat_eyes, "bg_alpha_face": a_face, "bg_alpha_eyes": a_eyes,
            "w_face": br["face"][2], "w_eyes": br["eyes"][2],
            "merge_face": merge_face, "eyes_planes": eyes_planes, "merge": merged,
            "ray_o": smp["ray_o"], "ray_d": smp["ray_d"], "ray_l": smp["ray_l"],
        }
    return out


def forward_hier(sd, opt: OracleOptions, batch_xy, shape_code, appea_code, gaze_code, rmats, tvecs, inv_inmats,
                 n_fine: int):
    """BASELINE config 3: coarse pass (both branches) -> FineSample on the face-branch weights
    (models/gaze_nerf.py:209 returns ori_batch_weight_face) -> second pass of both branch MLPs on the
    N_c + n_fine sorted samples -> composite.  The reference's own hier path is dead code, so this
    composes its working modules.  Returns fine feat/bg_alpha per branch + the fine-sample dict."""
    smp = sample_points(batch_xy, rmats, tvecs, inv_inmats, opt.num_sample_coarse, opt.world_z1, opt.world_z2)
    coarse = render_branches(sd, opt, smp["pts"], smp["z_dists"], smp["zvals"], shape_code, appea_code, gaze_code)
    fs = fine_sample(coarse["face"][2], smp["zvals"], smp["ray_o"], smp["ray_d"], smp["ray_l"], n_fine)
    fine = render_branches(sd, opt, fs["pts"], fs["z_dists"], fs["zvals"], shape_code, appea_code, gaze_code)
    return {"coarse": coarse, "fine_sample": fs, "fine": fine}


# --------------------------------------------------------------------------------------
# synthetic parameters / inputs (SURVEY §8d "Synthetic inputs")
# --------------------------------------------------------------------------------------
def synthetic_codes(batch: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    shape = torch.randn(batch, 179, generator=g) * 0.3
    appea = torch.randn(batch, 127, generator=g) * 0.3
    gaze = torch.rand(batch, 2, generator=g) - 0.5
    return shape, appea, gaze


def densify(sd: Dict[str, Tensor], bias_face: float, bias_eyes: float, scale: float = 30.0) -> Dict[str, Tensor]:
    """Non-vacuous "dense" weight variant (SURVEY §8c caveat 1): density weights * scale, density bias set per branch.
    With reference init the density is ~0 and every image is pure background, so parity would be vacuous."""
    sd = dict(sd)
    for br, b in (("face", bias_face), ("eyes", bias_eyes)):
        k = "fg_CD_predictor_%s.density_module." % br
        sd[k + "weight"] = sd[k + "weight"] * scale
        sd[k + "bias"] = torch.full_like(sd[k + "bias"], float(b))
    return sd


def calibrate_dense_bias(sd, opt: OracleOptions, batch_xy, shape_code, appea_code, gaze_code, rmats, tvecs, inv_inmats,
                         scale: float = 30.0, max_rays: int = 64):
    """Biases that centre the scaled density on its median over (a subsample of) these inputs, so about half of the
    sample points are opaque in both branches.  Same recipe as oracle/gen_golden.py:make_dense."""
    n_r = batch_xy.shape[2]
    step = max(1, n_r // max_rays)
    smp = sample_points(batch_xy[:, :, ::step].contiguous(), rmats, tvecs, inv_inmats, opt.num_sample_coarse, opt.world_z1,
                        opt.world_z2)
    pe = posenc(smp["pts"], opt.vp_n_freqs, True)
    shape_ext = torch.cat([shape_code, gaze_code], 1)
    out = []
    for br in ("face", "eyes"):
        k = "fg_CD_predictor_%s.density_module." % br
        sd0 = dict(sd)
        sd0[k + "weight"] = -sd[k + "weight"]  # ReLU(-raw) and ReLU(raw) together recover the signed raw density
        sd0[k + "bias"] = torch.zeros_like(sd[k + "bias"])
        sd1 = dict(sd)
        sd1[k + "bias"] = torch.zeros_like(sd[k + "bias"])
        _, neg = mlp_branch(sd0, "fg_CD_predictor_" + br, pe, shape_ext, appea_code)
        _, pos = mlp_branch(sd1, "fg_CD_predictor_" + br, pe, shape_ext, appea_code)
        raw = pos - neg
        out.append(-scale * float(raw.median()))
    return out[0], out[1]


# --------------------------------------------------------------------------------------
# GazeNeRFLoss data terms (losses/gazenerf_loss.py:294-352 + mask algebra :420-424), use_vgg_loss=False
# --------------------------------------------------------------------------------------
def data_loss_terms(pred: Dict[str, Tensor], gt: Tensor, face_mask: Tensor, full_eye: Tensor, left_eye: Tensor, right_eye: Tensor,
                    use_l1: bool = True, bg_value: float = 1.0) -> Dict[str, Tensor]:
    """Boolean-gather formulation exactly as the reference writes it."""
    head_m = torch.logical_and(face_mask >= 0.5, full_eye < 0.5).expand(-1, 3, -1, -1)
    face_m = torch.logical_and(face_mask >= 0.5, torch.logical_and(left_eye < 0.5, right_eye < 0.5)).expand(-1, 3, -1, -1)
    eyes_m = torch.logical_or(left_eye >= 0.5, right_eye >= 0.5).expand(-1, 3, -1, -1)
    nonhead_m = (face_mask < 0.5).expand(-1, 3, -1, -1)
    fn = torch.nn.functional.l1_loss if use_l1 else torch.nn.functional.mse_loss
    bg = pred["bg_img"]
    tv = pred["merge_img"][nonhead_m] - bg_value
    return {
        "bg_loss": torch.mean((bg - bg_value) * (bg - bg_value)),
        "eyes_loss": fn(pred["merge_img_eyes"][eyes_m], gt[eyes_m]),
        "face_loss": fn(pred["merge_img_face"][face_m], gt[face_m]),
        "nonhead_loss": torch.mean(tv * tv),
        "head_loss": fn(pred["merge_img"][head_m], gt[head_m]),
    }


def synthetic_loss_inputs(batch: int, size: int, seed: int = 0):
    """Random images / gt and blob-like masks (head ellipse, two eye boxes, full-eye box) for loss parity."""
    g = torch.Generator().manual_seed(seed)
    pred = {k: torch.rand(batch, 3, size, size, generator=g) for k in ("merge_img_face", "merge_img_eyes", "merge_img")}
    pred["bg_img"] = torch.rand(1, 3, size, size, generator=g)
    gt = torch.rand(batch, 3, size, size, generator=g)
    yy, xx = torch.meshgrid(torch.arange(size).float(), torch.arange(size).float(), indexing="ij")
    c = (size - 1) / 2.0
    face = ((((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2) < 1.0).float()
    face = (face * (0.6 + 0.4 * torch.rand(batch, 1, size, size, generator=g)))          # soft values around the 0.5 threshold
    box = lambda y0, y1, x0, x1: ((yy >= y0 * size) & (yy < y1 * size) & (xx >= x0 * size) & (xx < x1 * size)).float().expand(batch, 1, -1, -1).clone()
    left, right = box(0.38, 0.46, 0.30, 0.44), box(0.38, 0.46, 0.56, 0.70)
    full_eye = box(0.36, 0.48, 0.28, 0.72)
    return pred, gt, face, full_eye, left, right


# --------------------------------------------------------------------------------------
# dataset sample -> tensors (datasets/eth_xgaze.py:308-360 + trainer/gazenerf_trainer.py:250-337)
# --------------------------------------------------------------------------------------
def synthetic_hdf5_records(batch: int, size: int = 512, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random records in the HDF5 schema of dataset_pre_processing.py:260-380 (dtypes as stored: u8 images / masks, f64 parameters)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    c = (size - 1) / 2.0
    head = (((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2 < 1.0)
    head_mask = np.stack([np.roll(head, rng.randint(-size // 5, size // 5), axis=rng.randint(2)) for _ in range(batch)]).astype(np.uint8) * 255
    head_mask[:, :3, :] = 255                       # touches the border: cv2's erosion ignores out-of-image taps
    head_mask[rng.rand(batch, size, size) < 0.002] = 0   # pin-holes that the 5x5 erosion grows
    box = lambda y0, y1, x0, x1: ((yy >= y0 * size) & (yy < y1 * size) & (xx >= x0 * size) & (xx < x1 * size))
    left = np.repeat(box(0.38, 0.46, 0.30, 0.44)[None], batch, 0).astype(np.uint8) * 255
    right = np.repeat(box(0.38, 0.46, 0.56, 0.70)[None], batch, 0).astype(np.uint8) * 255
    inmat = np.tile(np.array([[1600.0, 0.0, 256.0], [0.0, 1600.0, 256.0], [0.0, 0.0, 1.0]]), (batch, 1, 1)) + rng.randn(batch, 3, 3) * [[30, 0, 5], [0, 30, 5], [0, 0, 0]]
    rmat = np.linalg.qr(rng.randn(batch, 3, 3))[0]
    return {
        "face_patch": rng.randint(0, 256, (batch, size, size, 3)).astype(np.uint8),   # BGR
        "head_mask": head_mask, "left_eye_mask": left, "right_eye_mask": right,
        "latent_codes_row0": rng.randn(306) * 0.5, "latent_codes": rng.randn(batch, 306) * 0.5,
        "pitchyaw_head": rng.rand(batch, 2) - 0.5, "c2w_Rmat": rmat, "c2w_Tvec": rng.randn(batch, 3) + [0.0, 0.0, 12.0], "inmat": inmat,
    }


def dataset_sample_tensors(rec: Dict[str, np.ndarray], featmap_size: int, opt: Optional[OracleOptions] = None) -> Dict[str, Tensor]:
    """What the reference's DataLoader item + ``prepare_data`` leave on the trainer, computed the way the reference computes it.

    Per item (datasets/eth_xgaze.py:326-360): ``image[:, :, [2,1,0]]`` then ``ToPILImage -> ToTensor`` (:12) = u8 HWC -> f32 CHW / 255;
    ``cv2.erode(head_mask, ones((3,3), u8), iterations=2)``; ``code = latent_codes[0]`` with ``[279:]`` from the item's row.
    Per batch (trainer/gazenerf_trainer.py:256-336): masks ``unsqueeze(1)``; the 306-d code split 100 | 79 | 100 | 27 and cast to float;
    ``inmat[:, :2, :] *= featmap_size / img_size`` and its closed-form inverse in f64, cast to float.
    Third-party ops are called, not restated: cv2.erode (OpenCV) and the uint8 -> float / 255 of torchvision's ToTensor."""
    import cv2

    opt = opt or OracleOptions()
    b, size = rec["face_patch"].shape[0], rec["face_patch"].shape[1]
    imgs, heads = [], []
    for i in range(b):
        rgb = rec["face_patch"][i][:, :, [2, 1, 0]]
        imgs.append(torch.from_numpy(np.ascontiguousarray(rgb)).permute(2, 0, 1).contiguous().to(torch.float32).div(255))
        heads.append(torch.from_numpy(cv2.erode(rec["head_mask"][i], np.ones((3, 3), dtype=np.uint8), iterations=2)))
    code = np.repeat(rec["latent_codes_row0"][None].copy(), b, 0)
    code[:, 279:] = rec["latent_codes"][:, 279:]
    code = torch.from_numpy(code)
    i0, i1, i2 = opt.iden_code_dims, opt.iden_code_dims + opt.expr_code_dims, opt.iden_code_dims + opt.expr_code_dims + opt.text_code_dims
    k = torch.from_numpy(rec["inmat"].copy())
    k[:, :2, :] *= featmap_size / size
    kinv = torch.zeros_like(k)
    kinv[:, 0, 0] = 1.0 / k[:, 0, 0]
    kinv[:, 1, 1] = 1.0 / k[:, 1, 1]
    kinv[:, 0, 2] = -(k[:, 0, 2] / k[:, 0, 0])
    kinv[:, 1, 2] = -(k[:, 1, 2] / k[:, 1, 1])
    kinv[:, 2, 2] = 1.0
    f = lambda t: t.type(torch.FloatTensor)
    return {
        "img_tensor": torch.stack(imgs, 0),
        "head_mask_tensor": torch.stack(heads, 0).unsqueeze(1),
        "left_eye_mask_tensor": torch.from_numpy(rec["left_eye_mask"]).unsqueeze(1),
        "right_eye_mask_tensor": torch.from_numpy(rec["right_eye_mask"]).unsqueeze(1),
        "base_iden": f(code[:, :i0]), "base_expr": f(code[:, i0:i1]), "base_text": f(code[:, i1:i2]), "base_illu": f(code[:, i2:]),
        "base_gaze_direction": f(torch.from_numpy(rec["pitchyaw_head"])),
        "cam_info": {"batch_Rmats": f(torch.from_numpy(rec["c2w_Rmat"])), "batch_Tvecs": f(torch.from_numpy(rec["c2w_Tvec"]).unsqueeze(-1)),
                     "batch_inv_inmats": f(kinv)},
    }
