"""Host-side input fixtures of the path: pixel grid, scaled inverse intrinsics, base + orbit cameras.

Mirror of RenderUtils.build_base_info / build_cam_info (utils/render_utils.py:20-99).  Pure host code (init-time only);
the intrinsics constants are the "inv_inmat" entry of configs/config_files/cam_inmat_info_32x32.json.
The view-sweep drivers (render_novel_views / render_novel_views_gaze / render_morphing_res, :101-324) keep
the reference's signatures and outputs (lists of HxWx3 uint8 arrays) but submit ALL views of a sweep as one batch through the
same kernels instead of 45 sequential batch-1 forwards, convert to uint8 on the device and copy back once.
"""
import math

import numpy as np
import torch

INV_INMAT_32 = [
    [0.007790804840624332, 0.0, -0.12553827464580536],
    [0.0, 0.007790804840624332, -0.12832458317279816],
    [0.0, 0.0, 1.0],
]


class RenderUtils(object):
    def __init__(self, view_num, device, opt) -> None:
        self.view_num = view_num
        self.device = device
        self.opt = opt
        self.build_base_info()
        self.build_cam_info()

    def build_base_info(self):
        s = self.opt.featmap_size
        idx = torch.arange(s * s)
        x = (idx % s).view(-1)
        y = torch.div(idx, s, rounding_mode="floor").view(-1)
        self.ray_xy = torch.stack([x, y], dim=0).float().unsqueeze(0).to(self.device)
        self.ray_uv = torch.stack([x.float() / float(s), y.float() / float(s)], dim=-1).unsqueeze(0).to(self.device)
        k = torch.as_tensor(INV_INMAT_32)
        k[:2, :2] /= s / 32.0
        self.inv_inmat = k.view(1, 3, 3).to(self.device)

    def build_cam_info(self):
        tv_z, tv_x = 0.5 + 11.5, 5.3
        center = np.zeros(3)
        radius = math.sqrt(np.sum((np.array([tv_x, 0.0, tv_z]) - center) ** 2) - np.sum((np.array([0.0, 0.0, tv_z]) - center) ** 2))
        up = np.array([0.0, -1.0, 0.0])
        self.cam_info_list = []
        for angle in np.linspace(0, 360.0, self.view_num):
            th = angle / 180.0 * 3.1415926535
            vp = np.array([math.cos(th) * radius, math.sin(th) * radius, tv_z])
            d1 = center - vp
            d2 = np.cross(up, d1)
            d3 = np.cross(d1, d2)
            d1, d2, d3 = (v / np.linalg.norm(v) for v in (d1, d2, d3))
            r = np.zeros((3, 3), dtype=np.float32)
            r[:, 0], r[:, 1], r[:, 2] = d2, d3, d1
            self.cam_info_list.append({
                "batch_Rmats": torch.from_numpy(r).view(1, 3, 3).to(self.device),
                "batch_Tvecs": torch.from_numpy(vp).view(1, 3, 1).float().to(self.device),
                "batch_inv_inmats": self.inv_inmat,
            })
        base_r = torch.eye(3).float().view(1, 3, 3)
        base_r[0, 1:, :] *= -1
        base_t = torch.zeros(3).float().view(1, 3, 1)
        base_t[0, 2, 0] = tv_z
        self.base_cam_info = {"batch_Rmats": base_r.to(self.device), "batch_Tvecs": base_t.to(self.device),
                              "batch_inv_inmats": self.inv_inmat}

    # ------------------------------------------------------------------------------------------------ view sweeps (batched)
    # gaze sweep tables of render_novel_views (utils/render_utils.py:104-197)
    _SWEEP_H = [-0.3] * 5 + [-0.2] * 3 + [-0.1] * 3 + [0.0] + [0.1] * 3 + [0.2] * 3 + [0.3] * 10 + [0.2] * 3 + [0.1, 0.0, -0.1] + [-0.2] * 3 + [-0.3] * 8
    _SWEEP_V = ([0.0, -0.1, -0.2, -0.2, -0.3, -0.3] + [-0.4] * 12 + [-0.3, -0.3, -0.2, -0.2, -0.1, 0.0, 0.1, 0.2, 0.2, 0.3, 0.3] + [0.4] * 10 +
                [0.3, 0.3, 0.2, 0.2, 0.1, 0.0])

    def _render_batch(self, net, shape_code, appea_code, gazes, cams, max_batch=48):
        """views -> list of uint8 HxWx3 arrays of merge_img.  shape/appea: [1,*] (broadcast) or [V,*]; gazes [V,2]; cams: V dicts."""
        V = gazes.shape[0]
        dev = gazes.device
        out = []
        for s0 in range(0, V, max_batch):
            s1 = min(V, s0 + max_batch)
            n = s1 - s0
            pick = lambda t: (t.expand(n, -1) if t.shape[0] == 1 else t[s0:s1]).contiguous()
            cam = {k: torch.cat([cams[i][k] for i in range(s0, s1)], 0) for k in ("batch_Rmats", "batch_Tvecs", "batch_inv_inmats")}
            with torch.no_grad():
                pred = net("test", self.ray_xy.expand(n, -1, -1), None, bg_code=None, shape_code=pick(shape_code), appea_code=pick(appea_code),
                           gaze_code=gazes[s0:s1].contiguous(), only_merge=True, **cam)
            img = pred["coarse_dict"]["merge_img"]
            u8 = (img.permute(0, 2, 3, 1) * 255).to(torch.uint8)   # == (x * 255).astype(np.uint8): truncation, values in (0, 255)
            out += list(u8.cpu().numpy())
        return out

    def render_novel_views(self, net, code_info, move_gaze=True):
        """utils/render_utils.py:101-221: orbit the head over view_num cameras while sweeping the gaze (or holding (0,-0.5))."""
        dev = code_info["gaze_code"].device
        n = self.view_num
        if move_gaze:
            assert n <= len(self._SWEEP_H), "the reference's gaze tables have 45 entries"
            gazes = torch.tensor(list(zip(self._SWEEP_H[:n], self._SWEEP_V[:n])), dtype=torch.float32, device=dev)
        else:
            gazes = torch.tensor([[0.0, -0.5]] * n, dtype=torch.float32, device=dev)
        res = self._render_batch(net, code_info["shape_code"], code_info["appea_code"], gazes, self.cam_info_list)
        code_info["gaze_code"][:, :] = gazes[-1]   # the reference mutates the caller's gaze_code in place; leave the last value
        return res

    def render_novel_views_gaze(self, net, code_info, cam_info):
        """utils/render_utils.py:223-289: a rectangular gaze sweep under a fixed camera."""
        horizontal, vertical, range_x, range_y = [-20, 20], [-50, 50], 4, 10
        g = []
        g += [(horizontal[0] / 100.0, j / 100.0) for j in range(vertical[0], vertical[1] + 1, range_y)]
        g += [(j / 100.0, vertical[1] / 100.0) for j in range(horizontal[0], horizontal[1] + 1, range_x)]
        g += [(horizontal[1] / 100.0, j / 100.0) for j in range(vertical[1], vertical[0] + 1, -range_y)]
        g += [(j / 100.0, vertical[0] / 100.0) for j in range(horizontal[1], horizontal[0] + 1, -range_x)]
        dev = code_info["gaze_code"].device
        gazes = torch.tensor(g, dtype=torch.float32, device=dev)
        res = self._render_batch(net, code_info["shape_code"], code_info["appea_code"], gazes, [cam_info] * len(g))
        code_info["gaze_code"][:, :] = gazes[-1]
        return res

    def render_morphing_res(self, net, code_info_1, code_info_2, nums):
        """utils/render_utils.py:291-324: linear interpolation of the shape / appearance codes under the base camera.
        (The reference's loop omits gaze_code and cannot run as written; here the gaze of code_info_1 is used when present, else 0.)"""
        tv = 1.0 - torch.arange(nums, dtype=torch.float32, device=code_info_1["shape_code"].device) / (nums - 1)
        tv = tv.view(-1, 1)
        shape = code_info_1["shape_code"] * tv + code_info_2["shape_code"] * (1 - tv)
        appea = code_info_1["appea_code"] * tv + code_info_2["appea_code"] * (1 - tv)
        gaze = code_info_1.get("gaze_code")
        gazes = (gaze if gaze is not None else torch.zeros(1, 2, device=shape.device)).expand(nums, -1).contiguous().float()
        return self._render_batch(net, shape, appea, gazes, [self.base_cam_info] * nums)
