"""Host-side input fixtures of the path: pixel grid, scaled inverse intrinsics, base + orbit cameras.

Mirror of RenderUtils.build_base_info / build_cam_info (utils/render_utils.py:20-99).  Pure host code (init-time only);
the intrinsics constants are the "inv_inmat" entry of configs/config_files/cam_inmat_info_32x32.json.
The view-sweep driver loops of the reference (render_novel_views*, :101-324) are callers, SURVEY §8(f) rank 2.
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
