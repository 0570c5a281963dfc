"""Per-step code / camera assembly of the reference's trainer as batched, launch-lean tensor functions.

Mirrors ``BaseTrainer.eulurangle2Rmat`` (trainer/base.py:92-124) and ``GazeNeRFTrainer.build_code_and_cam``
(trainer/gazenerf_trainer.py:338-405).  The reference builds three identity matrices and writes twelve slices in place per step;
here each rotation is assembled with one ``stack`` (same values, same autograd graph semantics), so that once the render itself takes
milliseconds these host-launched micro-ops do not dominate the step.  Pure torch (device-agnostic, differentiable): this is caller
code of the hot path, not part of libgnrf.
"""
from typing import Dict, Optional, Tuple

import torch


def eulurangle2Rmat(angles: torch.Tensor) -> torch.Tensor:
    """[B,3] Euler angles (x, y, z) -> [B,3,3] = Rz @ Ry @ Rx  (trainer/base.py:92-124)."""
    s, c = torch.sin(angles), torch.cos(angles)
    sx, sy, sz = s[:, 0], s[:, 1], s[:, 2]
    cx, cy, cz = c[:, 0], c[:, 1], c[:, 2]
    one, zero = torch.ones_like(sx), torch.zeros_like(sx)
    rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], 1).view(-1, 3, 3)
    ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], 1).view(-1, 3, 3)
    rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], 1).view(-1, 3, 3)
    return rz.bmm(ry.bmm(rx))


def build_code_and_cam(base: Dict[str, torch.Tensor], offsets: Dict[str, torch.Tensor], cam: Dict[str, torch.Tensor], pos: int, batch_size: int,
                       delta_eulur: Optional[torch.Tensor] = None, delta_tvecs: Optional[torch.Tensor] = None
                       ) -> Tuple[dict, dict, dict, Optional[dict]]:
    """trainer/gazenerf_trainer.py:338-405.

    base: ``iden, expr, text, illu, gaze`` (the per-sample 3DMM codes + gaze direction); offsets: the learnable ``iden, expr, appea``
    offset tables; cam: ``batch_Rmats, batch_Tvecs, batch_inv_inmats``; delta_*: learnable camera corrections (opt_cam) or None.
    Returns (code_info, opt_code_dict, cam_info, delta_cam_info) exactly as the reference does."""
    sl = slice(pos, pos + batch_size)
    shape_code = torch.cat([base["iden"] + offsets["iden"][sl], base["expr"] + offsets["expr"][sl]], dim=-1).float()
    appea_code = (torch.cat([base["text"], base["illu"]], dim=-1) + offsets["appea"][sl]).float()
    code_info = {"bg_code": None, "shape_code": shape_code, "appea_code": appea_code, "gaze_code": base["gaze"].float()}
    opt_code_dict = {"bg": None, "iden": offsets["iden"][sl], "expr": offsets["expr"][sl], "appea": offsets["appea"][sl]}
    if delta_eulur is None:
        return code_info, opt_code_dict, cam, None
    d_r = eulurangle2Rmat(delta_eulur[sl])
    cam_info = {"batch_Rmats": d_r.bmm(cam["batch_Rmats"]), "batch_Tvecs": d_r.bmm(cam["batch_Tvecs"]) + delta_tvecs[sl],
                "batch_inv_inmats": cam["batch_inv_inmats"]}
    return code_info, opt_code_dict, cam_info, {"delta_eulur": delta_eulur[sl], "delta_tvec": delta_tvecs[sl]}
