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


class GraphedTrainStep(object):
    """The reference's fitting step -- ``build_code_and_cam`` -> ``net("train", ...)`` -> ``GazeNeRFLoss.calc_total_loss`` ->
    ``loss.backward()`` -> ``optimizer.step()`` (trainer/gazenerf_trainer.py:479-528) -- for FIXED shapes captured ONCE into a CUDA
    graph and replayed per step: the ~370 kernel launches of a step (about a third of them small eager-PyTorch
    ops of the code / camera assembly, the loss glue and Adam) become one graph launch, and the stratified jitter is drawn on the
    device inside the graph (graph-safe Philox generator), so no host round trip is left in the step.

    ``step(base, cam, targets)`` copies the per-batch tensors into static device buffers, replays, and returns the static loss dict
    (valid until the next call).  Requirements: a CUDA optimizer built with ``capturable=True``; learnable offsets / camera deltas are
    used as whole tensors (``pos = 0``, ``batch_size`` rows).
    """

    def __init__(self, net, loss_fn, optimizer, xy, base: Dict[str, torch.Tensor], offsets: Dict[str, torch.Tensor], cam: Dict[str, torch.Tensor],
                 targets: Dict[str, torch.Tensor], delta_eulur: Optional[torch.Tensor] = None, delta_tvecs: Optional[torch.Tensor] = None,
                 jitter_u: Optional[torch.Tensor] = None, epoch: int = 0, warmup_steps: int = 3, post_backward=None):
        if not xy.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (libgnrf has no CPU path)")
        self.net, self.loss_fn, self.optimizer = net, loss_fn, optimizer
        self.offsets, self.delta_eulur, self.delta_tvecs = offsets, delta_eulur, delta_tvecs
        self.batch = int(xy.shape[0])
        self.epoch = epoch
        self.post_backward = post_backward   # e.g. a gradient all-reduce for data-parallel training; captured with the step
        clone = lambda d: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in d.items()}
        self.s_xy = xy.detach().clone()
        self.s_base, self.s_cam, self.s_tg = clone(base), clone(cam), clone(targets)
        self.s_jitter = jitter_u.detach().clone() if jitter_u is not None else None   # None: drawn inside the graph every replay
        from ._lib import lib
        L = lib()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup_steps):   # lazy initialisation (packed-weight caches, Adam state, workspaces) outside the capture
                self._one_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        net.invalidate_caches()             # the weight-packing kernels must be part of the captured step (weights change every replay)
        self.optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.gnrf_launch_count()
        with torch.cuda.graph(self.graph):
            self.static_losses = self._one_step()
        self.launches_per_replay = int(L.gnrf_launch_count() - n0)

    def _one_step(self):
        code_info, opt_code, cam_info, delta_cam = build_code_and_cam(self.s_base, self.offsets, self.s_cam, 0, self.batch, self.delta_eulur,
                                                                       self.delta_tvecs)
        extra = {"jitter_u": self.s_jitter} if self.s_jitter is not None else {}
        pred = self.net("train", self.s_xy, None, **code_info, **cam_info, **extra)
        tg = self.s_tg
        losses = self.loss_fn.calc_total_loss(delta_cam, opt_code, pred, tg["gt"], tg["head"], tg["full_eye"], tg["left_eye"], tg["right_eye"],
                                              None, None, self.epoch, 0)
        self.optimizer.zero_grad(set_to_none=True)
        losses["total_loss"].backward()
        if self.post_backward is not None:
            self.post_backward()
        self.optimizer.step()
        return {k: v.detach() for k, v in losses.items()}

    def step(self, base: Optional[Dict[str, torch.Tensor]] = None, cam: Optional[Dict[str, torch.Tensor]] = None,
             targets: Optional[Dict[str, torch.Tensor]] = None, jitter_u: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        for dst, src in ((self.s_base, base), (self.s_cam, cam), (self.s_tg, targets)):
            if src is not None:
                for k, v in src.items():
                    if torch.is_tensor(v):
                        dst[k].copy_(v, non_blocking=True)
        if jitter_u is not None:
            if self.s_jitter is None:
                raise RuntimeError("this step was captured with in-graph jitter; pass jitter_u at construction to feed it per step")
            self.s_jitter.copy_(jitter_u, non_blocking=True)
        self.graph.replay()
        return self.static_losses
