"""Drop-in ``GazeNeRFLoss`` (losses/gazenerf_loss.py:190-470) with the data terms fused into libgnrf.

Same constructor, ``calc_total_loss`` / ``calc_cam_loss`` / ``calc_code_loss`` signatures, loss-dict keys and weights as the
reference.  ``calc_data_loss`` differs on purpose: the reference's (losses/gazenerf_loss.py:294-308) takes the four BOOLEAN masks that
``calc_total_loss`` derives (:420-424); here it takes the raw float mask tensors and the boolean algebra happens inside the kernel.  The five data terms (bg, eyes, face, nonhead, head) come from ONE streaming kernel instead of boolean-mask
gathers (each ``res_img[mask]`` in the reference synchronises with the host to size its output), and their image gradients from one
more (csrc/loss.cu).  The perceptual (VGG16), gaze-angular and patch-GAN terms need networks whose pretrained weights the reference
downloads at run time (losses/gazenerf_loss.py:49-52); they are out of this path's scope: asking for them raises unless the caller
injects the callables (``vgg_loss_func`` / ``gaze_loss_func``), which are then applied exactly where the reference applies them.
"""
from __future__ import annotations

import torch

from . import _lib

TERMS = ("head_loss", "eyes_loss", "face_loss", "nonhead_loss", "bg_loss")


class _DataLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_face, img_eyes, img, bg_img, gt, face_mask, full_eye, left_eye, right_eye, use_l1, bg_value):
        L = _lib.lib()
        f = lambda t: t.detach().float().contiguous()
        ins = [f(t) for t in (img_face, img_eyes, img, bg_img, gt, face_mask, full_eye, left_eye, right_eye)]
        for t in ins:
            if not t.is_cuda:
                raise RuntimeError("GazeNeRFLoss data terms run in libgnrf: tensors must be CUDA tensors")
        B, _, H, W = ins[2].shape
        dev = ins[2].device
        terms = torch.empty(5, device=dev)
        sums = torch.empty(9, device=dev)
        ws = torch.empty(L.gnrf_data_loss_workspace_floats(), device=dev)
        with torch.cuda.device(dev):   # the launch goes to the tensors' device and ITS current stream
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.gnrf_data_loss_fwd(*[t.data_ptr() for t in ins], B, H * W, 1 if use_l1 else 0, float(bg_value), terms.data_ptr(),
                                            sums.data_ptr(), ws.data_ptr(), st), "gnrf_data_loss_fwd")
        ctx.ins, ctx.sums, ctx.cfg = ins, sums, (B, H * W, 1 if use_l1 else 0, float(bg_value))
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        L = _lib.lib()
        ins, sums = ctx.ins, ctx.sums
        B, HW, use_l1, bg_value = ctx.cfg
        g = [torch.empty_like(ins[0]), torch.empty_like(ins[1]), torch.empty_like(ins[2]), torch.empty_like(ins[3])]
        gt_ = g_terms.detach().float().contiguous()
        with torch.cuda.device(ins[0].device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.gnrf_data_loss_bwd(*[t.data_ptr() for t in ins], B, HW, use_l1, bg_value, sums.data_ptr(), gt_.data_ptr(),
                                            *[t.data_ptr() for t in g], st), "gnrf_data_loss_bwd")
        return g[0], g[1], g[2], g[3], None, None, None, None, None, None, None


class GazeNeRFLoss(object):
    def __init__(self, eye_loss_importance, vgg_importance, bg_type="white", use_vgg_loss=True, use_l1_loss=False, use_angular_loss=False,
                 use_patch_gan_loss=False, device=None, vgg_loss_func=None, gaze_loss_func=None) -> None:
        self.vgg_importance = vgg_importance
        self.eye_loss_importance = eye_loss_importance
        self.eye_region_importance = 1.0
        if bg_type == "white":
            self.bg_value = 1.0
        elif bg_type == "black":
            self.bg_value = 0.0
        else:
            raise ValueError("Error BG type. ")
        self.use_vgg_loss, self.use_l1_loss = use_vgg_loss, use_l1_loss
        self.use_angular_loss, self.use_patch_gan_loss = use_angular_loss, use_patch_gan_loss
        self.device = device
        self.vgg_loss_func, self.gaze_loss_func = vgg_loss_func, gaze_loss_func
        if use_vgg_loss and vgg_loss_func is None:
            raise NotImplementedError("use_vgg_loss=True needs torchvision's pretrained VGG16 (downloaded by the reference, "
                                      "losses/gazenerf_loss.py:49-52); pass vgg_loss_func=<callable(pred, target)> or use_vgg_loss=False")
        if use_angular_loss and gaze_loss_func is None:
            raise NotImplementedError("use_angular_loss=True needs the pretrained gaze estimator; pass gaze_loss_func=<callable>")
        if use_patch_gan_loss:
            raise NotImplementedError("the patch-GAN term (discriminator) is outside the render hot path")

    @staticmethod
    def calc_cam_loss(delta_cam_info):
        return {"delta_eular": torch.mean(delta_cam_info["delta_eulur"] * delta_cam_info["delta_eulur"]),
                "delta_tvec": torch.mean(delta_cam_info["delta_tvec"] * delta_cam_info["delta_tvec"])}

    def increase_eye_importance(self):
        if self.use_l1_loss:
            self.eye_region_importance += 1.0
            self.eye_loss_importance += 30.0
        else:
            self.eye_region_importance += 30.0
            self.eye_loss_importance += 30.0

    def calc_code_loss(self, opt_code_dict):
        iden_loss = torch.mean(opt_code_dict["iden"] * opt_code_dict["iden"])
        expr_loss = torch.mean(opt_code_dict["expr"] * opt_code_dict["expr"])
        appea_loss = torch.mean(opt_code_dict["appea"] * opt_code_dict["appea"])
        bg_code = opt_code_dict["bg"]
        bg_loss = torch.zeros((), dtype=iden_loss.dtype, device=iden_loss.device) if bg_code is None else torch.mean(bg_code * bg_code)
        return {"iden_code": iden_loss, "expr_code": expr_loss, "appea_code": appea_loss, "bg_code": bg_loss}

    def calc_data_loss(self, data_dict, gt_rgb, face_mask_tensor, full_eye_mask_tensor, left_eye_mask_tensor, right_eye_mask_tensor,
                       cam_ind=None, ldms=None, epoch=0, batch_num=0, discriminator=None):
        """Takes the raw mask tensors (the boolean algebra of calc_total_loss, :420-424, happens inside the kernel)."""
        terms = _DataLossFn.apply(data_dict["merge_img_face"], data_dict["merge_img_eyes"], data_dict["merge_img"], data_dict["bg_img"], gt_rgb,
                                  face_mask_tensor, full_eye_mask_tensor, left_eye_mask_tensor, right_eye_mask_tensor, self.use_l1_loss,
                                  self.bg_value)
        res = {"bg_loss": terms[4], "eyes_loss": terms[1], "face_loss": terms[2], "nonhead_loss": terms[3]}
        if epoch > -1:
            res["head_loss"] = terms[0]
        if self.use_vgg_loss or (self.use_angular_loss and epoch > -1):
            bg = self.bg_value
            face_m = ((face_mask_tensor >= 0.5) & (left_eye_mask_tensor < 0.5) & (right_eye_mask_tensor < 0.5)).expand(-1, 3, -1, -1)
            eyes_m = ((left_eye_mask_tensor >= 0.5) | (right_eye_mask_tensor >= 0.5)).expand(-1, 3, -1, -1)
            nonhead_m = (face_mask_tensor < 0.5).expand(-1, 3, -1, -1)
            fill = torch.full_like(gt_rgb, bg)
            masked_gt = torch.where(nonhead_m, fill, gt_rgb)
            if self.use_vgg_loss:
                res["vgg_face_loss"] = self.vgg_loss_func(data_dict["merge_img_face"], torch.where(face_m, gt_rgb, fill))
                res["vgg_eyes_loss"] = self.vgg_loss_func(data_dict["merge_img_eyes"], torch.where(eyes_m, gt_rgb, fill))
                res["vgg"] = self.vgg_loss_func(data_dict["merge_img"], masked_gt) * self.vgg_importance
            if self.use_angular_loss and epoch > -1:
                res["angular"] = (self.gaze_loss_func(data_dict["merge_img"], masked_gt, cam_ind, ldms) / 60000.0) * self.eye_loss_importance
        return res

    def calc_total_loss(self, delta_cam_info, opt_code_dict, pred_dict, gt_rgb, face_mask_tensor, full_eye_mask_tensor, left_eye_mask_tensor,
                        right_eye_mask_tensor, cam_ind=None, ldms=None, epoch=0, batch_num=0, discriminator=None):
        loss_dict = self.calc_data_loss(pred_dict["coarse_dict"], gt_rgb, face_mask_tensor, full_eye_mask_tensor, left_eye_mask_tensor,
                                        right_eye_mask_tensor, cam_ind, ldms, epoch, batch_num, discriminator)
        total_loss = 0.0
        for k in loss_dict:
            total_loss = total_loss + loss_dict[k]
        if delta_cam_info is not None:
            loss_dict.update(self.calc_cam_loss(delta_cam_info))
            total_loss = total_loss + (0.001 * loss_dict["delta_eular"] + 0.001 * loss_dict["delta_tvec"])
        loss_dict.update(self.calc_code_loss(opt_code_dict))
        total_loss = total_loss + (0.001 * loss_dict["iden_code"] + 1.0 * loss_dict["expr_code"] + 0.001 * loss_dict["appea_code"]
                                   + 0.01 * loss_dict["bg_code"])
        loss_dict["total_loss"] = total_loss
        return loss_dict
