"""Dataset sample -> device tensors: the step in front of the render path.

Mirrors what the reference does per training item on the CPU -- ``GazeDataset.__getitem__`` (datasets/eth_xgaze.py:308-360: BGR->RGB,
ToTensor, ``cv2.erode`` of the head mask, latent-code assembly) followed by ``GazeNeRFTrainer.prepare_data``
(trainer/gazenerf_trainer.py:250-337: code split, float casts, intrinsics rescale + inverse, ``.to(device)``) -- but uploads the raw
HDF5 records (u8 pixels, f64 parameters; schema dataset_pre_processing.py:260-380) once and does every transform in libgnrf
(csrc/data.cu).  No ``h5py`` here: the caller hands over the arrays it read (``hdf["face_patch"][idx]`` ...), batched on axis 0.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

_U8_KEYS = ("face_patch", "head_mask", "left_eye_mask", "right_eye_mask")
_F64_KEYS = ("latent_codes_row0", "latent_codes", "pitchyaw_head", "c2w_Rmat", "c2w_Tvec", "inmat")


class SampleStager(object):
    """Re-usable pinned-host + device staging for batches of raw samples of a fixed shape.

    ``stager(records)`` with ``records`` = dict of numpy arrays
        face_patch u8 [B,H,W,3] (BGR) · head_mask / left_eye_mask / right_eye_mask u8 [B,H,W] · latent_codes f64 [B,306] (the samples'
        own rows) · latent_codes_row0 f64 [306] (row 0 of the subject file) · pitchyaw_head f64 [B,2] · c2w_Rmat f64 [B,3,3] ·
        c2w_Tvec f64 [B,3] · inmat f64 [B,3,3]
    returns the tensors ``prepare_data`` leaves on the trainer, all fp32 on ``device``:
        img_tensor [B,3,H,W] · head_mask_tensor / left_eye_mask_tensor / right_eye_mask_tensor [B,1,H,W] ·
        base_iden [B,100] · base_expr [B,79] · base_text [B,100] · base_illu [B,27] · base_gaze_direction [B,2] ·
        cam_info = {batch_Rmats [B,3,3], batch_Tvecs [B,3,1], batch_inv_inmats [B,3,3]}
    (The trainer then overwrites base_expr with its fixed expression code, gazenerf_trainer.py:305-310 -- caller's business.)
    """

    def __init__(self, batch: int, img_size: int, featmap_size: int, device, erode_iterations: int = 2):
        self.B, self.P, self.S, self.device, self.erode = batch, img_size, featmap_size, torch.device(device), erode_iterations
        if self.device.type != "cuda":
            raise RuntimeError("SampleStager stages onto a CUDA device (libgnrf has no CPU path)")
        B, P = batch, img_size
        shapes = {"face_patch": (B, P, P, 3), "head_mask": (B, P, P), "left_eye_mask": (B, P, P), "right_eye_mask": (B, P, P)}
        fshapes = {"latent_codes_row0": (306,), "latent_codes": (B, 306), "pitchyaw_head": (B, 2), "c2w_Rmat": (B, 3, 3), "c2w_Tvec": (B, 3),
                   "inmat": (B, 3, 3)}
        self.pin = {k: torch.empty(s, dtype=torch.uint8).pin_memory() for k, s in shapes.items()}
        self.pin.update({k: torch.empty(s, dtype=torch.float64).pin_memory() for k, s in fshapes.items()})
        self.dev = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in self.pin.items()}

    def h2d_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.pin.values())

    def __call__(self, records: Dict[str, np.ndarray]) -> Dict[str, object]:
        L = _lib.lib()
        B, P = self.B, self.P
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            for k in _U8_KEYS + _F64_KEYS:
                a = np.ascontiguousarray(records[k])
                want = torch.uint8 if k in _U8_KEYS else torch.float64
                t = torch.from_numpy(a)
                if t.dtype != want or tuple(t.shape) != tuple(self.pin[k].shape):
                    raise ValueError("%s: expected %s %s, got %s %s" % (k, want, tuple(self.pin[k].shape), t.dtype, tuple(t.shape)))
                self.pin[k].copy_(t)
                self.dev[k].copy_(self.pin[k], non_blocking=True)
            f = lambda *s: torch.empty(s, device=self.device, dtype=torch.float32)
            img, head, left, right = f(B, 3, P, P), f(B, 1, P, P), f(B, 1, P, P), f(B, 1, P, P)
            d = self.dev
            _lib.check(L.gnrf_sample_images_to_device(d["face_patch"].data_ptr(), d["head_mask"].data_ptr(), d["left_eye_mask"].data_ptr(),
                                                      d["right_eye_mask"].data_ptr(), B, P, P, self.erode, img.data_ptr(), head.data_ptr(),
                                                      left.data_ptr(), right.data_ptr(), st), "gnrf_sample_images_to_device")
            iden, expr, text, illu, gaze = f(B, 100), f(B, 79), f(B, 100), f(B, 27), f(B, 2)
            R, T, Kinv = f(B, 3, 3), f(B, 3, 1), f(B, 3, 3)
            _lib.check(L.gnrf_sample_meta_to_device(d["latent_codes_row0"].data_ptr(), d["latent_codes"].data_ptr(), d["pitchyaw_head"].data_ptr(),
                                                    d["c2w_Rmat"].data_ptr(), d["c2w_Tvec"].data_ptr(), d["inmat"].data_ptr(), B, self.S, P,
                                                    iden.data_ptr(), expr.data_ptr(), text.data_ptr(), illu.data_ptr(), gaze.data_ptr(),
                                                    R.data_ptr(), T.data_ptr(), Kinv.data_ptr(), st), "gnrf_sample_meta_to_device")
        return {"img_tensor": img, "head_mask_tensor": head, "left_eye_mask_tensor": left, "right_eye_mask_tensor": right,
                "base_iden": iden, "base_expr": expr, "base_text": text, "base_illu": illu, "base_gaze_direction": gaze,
                "cam_info": {"batch_Rmats": R, "batch_Tvecs": T, "batch_inv_inmats": Kinv}}


def sample_to_device(records: Dict[str, np.ndarray], featmap_size: int, device, erode_iterations: int = 2,
                     stager: Optional[SampleStager] = None) -> Dict[str, object]:
    """One-shot convenience wrapper around SampleStager (allocates the staging buffers for this batch shape)."""
    B, P = records["face_patch"].shape[0], records["face_patch"].shape[1]
    st = stager or SampleStager(B, P, featmap_size, device, erode_iterations)
    return st(records)
