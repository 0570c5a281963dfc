"""Dataset sample -> device tensors (gazenerf_b200/data.py, csrc/data.cu) against the oracle restatement of
datasets/eth_xgaze.py:308-360 + trainer/gazenerf_trainer.py:250-337.  Byte / integer work (pixels, masks, erosion): bit-exact;
f64 -> f32 parameter casts: bit-exact (same rounding)."""
import numpy as np
import pytest
import torch

from oracle import gazenerf_oracle as O


def test_oracle_sample_transform_against_the_library_ops_the_reference_calls():
    """CPU: the oracle's image transform equals torchvision's ToPILImage -> ToTensor (datasets/eth_xgaze.py:12), its erosion is
    cv2.erode(iterations=2) == one 5x5 minimum that ignores out-of-image taps, and the intrinsics inverse really inverts."""
    from torchvision import transforms
    rec = O.synthetic_hdf5_records(2, size=64, seed=3)
    out = O.dataset_sample_tensors(rec, featmap_size=8)
    trans = transforms.Compose([transforms.ToPILImage(), transforms.ToTensor()])
    for i in range(2):
        assert torch.equal(out["img_tensor"][i], trans(rec["face_patch"][i][:, :, [2, 1, 0]]))
    m = rec["head_mask"][0].astype(np.int32)
    pad = np.pad(m, 2, constant_values=255)
    win = np.stack([pad[dy:dy + 64, dx:dx + 64] for dy in range(5) for dx in range(5)]).min(0)
    assert np.array_equal(out["head_mask_tensor"][0, 0].numpy(), win.astype(np.uint8))
    assert 0 < int((out["head_mask_tensor"] > 0).sum()) < int((torch.from_numpy(rec["head_mask"]) > 0).sum())   # it eroded something
    k = rec["inmat"].copy()
    k[:, :2, :] *= 8 / 64
    assert np.allclose(np.einsum("bij,bjk->bik", out["cam_info"]["batch_inv_inmats"].double().numpy(), k), np.eye(3)[None], atol=1e-6)
    assert out["base_iden"].shape == (2, 100) and out["base_expr"].shape == (2, 79) and out["base_illu"].shape == (2, 27)
    assert np.array_equal(out["base_illu"].numpy(), rec["latent_codes"][:, 279:].astype(np.float32))
    assert np.array_equal(out["base_iden"][1].numpy(), rec["latent_codes_row0"][:100].astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("B,size", [(2, 512), (3, 64)])
def test_sample_to_device_vs_oracle(B, size):
    import gazenerf_b200 as G
    dev = torch.device("cuda:0")
    rec = O.synthetic_hdf5_records(B, size=size, seed=B)
    ref = O.dataset_sample_tensors(rec, featmap_size=size // 8)
    stager = G.SampleStager(B, size, size // 8, dev)
    for _ in range(2):   # staging buffers are re-used
        got = stager(rec)
    torch.cuda.synchronize()
    assert stager.h2d_bytes() < 0.3 * sum(v.numel() * 4 for k, v in got.items() if torch.is_tensor(v))   # raw records, not fp32 tensors
    for k in ("img_tensor", "base_iden", "base_expr", "base_text", "base_illu", "base_gaze_direction"):
        assert got[k].dtype == torch.float32 and torch.equal(got[k].cpu(), ref[k]), k
    for k in ("head_mask_tensor", "left_eye_mask_tensor", "right_eye_mask_tensor"):
        assert got[k].shape == ref[k].shape and torch.equal(got[k].cpu(), ref[k].float()), k     # mask VALUES kept (0 / 255)
    for k, v in ref["cam_info"].items():
        assert got["cam_info"][k].shape == v.shape and torch.equal(got["cam_info"][k].cpu(), v), k
    # the tensors are what the drop-in loss consumes: thresholds at 0.5 behave like the reference's `mask >= 0.5`
    assert torch.equal((got["head_mask_tensor"] >= 0.5).cpu(), ref["head_mask_tensor"] >= 0.5)
