"""Pre-split bf16-plane GEMM kernels of the training path (csrc/lin_hl.cu) against fp64 products of the SAME operands, through the C ABI.

What they replace: nn.Conv2d(k=1) forward / autograd backward of models/mlp_nerf.py:95-119 on channel-major activations.
Tolerances (rel-L2 of the whole tensor):
  planes = 2 (bf16x3: hi*hi + lo*hi + hi*lo, fp32 accumulate; outputs re-split to 16 significand bits)   3e-5
  planes = 1 (single-pass bf16 on the hi planes; outputs rounded to bf16)                                   4e-3 for plane outputs
                                                                                                          2e-5 for fp32 outputs
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _planes(x: torch.Tensor, planes: int) -> torch.Tensor:
    """fp32 [img][rows][HW] -> bf16 [planes][img][rows][HW] (hi | lo)."""
    hi = x.to(torch.bfloat16)
    if planes == 1:
        return hi[None].contiguous()
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def _value(pl: torch.Tensor) -> torch.Tensor:
    return pl.double().sum(0)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _w_eff(W: torch.Tensor, planes: int) -> torch.Tensor:
    hi = W.to(torch.bfloat16).float()
    if planes == 1:
        return hi.double()
    lo = (W - hi).to(torch.bfloat16).float()
    return hi.double() + lo.double()


def _bits(x: torch.Tensor) -> torch.Tensor:
    """fp32 [img][rows][HW] -> int32 [img][rows][HW/32]: bit (p % 32) of word p / 32 = x > 0."""
    b = (x > 0).reshape(*x.shape[:-1], x.shape[-1] // 32, 32).to(torch.int64)
    w = (b << torch.arange(32, device=x.device)).sum(-1)
    return (w - ((w >> 31) << 32)).to(torch.int32).contiguous()


def _lin(L, W, bias, x_pl, planes, n_img, HW, act=0, bias_img=None, hl_rows=None, mask_bits=None, mask_rows=0, transposed=False,
         sign_out=None, act_rows=0):
    from gazenerf_b200 import _lib
    dev = x_pl.device
    st = torch.cuda.current_stream().cuda_stream
    N, K = (W.shape[1], W.shape[0]) if transposed else W.shape
    pk = torch.empty((L.gnrf_lin_hl_packed_bytes(N, K, planes),), device=dev, dtype=torch.uint8)
    _lib.check(L.gnrf_lin_hl_pack(W.data_ptr(), bias.data_ptr() if bias is not None else None, N, K, 1 if transposed else 0, planes,
                                  pk.data_ptr(), st), "pack")
    hl_rows = N if hl_rows is None else hl_rows
    out = torch.full((planes, n_img, max(hl_rows, 1), HW), float("nan"), device=dev, dtype=torch.bfloat16)
    out32 = torch.full((n_img, max(N - hl_rows, 1), HW), float("nan"), device=dev, dtype=torch.float32)
    K_rows = x_pl.shape[2]
    _lib.check(L.gnrf_lin_hl(pk.data_ptr(), N, K, planes, x_pl.data_ptr(), K_rows * HW, x_pl.stride(0),
                             bias_img.data_ptr() if bias_img is not None else None, act, act_rows,
                             out.data_ptr(), out.shape[2] * HW, out.stride(0), hl_rows, out32.data_ptr(), out32.shape[1] * HW,
                             mask_bits.data_ptr() if mask_bits is not None else None,
                             mask_bits.shape[1] * (HW // 32) if mask_bits is not None else 0, mask_rows,
                             sign_out.data_ptr() if sign_out is not None else None,
                             sign_out.shape[1] * (HW // 32) if sign_out is not None else 0, sign_out.shape[1] if sign_out is not None else 0,
                             n_img, HW, st), "gnrf_lin_hl")
    torch.cuda.synchronize()
    return out, out32


@pytest.mark.parametrize("planes", [2, 1])
@pytest.mark.parametrize("N,K,HW,n_img,act", [(384, 384, 512, 2, 1), (384, 63, 256, 1, 1), (384, 447, 768, 1, 1), (192, 384, 256, 2, 1),
                                               (160, 96, 1024, 1, 0),
                                               # edges: one tile per image and more images than M-tiles; the smallest K (one stage);
                                               # K = 33 (second K-block holds ONE valid row, the rest zero-filled by the TMA unit);
                                               # N = 100 (rows 100..127 of the only M-tile clipped by the tensor store); 5 M-tiles
                                               (128, 64, 256, 3, 1), (128, 32, 256, 1, 0), (256, 33, 512, 1, 1), (100, 128, 256, 2, 0),
                                               (600, 64, 256, 1, 1)])
def test_lin_hl_forward(planes, N, K, HW, n_img, act):
    from gazenerf_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(N * 7 + K)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = (0.1 * torch.randn(N, generator=g)).to(dev)
    bimg = (0.1 * torch.randn(n_img, N, generator=g)).to(dev)
    x = torch.randn(n_img, K, HW, generator=g).to(dev)
    x_pl = _planes(x, planes)
    sign = torch.full((n_img, N, HW // 32), 0x55555555, device=dev, dtype=torch.int32)
    out, _ = _lin(L, W, b, x_pl, planes, n_img, HW, act=act, bias_img=bimg, sign_out=sign)
    ref = torch.einsum("nk,ikp->inp", _w_eff(W, planes), _value(x_pl)) + b.double()[None, :, None] + bimg.double()[:, :, None]
    if act:
        ref = ref.clamp_min(0)
    assert torch.isfinite(out.float()).all()
    assert torch.equal(sign, _bits(out[0].float()))   # the sign bits describe exactly what was stored
    err = _rel(_value(out), ref)
    assert err < (3e-5 if planes == 2 else 4e-3), err
    if planes == 2:   # the lo plane is the bf16 rounding of the residual: |lo| <= ulp(hi)/2
        hi, lo = out[0].float(), out[1].float()
        assert bool((lo.abs() <= hi.abs() * 2 ** -8 + 1e-30).all())


@pytest.mark.parametrize("planes", [2, 1])
def test_lin_hl_fp32_rows_and_mask(planes):
    """RGB_layer_0 + density row (385 outputs: rows < 384 planes, row 384 fp32) and the skip layer's input gradient
    (447 outputs: rows < 384 masked by the saved activation's sign and written as planes, 63 positional-encoding rows fp32)."""
    from gazenerf_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(5)
    HW, n_img = 512, 2
    W = (torch.randn(385, 384, generator=g) / 20).to(dev)
    b = (0.1 * torch.randn(385, generator=g)).to(dev)
    x_pl = _planes(torch.randn(n_img, 384, HW, generator=g).to(dev), planes)
    out, out32 = _lin(L, W, b, x_pl, planes, n_img, HW, act=0, hl_rows=384)
    ref = torch.einsum("nk,ikp->inp", _w_eff(W, planes), _value(x_pl)) + b.double()[None, :, None]
    assert _rel(_value(out), ref[:, :384]) < (3e-5 if planes == 2 else 4e-3)
    assert _rel(out32, ref[:, 384:]) < 2e-5
    # folded RGB head: 193 fp32 outputs, ReLU on the first 192 rows only (row 192 = raw density)
    _, o193 = _lin(L, W[:193].contiguous(), b[:193].contiguous(), x_pl, planes, n_img, HW, act=1, hl_rows=0, act_rows=192)
    r193 = ref[:, :193].clone()
    r193[:, :192] = r193[:, :192].clamp_min(0)
    assert _rel(o193, r193) < 2e-5 and float(o193[:, 192].min()) < 0
    # all-fp32 output (hl_rows = 0)
    _, o32 = _lin(L, W[:192].contiguous(), b[:192].contiguous(), x_pl, planes, n_img, HW, act=1, hl_rows=0)
    assert _rel(o32, ref[:, :192].clamp_min(0)) < 2e-5
    # input gradient of the skip layer: forward weight [384][447] used transposed, mask on the first 384 output rows
    Wf = (torch.randn(384, 447, generator=g) / 20).to(dev)
    gy_pl = _planes(torch.randn(n_img, 384, HW, generator=g).to(dev), planes)
    saved = torch.randn(n_img, 448, HW, generator=g).clamp_min(0).to(dev)
    saved[:, :, ::7] = 0.0
    out, out32 = _lin(L, Wf, None, gy_pl, planes, n_img, HW, hl_rows=384, mask_bits=_bits(saved), mask_rows=384, transposed=True)
    ref = torch.einsum("nk,inp->ikp", _w_eff(Wf, planes), _value(gy_pl))
    ref_m = ref[:, :384] * (saved[:, :384] > 0)
    assert _rel(_value(out), ref_m) < (3e-5 if planes == 2 else 4e-3)
    assert bool((_value(out)[saved[:, :384] <= 0] == 0).all())
    assert _rel(out32, ref[:, 384:]) < 2e-5


@pytest.mark.parametrize("planes", [2, 1])
@pytest.mark.parametrize("N,K,HW,n_img,db_sum", [(384, 384, 2048, 2, 1), (384, 447, 1024, 1, 0), (384, 63, 512, 2, 0), (192, 384, 1024, 2, 0),
                                                  (385, 384, 4096, 1, 1),
                                                  # edges: a single 32-point K-block per image (split-K clamps to 1); one X row; 193 rows
                                                  # (the folded RGB head); three column chunks
                                                  (128, 128, 32, 2, 0), (128, 1, 256, 1, 1), (193, 384, 512, 2, 0), (128, 800, 256, 1, 1)])
def test_wgrad_hl(planes, N, K, HW, n_img, db_sum):
    from gazenerf_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cpu").manual_seed(N + K + HW)
    dy_pl = _planes(torch.randn(n_img, N, HW, generator=g).to(dev), planes)
    x_pl = _planes(torch.randn(n_img, K, HW, generator=g).to(dev), planes)
    ws = torch.empty((L.gnrf_wgrad_hl_workspace_bytes(N, K, n_img, HW),), device=dev, dtype=torch.uint8)
    dW = torch.full((N, K), float("nan"), device=dev)
    db = torch.full((N,) if db_sum else (n_img, N), float("nan"), device=dev)
    _lib.check(L.gnrf_wgrad_hl(dy_pl.data_ptr(), N * HW, dy_pl.stride(0), x_pl.data_ptr(), K * HW, x_pl.stride(0), planes, N, K, n_img, HW,
                               dW.data_ptr(), db.data_ptr(), db_sum, ws.data_ptr(), ws.numel(), st), "gnrf_wgrad_hl")
    torch.cuda.synchronize()
    dyv, xv = _value(dy_pl), _value(x_pl)
    ref = torch.einsum("inp,ikp->nk", dyv, xv)
    ref_b = dyv.sum(2)
    tol = 3e-5 if planes == 2 else 2e-5   # planes = 1: exact bf16 products, fp32 accumulation
    assert _rel(dW, ref) < tol, _rel(dW, ref)
    assert _rel(db, ref_b.sum(0) if db_sum else ref_b) < tol


def test_lin_hl_strided_views():
    """Operands addressed as row windows of a larger plane tensor (buf0 = [layer-4 output | positional encoding]) and written into one."""
    from gazenerf_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cpu").manual_seed(11)
    HW, n_img, H, PE = 512, 2, 384, 63
    buf = torch.zeros(2, n_img, H + 64, HW, device=dev, dtype=torch.bfloat16)
    pe = torch.randn(n_img, PE, HW, generator=g).to(dev)
    buf[:, :, H:H + PE] = _planes(pe, 2)
    buf[:, :, H + PE] = float("nan")   # the pad row is never read (the tensor map ends at K rows)
    W = (torch.randn(H, PE, generator=g) / 8).to(dev)
    pk = torch.empty((L.gnrf_lin_hl_packed_bytes(H, PE, 2),), device=dev, dtype=torch.uint8)
    _lib.check(L.gnrf_lin_hl_pack(W.data_ptr(), None, H, PE, 0, 2, pk.data_ptr(), st), "pack")
    x_ptr = buf.data_ptr() + H * HW * 2
    _lib.check(L.gnrf_lin_hl(pk.data_ptr(), H, PE, 2, x_ptr, (H + 64) * HW, buf.stride(0), None, 1, 0, buf.data_ptr(), (H + 64) * HW,
                             buf.stride(0), H, None, 0, None, 0, 0, None, 0, 0, n_img, HW, st), "gnrf_lin_hl")
    torch.cuda.synchronize()
    ref = torch.einsum("nk,ikp->inp", _w_eff(W, 2), _value(_planes(pe, 2))).clamp_min(0)
    assert _rel(_value(buf[:, :, :H]), ref) < 3e-5
    assert _rel(_value(buf[:, :, H:H + PE]), pe) < 1e-5
