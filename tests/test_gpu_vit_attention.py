"""SURVEY 8f-3 on the GPU: the flash-style attention of the token transformer (th_vit_attention) against the
reference's formula (vision_transformer.py:267-275) evaluated by plain torch in float64 and in float32, and the
plugin's ViT forward (reference modules + this kernel) against the reference module run whole."""
import pytest
import torch

from transhuman_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_attention(qkv, H, scale, dtype):
    B, N, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.to(dtype).reshape(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B, N, C)


# (3, 6000) and (3, 4321) fill the machine with pairs of query tiles (k_attn_tc<2>); the others run k_attn_tc<1>
@pytest.mark.parametrize("B,N", [(3, 300), (1, 1), (2, 129), (3, 1500), (1, 64), (2, 193), (1, 128), (1, 65),
                                 (3, 6000), (3, 4321)])
def test_vit_attention_matches_float64(B, N):
    H, D = 3, 64
    g = torch.Generator("cpu").manual_seed(N)
    qkv = (torch.randn((B, N, 3 * H * D), generator=g) * 1.5).to(DEV)
    scale = D ** -0.5
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        want = _ref_attention(qkv, H, scale, torch.float64)
        f32 = _ref_attention(qkv, H, scale, torch.float32)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    got = ops.vit_attention(qkv, H, scale)
    torch.cuda.synchronize()
    assert got.shape == (B, N, H * D)
    err = (got.double() - want).abs().max().item()
    err32 = (f32.double() - want).abs().max().item()
    print(f"[vit_attention] B={B} N={N}: max-abs vs float64 {err:.2e} (torch fp32: {err32:.2e})")
    # hi/lo operands carry 22 bits: the same accuracy class as torch's own fp32 evaluation
    assert err <= max(3e-6, 4 * err32)


def test_vit_attention_peaked_softmax_and_large_logits():
    """Rows whose softmax is a one-hot (logit gaps of hundreds) and large |q k|: the online rescaling and the
    exp2 of the shifted logits must stay exact where the reference's softmax is."""
    H, D, B, N = 3, 64, 1, 200
    g = torch.Generator("cpu").manual_seed(7)
    qkv = torch.randn((B, N, 3 * H * D), generator=g)
    qkv[:, :, :H * D] *= 12.0           # q
    qkv[:, :, H * D:2 * H * D] *= 6.0   # k
    qkv = qkv.to(DEV)
    scale = D ** -0.5
    want = _ref_attention(qkv, H, scale, torch.float64)
    got = ops.vit_attention(qkv, H, scale)
    assert torch.isfinite(got).all()
    # logits ~ +-600: one ulp of the fp32 logit (6e-5) moves a softmax weight by 6e-5 relative
    assert (got.double() - want).abs().max().item() <= 5e-4


def test_vit_attention_rejects_other_head_dims():
    qkv = torch.zeros((1, 8, 3 * 3 * 32), device=DEV)
    with pytest.raises(Exception):
        ops.vit_attention(qkv, 3, 1.0)


@pytest.mark.parametrize("n_out,n_in", [(576, 192), (192, 192), (768, 192), (192, 768), (64, 64), (260, 128)])
@pytest.mark.parametrize("rows", [1, 300, 5000])
def test_tc_linear_matches_float64(n_out, n_in, rows):
    """th_linear (the ViT's Linear layers on the tcgen05 GEMM, fp16 hi/lo three-product scheme) against float64:
    chunking of n_out into 256 / 128 columns with a zero-padded, column-guarded last chunk; bias; row guard."""
    g = torch.Generator("cpu").manual_seed(n_out * 7 + rows)
    w = torch.randn((n_out, n_in), generator=g) * 0.05
    b = torch.randn((n_out,), generator=g) * 0.1
    x = torch.randn((rows, n_in), generator=g).to(DEV)
    lin = ops.PackedLinear(w, b, device=DEV)
    guard = torch.full((rows + 1, n_out), 7.0, device=DEV)          # nothing may be written past the last row
    y = lin(x)
    want = x.double() @ w.double().t().to(DEV) + b.double().to(DEV)
    err = (y.double() - want).abs().max().item()
    tol = 1e-5 * max(1.0, want.abs().max().item())   # 22-bit operands, fp32 accumulation over up to 768 terms
    assert y.shape == (rows, n_out) and err <= tol, err
    y3 = lin(x.reshape(1, rows, n_in), relu=True)
    assert y3.shape == (1, rows, n_out) and (y3.double() - want.clamp(min=0)[None]).abs().max().item() <= tol
    nob = ops.PackedLinear(w, None, device=DEV)(x)
    assert (nob.double() - (want - b.double().to(DEV))).abs().max().item() <= tol
    assert bool((guard == 7.0).all())
