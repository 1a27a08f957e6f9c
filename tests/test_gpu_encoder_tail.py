"""SURVEY 8f-2 on the GPU: the encoder's tail (encoder.py:133-146: bilinear upsampling of the three latents,
concatenation with upsample_color(images), reduction_layer) evaluated inside the kernels that consume it.

Reference = plain torch fp32 on the same device (TF32 off): F.interpolate(align_corners=True) + cat + conv2d build
pixel_feat_map / holder_feat_map, which then go through the map-based entry points (th_premap_features,
th_paint_group) that are themselves tested against the genuine reference's outputs.  The latents-based entry points
must reproduce them without those maps."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from transhuman_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _tail(V, H, W, seed, scale=1.0):
    g = torch.Generator("cpu").manual_seed(seed)
    lat = [torch.randn((V, c, max(1, (H + d - 1) // d), max(1, (W + d - 1) // d)), generator=g) * scale
           for c, d in ((64, 2), (64, 4), (128, 8))]          # ResNet strides of conv1 / layer1 / layer2
    img = torch.rand((V, 3, H, W), generator=g)
    wc, bc = torch.randn((128, 3, 1, 1), generator=g) * 0.5, torch.randn((128,), generator=g) * 0.1
    wr, br = torch.randn((192, 384, 1, 1), generator=g) * 0.05, torch.randn((192,), generator=g) * 0.1
    return [l.to(DEV) for l in lat], img.to(DEV), wc.to(DEV), bc.to(DEV), wr.to(DEV), br.to(DEV)


def _maps(lat, img, wc, bc, wr, br):
    """encoder.py:133-146 in torch on the device, fp32 without TF32."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        H, W = img.shape[-2:]
        up = [F.interpolate(l, (H, W), mode="bilinear", align_corners=True) for l in lat]
        pixel = torch.cat(up + [F.conv2d(img, wc, bc)], dim=1)
        holder = F.conv2d(pixel, wr, br)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return pixel.contiguous(), holder.contiguous()


@pytest.mark.parametrize("V,H,W", [(3, 64, 64), (2, 40, 56), (1, 21, 21)])
def test_premap_from_latents_matches_premap_of_the_maps(V, H, W):
    fr = synth.make_frame(H=8, W=8, n_class=100, V=V, feat_hw=16, seed=5)
    wts = ops.PackedWeights(fr["weights"], V, device=DEV)
    lat, img, wc, bc, wr, br = _tail(V, H, W, seed=V * 100 + H)
    pixel, _ = _maps(lat, img, wc, bc, wr, br)
    want = ops.premap_features(pixel, wts)
    got = ops.premap_from_latents(ops.EncoderTail(lat, img, wc, bc), wts)
    torch.cuda.synchronize()
    assert got.shape == (V, H, W, 512)
    # same GEMM, same operand split; the A rows differ by the rounding of the interpolation / colour convolution
    assert (got - want).abs().max().item() <= 4e-6 * want.abs().max().item()


def test_paint_group_latents_matches_paint_group_of_the_holder_map():
    V, n_class, hw = 3, 300, 96
    fr = synth.make_frame(H=8, W=8, n_class=n_class, V=V, feat_hw=hw, seed=9)
    lat, img, wc, bc, wr, br = _tail(V, hw, hw, seed=77)
    _, holder = _maps(lat, img, wc, bc, wr, br)
    cl = ops.ClusterIndex(pc2voxel_ind=fr["pc2voxel_ind"], device=DEV)
    g = torch.Generator("cpu").manual_seed(3)
    viz = (torch.rand((V, synth.N_VERTS), generator=g) < 0.6).to(DEV)
    uv = ops.uv_scale_for(hw, hw, hw, hw)
    cams = (_t(fr["input_R"]), _t(fr["input_T"]).reshape(V, 3), _t(fr["input_K"]))
    enc = ops.EncoderTail(lat, img, wc, bc)
    for verts, vz in ((_t(fr["tar_smpl_vertice"]), viz), (_t(fr["tar_smpl_vertice"]) * 3.0, None)):   # second: border
        want = ops.paint_group(holder, uv, verts, *cams, vz, cl)
        got = ops.paint_group_latents(enc, wr, br, uv, verts, *cams, vz, cl)
        torch.cuda.synchronize()
        assert got.shape == (V, n_class, 192)
        assert (got - want).abs().max().item() <= 5e-6 * max(1.0, want.abs().max().item())
    # an all-invisible cluster gives exact zeros like the reference (every painted vertex is 0 there)
    vz = viz.clone()
    members = cl.members_host[cl.start_host[5]:cl.start_host[6]]
    vz[:, torch.from_numpy(members.astype(np.int64)).to(DEV)] = False
    got = ops.paint_group_latents(enc, wr, br, uv, _t(fr["tar_smpl_vertice"]), *cams, vz, cl)
    assert torch.all(got[:, 5] == 0)


def test_encoder_tail_rejects_bad_shapes():
    lat, img, wc, bc, wr, br = _tail(2, 16, 16, seed=1)
    with pytest.raises(AssertionError):
        ops.EncoderTail([lat[0], lat[2], lat[1]], img, wc, bc)
    with pytest.raises(ValueError):
        ops.EncoderTail([l.cpu() for l in lat], img, wc, bc)
