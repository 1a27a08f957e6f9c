#!/usr/bin/env python
"""Small invocations of the prologue kernels (encoder tail from the latents, paint + group, flash attention in both
its variants) for compute-sanitizer; checks each against torch so that a sanitizer-instrumented run is also a
correctness run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F


def main():
    import __graft_entry__ as entry
    entry.build()
    from transhuman_b200 import ops, synth
    dev = "cuda:0"
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator("cpu").manual_seed(0)
    V, H, W = 2, 40, 56
    lat = [torch.randn((V, c, (H + d - 1) // d, (W + d - 1) // d), generator=g).to(dev) for c, d in ((64, 2), (64, 4), (128, 8))]
    img = torch.rand((V, 3, H, W), generator=g).to(dev)
    wc, bc = (torch.randn((128, 3, 1, 1), generator=g) * 0.5).to(dev), torch.randn((128,), generator=g).to(dev)
    wr, br = (torch.randn((192, 384, 1, 1), generator=g) * 0.05).to(dev), torch.randn((192,), generator=g).to(dev)
    up = [F.interpolate(l, (H, W), mode="bilinear", align_corners=True) for l in lat]
    pixel = torch.cat(up + [F.conv2d(img, wc, bc)], dim=1).contiguous()
    holder = F.conv2d(pixel, wr, br).contiguous()
    fr = synth.make_frame(H=8, W=8, n_class=100, V=V, feat_hw=16, seed=5)
    wts = ops.PackedWeights(fr["weights"], V, device=dev)
    enc = ops.EncoderTail(lat, img, wc, bc)
    a, b = ops.premap_features(pixel, wts), ops.premap_from_latents(enc, wts)
    print("[sanitize] premap_from_latents vs premap_features:", float((a - b).abs().max()))
    assert float((a - b).abs().max()) <= 4e-6 * float(a.abs().max())

    def t(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    cl = ops.ClusterIndex(pc2voxel_ind=fr["pc2voxel_ind"], device=dev)
    uv = ops.uv_scale_for(H, W, H, W)
    cams = (t(fr["input_R"]), t(fr["input_T"]).reshape(V, 3), t(fr["input_K"]))
    viz = (torch.rand((V, synth.N_VERTS), generator=g) < 0.6).to(dev)
    pa = ops.paint_group(holder, uv, t(fr["tar_smpl_vertice"]), *cams, viz, cl)
    pb = ops.paint_group_latents(enc, wr, br, uv, t(fr["tar_smpl_vertice"]), *cams, viz, cl)
    print("[sanitize] paint_group_latents vs paint_group:", float((pa - pb).abs().max()))
    assert float((pa - pb).abs().max()) <= 5e-6 * max(1.0, float(pa.abs().max()))
    for B, N in ((1, 200), (10, 4100)):        # k_attn_tc<1>, k_attn_tc<2> (>= 148 CTAs of query-tile pairs)
        qkv = torch.randn((B, N, 576), generator=g).to(dev)
        got = ops.vit_attention(qkv, 3, 0.125)
        q, k, v = qkv.reshape(B, N, 3, 3, 64).permute(2, 0, 3, 1, 4)
        want = ((q @ k.transpose(-2, -1)) * 0.125).softmax(dim=-1) @ v
        want = want.transpose(1, 2).reshape(B, N, 192)
        print(f"[sanitize] vit_attention B={B} N={N}:", float((got - want).abs().max()))
        assert float((got - want).abs().max()) <= 2e-5
    torch.cuda.synchronize()
    print("[sanitize] ok")


if __name__ == "__main__":
    main()
