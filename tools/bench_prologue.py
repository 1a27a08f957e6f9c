#!/usr/bin/env python
"""CUDA-event timing of the per-frame prologue kernels at the C2 frame size (V = 3 views of 512 x 512; ResNet-18
latents 256^2 / 128^2 / 64^2), map-based against latents-based entry points, next to the torch ops of the encoder's
tail they replace (encoder.py:133-146).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F


QUICK = "--quick" in sys.argv      # one warm-up + one timed call per op (for an ncu launch list)


def timed(fn, iters=20, warm=3):
    if QUICK:
        iters, warm = 1, 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    import __graft_entry__ as entry
    entry.build()
    from transhuman_b200 import ops, synth
    dev = "cuda:0"
    V, H = 3, 512
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator("cpu").manual_seed(0)
    lat = [torch.randn((V, c, H // d, H // d), generator=g).to(dev).contiguous(memory_format=torch.channels_last)
           for c, d in ((64, 2), (64, 4), (128, 8))]
    img = torch.rand((V, 3, H, H), generator=g).to(dev)
    wc, bc = (torch.randn((128, 3, 1, 1), generator=g) * 0.5).to(dev), torch.randn((128,), generator=g).to(dev)
    wr, br = (torch.randn((192, 384, 1, 1), generator=g) * 0.05).to(dev), torch.randn((192,), generator=g).to(dev)
    res = {}

    def torch_tail():
        up = [F.interpolate(l, (H, H), mode="bilinear", align_corners=True) for l in lat]
        pixel = torch.cat(up + [F.conv2d(img, wc, bc)], dim=1)
        return pixel, F.conv2d(pixel, wr, br)

    res["torch_encoder_tail_ms"] = timed(torch_tail, iters=10)
    pixel, holder = torch_tail()
    fr = synth.make_frame(H=8, W=8, n_class=300, V=V, feat_hw=16, seed=5)
    wts = ops.PackedWeights(fr["weights"], V, device=dev)
    enc = ops.EncoderTail(lat, img, wc, bc)
    out = torch.empty((V, H, H, 512), device=dev)
    res["premap_features_ms"] = timed(lambda: ops.premap_features(pixel, wts, out=out))
    a = out.clone()
    res["premap_from_latents_ms"] = timed(lambda: ops.premap_from_latents(enc, wts, out=out))
    res["premap_max_abs_diff"] = float((a - out).abs().max())
    res["premap_scale"] = float(a.abs().max())

    def t(x):
        return torch.from_numpy(np.ascontiguousarray(x)).to(dev)

    for n_tok in (300, 1500, 6000):
        frt = synth.make_frame(H=8, W=8, n_class=n_tok, V=V, feat_hw=16, seed=5, with_feature_maps=False)
        cl = ops.ClusterIndex(pc2voxel_ind=frt["pc2voxel_ind"], device=dev)
        uv = ops.uv_scale_for(H, H, H, H)
        cams = (t(frt["input_R"]), t(frt["input_T"]).reshape(V, 3), t(frt["input_K"]))
        verts = t(frt["tar_smpl_vertice"])
        viz = (torch.rand((V, synth.N_VERTS), generator=g) < 0.6).to(dev)
        res[f"paint_group_{n_tok}_ms"] = timed(lambda: ops.paint_group(holder, uv, verts, *cams, viz, cl))
        res[f"paint_group_latents_{n_tok}_ms"] = timed(
            lambda: ops.paint_group_latents(enc, wr, br, uv, verts, *cams, viz, cl))
        d = (ops.paint_group(holder, uv, verts, *cams, viz, cl) -
             ops.paint_group_latents(enc, wr, br, uv, verts, *cams, viz, cl)).abs().max()
        res[f"paint_group_{n_tok}_max_abs_diff"] = float(d)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
