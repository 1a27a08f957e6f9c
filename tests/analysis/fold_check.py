#!/usr/bin/env python
"""Round-2 groundwork, CPU only: how far does "transform, then interpolate"
move the results?  alpha_res_0 / rgb_res_0 / rgb_res_1 are 1x1 convolutions of a bilinear blend of
feature-map rows (cross_transformer.py:315, 333, 343); here they are applied ONCE to the whole
(V,384,H,W) maps and the transformed maps are sampled instead.  Exact in real arithmetic; this prints
the fp32 deviation of raw and of the composited image against the oracle on a small synthetic frame."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import transhuman_oracle as orc  # noqa: E402  (checker only)
from transhuman_b200 import synth  # noqa: E402


def main(H=24, W=24, S=16, seed=11):
    fr = synth.make_frame(H=H, W=W, n_class=300, V=3, feat_hw=32, seed=seed, alpha_bias_shift=-15.0)
    tf = orc.to_torch_frame(fr)
    tokens = orc.build_tokens(tf)
    want = orc.render(tf, S, tokens=tokens)
    w = tf["weights"]
    fmap = tf["pixel_feat_map"]                                   # (V,384,h,w)

    def conv_map(name):                                           # 1x1 conv of the whole map, bias kept aside
        return F.conv2d(fmap, w[name + ".weight"][:, :, None, None])

    maps = {n: conv_map(n) for n in ("alpha_res_0", "rgb_res_0", "rgb_res_1")}
    image_shape = fmap.shape[-2:]
    ray_o, ray_d = tf["ray_o"][None], tf["ray_d"][None]
    pts, z_vals = orc.get_sampling_points(ray_o, ray_d, tf["near"][None], tf["far"][None], S)
    xyz = pts.clone().flatten(1, 2)
    pts_s = orc.world2smpl(pts, tf["Rh"][None], tf["Th"][None]).flatten(1, 2)
    viewdir = orc.view_embed(ray_d)[:, :, None].repeat(1, 1, S, 1).contiguous().view(1, -1, 27)
    tok_xyz, tok_blend = tokens
    V = fmap.shape[0]

    def sample(m):
        return orc.get_pixel_aligned_feature(xyz, tf["input_R"], tf["input_T"], tf["input_K"], m, image_shape)

    b = lambda n: w[n + ".bias"][None, :, None]
    rep = orc.human_representation(pts_s[0], tok_xyz, tok_blend, tf["holder"], K=7)
    net_ske = F.relu(orc._conv(w, "fc_0", rep))
    net_pix = F.relu(sample(maps["alpha_res_0"]) + b("alpha_res_0"))
    net = orc.cross_attention(w, net_ske, net_pix)
    net = F.relu(orc._conv(w, "fc_1", net))
    inter = F.relu(orc._conv(w, "fc_2", net))
    alpha = orc.alpha_forward(w, inter, V)
    feats = orc._conv(w, "feature_fc", inter) + sample(maps["rgb_res_0"]) + b("rgb_res_0")
    vd = viewdir.unsqueeze(1).expand(-1, V, *viewdir.shape[1:]).reshape(-1, *viewdir.shape[1:]).transpose(1, 2)
    n2 = F.relu(orc._conv(w, "view_fc", torch.cat((feats, vd), dim=1)))
    n2 = n2 + sample(maps["rgb_res_1"]) + b("rgb_res_1")
    n2 = n2.reshape(-1, V, *n2.shape[1:]).mean(dim=1)
    rgb = orc._conv(w, "rgb_fc", F.relu(orc._conv(w, "fc_4", n2)))
    raw = torch.cat((rgb, alpha), dim=1).transpose(1, 2).reshape(-1, S, 4)
    rgb_map, acc, _, depth = orc.raw2outputs(raw, z_vals.view(-1, S), ray_d.view(-1, 3), False)
    d_raw = (raw - want["raw"]).abs()
    rel = (d_raw / want["raw"].abs().clamp_min(1.0)).max().item()
    print(f"frame {H}x{W}x{S}: raw max-abs {d_raw.max().item():.3e} (rel. to max(1,|raw|) {rel:.3e}; bar 2e-5), "
          f"rgb_map {(rgb_map - want['rgb_map'][0]).abs().max().item():.3e}, "
          f"acc_map {(acc - want['acc_map'][0]).abs().max().item():.3e} (bar 1e-4)")


if __name__ == "__main__":
    main()
    main(H=16, W=16, S=32, seed=3)
