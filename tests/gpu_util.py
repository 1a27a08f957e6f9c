"""Helpers shared by the GPU parity tests, smoke() and bench.py: move a synthetic
frame (numpy, transhuman_b200.synth) to the device as an ops.Frame."""
from __future__ import annotations

import numpy as np
import torch

from transhuman_b200 import ops


def frame_to_device(fr: dict, tokens, device="cuda:0", simt_mlp: bool = False, white_bkgd: bool = False,
                    weights=None, premapped=None):
    """tokens = (tok_xyz (N_c,3) fp32, tok_blend (N_c,4,4)) as torch CPU tensors.
    Returns (ops.Frame, (ray_o, ray_d, near, far) on the device).  ``premapped=None`` picks the path the
    Renderer plugin uses: pre-mapped feature maps (th_premap_features) wherever the layer-chained schedule
    exists (V <= 3, tensor cores), the plain channel-last maps otherwise."""
    if premapped is None:
        premapped = (not simt_mlp) and fr["V"] <= 3
    dev = torch.device(device)
    tok_xyz, tok_blend = tokens

    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    V = fr["V"]
    if weights is None:
        weights = ops.PackedWeights(fr["weights"], V, device=dev)
    feat = (ops.premap_features(t(fr["pixel_feat_map"]), weights) if premapped
            else ops.nchw_to_nhwc(t(fr["pixel_feat_map"])))
    hw = fr["feat_hw"]
    frame = ops.Frame(
        holder=t(fr["holder"]), tok_xyz=tok_xyz.float().to(dev),
        tok_rot=tok_blend[:, :3, :3].float().contiguous().to(dev), verts=t(fr["tar_smpl_vertice"]),
        feat_nhwc=feat, cam_R=t(fr["input_R"]), cam_T=t(fr["input_T"]).reshape(V, 3), cam_K=t(fr["input_K"]),
        Rh=t(fr["Rh"]), Th=t(fr["Th"]).reshape(3), weights=weights,
        uv_scale=ops.uv_scale_for(hw, hw, hw, hw), simt_mlp=simt_mlp, white_bkgd=white_bkgd, premapped=premapped)
    rays = (t(fr["ray_o"]), t(fr["ray_d"]), t(fr["near"]), t(fr["far"]))
    return frame, rays


def assert_maps_close(got: dict, want: dict, oracle_raw, z_vals, ray_d, S: int, far: float, name: str,
                      white_bkgd: bool = False, tol: float = 1e-4):
    """rgb / acc <= tol max-abs, depth <= tol * far against ``want`` (golden or oracle maps, (1,N,.)).

    Knife edge: the last sample's interval is 1e10, so alpha_S is a step function of
    sign(alpha_raw_S) (nerf_net_utils.py:31-34).  A ray whose ORACLE alpha_raw_S lies within the raw tolerance
    (2e-5 x the frame's raw scale) of 0 may legitimately land on the other side of the step; such a ray is not
    skipped: it must match the oracle composite under one of the two step hypotheses (last sample fully opaque /
    fully transparent).  The count is printed.  ``got`` must carry ``raw`` (want_raw=True): in the progressive
    branch the oracle never evaluated the colour of a sample with alpha_raw <= 0, so the opaque hypothesis takes
    that one sample's colour from the GPU."""
    from oracle import transhuman_oracle as orc
    raw = oracle_raw.reshape(-1, S, 4).clone().float()
    N = raw.shape[0]
    scale = max(1.0, float(raw.abs().max()))
    eps = 2e-5 * scale
    a_last = raw[:, -1, 3]
    edge = (a_last.abs() < eps) & (a_last != 0)          # raw == 0 exactly is a masked-out sample, not an edge
    g = {k: got[k].detach().cpu().reshape(N, -1) for k in ("rgb_map", "acc_map", "depth_map")}
    w = {k: want[k].detach().cpu().reshape(N, -1) for k in ("rgb_map", "acc_map", "depth_map")}
    tols = {"rgb_map": tol, "acc_map": tol, "depth_map": tol * far}

    def ok_rows(a: dict, b: dict):
        good = torch.ones(a["rgb_map"].shape[0], dtype=torch.bool)
        for k, t in tols.items():
            good &= (a[k] - b[k]).abs().amax(dim=1) <= t
        return good

    good = ok_rows(g, w)
    worst = {k: float((g[k] - w[k])[~edge].abs().max()) if (~edge).any() else 0.0 for k in tols}
    assert bool(good[~edge].all()), f"{name}: {int((~good & ~edge).sum())} rays off; worst {worst}"
    n_edge = int(edge.sum())
    if n_edge:
        ge = {k: v[edge] for k, v in g.items()}
        matched = good[edge].clone()
        graw = got["raw"].detach().cpu().reshape(N, S, 4)[edge]
        for sign in (1.0, -1.0):
            r = raw[edge].clone()
            r[:, -1, 3] = sign * max(eps, 1e-6)
            if sign > 0:
                unset = (r[:, -1, :3] == 0).all(dim=1)
                r[unset, -1, :3] = graw[unset, -1, :3]
            rgb, acc, _, depth = orc.raw2outputs(r, z_vals.reshape(-1, S)[edge], ray_d.reshape(-1, 3)[edge], white_bkgd)
            matched |= ok_rows(ge, {"rgb_map": rgb, "acc_map": acc[:, None], "depth_map": depth[:, None]})
        assert bool(matched.all()), f"{name}: {int((~matched).sum())} knife-edge rays match neither step hypothesis"
    print(f"[knife-edge] {name}: {n_edge} of {N} rays have |alpha_raw_S| < {eps:.1e}; "
          f"all matched under a step hypothesis; worst elsewhere {worst}")
    return n_edge
