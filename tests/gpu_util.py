"""Helpers shared by the GPU parity tests, smoke() and bench.py: move a synthetic
frame (numpy, transhuman_b200.synth) to the device as an ops.Frame."""
from __future__ import annotations

import numpy as np
import torch

from transhuman_b200 import ops


def frame_to_device(fr: dict, tokens, device="cuda:0", simt_mlp: bool = False, white_bkgd: bool = False,
                    weights=None, premapped: bool = False):
    """tokens = (tok_xyz (N_c,3) fp32, tok_blend (N_c,4,4)) as torch CPU tensors.
    Returns (ops.Frame, (ray_o, ray_d, near, far) on the device)."""
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
