"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.npz`` from the genuine
reference modules.

Run in the build container (needs ``/root/reference``):

    python oracle/make_golden.py

For each case a synthetic frame (``transhuman_b200.synth.make_frame``, fully
determined by its arguments) is fed to the *reference's own*
``Renderer.render`` / ``Renderer.render_fast`` / ``Network.forward`` /
``get_human_representation`` / ``raw2outputs`` (imported in place under
``oracle/ref_shim.py``; the encoder and ViT prologue -- out of scope, SURVEY
section 8 -- are replaced by modules returning the synthetic feature maps and
tokens).  Only the *outputs* and the frame arguments are stored; tests rebuild
the inputs from the arguments, so the fixtures stay small.

The reference cannot travel to the GPU box; these fixtures can.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import transhuman_oracle as orc  # noqa: E402
from transhuman_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> make_frame kwargs + S (+ mode).  Kept small: the reference runs on CPU.
CASES = {
    "tiny_dense": dict(frame=dict(H=12, W=12, n_class=300, V=3, feat_hw=24, seed=1), S=16, mode="dense"),
    "tiny_culled": dict(frame=dict(H=16, W=16, n_class=300, V=3, feat_hw=24, seed=2), S=16, mode="culled",
                        signed=True),
    "tiny_rotated": dict(frame=dict(H=10, W=10, n_class=100, V=3, feat_hw=16, seed=3, rotate_rh=True,
                                    posed=True), S=8, mode="dense"),
    "oneshot_v1": dict(frame=dict(H=10, W=10, n_class=300, V=1, feat_hw=16, seed=4), S=8, mode="culled"),
    # > 2400 surviving rays, so the reference takes its chunked + masked +
    # progressive branch (if_clight_renderer.py:551, 574-586)
    "culled_144x144x24": dict(frame=dict(H=144, W=144, n_class=300, V=3, feat_hw=32, seed=5), S=24, mode="culled",
                              signed=True),
    "c1_64x64x32": dict(frame=dict(H=64, W=64, n_class=300, V=3, feat_hw=64, seed=0), S=32, mode="culled",
                        signed=True),
    # token counts of BASELINE configs[2] (kmeans_dict_1500) and configs[4] (6000 tokens), small ray sets
    "tokens1500_dense": dict(frame=dict(H=12, W=12, n_class=1500, V=3, feat_hw=24, seed=6), S=32, mode="dense"),
    "tokens6000_culled": dict(frame=dict(H=14, W=14, n_class=6000, V=3, feat_hw=24, seed=7), S=16, mode="culled",
                              signed=True),
}


class _FakeEncoder(nn.Module):
    """Stands in for ``SpatialEncoder`` (encoder.py:97-155): returns the
    synthetic maps with the reference's scale convention."""

    def __init__(self, pixel_feat_map):
        super().__init__()
        self.pixel = pixel_feat_map
        self.holder_map = pixel_feat_map[:, :192].contiguous()

    def forward(self, images):
        sc = np.array([self.pixel.shape[-1], self.pixel.shape[-2]])
        sc = sc / (sc - 1) * 2.0
        return self.holder_map, sc, self.pixel, sc


class _FakeViT(nn.Module):
    """Stands in for ``vit_tiny`` (vision_transformer.py:309-383): returns the
    synthetic tokens."""

    def __init__(self, holder):
        super().__init__()
        self.holder = holder
        self.embed_dim = 192

    def forward(self, tokens, pe, mask=None):
        assert tokens.shape == self.holder.shape, (tokens.shape, self.holder.shape)
        return self.holder


def build_reference(frame: dict, S: int, device: str = "cpu", knn=None, fake_prologue: bool = True):
    """Instantiate the genuine reference ``Network`` + ``Renderer`` for a
    synthetic frame.  Returns ``(ns, net, renderer, batch)``.  ``device="cuda"`` runs the reference in
    torch-CUDA (GPU box, from the copy under baseline/_ref); ``fake_prologue=False`` keeps the genuine
    encoder + ViT (random init) and feeds them ``frame["input_imgs"]``."""
    cwd = ref_shim.make_scratch_cwd(
        smpl_pkl={"v_template": synth.make_body(0), "f": np.zeros((1, 3), dtype=np.int64)},
        kmeans={frame["n_class"]: frame["pc2voxel_ind"]})
    ns = ref_shim.load_reference(knn or orc.knn_points, cwd,
                                 opts=dict(N_samples=S, num_class=frame["n_class"], perturb=0,
                                           rasterize=True), device=device)
    tf = orc.to_torch_frame(frame)
    torch.manual_seed(0)
    # Network.__init__ leaves cfg.img_feat_size = 384 behind (cross_transformer.py:123), which would size the NEXT
    # encoder's reduction_layer for 512 channels (encoder.py:85): restore the YAML value before every construction
    ns.cfg.img_feat_size = 256
    net = ns.cross_transformer.Network()
    sd = net.state_dict()
    for name, arr in tf["weights"].items():
        assert name in sd, name
        sd[name].copy_(arr.view(sd[name].shape))
    dev = torch.device(device)
    if fake_prologue:
        net.encoder = _FakeEncoder(tf["pixel_feat_map"].to(dev))
        net.ViT = _FakeViT(tf["holder"].to(dev))
    net.to(dev)
    net.train()  # run.py:29 -- inference runs with net.training == True
    renderer = ns.renderer_mod.Renderer(net)
    V, hw = frame["V"], frame["feat_hw"]
    imgs = tf["input_imgs"][None] if "input_imgs" in tf else torch.zeros((1, V, 3, hw, hw))
    batch = {
        "ray_o": tf["ray_o"][None], "ray_d": tf["ray_d"][None],
        "near": tf["near"][None], "far": tf["far"][None],
        "tar_smpl_vertice": tf["tar_smpl_vertice"][None],
        "tar_smpl_vertice_smplcoord": tf["tar_smpl_vertice_smplcoord"][None],
        "Rh": tf["Rh"][None], "Th": tf["Th"][None],
        "blend_mtx": tf["blend_mtx"][None],
        "input_imgs": [imgs],
        "input_R": [tf["input_R"][None]], "input_T": [tf["input_T"][None]], "input_K": [tf["input_K"][None]],
        "input_smpl_vertice": [tf["tar_smpl_vertice"][None]],
        "input_vizmaps": [torch.ones((1, V, synth.N_VERTS), dtype=torch.bool)],
        "input_blend_mtx": [tf["blend_mtx"][None]],
        "input_smpl_vertice_smplcoord": [tf["tar_smpl_vertice_smplcoord"][None]],
    }
    if device != "cpu":
        batch = {k: ([x.to(dev) for x in v] if isinstance(v, list) else v.to(dev)) for k, v in batch.items()}
    return ns, net, renderer, batch


def signed_shift(frame: dict, S: int) -> float:
    """alpha_fc bias shift that makes ~half of alpha_raw <= 0 on the culled
    points (exercises the progressive RGB branch, cross_transformer.py:291-311)."""
    tf = orc.to_torch_frame(frame)
    out = orc.render_fast(tf, S, train_branch_max_rays=0)
    m = out["valid_pts_mask"][0]
    a = out["raw"][..., 3][m]
    return -float(a.median()) if a.numel() else 0.0


def make_case(name: str, spec: dict) -> dict:
    S = spec["S"]
    kw = dict(spec["frame"])
    if spec.get("signed"):
        kw["alpha_bias_shift"] = signed_shift(synth.make_frame(**kw), S)
    frame = synth.make_frame(**kw)
    ns, net, renderer, batch = build_reference(frame, S)
    tf = orc.to_torch_frame(frame)
    with torch.no_grad():
        if spec["mode"] == "dense":
            # Renderer.render takes the un-chunked train branch when N <= 2400
            # (if_clight_renderer.py:551); both branches are pts_mask=None.
            ret = renderer.render(dict(batch), is_train=False)
        else:
            ret = renderer.render_fast(dict(batch), is_train=False)
        # stage outputs from the reference's own methods
        pts, z_vals = renderer.get_sampling_points(batch["ray_o"], batch["ray_d"], batch["near"], batch["far"])
        pts_s = renderer.world2smpl(pts, batch["Rh"], batch["Th"]).flatten(1, 2)
        tok_xyz = renderer.voxelization(renderer.dict_voxel2pc_ind, batch["tar_smpl_vertice_smplcoord"][0])
        tok_blend = renderer.voxelization(renderer.dict_voxel2pc_ind, batch["blend_mtx"][0])
        sub = slice(0, min(pts_s.shape[1], 2048))
        rep, _ = net.get_human_representation(pts_s[:, sub], tok_xyz[None], tok_blend[None], tf["holder"])
        d2, idx, _ = orc.knn_points(pts_s[:, sub], tok_xyz[None], K=7)
        pix = renderer.get_pixel_aligned_feature(
            batch, pts.flatten(1, 2)[:, sub], tf["pixel_feat_map"],
            np.array([frame["feat_hw"]] * 2) / (np.array([frame["feat_hw"]] * 2) - 1) * 2.0, t=0)
        viewdir = ns.embedder.view_embedder(batch["ray_d"] / torch.norm(batch["ray_d"], dim=2, keepdim=True))
        vd = viewdir[:, :, None].repeat(1, 1, S, 1).contiguous().view(1, -1, 27)[:, sub]
        dp = {"pts_smplcoord": pts_s[:, sub], "obs_smpl_smplcoord": tok_xyz[None], "blend_mtx": tok_blend[None]}
        raw_dense = net(pixel_feat=pix, sincos_viewdir=vd, DPaRF_param_dict=dict(dp), holder=tf["holder"])
        cullmask = orc.cull_mask(pts.flatten(1, 2), batch["tar_smpl_vertice"])
        # the masked + progressive network branch on the same sub-slice: use the
        # cull mask where it has survivors, else a fixed stride pattern
        m_sub = cullmask[:, sub].clone()
        if m_sub.sum() < 8:
            m_sub[:, ::3] = True
        raw_masked = net(pixel_feat=pix, sincos_viewdir=vd, DPaRF_param_dict=dict(dp), holder=tf["holder"],
                         pts_mask=m_sub)
    out = {
        "frame_kwargs": np.array(repr(kw)), "S": np.int32(S), "mode": np.array(spec["mode"]),
        "rgb_map": ret["rgb_map"][0].numpy(), "acc_map": ret["acc_map"][0].numpy(),
        "depth_map": ret["depth_map"][0].numpy(),
        "z_vals_sub": z_vals[0, :8].numpy(), "pts_sub": pts[0, :8].numpy(),
        "pts_smpl_sub": pts_s[0, sub][:64].numpy(),
        "tok_xyz": tok_xyz.numpy(), "tok_rot": tok_blend[:, :3, :3].float().numpy(),
        "knn_idx": idx[0].numpy().astype(np.int16), "knn_d2": d2[0].numpy(),
        "human_rep_sub": rep[:, :, :64].numpy(),
        "pixel_feat_sub": pix[:, :, :48].numpy(),
        "viewdir_sub": viewdir[0, :64].numpy(),
        "raw_dense_sub": raw_dense[0].numpy(),
        "raw_masked_sub": raw_masked[0].numpy(), "mask_sub": np.packbits(m_sub[0].numpy()),
        "n_rays_surviving": np.int64((cullmask.view(1, -1, S).sum(-1) > 0).sum().item()),
        "cull_mask": np.packbits(cullmask[0].numpy()),
        "n_cull": np.int64(cullmask.sum().item()),
    }
    return out


def make_prologue_case() -> dict:
    """Token prologue and ray generation (8f-4) through the reference's own functions:
    ``paint_neural_human`` + ``can_body_grouping`` on a synthetic holder map with random visibility, and
    ``get_rays`` / ``get_near_far`` of lib/utils/if_nerf/if_nerf_data_utils.py for a small target camera."""
    import importlib
    kw = dict(H=8, W=8, n_class=100, V=2, feat_hw=24, seed=9)
    frame = synth.make_frame(**kw)
    ns, net, renderer, batch = build_reference(frame, 8)
    g = torch.Generator().manual_seed(3)
    viz = torch.rand((1, 2, synth.N_VERTS), generator=g) > 0.3
    batch["input_vizmaps"] = [viz]
    tf = orc.to_torch_frame(frame)
    hm = tf["pixel_feat_map"][:, :192].contiguous()
    sc = np.array([24, 24])
    sc = sc / (sc - 1) * 2.0
    with torch.no_grad():
        _, big = renderer.paint_neural_human(batch, 0, hm, sc)
        grouped = renderer.can_body_grouping(big)
        tok_xyz = renderer.voxelization(renderer.dict_voxel2pc_ind, batch["tar_smpl_vertice_smplcoord"][0])
        tok_blend = renderer.voxelization(renderer.dict_voxel2pc_ind, batch["blend_mtx"][0])
    du = importlib.import_module("lib.utils.if_nerf.if_nerf_data_utils")
    Hc = 24
    K = frame["target_K"].copy()
    K[0, 0] = K[1, 1] = 30.0
    K[0, 2] = K[1, 2] = Hc / 2
    R, T = frame["target_R"], frame["target_T"]
    v = frame["tar_smpl_vertice"]
    bounds = np.stack([v.min(0) - 0.05, v.max(0) + 0.05]).astype(np.float32)
    ray_o, ray_d = du.get_rays(Hc, Hc, K, R, T)
    ray_o = ray_o.reshape(-1, 3).astype(np.float32)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    near, far, mask_at_box = du.get_near_far(bounds, ray_o, ray_d)        # clamps ray_d in place
    return {
        "frame_kwargs": np.array(repr(kw)), "viz": np.packbits(viz[0].numpy()),
        "painted_sub": big[:, :256].numpy(), "grouped": grouped.numpy(),
        "tok_xyz": tok_xyz.numpy(), "tok_blend": tok_blend.numpy(),
        "cam_H": np.int32(Hc), "cam_K": K, "cam_R": R, "cam_T": T, "bounds": bounds,
        "ray_o": ray_o, "ray_d": ray_d, "near": near.astype(np.float32), "far": far.astype(np.float32),
        "mask_at_box": np.packbits(mask_at_box),
    }


def main(names=None):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if not names or "prologue_v2_100" in names:
        out = make_prologue_case()
        path = os.path.join(GOLDEN_DIR, "prologue_v2_100.npz")
        np.savez_compressed(path, **out)
        print(f"prologue_v2_100: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    for name, spec in CASES.items():
        if names and name not in names:
            continue
        out = make_case(name, spec)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), "
              f"cull {int(out['n_cull'])} pts, rgb max {out['rgb_map'].max():.4f}")


if __name__ == "__main__":
    main(sys.argv[1:])
