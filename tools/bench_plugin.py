"""bench.py's `plugin` record: the drop-in `Renderer` end to end -- `render_fast(batch)` / `render(batch)` exactly as
run.py:52,109 / if_nerf_clight.py:45 call them -- with reference-shaped `encoder` / `ViT` modules (the genuine
`Network()` with random weights when the reference tree is present, torchvision ResNet-18-sized stand-ins otherwise),
and the per-stage prologue times (pack, encoder, paint, group, ViT, pre-map) at 300 / 1500 / 6000 tokens.

This is what a frame costs through the boundary; the bench line's `value` is the query path alone."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _batch(fr, imgs, device):
    def t(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(device)

    from transhuman_b200 import synth
    V = fr["V"]
    return {
        "ray_o": t(fr["ray_o"])[None], "ray_d": t(fr["ray_d"])[None], "near": t(fr["near"])[None],
        "far": t(fr["far"])[None], "tar_smpl_vertice": t(fr["tar_smpl_vertice"])[None],
        "tar_smpl_vertice_smplcoord": t(fr["tar_smpl_vertice_smplcoord"])[None],
        "Rh": t(fr["Rh"])[None], "Th": t(fr["Th"])[None], "blend_mtx": t(fr["blend_mtx"])[None],
        "input_imgs": [imgs.to(device)], "input_R": [t(fr["input_R"])[None]], "input_T": [t(fr["input_T"])[None]],
        "input_K": [t(fr["input_K"])[None]], "input_smpl_vertice": [t(fr["tar_smpl_vertice"])[None]],
        "input_vizmaps": [torch.ones((1, V, synth.N_VERTS), dtype=torch.bool, device=device)],
    }


def _network(fr, device):
    """(net, description): the genuine reference Network when its tree is present."""
    from oracle import ref_shim
    from oracle import transhuman_oracle as orc
    if ref_shim.reference_available():
        from transhuman_b200 import synth
        cwd = ref_shim.make_scratch_cwd(
            smpl_pkl={"v_template": synth.make_body(0), "f": np.zeros((1, 3), dtype=np.int64)},
            kmeans={n: synth.cluster_body(synth.make_body(0), n) for n in (300,)})
        ns = ref_shim.load_reference(orc.knn_points, cwd, opts=dict(perturb=0, rasterize=True), device="cuda")
        torch.manual_seed(0)
        ns.cfg.img_feat_size = 256      # Network.__init__ overwrites it (cross_transformer.py:123); see make_golden.py
        net = ns.cross_transformer.Network()
        sd = net.state_dict()
        for name, arr in fr["weights"].items():
            sd[name].copy_(torch.from_numpy(arr).view(sd[name].shape))
        return net.to(device).train(), "genuine reference Network() (ResNet-18 SpatialEncoder + vit_tiny), random init"
    raise RuntimeError("reference tree not present")


class _Cfg:
    N_samples = 64
    num_class = 300
    KNN = 7
    KNN_DIST_ALPHA = 0.5
    white_bkgd = False
    perturb = 0.
    rasterize = True
    time_steps = 1


def make(n_tok: int, device, size: int = 512):
    """(Renderer, batch) of one synthetic frame with the genuine Network (tools/profile_plugin_frame.py)."""
    from transhuman_b200 import synth
    from transhuman_b200.renderer import Renderer
    g = torch.Generator().manual_seed(7)
    imgs = torch.rand((1, 3, 3, size, size), generator=g)
    fr = synth.make_frame(H=size, W=size, n_class=n_tok, V=3, feat_hw=size, seed=0, with_feature_maps=False)
    net, _ = _network(fr, device)
    cfg = _Cfg()
    cfg.num_class = n_tok
    r = Renderer(net, cfg=cfg, pc2voxel_ind=fr["pc2voxel_ind"], vertex_can=synth.make_body(0))
    return r, _batch(fr, imgs, device)


def run(device, size: int = 512, tokens=(300, 1500, 6000)) -> dict:
    from transhuman_b200 import synth
    from transhuman_b200.renderer import Renderer
    out = {}
    g = torch.Generator().manual_seed(7)
    imgs = torch.rand((1, 3, 3, size, size), generator=g)
    net = desc = None
    for n_tok in tokens:
        fr = synth.make_frame(H=size, W=size, n_class=n_tok, V=3, feat_hw=size, seed=0, with_feature_maps=False)
        if net is None:
            net, desc = _network(fr, device)
        cfg = _Cfg()
        cfg.num_class = n_tok
        r = Renderer(net, cfg=cfg, pc2voxel_ind=fr["pc2voxel_ind"], vertex_can=synth.make_body(0))
        batch = _batch(fr, imgs, device)
        rec = {}
        with torch.no_grad():
            for mode, fn in (("render_fast", r.render_fast), ("render", r.render)):
                if mode == "render" and n_tok != tokens[0]:
                    continue
                fn(batch)                                            # warm-up (cuDNN autotune, weight pack)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    ret = fn(batch)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                rec[mode] = {"ms_per_frame": ms, "rays_per_s": size * size / (ms * 1e-3)}
                if mode == "render_fast":
                    rec[mode]["counters"] = list(r.last_counters)
            r.profile = True
            r.render_fast(batch)
            rec["prologue_ms"] = {k: round(v, 3) for k, v in r.last_prologue_ms.items()}
            rec["prologue_ms_total"] = round(sum(r.last_prologue_ms.values()), 3)
            assert torch.isfinite(ret["rgb_map"]).all()
        out[f"{n_tok}_tokens"] = rec
    out["net"] = desc
    out["frame"] = f"{size}x{size} rays, 64 samples, V=3 input views of {size}x{size}"
    return out
