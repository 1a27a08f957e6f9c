#!/usr/bin/env python
"""Timing of the other BASELINE.json configs on one GPU (not the bench line):
C3 = 512x512 x 128 samples, 1500 tokens (dense and culled); C5 = 256^3 density grid, 6000 tokens."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def c3():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--samples", "128", "--tokens", "1500",
                          "--no-cpu-baseline", "--steps", "2"], capture_output=True, text=True).stdout
    for l in out.splitlines():
        if l.startswith("{"):
            d = json.loads(l)
            return {"rays_per_s": d["value"], "ms_per_frame": d["ms_per_step"], "by_category": d["ms_per_step_by_category"],
                    "culled": d.get("culled")}
    return {"error": out[-400:]}


def c5(n=256, tokens=6000):
    import __graft_entry__ as entry
    entry.build()
    from tests.gpu_util import frame_to_device
    from transhuman_b200 import ops, synth
    from transhuman_b200.renderer import segment_mean
    fr = synth.make_frame(H=8, W=8, n_class=tokens, V=3, feat_hw=128, seed=5)
    pc2 = torch.from_numpy(fr["pc2voxel_ind"]).long()
    tokens_t = (segment_mean(torch.from_numpy(fr["tar_smpl_vertice_smplcoord"]), pc2, tokens).float(),
                segment_mean(torch.from_numpy(fr["blend_mtx"]), pc2, tokens))
    frame, _ = frame_to_device(fr, tokens_t, "cuda:0")
    v = fr["tar_smpl_vertice"]
    lo, hi = v.min(0) - 0.05, v.max(0) + 0.05
    ax = [torch.linspace(float(lo[i]), float(hi[i]), n) for i in range(3)]
    pts = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3).cuda().contiguous()
    res = {}
    for it in range(3):
        torch.cuda.synchronize()
        ops.profile_start()
        t0 = time.perf_counter()
        alpha, mask = ops.query_density(frame, pts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        prof = ops.profile_stop()
        res = {"grid": n, "tokens": tokens, "points": pts.shape[0], "inside_radius": int(mask.sum().item()),
               "seconds": dt, "points_per_s": pts.shape[0] / dt, "by_category_ms": {k: round(v[0], 2) for k, v in prof.items()}}
    return res


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c5"]
    out = {}
    if "c3" in which:
        out["c3"] = c3()
    if "c5" in which:
        out["c5"] = c5()
    print(json.dumps(out))
