"""The CPU oracle against the genuine reference modules imported in place from
/root/reference (build container only; skipped where the tree is absent, e.g.
on the GPU box).  This is what pins oracle/transhuman_oracle.py; the committed
fixtures under tests/golden carry the same outputs to boxes without the tree."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle import transhuman_oracle as orc
from transhuman_b200 import synth

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="no /root/reference here")]


@pytest.fixture(scope="module")
def ref():
    from oracle.make_golden import build_reference
    frame = synth.make_frame(H=20, W=20, n_class=300, V=3, feat_hw=20, seed=7)
    ns, net, renderer, batch = build_reference(frame, 16)
    return frame, ns, net, renderer, batch


def test_render_dense(ref):
    frame, ns, net, renderer, batch = ref
    with torch.no_grad():
        r = renderer.render(dict(batch), is_train=False)
    o = orc.render(orc.to_torch_frame(frame), 16)
    for k, tol in (("rgb_map", 2e-6), ("acc_map", 2e-6), ("depth_map", 1e-5)):
        assert (r[k] - o[k]).abs().max().item() <= tol, k
    assert r["rgb_map"].abs().max() > 0.1


def test_render_fast_culled(ref):
    frame, ns, net, renderer, batch = ref
    with torch.no_grad():
        r = renderer.render_fast(dict(batch), is_train=False)
    o = orc.render_fast(orc.to_torch_frame(frame), 16)
    for k, tol in (("rgb_map", 2e-6), ("acc_map", 2e-6), ("depth_map", 1e-5)):
        assert (r[k] - o[k]).abs().max().item() <= tol, k
    assert (r["acc_map"] > 0).sum() > 0


def test_raw2outputs(ref):
    frame, ns, *_ = ref
    g = torch.Generator().manual_seed(3)
    raw = torch.randn((50, 16, 4), generator=g) * 3
    z = torch.sort(torch.rand((50, 16), generator=g) + 1.0, dim=1)[0]
    d = torch.randn((50, 3), generator=g)
    a = ns.nerf_net_utils.raw2outputs(raw, z, d, 0, False)
    b = orc.raw2outputs(raw, z, d)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[1]) and torch.equal(a[4], b[3])
    a = ns.nerf_net_utils.raw2outputs(raw, z, d, 0, True)
    assert torch.equal(a[0], orc.raw2outputs(raw, z, d, white_bkgd=True)[0])


def test_embedders_and_pe(ref):
    frame, ns, net, *_ = ref
    g = torch.Generator().manual_seed(4)
    x = torch.randn((1000, 3), generator=g) * 0.3
    assert torch.equal(net.PE_relative(x), orc.positional_encoding(x))
    v = torch.randn((1, 100, 3), generator=g)
    assert torch.equal(ns.embedder.view_embedder(v / torch.norm(v, dim=2, keepdim=True)), orc.view_embed(v))


def test_voxelization_on_real_kmeans_dict(ref):
    """Token construction on the reference's own k-means dictionaries: their keys
    are exactly arange(N_c), so the pc2voxel form is equivalent."""
    import os
    frame, ns, net, renderer, batch = ref
    for n in (300, 1500):
        d = np.load(os.path.join(ref_shim.REFERENCE_ROOT, "kmeans_dict", f"kmeans_dict_{n}.npy"),
                    allow_pickle=True).item()
        pc2voxel, v2pc = d["pc2voxel_ind"], d["dict_voxel2pc_ind"]
        assert list(v2pc.keys()) == list(range(n))
        x = torch.from_numpy(frame["tar_smpl_vertice_smplcoord"])
        a = renderer.voxelization(v2pc, x)
        b = orc.voxelization(torch.from_numpy(pc2voxel.astype(np.int64)), x, n)
        assert torch.equal(a, b)
        c = synth.segment_mean(frame["blend_mtx"], pc2voxel, n)
        bm = renderer.voxelization(v2pc, torch.from_numpy(frame["blend_mtx"]))
        assert np.abs(c - bm.numpy()).max() < 1e-12


def test_prologue_paint_group_and_rays_bit_equal(ref):
    """SURVEY 8f-1 / 8f-4 restatements against the reference's own functions, bit for bit."""
    import importlib
    frame, ns, net, renderer, batch = ref
    tf = orc.to_torch_frame(frame)
    g = torch.Generator().manual_seed(3)
    viz = torch.rand((1, 3, synth.N_VERTS), generator=g) > 0.3
    b2 = dict(batch)
    b2["input_vizmaps"] = [viz]
    hm = tf["pixel_feat_map"][:, :192].contiguous()
    hw = frame["feat_hw"]
    sc = np.array([hw, hw])
    sc = sc / (sc - 1) * 2.0
    with torch.no_grad():
        _, big = renderer.paint_neural_human(b2, 0, hm, sc)
        grouped = renderer.can_body_grouping(big)
    mine = orc.paint_neural_human(tf["tar_smpl_vertice"], tf["input_R"], tf["input_T"], tf["input_K"], hm, (hw, hw), viz[0])
    assert torch.equal(mine, big)
    lists = list(renderer.dict_voxel2pc_ind.values())
    assert torch.equal(orc.can_body_grouping(lists, mine), grouped)
    du = importlib.import_module("lib.utils.if_nerf.if_nerf_data_utils")
    K, R, T = frame["target_K"], frame["target_R"], frame["target_T"]
    ro, rd = du.get_rays(20, 20, K, R, T)
    ro2, rd2 = orc.get_rays_np(20, 20, K, R, T)
    assert np.array_equal(ro, ro2) and np.array_equal(rd, rd2)
    v = frame["tar_smpl_vertice"]
    bounds = np.stack([v.min(0) - 0.05, v.max(0) + 0.05]).astype(np.float32)
    f32 = lambda a: a.reshape(-1, 3).astype(np.float32).copy()
    a = du.get_near_far(bounds, f32(ro), f32(rd))
    b = orc.get_near_far_np(bounds, f32(ro), f32(rd))
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and 0 < a[2].sum() < a[2].size
