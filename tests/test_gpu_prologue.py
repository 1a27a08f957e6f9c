"""SURVEY 8f rows 1 and 4 on the GPU: the paint + group kernel, the bit-exact cluster means and the ray generator,
through the C ABI, against the oracle restatements and the fixture generated from the genuine reference
(tests/golden/prologue_v2_100.npz: paint_neural_human, can_body_grouping, voxelization, get_rays, get_near_far)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import transhuman_oracle as orc
from tests.conftest import GOLDEN_CASES, GOLDEN_CASES_MANY_TOKENS, GOLDEN_DIR, load_golden
from transhuman_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def pro():
    z = np.load(os.path.join(GOLDEN_DIR, "prologue_v2_100.npz"))
    kw = ast.literal_eval(str(z["frame_kwargs"]))
    fr = synth.make_frame(**kw)
    return z, kw, fr


def test_paint_group_matches_reference_golden(pro):
    z, kw, fr = pro
    V, n = kw["V"], kw["n_class"]
    viz = np.unpackbits(z["viz"])[: V * synth.N_VERTS].reshape(V, -1).astype(bool)
    hm = _t(fr["pixel_feat_map"][:, :192])
    cl = ops.ClusterIndex(pc2voxel_ind=fr["pc2voxel_ind"], device=DEV)
    hw = fr["feat_hw"]
    got, painted = ops.paint_group(hm, ops.uv_scale_for(hw, hw, hw, hw), _t(fr["tar_smpl_vertice"]), _t(fr["input_R"]),
                                   _t(fr["input_T"]).reshape(V, 3), _t(fr["input_K"]), _t(viz), cl, want_painted=True)
    # per-vertex samples: same taps and weights as ATen's grid_sample, FMA order may differ by an ulp
    assert (painted[:, :256].cpu().numpy() - z["painted_sub"]).__abs__().max() <= 2e-6
    assert np.all(painted.cpu().numpy()[~viz] == 0)
    # cluster means: same summation order as torch-CPU -> the only difference is the per-vertex ulp above
    assert np.abs(got.cpu().numpy() - z["grouped"]).max() <= 2e-6
    # given identical per-vertex values the grouping is BIT-equal to the reference's voxelization
    tf = orc.to_torch_frame(fr)
    big = orc.paint_neural_human(tf["tar_smpl_vertice"], tf["input_R"], tf["input_T"], tf["input_K"],
                                 tf["pixel_feat_map"][:, :192].contiguous(), (hw, hw), torch.from_numpy(viz))
    for v in range(V):
        g = ops.group_mean(big[v].to(DEV), cl)
        assert np.array_equal(g.cpu().numpy(), z["grouped"][v])


def test_paint_without_vizmap_and_border(pro):
    z, kw, fr = pro
    V = kw["V"]
    hw = fr["feat_hw"]
    tf = orc.to_torch_frame(fr)
    verts = tf["tar_smpl_vertice"] * 3.0          # many vertices project outside the views: border clamp
    hm = tf["pixel_feat_map"][:, :192].contiguous()
    want = orc.paint_neural_human(verts, tf["input_R"], tf["input_T"], tf["input_K"], hm, (hw, hw), None)
    cl = ops.ClusterIndex(pc2voxel_ind=fr["pc2voxel_ind"], device=DEV)
    _, painted = ops.paint_group(hm.to(DEV), ops.uv_scale_for(hw, hw, hw, hw), verts.to(DEV), _t(fr["input_R"]),
                                 _t(fr["input_T"]).reshape(V, 3), _t(fr["input_K"]), None, cl, want_painted=True)
    assert (painted.cpu() - want).abs().max().item() <= 2e-6


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_CASES_MANY_TOKENS)
def test_token_means_bit_equal_to_reference_voxelization(name):
    """tok_xyz / tok_rot of every golden frame (100 ... 6000 tokens) were produced by the reference's voxelization
    loop: th_group_mean reproduces them bit for bit (fp32 (n,3) 'row' order; fp64 (n,4,4) 'outer' order)."""
    kw, S, mode, g = load_golden(name)
    fr = synth.make_frame(**kw)
    cl = ops.ClusterIndex(pc2voxel_ind=fr["pc2voxel_ind"], device=DEV)
    xyz = ops.group_mean(_t(fr["tar_smpl_vertice_smplcoord"]), cl)
    assert np.array_equal(xyz.cpu().numpy(), g["tok_xyz"])
    blend = ops.group_mean(_t(fr["blend_mtx"]), cl)
    assert blend.dtype == torch.float64
    assert np.array_equal(blend[:, :3, :3].float().cpu().numpy(), g["tok_rot"])


def test_group_mean_orders_on_large_clusters():
    """Clusters of 1 ... 700 members cross the 16-row (and 256-row) cascade boundaries of both orders."""
    gen = torch.Generator().manual_seed(0)
    sizes = [1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 700]
    pc2 = np.concatenate([np.full(s, i) for i, s in enumerate(sizes)])
    cl = ops.ClusterIndex(pc2voxel_ind=pc2, device=DEV)
    for shape, dt in (((3,), torch.float32), ((192,), torch.float32), ((4, 4), torch.float64), ((3,), torch.float64)):
        x = torch.randn((len(pc2),) + shape, generator=gen, dtype=dt)
        want = torch.stack([x[torch.from_numpy(np.nonzero(pc2 == i)[0])].mean(0) for i in range(len(sizes))])
        got = ops.group_mean(x.to(DEV), cl)
        assert torch.equal(got.cpu(), want), (shape, dt)


def test_generate_rays_matches_reference_golden(pro):
    z, kw, fr = pro
    H = int(z["cam_H"])
    out = ops.generate_rays(H, H, _t(z["cam_K"]), _t(z["cam_R"]), _t(z["cam_T"]), _t(z["bounds"]))
    m = np.unpackbits(z["mask_at_box"])[: H * H].astype(bool)
    # rays: 3-term dot products whose order numpy leaves to BLAS -> 1e-6 relative
    assert np.abs(out["ray_o"].cpu().numpy() - z["ray_o"]).max() <= 2e-6
    assert np.abs(out["ray_d"].cpu().numpy() - z["ray_d"]).max() <= 2e-6
    assert np.array_equal(out["mask_at_box"].cpu().numpy().astype(bool), m)
    assert out["count"] == int(m.sum()) and 0 < out["count"] < H * H
    assert np.abs(out["near_c"].cpu().numpy() - z["near"]).max() <= 5e-6
    assert np.abs(out["far_c"].cpu().numpy() - z["far"]).max() <= 5e-6
    assert torch.equal(out["ray_o_c"], out["ray_o"][torch.from_numpy(m).to(DEV)])
    assert torch.equal(out["ray_d_c"], out["ray_d"][torch.from_numpy(m).to(DEV)])


def test_near_far_bit_exact_given_the_same_rays():
    """get_near_far is float64 arithmetic on float32 rays: fed the reference's own rays the kernel's near / far /
    mask are bit-equal (the ray generator's own 1-ulp freedom is the only difference above)."""
    fr = synth.make_frame(H=96, W=96, n_class=100, V=1, feat_hw=8, seed=2, with_feature_maps=False)
    v = fr["tar_smpl_vertice"]
    bounds = np.stack([v.min(0) - 0.05, v.max(0) + 0.05]).astype(np.float32)
    want = orc.test_split_rays(96, 96, fr["target_K"], fr["target_R"], fr["target_T"], bounds)
    ro, rd = orc.get_rays_np(96, 96, fr["target_K"], fr["target_R"], fr["target_T"])
    ro = np.ascontiguousarray(ro.reshape(-1, 3).astype(np.float32))
    rd = np.ascontiguousarray(rd.reshape(-1, 3).astype(np.float32))
    rd[::97, 1] = 3e-6                                   # exercises the |d| < 1e-5 clamp (mutates ray_d like the reference)
    rd_ref = rd.copy()
    near, far, mask = orc.get_near_far_np(bounds, ro.copy(), rd_ref)
    got = ops.near_far(_t(ro), _t(rd), _t(bounds))
    assert np.array_equal(got["mask_at_box"].cpu().numpy().astype(bool), mask)
    assert np.array_equal(got["near"].cpu().numpy()[mask], near.astype(np.float32))
    assert np.array_equal(got["far"].cpu().numpy()[mask], far.astype(np.float32))
    assert np.array_equal(got["ray_d"].cpu().numpy(), rd_ref)       # the in-place clamp
    assert 0 < mask.sum() < mask.size and int(want["mask_at_box"].sum()) > 0
