"""The CPU oracle against the committed golden fixtures (generated from the
genuine reference modules by oracle/make_golden.py).  Pins the oracle; runs on
CPU anywhere (the reference tree is NOT needed)."""
import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN_CASES, GOLDEN_CASES_MANY_TOKENS, load_golden
from oracle import transhuman_oracle as orc
from transhuman_b200 import synth


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_CASES_MANY_TOKENS)
def test_oracle_matches_reference_golden(name):
    kw, S, mode, g = load_golden(name)
    if name == "c1_64x64x32":
        pytest.importorskip("torch")
    frame = synth.make_frame(**kw)
    tf = orc.to_torch_frame(frame)
    tok_xyz, tok_blend = orc.build_tokens(tf)
    # token construction == reference voxelization (if_clight_renderer.py:356-371)
    np.testing.assert_array_equal(tok_xyz.numpy(), g["tok_xyz"])
    np.testing.assert_array_equal(tok_blend[:, :3, :3].float().numpy(), g["tok_rot"])
    out = orc.render(tf, S, tokens=(tok_xyz, tok_blend)) if mode == "dense" else \
        orc.render_fast(tf, S, tokens=(tok_xyz, tok_blend))
    # same torch ops in the same order as the reference -> tight tolerance
    np.testing.assert_allclose(out["rgb_map"][0].numpy(), g["rgb_map"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(out["acc_map"][0].numpy(), g["acc_map"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(out["depth_map"][0].numpy(), g["depth_map"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("name", GOLDEN_CASES[:4] + GOLDEN_CASES_MANY_TOKENS)
def test_oracle_stages_match_reference_golden(name):
    kw, S, mode, g = load_golden(name)
    frame = synth.make_frame(**kw)
    tf = orc.to_torch_frame(frame)
    pts, z = orc.get_sampling_points(tf["ray_o"][None], tf["ray_d"][None], tf["near"][None], tf["far"][None], S)
    np.testing.assert_array_equal(z[0, :8].numpy(), g["z_vals_sub"])
    np.testing.assert_array_equal(pts[0, :8].numpy(), g["pts_sub"])
    pts_s = orc.world2smpl(pts, tf["Rh"][None], tf["Th"][None]).flatten(1, 2)
    np.testing.assert_array_equal(pts_s[0, :64].numpy(), g["pts_smpl_sub"])
    n = g["knn_idx"].shape[0]
    tok_xyz, tok_blend = orc.build_tokens(tf)
    rep, idx, dist, w, deformed = orc.human_representation(pts_s[0, :n], tok_xyz, tok_blend, tf["holder"],
                                                           return_knn=True)
    np.testing.assert_array_equal(idx.numpy(), g["knn_idx"].astype(np.int64))
    np.testing.assert_array_equal((dist ** 2).numpy().shape, g["knn_d2"].shape)
    np.testing.assert_allclose(rep[:, :, :64].numpy(), g["human_rep_sub"], atol=1e-6, rtol=0)
    pix = orc.get_pixel_aligned_feature(pts.flatten(1, 2)[:, :n], tf["input_R"], tf["input_T"], tf["input_K"],
                                        tf["pixel_feat_map"], tf["pixel_feat_map"].shape[-2:])
    np.testing.assert_allclose(pix[:, :, :48].numpy(), g["pixel_feat_sub"], atol=1e-6, rtol=0)
    vd = orc.view_embed(tf["ray_d"][None])
    np.testing.assert_allclose(vd[0, :64].numpy(), g["viewdir_sub"], atol=1e-7, rtol=0)
    vdp = vd[:, :, None].repeat(1, 1, S, 1).contiguous().view(1, -1, 27)[:, :n]
    raw = orc.network_forward(tf["weights"], pix, vdp, pts_s[:, :n], tok_xyz, tok_blend, tf["holder"])
    np.testing.assert_allclose(raw[0].numpy(), g["raw_dense_sub"], atol=2e-5, rtol=1e-5)
    m_sub = torch.from_numpy(np.unpackbits(g["mask_sub"])[:n].astype(bool))[None]
    rawm = orc.network_forward(tf["weights"], pix, vdp, pts_s[:, :n], tok_xyz, tok_blend, tf["holder"],
                               pts_mask=m_sub)
    np.testing.assert_allclose(rawm[0].numpy(), g["raw_masked_sub"], atol=2e-5, rtol=1e-5)
    assert np.all(rawm[0].numpy()[~m_sub[0].numpy()] == 0)
    mask = orc.cull_mask(pts.flatten(1, 2), tf["tar_smpl_vertice"][None])
    np.testing.assert_array_equal(np.packbits(mask[0].numpy()), g["cull_mask"])
    assert int(mask.sum()) == int(g["n_cull"])


def test_knn_tie_rule_and_order():
    """Ties go to the lower index; output sorted by (d2, idx)."""
    q = torch.tensor([[[0.0, 0.0, 0.0]]])
    p = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0.5, 0, 0], [-1.0, 0, 0], [0, 0, 0.5]]])
    d, i, _ = orc.knn_points(q, p, K=4)
    assert i[0, 0].tolist() == [3, 5, 0, 1]
    assert d[0, 0].tolist() == [0.25, 0.25, 1.0, 1.0]
    d1, i1, _ = orc.knn_points(q, p, K=1)
    assert i1[0, 0, 0].item() == 3


def test_raw2outputs_masked_points_are_transparent():
    """raw == 0 leaves transmittance untouched: 1 - 0 + 1e-10 == 1 in fp32
    (SURVEY 3.5-8)."""
    raw = torch.zeros((3, 8, 4))
    raw[:, 4, 3] = 50.0
    raw[:, 4, :3] = 1.0
    z = torch.linspace(1, 2, 8)[None].repeat(3, 1)
    d = torch.tensor([[0, 0, 1.0]]).repeat(3, 1)
    rgb, acc, w, depth = orc.raw2outputs(raw, z, d)
    assert torch.all(w[:, :4] == 0) and torch.all(w[:, 5:7] == 0)
    assert torch.allclose(acc, w[:, 4])


def test_prologue_golden_matches_oracle():
    """tests/golden/prologue_v2_100.npz (the reference's own paint / grouping / ray functions) against the oracle
    restatements -- the fixture that carries SURVEY 8f-1 / 8f-4 to the GPU box."""
    import ast
    import os
    from tests.conftest import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "prologue_v2_100.npz"))
    kw = ast.literal_eval(str(z["frame_kwargs"]))
    fr = synth.make_frame(**kw)
    tf = orc.to_torch_frame(fr)
    viz = torch.from_numpy(np.unpackbits(z["viz"])[: 2 * synth.N_VERTS].reshape(2, -1).astype(bool))
    hm = tf["pixel_feat_map"][:, :192].contiguous()
    big = orc.paint_neural_human(tf["tar_smpl_vertice"], tf["input_R"], tf["input_T"], tf["input_K"], hm, (24, 24), viz)
    assert np.array_equal(big[:, :256].numpy(), z["painted_sub"])
    pc2 = tf["pc2voxel_ind"].long()
    lists = [torch.nonzero(pc2 == c)[:, 0] for c in range(kw["n_class"])]
    assert np.array_equal(orc.can_body_grouping(lists, big).numpy(), z["grouped"])
    assert np.array_equal(orc.voxelization(pc2, tf["tar_smpl_vertice_smplcoord"], 100).numpy(), z["tok_xyz"])
    assert np.array_equal(orc.voxelization(pc2, tf["blend_mtx"], 100).numpy(), z["tok_blend"])
    H = int(z["cam_H"])
    r = orc.test_split_rays(H, H, z["cam_K"], z["cam_R"], z["cam_T"], z["bounds"])
    m = np.unpackbits(z["mask_at_box"])[: H * H].astype(bool)
    assert np.array_equal(r["mask_at_box"], m) and np.array_equal(r["near"], z["near"]) and np.array_equal(r["far"], z["far"])
    assert np.array_equal(r["ray_d_all"], z["ray_d"]) and np.array_equal(r["ray_o_all"], z["ray_o"])
