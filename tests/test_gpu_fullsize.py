"""BASELINE.json configurations at full size on the B200.  The CPU oracle cannot
render 16.7 M points in test time, so these check (a) random ray / point samples
of the full-size result against the oracle run on just those rays, and (b)
size-independent properties: chunk independence, culled rays exactly zero,
dense == masked where every sample is inside the cull radius, density query ==
alpha channel of the ray render, fp32-SIMT path == tensor-core path."""
import numpy as np
import pytest
import torch

from oracle import transhuman_oracle as orc
from tests.gpu_util import assert_maps_close, frame_to_device
from transhuman_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frame(n_class, H, feat_hw, seed=0, shift=-15.0, premapped=None):
    fr = synth.make_frame(H=H, W=H, n_class=n_class, V=3, feat_hw=feat_hw, seed=seed, alpha_bias_shift=shift)
    tf = orc.to_torch_frame(fr)
    tokens = orc.build_tokens(tf)
    frame, rays = frame_to_device(fr, tokens, DEV, premapped=premapped)     # default: the pre-mapped path
    return fr, tf, tokens, frame, rays


def _check(got_sel, want, sub_tf, S, name):
    _, z = orc.get_sampling_points(sub_tf["ray_o"][None], sub_tf["ray_d"][None], sub_tf["near"][None],
                                   sub_tf["far"][None], S)
    assert_maps_close(got_sel, want, want["raw"], z[0], sub_tf["ray_d"], S, 3.5, name)


def _oracle_on_rays(tf, tokens, sel, S, culled):
    sub = dict(tf)
    for k in ("ray_o", "ray_d", "near", "far"):
        sub[k] = tf[k][sel]
    if culled:
        return orc.render_fast(sub, S, tokens=tokens, train_branch_max_rays=0), sub
    return orc.render(sub, S, tokens=tokens), sub


@pytest.mark.parametrize("n_class,S", [(300, 64), (1500, 128)])
def test_config_512_sampled_parity(n_class, S):
    """configs[1] (512x512x64, 300 tokens) and configs[2] (512x512x128, 1500 tokens):
    full-size culled + a dense band, spot-checked against the oracle."""
    fr, tf, tokens, frame, rays = _frame(n_class, 512, 128)     # 128x128 input views keep host memory small
    N = 512 * 512
    got = ops.render_rays(frame, *rays, S, mode=ops.TH_RENDER_MASKED, want_mask=True, want_raw=True)
    n_in, n_rays, n_eval = got["counters"]
    assert 0.005 < n_in / (N * S) < 0.2 and n_eval == n_in
    pm = got["pts_mask"].bool()                      # uint8.any() stays uint8 and ~ would be a bitwise not
    alive = torch.nonzero(pm.any(dim=1))[:, 0].cpu()
    dead = torch.nonzero(~pm.any(dim=1))[:, 0]
    assert len(alive) == n_rays
    assert torch.all(got["rgb_map"][dead] == 0) and torch.all(got["acc_map"][dead] == 0)
    g = torch.Generator().manual_seed(0)
    sel = alive[torch.randperm(len(alive), generator=g)[:96]]
    want, sub = _oracle_on_rays(tf, tokens, sel, S, culled=True)
    assert torch.equal(got["pts_mask"][sel.to(DEV)].cpu().bool(), want["valid_pts_mask"][0])   # exact cull
    _check({k: got[k][sel.to(DEV)] for k in ("rgb_map", "acc_map", "depth_map", "raw")}, want, sub, S,
           f"512^2 culled, {n_class} tokens")
    del got
    # dense band of 2048 rays: sampled parity + chunk independence against the same rays alone
    band = slice(N // 2, N // 2 + 2048)
    dense = ops.render_rays(frame, *(r[band] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    sel2 = torch.arange(0, 2048, 64)
    want2, sub2 = _oracle_on_rays(tf, tokens, torch.arange(N // 2, N // 2 + 2048)[sel2], S, culled=False)
    _check({k: dense[k][sel2.to(DEV)] for k in ("rgb_map", "acc_map", "depth_map", "raw")}, want2, sub2, S,
           f"512^2 dense band, {n_class} tokens")
    part = ops.render_rays(frame, *(r[N // 2 + 512:N // 2 + 1024] for r in rays), S, mode=ops.TH_RENDER_DENSE)
    assert torch.equal(part["rgb_map"], dense["rgb_map"][512:1024])


def test_simt_and_tensor_core_paths_agree_at_size():
    """Three independent evaluations of the per-point network on 524,288 points: fp32 CUDA cores with the
    UNFOLDED layers, tcgen05 with the folded layers on the plain maps, tcgen05 on the pre-mapped maps."""
    fr, tf, tokens, frame_pre, rays = _frame(300, 256, 64, seed=2)
    frame, _ = frame_to_device(fr, tokens, DEV, premapped=False, weights=frame_pre.weights)
    S = 64
    sel = slice(30000, 30000 + 8192)
    p = ops.render_rays(frame_pre, *(r[sel] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    a = ops.render_rays(frame, *(r[sel] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    frame.set_flag(ops.TH_FLAG_SIMT_MLP, True)
    b = ops.render_rays(frame, *(r[sel] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    frame.set_flag(ops.TH_FLAG_SIMT_MLP, False)
    scale = max(1.0, b["raw"].abs().max().item())
    last = b["raw"][:, -1, 3]
    keep = ~(last.abs() < 2e-5 * scale)
    print(f"[simt vs tc] {int((~keep).sum())} of {keep.numel()} rays on the alpha_S step excluded")
    for name, x in (("tc", a), ("premapped", p)):
        assert (x["raw"] - b["raw"]).abs().max().item() <= 2e-5 * scale, name
        assert (x["rgb_map"] - b["rgb_map"])[keep].abs().max().item() <= 1e-4, name


def test_chain_kernel_is_insensitive_to_role_timing(monkeypatch):
    """The layer-chained kernel synchronises its warp roles with mbarriers and counters only.
    TH_CHAIN_DBG=240 delays the loader, the MMA issuer, the epilogue and the mix warps by random
    amounts (15 units per cluster here): the arithmetic is deterministic, so the result must be
    bit-identical.  (Caught a per-k-block counter that assumed the mix warps ran in lockstep.)"""
    fr, tf, tokens, frame, rays = _frame(300, 256, 64, seed=2)
    S = 64
    sel = slice(20000, 20000 + 8192)
    a = ops.render_rays(frame, *(r[sel] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    for bits in ("128", "240"):
        monkeypatch.setenv("TH_CHAIN_DBG", bits)
        b = ops.render_rays(frame, *(r[sel] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
        monkeypatch.delenv("TH_CHAIN_DBG")
        assert torch.equal(a["raw"], b["raw"]) and torch.equal(a["rgb_map"], b["rgb_map"])
    # the opt-in schedule that issues the tail of the network one unit late (software pipelining across units,
    # alternating scratch slot sets): same arithmetic in a different order of jobs -> bit-identical, also under jitter
    # ("2" = deep deferral: everything behind the scores -- fc_1' ... fc_4' -- one unit late, so that the tensor pipe
    # never waits for the in-place mix; the attention table alternates with the unit's parity)
    for mode in ("1", "2"):
        monkeypatch.setenv("TH_CHAIN_DEFER", mode)
        for bits in ("0", "240", "128"):
            monkeypatch.setenv("TH_CHAIN_DBG", bits)
            c = ops.render_rays(frame, *(r[sel] for r in rays), S, mode=ops.TH_RENDER_DENSE, want_raw=True)
            assert torch.equal(a["raw"], c["raw"]) and torch.equal(a["rgb_map"], c["rgb_map"]), (mode, bits)
    monkeypatch.delenv("TH_CHAIN_DBG")
    monkeypatch.delenv("TH_CHAIN_DEFER")


def test_config_grid_6000_tokens_density():
    """configs[4] shape: dense voxel grid, 6000 tokens, alpha only (reduced to 96^3
    so the oracle spot check stays small; the kernel path is the same)."""
    fr, tf, tokens, frame, rays = _frame(6000, 64, 64, seed=1, shift=0.0)
    grid = torch.from_numpy(synth.make_grid_points(fr, 96).reshape(-1, 3))
    alpha, mask = ops.query_density(frame, grid.to(DEV))
    mask = mask.bool()
    assert 0.01 < mask.float().mean().item() < 0.6
    assert torch.all(alpha[~mask] == 0)
    inside = torch.nonzero(mask.cpu())[:, 0]
    g = torch.Generator().manual_seed(3)
    sel = inside[torch.randperm(len(inside), generator=g)[:512]]
    far = torch.nonzero(~mask.cpu())[:, 0][:512]
    pts = torch.cat([grid[sel], grid[far]])
    walpha, wmask = orc.query_density(tf, pts, tokens=tokens)
    assert torch.equal(wmask, torch.cat([torch.ones(512, dtype=torch.bool), torch.zeros(512, dtype=torch.bool)]))
    got = torch.cat([alpha.cpu()[sel], alpha.cpu()[far]])
    assert (got - walpha).abs().max().item() <= 2e-5 * max(1.0, walpha.abs().max().item())
    # the same points as a ray bundle of one sample each give the same alpha channel
    o = grid[sel].to(DEV)
    d = torch.zeros_like(o)
    d[:, 2] = 1.0
    z = torch.zeros(len(sel), device=DEV)
    r = ops.render_rays(frame, o, d, z, z, 1, mode=ops.TH_RENDER_DENSE, want_raw=True)
    assert (r["raw"][:, 0, 3].cpu() - alpha.cpu()[sel]).abs().max().item() <= 2e-5 * max(1.0, walpha.abs().max().item())


def test_knn_exact_with_many_tokens():
    """Bit-exact K-NN at 1500 and 6000 tokens (configs[2], configs[4])."""
    for n_class in (1500, 6000):
        fr = synth.make_frame(H=8, W=8, n_class=n_class, V=1, feat_hw=8, seed=5)
        tf = orc.to_torch_frame(fr)
        tokens = orc.build_tokens(tf)
        frame, rays = frame_to_device(fr, tokens, DEV)
        g = torch.Generator().manual_seed(n_class)
        body = tf["tar_smpl_vertice_smplcoord"]
        ps = body[torch.randint(0, body.shape[0], (6000,), generator=g)] + torch.randn((6000, 3), generator=g) * 0.03
        wd2, widx, _ = orc.knn_points(ps[None], tokens[0][None], K=7)
        # through the token grid (what the fused path uses at these token counts) and by the scan over all tokens
        for grid in (True, False):
            idx, d2, rep = ops.knn_dparf(frame, ps.to(DEV), token_grid=grid)
            assert torch.equal(idx.cpu(), widx[0]) and torch.equal(d2.cpu(), wd2[0]), (n_class, grid)


@pytest.mark.parametrize("n_class,K", [(300, 7), (300, 12), (1500, 1), (6000, 3)])
def test_token_grid_knn_is_exact_everywhere(n_class, K):
    """The grid search must return the scan's (d2, index)-ordered result for EVERY point: near the body (one or two
    shells), in sparse neighbourhoods (three shells, then the fall-back scan), far outside the grid box (fall-back),
    exactly on tokens (d2 = 0) and with duplicated tokens (equal distances: the lower index first)."""
    fr = synth.make_frame(H=8, W=8, n_class=n_class, V=1, feat_hw=8, seed=8)
    tf = orc.to_torch_frame(fr)
    tok_xyz, tok_blend = orc.build_tokens(tf)
    tok_xyz = tok_xyz.clone()
    tok_xyz[5] = tok_xyz[200]                       # duplicates: ties between indices 5 and 200, 17 and 18
    tok_xyz[18] = tok_xyz[17]
    frame, _ = frame_to_device(fr, (tok_xyz, tok_blend), DEV)
    frame.c.knn = K
    g = torch.Generator().manual_seed(K)
    body = tf["tar_smpl_vertice_smplcoord"]
    near = body[torch.randint(0, body.shape[0], (3000,), generator=g)] + torch.randn((3000, 3), generator=g) * 0.02
    shell = body[torch.randint(0, body.shape[0], (1500,), generator=g)] + torch.randn((1500, 3), generator=g) * 0.15
    far = torch.randn((500, 3), generator=g) * 3.0
    on = tok_xyz[torch.randint(0, n_class, (200,), generator=g)].float()
    ps = torch.cat([near, shell, far, on, tok_xyz[[5, 17, 18, 200]].float()])
    wd2, widx, _ = orc.knn_points(ps[None], tok_xyz.float()[None], K=K)
    idx, d2, _ = ops.knn_dparf(frame, ps.to(DEV), token_grid=True)
    assert torch.equal(d2.cpu(), wd2[0])
    assert torch.equal(idx.cpu(), widx[0])
