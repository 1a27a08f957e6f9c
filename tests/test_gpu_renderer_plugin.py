"""The drop-in boundary end to end on the GPU: `transhuman_b200.renderer.Renderer`
and `mesh_renderer.Renderer` driven exactly like the reference drives its own
(`Renderer(net)`, `render_fast(batch)`, `render(batch)`; run.py:52,109,156), with a
small stand-in `net` (encoder / ViT / state_dict), checked against the CPU oracle
fed with the same prologue outputs."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import transhuman_oracle as orc
from transhuman_b200 import synth
from transhuman_b200.mesh_renderer import Renderer as MeshRenderer
from transhuman_b200.renderer import Renderer, segment_mean

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Encoder(nn.Module):
    """Stand-in for SpatialEncoder (encoder.py:97-155): same outputs and scales."""

    def __init__(self):
        super().__init__()
        self.pix = nn.Conv2d(3, 384, 3, padding=1)
        self.red = nn.Conv2d(384, 192, 1)

    def forward(self, x):
        p = torch.tanh(self.pix(x)) * 2.0
        h = self.red(p)
        sc = np.array([p.shape[-1], p.shape[-2]])
        sc = sc / (sc - 1) * 2.0
        return h, sc, p, sc


class _ViT(nn.Module):
    """Stand-in for vit_tiny: a token-wise map that also uses the positional input."""

    def __init__(self):
        super().__init__()
        self.lin = nn.Linear(192, 192)
        self.pe = nn.Linear(3, 192)
        self.embed_dim = 192

    def forward(self, tokens, pe, mask=None):
        return torch.tanh(self.lin(tokens) + self.pe(pe))


class _Net(nn.Module):
    def __init__(self, weights):
        super().__init__()
        self.encoder = _Encoder()
        self.ViT = _ViT()
        for name, arr in weights.items():           # reference state_dict names (Conv1d weights are (out,in,1))
            t = torch.from_numpy(arr)
            self.register_buffer(name.replace(".", "__"), t.view(*t.shape, 1) if name.endswith("weight") else t)
        self._names = list(weights)

    def state_dict(self, *a, **k):
        sd = super().state_dict(*a, **k)
        for name in self._names:
            sd[name] = sd.pop(name.replace(".", "__"))
        return sd


class _Cfg:
    N_samples = 24
    num_class = 300
    KNN = 7
    KNN_DIST_ALPHA = 0.5
    white_bkgd = False
    perturb = 0.
    rasterize = True
    time_steps = 1
    voxel_size = [0.005, 0.005, 0.005]
    mesh_th = 5


def _setup(H, seed, shift):
    fr = synth.make_frame(H=H, W=H, n_class=300, V=3, feat_hw=32, seed=seed, with_feature_maps=False,
                          alpha_bias_shift=shift)
    torch.manual_seed(seed)
    net = _Net(fr["weights"]).to(DEV)
    net.train()                                        # run.py:29

    def t(a, dev=DEV):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    g = torch.Generator().manual_seed(seed)
    imgs = torch.rand((1, 3, 3, 32, 32), generator=g)
    viz = torch.rand((1, 3, synth.N_VERTS), generator=g) > 0.3
    batch = {
        "ray_o": t(fr["ray_o"])[None], "ray_d": t(fr["ray_d"])[None], "near": t(fr["near"])[None],
        "far": t(fr["far"])[None], "tar_smpl_vertice": t(fr["tar_smpl_vertice"])[None],
        "tar_smpl_vertice_smplcoord": t(fr["tar_smpl_vertice_smplcoord"])[None],
        "Rh": t(fr["Rh"])[None], "Th": t(fr["Th"])[None], "blend_mtx": t(fr["blend_mtx"])[None],
        "input_imgs": [imgs.to(DEV)], "input_R": [t(fr["input_R"])[None]], "input_T": [t(fr["input_T"])[None]],
        "input_K": [t(fr["input_K"])[None]], "input_smpl_vertice": [t(fr["tar_smpl_vertice"])[None]],
        "input_vizmaps": [viz.to(DEV)],
    }
    renderer = Renderer(net, cfg=_Cfg(), pc2voxel_ind=fr["pc2voxel_ind"], vertex_can=synth.make_body(0))
    return fr, net, batch, renderer


def _oracle_frame(fr, renderer, batch):
    """Oracle inputs = the prologue outputs the plugin computed (tokens, feature maps)."""
    frame = renderer.prepare_frame(batch)
    tf = orc.to_torch_frame(fr)
    tf["holder"] = frame.holder.cpu()
    # the plugin keeps only the PRE-MAPPED maps; the oracle takes the encoder's own output
    with torch.no_grad():
        images = batch["input_imgs"][0].reshape(-1, *batch["input_imgs"][0].shape[2:])
        tf["pixel_feat_map"] = renderer.net.encoder(images)[2].contiguous().cpu()
    tokens = (frame.tok_xyz.cpu(), torch.cat([frame.tok_rot.cpu().double(), torch.zeros(300, 3, 1, dtype=torch.float64)],
                                             dim=2))
    tok_blend = torch.zeros((300, 4, 4), dtype=torch.float64)
    tok_blend[:, :3, :3] = frame.tok_rot.cpu().double()
    return tf, (tokens[0], tok_blend)


def test_renderer_render_and_render_fast_match_oracle():
    fr, net, batch, renderer = _setup(28, 31, 5.0)
    tf, tokens = _oracle_frame(fr, renderer, batch)
    S = _Cfg.N_samples
    # token construction: float64 segment mean vs the reference's per-cluster mean
    ref_xyz = orc.voxelization(tf["pc2voxel_ind"].long(), tf["tar_smpl_vertice_smplcoord"], 300)
    assert (tokens[0] - ref_xyz).abs().max().item() <= 2e-7
    out = renderer.render(batch)
    want = orc.render(tf, S, tokens=tokens)
    assert out["rgb_map"].shape == (1, 28 * 28, 3) and out["acc_map"].shape == (1, 28 * 28)
    last = want["raw"][:, -1, 3]
    keep = ~(last.abs() < 1e-3)
    assert (out["rgb_map"][0].cpu() - want["rgb_map"][0])[keep].abs().max().item() <= 1e-4
    assert (out["acc_map"][0].cpu() - want["acc_map"][0])[keep].abs().max().item() <= 1e-4
    assert (out["depth_map"][0].cpu() - want["depth_map"][0])[keep].abs().max().item() <= 3.5e-4
    outf = renderer.render_fast(batch)
    wantf = orc.render_fast(tf, S, tokens=tokens)          # <= 2400 surviving rays: the reference's un-masked branch
    assert renderer.last_counters[1] == int((wantf["valid_pts_mask"][0].any(dim=1)).sum())
    lastf = wantf["raw"][:, -1, 3]
    keepf = ~((lastf.abs() < 1e-3) & (lastf != 0))
    assert (outf["rgb_map"][0].cpu() - wantf["rgb_map"][0])[keepf].abs().max().item() <= 1e-4
    assert (outf["depth_map"][0].cpu() - wantf["depth_map"][0])[keepf].abs().max().item() <= 3.5e-4
    assert outf["rgb_map"].abs().max() > 0.05 and out["acc_map"].max() > 0.5
    # keys and batch are left as the caller passed them
    assert set(outf) == {"rgb_map", "acc_map", "depth_map"} and batch["ray_o"].shape == (1, 28 * 28, 3)


def test_renderer_is_forward_only():
    fr, net, batch, renderer = _setup(8, 3, 0.0)
    renderer.cfg.perturb = 1.0
    with pytest.raises(NotImplementedError):
        renderer.render(batch)
    renderer.cfg.perturb = 0.
    with torch.no_grad():
        renderer.render(batch)


def test_mesh_renderer_cube():
    fr, net, batch, _ = _setup(8, 5, 0.0)
    mr = MeshRenderer(net, cfg=_Cfg(), pc2voxel_ind=fr["pc2voxel_ind"], vertex_can=synth.make_body(0))
    grid = synth.make_grid_points(fr, 40)
    mb = dict(batch)
    mb["pts"] = torch.from_numpy(grid)[None].to(DEV)
    mb["can_bounds"] = torch.tensor([[grid.reshape(-1, 3).min(0), grid.reshape(-1, 3).max(0)]])
    ret = mr.render(mb)
    assert ret["cube"].shape == (60, 60, 60) and np.all(ret["cube"][:10] == 0)
    # the mesh step (if_mesh_renderer.py:98-109) through th_marching_cubes -- mcubes is not installed here -- against
    # the numpy oracle on the same cube, with the reference's index -> world transform
    from oracle import marching_cubes as omc
    cfg = _Cfg()
    v, f = ret["mesh"].vertices, ret["mesh"].faces
    wv, wf = omc.marching_cubes(ret["cube"], cfg.mesh_th)
    assert len(f) > 50 and np.array_equal(f, wf)
    lb = grid.reshape(-1, 3).min(0) - 10 * np.array(cfg.voxel_size)
    assert np.allclose(v, wv.astype(np.float64) * np.array(cfg.voxel_size) + lb, atol=1e-12)
    tf, tokens = _oracle_frame(fr, mr, mb)
    walpha, wmask = orc.query_density(tf, torch.from_numpy(grid.reshape(-1, 3)), tokens=tokens)
    got = torch.from_numpy(ret["cube"][10:-10, 10:-10, 10:-10]).reshape(-1)
    assert int(wmask.sum()) > 100
    assert (got - walpha).abs().max().item() <= 2e-5 * max(1.0, walpha.abs().max().item())
