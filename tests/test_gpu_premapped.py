"""GPU tests of the pre-mapped feature-map path (TH_FLAG_PREMAPPED, the Renderer plugin's default):
the tcgen05 pre-map GEMM over the NCHW maps against a float64 convolution, and the fused render /
density query on pre-mapped maps against the oracle and against the plain-map path.  The packed
matrices it uses are also checked on the CPU (tests/test_cabi_host.py, tests/test_chain_program.py)."""
import struct

import numpy as np
import pytest
import torch

from oracle import transhuman_oracle as orc
from tests.gpu_util import frame_to_device
from transhuman_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frame(**kw):
    fr = synth.make_frame(H=20, W=20, n_class=300, V=3, feat_hw=28, seed=21, alpha_bias_shift=-12.0, **kw)
    tf = orc.to_torch_frame(fr)
    return fr, tf, orc.build_tokens(tf)


def test_premap_kernel_matches_conv():
    fr, tf, _ = _frame()
    wts = ops.PackedWeights(fr["weights"], 3, device=DEV)
    offs = struct.unpack_from("<54Q", wts.host, 16)
    w_pre = torch.from_numpy(wts.host[offs[43]:offs[43] + 4 * 512 * 384].view(np.float32).reshape(512, 384).copy())
    b_pre = torch.from_numpy(wts.host[offs[44]:offs[44] + 4 * 512].view(np.float32).copy())
    fmap = tf["pixel_feat_map"]
    got = ops.premap_features(fmap.to(DEV), wts).cpu()
    want = torch.nn.functional.conv2d(fmap.double(), w_pre.double()[:, :, None, None], b_pre.double()).permute(0, 2, 3, 1)
    # fp16 hi/lo split operands (22 bits) with fp32 accumulation: ~1e-6 of the map's scale
    assert (got.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()


def test_premap_ragged_map_size():
    """H*W not a multiple of the 256-row super-tile: the GEMM's row guard."""
    fr = synth.make_frame(H=8, W=8, n_class=100, V=2, feat_hw=21, seed=4)
    tf = orc.to_torch_frame(fr)
    wts = ops.PackedWeights(fr["weights"], 2, device=DEV)
    offs = struct.unpack_from("<54Q", wts.host, 16)
    w_pre = torch.from_numpy(wts.host[offs[43]:offs[43] + 4 * 512 * 384].view(np.float32).reshape(512, 384).copy())
    b_pre = torch.from_numpy(wts.host[offs[44]:offs[44] + 4 * 512].view(np.float32).copy())
    fmap = tf["pixel_feat_map"]
    got = ops.premap_features(fmap.to(DEV), wts).cpu()
    want = torch.nn.functional.conv2d(fmap.double(), w_pre.double()[:, :, None, None], b_pre.double()).permute(0, 2, 3, 1)
    assert got.shape == (2, 21, 21, 512)
    assert (got.double() - want).abs().max().item() <= 4e-6 * want.abs().max().item()


@pytest.mark.parametrize("mode", ["dense", "culled"])
def test_premapped_render_matches_oracle_and_default_path(mode):
    fr, tf, tokens = _frame()
    S = 16
    want = orc.render(tf, S, tokens=tokens) if mode == "dense" else \
        orc.render_fast(tf, S, tokens=tokens, train_branch_max_rays=0)
    m = ops.TH_RENDER_DENSE if mode == "dense" else ops.TH_RENDER_MASKED
    f0, rays = frame_to_device(fr, tokens, DEV, premapped=False)
    f1, _ = frame_to_device(fr, tokens, DEV, premapped=True)
    base = ops.render_rays(f0, *rays, S, mode=m, want_raw=True)
    got = ops.render_rays(f1, *rays, S, mode=m, want_raw=True)
    torch.cuda.synchronize()
    assert (got["rgb_map"].cpu() - want["rgb_map"][0]).abs().max().item() <= 1e-4
    assert (got["acc_map"].cpu() - want["acc_map"][0]).abs().max().item() <= 1e-4
    scale = max(1.0, base["raw"].abs().max().item())
    assert (got["raw"] - base["raw"]).abs().max().item() <= 1e-5 * scale


def test_premapped_density_query():
    fr, tf, tokens = _frame()
    f0, _ = frame_to_device(fr, tokens, DEV, premapped=False)
    f1, _ = frame_to_device(fr, tokens, DEV, premapped=True)
    v = torch.from_numpy(fr["tar_smpl_vertice"]).to(DEV)
    pts = (v[::7] + 0.02 * torch.randn((v[::7].shape[0], 3), device=DEV, generator=torch.Generator(DEV).manual_seed(1)))
    a0, m0 = ops.query_density(f0, pts.contiguous())
    a1, m1 = ops.query_density(f1, pts.contiguous())
    assert torch.equal(m0, m1)
    assert (a0 - a1).abs().max().item() <= 1e-5 * max(1.0, a0.abs().max().item())


@pytest.mark.parametrize("mode", ["dense", "culled"])
def test_tmem_side_mix_matches_in_place_mix_and_oracle(mode, monkeypatch):
    """Both forms of the attention mix of the pre-mapped chain program (csrc/mlp_chain.cu): in place on the fp16
    operands by the mix warps (default) and on the accumulator side in TMEM (TH_CHAIN_MIX=tmem, fp32 combination of
    kept accumulators in the EPI_MIX epilogue)."""
    fr, tf, tokens = _frame()
    S = 16
    want = orc.render(tf, S, tokens=tokens) if mode == "dense" else \
        orc.render_fast(tf, S, tokens=tokens, train_branch_max_rays=0)
    m = ops.TH_RENDER_DENSE if mode == "dense" else ops.TH_RENDER_MASKED
    f1, rays = frame_to_device(fr, tokens, DEV, premapped=True)
    a = ops.render_rays(f1, *rays, S, mode=m, want_raw=True)
    monkeypatch.setenv("TH_CHAIN_MIX", "tmem")
    b = ops.render_rays(f1, *rays, S, mode=m, want_raw=True)
    monkeypatch.setenv("TH_CHAIN_DBG", "240")                 # and under role jitter: bit-identical
    c = ops.render_rays(f1, *rays, S, mode=m, want_raw=True)
    monkeypatch.delenv("TH_CHAIN_DBG")
    monkeypatch.delenv("TH_CHAIN_MIX")
    torch.cuda.synchronize()
    scale = max(1.0, want["raw"].abs().max().item())
    for name, got in (("in place", a), ("tmem", b)):
        assert (got["raw"].cpu() - want["raw"]).abs().max().item() <= 2e-5 * scale, name
        assert (got["rgb_map"].cpu() - want["rgb_map"][0]).abs().max().item() <= 1e-4, name
    assert torch.equal(b["raw"], c["raw"])
    assert (a["raw"] - b["raw"]).abs().max().item() <= 1e-5 * scale


@pytest.mark.parametrize("mode", ["dense", "culled"])
def test_pipelined_feature_kernel_is_bit_identical(mode, monkeypatch):
    """k_features<7, IMG, PRE, PIPE> (csrc/geometry.cu: phase 2 as a rolling software pipeline, the default for the
    id-list launches of culled rays) issues the same arithmetic per channel as the plain form: the rendered frame
    must be bit-identical whichever form a launch uses."""
    fr, tf, tokens = _frame()
    S = 16
    m = ops.TH_RENDER_DENSE if mode == "dense" else ops.TH_RENDER_MASKED
    f1, rays = frame_to_device(fr, tokens, DEV, premapped=True)
    out = {}
    for form in ("0", "1"):
        monkeypatch.setenv("TH_FEAT_PIPE", form)
        out[form] = ops.render_rays(f1, *rays, S, mode=m, want_raw=True)
    monkeypatch.delenv("TH_FEAT_PIPE")
    torch.cuda.synchronize()
    for k in ("raw", "rgb_map", "acc_map", "depth_map"):
        assert torch.equal(out["0"][k], out["1"][k]), k
    assert out["0"]["raw"].abs().max().item() > 0


@pytest.mark.parametrize("premapped", [True, False])
@pytest.mark.parametrize("white", [False, True])
@pytest.mark.parametrize("S", [16, 64, 128, 24])
def test_fused_compositing_equals_integrate_kernel(S, white, premapped, monkeypatch):
    """Dense rays on the chain schedule with S | 128: raw2outputs runs inside the chain kernel's fc_4' epilogue
    (CompositeArgs, csrc/mlp_chain.cu) -- no raw tensor, no k_integrate.  It must give the bits of the two-kernel
    form (TH_FUSE_INTEGRATE=0), with and without the raw output requested; S = 24 (not a divisor of 128) takes the
    two-kernel form either way.  400 rays x S is not a multiple of the 256-point unit: the ragged last tile."""
    fr, tf, tokens = _frame()
    f1, rays = frame_to_device(fr, tokens, DEV, premapped=premapped, white_bkgd=white)
    a = ops.render_rays(f1, *rays, S, mode=ops.TH_RENDER_DENSE)
    b = ops.render_rays(f1, *rays, S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    n_fused = ops.launch_count(reset=True)
    ops.render_rays(f1, *rays, S, mode=ops.TH_RENDER_DENSE)
    n_fused = ops.launch_count(reset=True)
    monkeypatch.setenv("TH_FUSE_INTEGRATE", "0")
    c = ops.render_rays(f1, *rays, S, mode=ops.TH_RENDER_DENSE, want_raw=True)
    n_plain = ops.launch_count(reset=True)
    monkeypatch.delenv("TH_FUSE_INTEGRATE")
    torch.cuda.synchronize()
    for k in ("rgb_map", "acc_map", "depth_map"):
        assert torch.equal(a[k], c[k]), k
        assert torch.equal(b[k], c[k]), k
    assert torch.equal(b["raw"], c["raw"])
    assert n_plain - n_fused == (1 if 128 % S == 0 else 0)   # k_integrate is gone from the fused form
    want = orc.render(tf, S, tokens=tokens, white_bkgd=white) if S == 16 else None
    if want is not None:
        assert (a["rgb_map"].cpu() - want["rgb_map"][0]).abs().max().item() <= 1e-4
