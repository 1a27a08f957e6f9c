"""Host-side checks of the C ABI that need no GPU: the library loads, exports
every symbol include/transhuman_b200.h declares, packs weights correctly, and
the product path refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest
import torch

from tests.conftest import ROOT
from transhuman_b200 import _lib, ops, synth


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "transhuman_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(th_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert b"sm_100a" in lib.th_version()


def test_struct_layouts_match_header():
    # ThFrame: 11 pointers, 6 int32, 4 float, 1 uint32 -> 88 + 44 = 132 -> padded to 136
    assert C.sizeof(_lib.ThFrame) == 136
    assert C.sizeof(_lib.ThRays) == 56
    assert C.sizeof(_lib.ThOut) == 48
    assert C.sizeof(_lib.ThWeightsF32) == 32 * 8


def test_pack_weights_folds_layers(lib):
    w = synth.make_weights(seed=5)
    for V in (1, 3):
        pw_bytes = lib.th_packed_weights_bytes(V)
        assert pw_bytes > 3_000_000
        wstruct = _lib.ThWeightsF32()
        keep = []
        for cname, rname in _lib.WEIGHT_FIELDS:
            for suf, key in (("w", rname + ".weight"), ("b", rname + ".bias")):
                a = np.ascontiguousarray(w[key].reshape(-1))
                keep.append(a)
                setattr(wstruct, f"{cname}_{suf}", a.ctypes.data)
        blob = np.zeros(pw_bytes, dtype=np.uint8)
        assert lib.th_pack_weights(C.byref(wstruct), V, blob.ctypes.data, pw_bytes) == 0
        # too small a buffer -> TH_EWORKSPACE with a message
        assert lib.th_pack_weights(C.byref(wstruct), V, blob.ctypes.data, 16) == -2
        assert b"need" in lib.th_last_error()
        magic, nviews, total = struct.unpack_from("<IiQ", blob, 0)
        assert magic == 0x35574854 and nviews == V and total == pw_bytes
        offs = struct.unpack_from("<54Q", blob, 16)
        f32 = lambda off, n: blob[off:off + 4 * n].view(np.float32)
        # fc_0: K padded 255 -> 256 with a zero column
        fc0 = f32(offs[0], 256 * 256).reshape(256, 256)
        assert np.array_equal(fc0[:, :255], w["fc_0.weight"]) and np.all(fc0[:, 255] == 0)
        # W_v = [value_embed_1 | value_embed_0], b = b1 + b0
        wv = f32(offs[8], 256 * 512).reshape(256, 512)
        assert np.array_equal(wv[:, :256], w["spatial_key_value_1.value_embed.weight"])
        assert np.array_equal(wv[:, 256:], w["spatial_key_value_0.value_embed.weight"])
        # fc_3 with the view mean folded into K
        w3 = f32(offs[14], 256 * 256 * V).reshape(256, 256 * V)
        for v in range(V):
            np.testing.assert_allclose(w3[:, 256 * v:256 * (v + 1)], w["fc_3.weight"] / V, rtol=1e-7)
        # W_t = [fc_4/V ... | fc_4 @ rgb_res_1], b_t = fc_4 @ b_r1 + b_4
        ld = 128 * V + 384
        wt = f32(offs[22], 128 * ld).reshape(128, ld)
        want = w["fc_4.weight"].astype(np.float64) @ w["rgb_res_1.weight"].astype(np.float64)
        np.testing.assert_allclose(wt[:, 128 * V:], want, rtol=0, atol=1e-7)
        bt = f32(offs[23], 128)
        wantb = w["fc_4.weight"].astype(np.float64) @ w["rgb_res_1.bias"].astype(np.float64) + w["fc_4.bias"]
        np.testing.assert_allclose(bt, wantb, rtol=0, atol=1e-7)
        # fp16 hi/lo tile images (128B-swizzled, per 64-wide k-block) reconstruct fc_0 to ~2^-22
        # layers folded across a missing non-linearity (tensor-core schedule)
        f64 = lambda k: w[k].astype(np.float64)
        fc1f = f32(offs[26], 256 * 512).reshape(256, 512)
        np.testing.assert_allclose(fc1f[:, :256], f64("fc_1.weight") @ f64("spatial_key_value_1.value_embed.weight"),
                                   rtol=0, atol=1e-7)
        np.testing.assert_allclose(fc1f[:, 256:], f64("fc_1.weight") @ f64("spatial_key_value_0.value_embed.weight"),
                                   rtol=0, atol=1e-7)
        bv = f64("spatial_key_value_1.value_embed.bias") + f64("spatial_key_value_0.value_embed.bias")
        np.testing.assert_allclose(f32(offs[27], 256), f64("fc_1.weight") @ bv + f64("fc_1.bias"), rtol=0, atol=1e-7)
        gvf = f32(offs[28], 128 * 704).reshape(128, 704)
        v1, v2 = f64("view_fc.weight")[:, :256], w["view_fc.weight"][:, 256:]
        np.testing.assert_allclose(gvf[:, :256], v1 @ f64("feature_fc.weight"), rtol=0, atol=1e-7)
        np.testing.assert_allclose(gvf[:, 256:640], v1 @ f64("rgb_res_0.weight"), rtol=0, atol=1e-7)
        assert np.array_equal(gvf[:, 640:667], v2) and np.all(gvf[:, 667:] == 0)
        np.testing.assert_allclose(f32(offs[29], 128),
                                   v1 @ (f64("feature_fc.bias") + f64("rgb_res_0.bias")) + f64("view_fc.bias"),
                                   rtol=0, atol=1e-7)
        h_fc0 = offs[30]
        # per k-block: [half 0: hi (128 rows) | lo] [half 1: hi | lo]  (one contiguous copy per CTA of a pair)
        img = blob[h_fc0:h_fc0 + 4 * 65536].view(np.float16).astype(np.float32).reshape(4, 2, 2, 128, 64)
        img = img.transpose(0, 2, 1, 3, 4).reshape(4, 2, 256, 64)          # -> (kb, plane, row, 64)
        n = np.arange(256)[:, None]
        kk = np.arange(64)[None, :]
        col = (((kk >> 3) ^ (n & 7)) << 3) + (kk & 7)          # position of element kk inside the swizzled row
        rec = np.concatenate([np.take_along_axis(img[kb, 0], col, 1) + np.take_along_axis(img[kb, 1], col, 1)
                              for kb in range(4)], axis=1)
        # the image holds W * 2^e, max|W| 2^e in (2^13, 2^14]; the header lists (image offset, 2^-e) per matrix
        hdr_tail = 16 + 57 * 8
        img_off = struct.unpack_from("<24Q", blob, hdr_tail)
        inv_scale = struct.unpack_from("<24f", blob, hdr_tail + 24 * 8)
        n_img = struct.unpack_from("<i", blob, hdr_tail + 24 * 8 + 24 * 4)[0]
        assert 16 <= n_img <= 24 and h_fc0 in img_off[:n_img]
        inv = inv_scale[img_off.index(h_fc0)]
        assert 2.0 ** 13 < np.abs(fc0).max() / inv <= 2.0 ** 14 and np.log2(inv) == np.round(np.log2(inv))
        assert np.abs(rec * inv - fc0).max() <= 2.0 ** -21 * np.abs(fc0).max()
        # ... and keeps 22 bits for SMALL weights too (unscaled, the fp16 lo plane flushes below 6e-8)
        small = np.abs(fc0) < 1e-3 * np.abs(fc0).max()
        small &= fc0 != 0
        assert small.any() and np.all(np.abs(rec * inv - fc0)[small] <= 2.0 ** -20 * np.abs(fc0[small]) + 2.0 ** -38)


def _pack(lib, w, V):
    wstruct = _lib.ThWeightsF32()
    keep = []
    for cname, rname in _lib.WEIGHT_FIELDS:
        for suf, key in (("w", rname + ".weight"), ("b", rname + ".bias")):
            a = np.ascontiguousarray(w[key].reshape(-1))
            keep.append(a)
            setattr(wstruct, f"{cname}_{suf}", a.ctypes.data)
    blob = np.zeros(lib.th_packed_weights_bytes(V), dtype=np.uint8)
    assert lib.th_pack_weights(C.byref(wstruct), V, blob.ctypes.data, blob.size) == 0
    return blob


# order of the (w, b) offset pairs in PackedHeader (kernels.cuh), then the image offsets
_MATS = ["fc0", "ar0", "k0", "k1", "v", "fc1", "fc2", "fc3m", "afc", "f", "view", "t", "rgb", "fc1f", "gvf",
         "pre", "gvfp", "tp", "xid"]


def _matrices(blob, V):
    offs = struct.unpack_from("<54Q", blob, 16)
    # the image offsets of the first 13 matrices sit between gvf and pre in the header
    pair = {name: (offs[2 * i], offs[2 * i + 1]) for i, name in enumerate(_MATS[:15])}
    pair.update({name: (offs[43 + 2 * i], offs[43 + 2 * i + 1]) for i, name in enumerate(_MATS[15:])})
    shapes = {"fc0": (256, 256), "ar0": (256, 384), "k0": (128, 256), "k1": (128, 256), "v": (256, 512),
              "fc1": (256, 256), "fc2": (256, 256), "fc3m": (256, 256 * V), "afc": (1, 256), "f": (256, 640),
              "view": (128, 320), "t": (128, 128 * V + 384), "rgb": (3, 128), "fc1f": (256, 512), "gvf": (128, 704),
              "pre": (512, 384), "gvfp": (128, 448), "tp": (128, 128 * V + 128), "xid": (256, 256)}
    # fc_1' cut into S / X parts and halves of 128 rows: (w, b) pairs after img_off / img_inv_scale / n_img
    tail = struct.unpack_from("<12Q", blob, 16 + 57 * 8 + 24 * 8 + 24 * 4 + 8)
    for i, name in enumerate(["fc1s0", "fc1s1", "fc1x0", "fc1x1"]):
        pair[name] = (tail[2 * i], tail[2 * i + 1])
        shapes[name] = (128, 256)
    out = {}
    for name, (ow, ob) in pair.items():
        n, k = shapes[name]
        out[name] = (torch.from_numpy(blob[ow:ow + 4 * n * k].view(np.float32).reshape(n, k).astype(np.float64)),
                     torch.from_numpy(blob[ob:ob + 4 * n].view(np.float32).astype(np.float64)))
    return out


@pytest.mark.parametrize("V", [1, 3])
def test_packed_programs_reproduce_the_network(lib, V):
    """The layer programs the tensor-core schedules run, evaluated in float64 straight from the packed blob
    -- the default one (folded fc_1' and view_fc') and the pre-mapped one (DESIGN.md section 5, round-2 item 1:
    alpha_res_0 / rgb_res_0 / rgb_res_1 applied to the maps, identity blocks in view_fc' and fc_4') --
    against the oracle's MLP in float64.  Checks every fold, transpose, 1/V and bias of th_pack_weights."""
    from oracle import transhuman_oracle as orc
    w = synth.make_weights(seed=7)
    m = _matrices(_pack(lib, w, V), V)
    g = torch.Generator().manual_seed(3)
    P = 64
    rep = torch.randn((V, 255, P), generator=g, dtype=torch.float64)
    pix = torch.randn((V, 384, P), generator=g, dtype=torch.float64)
    vd = torch.randn((1, P, 27), generator=g, dtype=torch.float64)
    w64 = {k: torch.from_numpy(np.asarray(v)).double() for k, v in w.items()}
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        want = orc.mlp_forward(w64, rep, pix, vd, progressive=False)[0]            # (P,4) = rgb, alpha
    finally:
        torch.set_default_dtype(old)

    lin = lambda name, x: x @ m[name][0].T + m[name][1]                          # x (rows, K)
    relu = torch.relu
    rep_r = torch.cat([rep.permute(0, 2, 1), torch.zeros((V, P, 1), dtype=torch.float64)], -1)   # (V,P,256)
    pix_r = pix.permute(0, 2, 1)                                                   # (V,P,384)
    vd_r = torch.cat([vd[0], torch.zeros((P, 37), dtype=torch.float64)], -1)       # (P,64)

    def program(premapped):
        S = relu(lin("fc0", rep_r))
        if premapped:
            pre = lin("pre", pix_r)                                                # what the blended pre-mapped map holds
            X = lin("xid", relu(pre[..., :256]))
            P2, R = pre[..., 256:384], pre[..., 384:].sum(0)
        else:
            X = relu(lin("ar0", pix_r))
        KP, KS = lin("k0", X), lin("k1", S)
        A = torch.softmax(torch.einsum("ipc,jpc->pij", KP, KS) / np.sqrt(128.0), dim=1)   # over i
        XT = torch.einsum("pij,ipc->jpc", A, X)
        inter = relu(lin("fc2", relu(lin("fc1f", torch.cat([S, XT], -1)))))
        alpha = lin("afc", relu(lin("fc3m", torch.cat(list(inter), -1))))
        if premapped:
            G = relu(lin("gvfp", torch.cat([inter, P2, vd_r.expand(V, P, 64)], -1)))
            T = relu(lin("tp", torch.cat(list(G) + [R], -1)))
        else:
            G = relu(lin("gvf", torch.cat([inter, pix_r, vd_r.expand(V, P, 64)], -1)))
            T = relu(lin("t", torch.cat(list(G) + [pix_r.mean(0)], -1)))
        return torch.cat([lin("rgb", T), alpha], -1)

    scale = want.abs().max().item()
    for premapped in (False, True):
        got = program(premapped)
        assert (got - want).abs().max().item() <= 2e-6 * max(1.0, scale), premapped


def test_workspace_bytes(lib):
    a = lib.th_workspace_bytes(1000, 3, 6890)
    b = lib.th_workspace_bytes(100000, 3, 6890)
    c = lib.th_workspace_bytes(262144 * 64, 3, 6890)
    assert 0 < a < b < c < 8 << 30


def test_frame_workspace_bytes_follows_the_schedule(lib):
    """The layer-chained schedule on pre-mapped maps keeps no per-point activation buffers: 8.25 KB per point of a
    chunk + the fixed per-CTA scratch instead of 26.9 KB per point; every other flag
    combination gets th_workspace_bytes' figure."""
    NP = 262144 * 64
    f = _lib.ThFrame()
    f.n_views, f.n_verts = 3, 6890
    full = lib.th_workspace_bytes(NP, 3, 6890)
    for flags in (0, ops.TH_FLAG_SIMT_MLP, ops.TH_FLAG_LAYERWISE, ops.TH_FLAG_PREMAPPED | ops.TH_FLAG_SIMT_MLP):
        f.flags = flags
        assert lib.th_frame_workspace_bytes(C.byref(f), NP, 1) == full
    f.flags = ops.TH_FLAG_PREMAPPED
    compact = lib.th_frame_workspace_bytes(C.byref(f), NP, 1)
    chunk = 284160
    per_point = (3 * (256 + 384) + 128 + 64) * 4            # rep, [X | P2] per view; R and the view direction per point
    scratch = 148 * (2 * 3 + 3 // 2 + 0.5) * 128 * 1024       # 148 CTAs x (2 V + V / 2) activation tiles of 128 KB
    fixed = NP * (1 + 1 + 4 + 16)                             # ray flags, mask, id list, raw
    assert 0 <= compact - (chunk * per_point + scratch + fixed) < 32 << 20, compact   # + the cull / token grids
    assert compact < 0.4 * full
    assert lib.th_frame_workspace_bytes(C.byref(f), NP, 0) < compact   # without the cull / token grids
    f.n_views = 4                                              # V = 4: no chain schedule -> the full carve
    assert lib.th_frame_workspace_bytes(C.byref(f), NP, 1) == lib.th_workspace_bytes(NP, 4, 6890)


def test_no_cpu_fallback():
    """CPU tensors are rejected; nothing in the package imports the oracle."""
    with pytest.raises(ValueError, match="CUDA"):
        ops.view_embed(torch.zeros((4, 3)))
    pkg = os.path.join(ROOT, "transhuman_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    if not torch.cuda.is_available():
        lib = _lib.load()
        out = np.zeros(27, dtype=np.float32)
        # a compute entry point without a device reports TH_ECUDA, it does not compute on the host
        rc = lib.th_view_embed(C.c_void_p(out.ctypes.data), 1, C.c_void_p(out.ctypes.data), None)
        assert rc == -3 and lib.th_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.TransHumanLibraryError, match="no CPU fallback"):
        _lib.load()


def test_synth_frame_is_deterministic():
    a = synth.make_frame(H=8, W=8, n_class=100, V=2, feat_hw=8, seed=3)
    b = synth.make_frame(H=8, W=8, n_class=100, V=2, feat_hw=8, seed=3)
    for k in ("ray_o", "ray_d", "holder", "pixel_feat_map", "tar_smpl_vertice", "pc2voxel_ind"):
        assert np.array_equal(a[k], b[k])
    assert sorted(np.unique(a["pc2voxel_ind"])) == list(range(100))
    v = a["tar_smpl_vertice"]
    assert v.shape == (6890, 3) and abs(v[:, 0]).max() < 0.95 and v[:, 1].min() > -1.25


def test_prologue_entry_points_validate_their_arguments(lib):
    """Argument errors of the SURVEY 8f entry points come back as TH_EINVAL (-1) with a message -- no device needed,
    nothing is launched -- and the workspace queries are pure host arithmetic."""
    enc = _lib.ThEncoderTail()
    dummy = np.zeros(64, dtype=np.float32)
    p = C.c_void_p(dummy.ctypes.data)
    assert lib.th_premap_from_latents(C.byref(enc), p, p, p, 1 << 20, None) == -1        # null latents
    assert b"th_premap_from_latents" in lib.th_last_error() or b"enc_views" in lib.th_last_error()
    for i in range(3):
        enc.latent[i], enc.lat_h[i], enc.lat_w[i] = dummy.ctypes.data, 4, 4
    enc.images, enc.color_w, enc.color_b = dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data
    enc.n_views, enc.h, enc.w = 9, 8, 8                                                   # more than TH_MAX_VIEWS
    assert lib.th_premap_from_latents(C.byref(enc), p, p, p, 1 << 20, None) == -1
    enc.n_views = 3
    need = lib.th_premap_from_latents_workspace_bytes(C.byref(enc))
    assert need >= 3 * 16 * 512 * 4 and need % 256 == 0                                   # three 4 x 4 levels of 512 floats
    assert lib.th_premap_from_latents(C.byref(enc), p, p, p, need - 256, None) == -1      # workspace too small
    assert lib.th_paint_group_latents_workspace_bytes(3, 6890, 300) >= 3 * 6890 * 384 * 4
    assert lib.th_paint_group_latents(C.byref(enc), p, p, 1.0, 1.0, p, 6890, p, p, p, None, p, p, 300, p, p, 16, None) == -1
    # attention: vit_tiny's head size only; workspace = the operand images (query tiles in pairs of 128, key tiles of 64)
    assert lib.th_vit_attention(p, 1, 8, 3, 32, 1.0, p, p, 1 << 20, None) == -1
    assert b"head_dim" in lib.th_last_error()
    assert lib.th_vit_attention(p, 1, 0, 3, 64, 1.0, p, p, 1 << 20, None) == -1
    ws = lib.th_vit_attention_workspace_bytes(3, 6000, 3)
    assert ws == 9 * (24 * 2 * 32768 + 94 * 32768)
    assert lib.th_vit_attention(p, 3, 6000, 3, 64, 0.125, p, C.c_void_p(256), ws - 256, None) == -1
