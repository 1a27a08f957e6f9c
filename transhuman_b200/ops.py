"""Torch-tensor front end of the C ABI.

PyTorch is plumbing here: it owns device memory and the CUDA stream; every
computation happens in ``libtranshuman_b200.so``.  All tensors must be CUDA,
contiguous, fp32 unless stated.  No function here has a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import numpy as np
import torch

from . import _lib
from ._lib import (TH_FLAG_LAYERWISE, TH_FLAG_PREMAPPED, TH_FLAG_SIMT_MLP, TH_FLAG_WHITE_BKGD, TH_RENDER_DENSE, TH_RENDER_FAST, TH_RENDER_MASKED,
                   ThFrame, ThOut, ThRays, ThWeightsF32)

__all__ = ["PackedWeights", "Frame", "render_rays", "query_density", "sample_points", "cull_knn1", "cull_grid",
           "world2smpl", "view_embed", "pixel_gather", "knn_dparf", "mlp_raw", "integrate", "nchw_to_nhwc",
           "premap_features", "vit_attention", "PackedLinear", "marching_cubes", "EncoderTail", "premap_from_latents", "paint_group_latents", "ClusterIndex", "paint_group",
           "group_mean", "generate_rays", "near_far",
           "launch_count", "TH_RENDER_DENSE", "TH_RENDER_MASKED", "TH_RENDER_FAST"]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (the query path has no CPU implementation)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_workspaces: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces[key] = None
        ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def release_workspaces():
    _workspaces.clear()


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().th_launch_count(1 if reset else 0))


PROFILE_CATEGORIES = ("cull", "features", "gemm", "pointwise", "integrate", "premap", "prologue")


def profile_start():
    _lib.check(_lib.load().th_profile_start(), "th_profile_start")


def profile_stop() -> dict:
    """-> {category: (milliseconds, launches)} summed since profile_start()."""
    ncat = len(PROFILE_CATEGORIES)
    ms = (C.c_double * ncat)()
    n = (C.c_int64 * ncat)()
    _lib.check(_lib.load().th_profile_stop(ms, n, ncat), "th_profile_stop")
    return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(PROFILE_CATEGORIES)}


class PackedWeights:
    """The per-point network's 16 Conv1d layers packed for the kernels
    (``th_pack_weights``).  ``state`` maps reference state_dict names
    (``fc_0.weight`` ... ``rgb_fc.bias``, cross_transformer.py:97-126) to arrays
    or tensors; extra keys (encoder, ViT, the dead ``xyzc_net``) are ignored."""

    def __init__(self, state: Mapping, n_views: int, device="cuda"):
        lib = _lib.load()
        self.n_views = int(n_views)
        keep = []
        w = ThWeightsF32()
        for cname, rname in _lib.WEIGHT_FIELDS:
            for suf, key in (("w", rname + ".weight"), ("b", rname + ".bias")):
                if key not in state:
                    raise KeyError(f"state dict lacks {key}")
                a = state[key]
                if isinstance(a, torch.Tensor):
                    a = a.detach().cpu().numpy()
                a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
                keep.append(a)
                setattr(w, f"{cname}_{suf}", a.ctypes.data)
        nbytes = lib.th_packed_weights_bytes(self.n_views)
        if nbytes == 0:
            raise ValueError(f"unsupported view count {n_views}")
        host = np.zeros(nbytes, dtype=np.uint8)
        _lib.check(lib.th_pack_weights(C.byref(w), self.n_views, host.ctypes.data, nbytes), "th_pack_weights")
        self.host = host
        self.blob = torch.from_numpy(host).to(device)


class Frame:
    """Per-frame state of the query path (``ThFrame``): what the torch prologue
    (encoder, painting, grouping, ViT -- out of scope) hands to the kernels."""

    def __init__(self, *, holder, tok_xyz, tok_rot, verts, feat_nhwc, cam_R, cam_T, cam_K, Rh, Th,
                 weights: PackedWeights, uv_scale, knn: int = 7, knn_dist_alpha: float = 0.5,
                 cull_radius: float = 0.1, white_bkgd: bool = False, simt_mlp: bool = False,
                 layerwise: bool = False, premapped: bool = False):
        self.holder = _f32(holder, "holder")
        V, n_tok, c = self.holder.shape
        assert c == 192, "token width must be 192"
        self.tok_xyz = _f32(tok_xyz, "tok_xyz").view(n_tok, 3)
        self.tok_rot = _f32(tok_rot, "tok_rot").view(n_tok, 3, 3)
        self.verts = _f32(verts, "verts").view(-1, 3)
        self.feat = _f32(feat_nhwc, "feat_nhwc")
        # premapped: ``feat_nhwc`` is the output of ``premap_features`` (512 channels)
        assert self.feat.dim() == 4 and self.feat.shape[0] == V and self.feat.shape[3] == (512 if premapped else 384), \
            "feature maps must be (V,H,W,384) channel-last"
        self.cam_R = _f32(cam_R, "cam_R").view(V, 3, 3)
        self.cam_T = _f32(cam_T, "cam_T").view(V, 3)
        self.cam_K = _f32(cam_K, "cam_K").view(V, 3, 3)
        self.Rh = _f32(Rh, "Rh").view(3, 3)
        self.Th = _f32(Th, "Th").view(3)
        assert weights.n_views == V, "weights were packed for a different view count"
        self.weights = weights
        self.V, self.n_tok = V, n_tok
        f = ThFrame()
        f.tok_feat, f.tok_xyz, f.tok_rot = self.holder.data_ptr(), self.tok_xyz.data_ptr(), self.tok_rot.data_ptr()
        f.verts, f.feat = self.verts.data_ptr(), self.feat.data_ptr()
        f.cam_R, f.cam_T, f.cam_K = self.cam_R.data_ptr(), self.cam_T.data_ptr(), self.cam_K.data_ptr()
        f.Rh, f.Th = self.Rh.data_ptr(), self.Th.data_ptr()
        f.weights = weights.blob.data_ptr()
        f.n_views, f.n_tok, f.n_verts = V, n_tok, self.verts.shape[0]
        f.feat_h, f.feat_w = self.feat.shape[1], self.feat.shape[2]
        f.knn = knn
        f.uv_scale_x, f.uv_scale_y = float(uv_scale[0]), float(uv_scale[1])
        f.knn_dist_alpha, f.cull_radius = knn_dist_alpha, cull_radius
        f.flags = ((TH_FLAG_WHITE_BKGD if white_bkgd else 0) | (TH_FLAG_SIMT_MLP if simt_mlp else 0) |
                   (TH_FLAG_LAYERWISE if layerwise else 0) | (TH_FLAG_PREMAPPED if premapped else 0))
        self.c = f

    @property
    def device(self):
        return self.holder.device

    def set_flag(self, flag: int, on: bool):
        self.c.flags = (self.c.flags | flag) if on else (self.c.flags & ~flag)


def uv_scale_for(feat_h: int, feat_w: int, image_h: int, image_w: int):
    """``feat_scale / image_shape`` of the reference, evaluated in float64
    (encoder.py:149-150, if_clight_renderer.py:193): x is scaled by index 0 =
    (W/(W-1)*2)/H_img and y by index 1 = (H/(H-1)*2)/W_img."""
    fs = np.array([feat_w, feat_h]) / (np.array([feat_w, feat_h]) - 1) * 2.0
    sc = fs / np.array([image_h, image_w])
    return np.float32(sc[0]), np.float32(sc[1])


def t_vals_for(S: int, device) -> torch.Tensor:
    """``torch.linspace(0., 1., steps=S)`` evaluated on the CPU like the
    reference (if_clight_renderer.py:273), then moved to the device."""
    return torch.linspace(0., 1., steps=S).to(device)


def _rays(ray_o, ray_d, near, far, S, t_vals=None):
    ray_o = _f32(ray_o, "ray_o").view(-1, 3)
    ray_d = _f32(ray_d, "ray_d").view(-1, 3)
    near = _f32(near, "near").view(-1)
    far = _f32(far, "far").view(-1)
    N = ray_o.shape[0]
    assert ray_d.shape[0] == N and near.shape[0] == N and far.shape[0] == N
    t = t_vals_for(S, ray_o.device) if t_vals is None else _f32(t_vals, "t_vals")
    r = ThRays()
    r.ray_o, r.ray_d, r.near_, r.far_, r.t_vals = (x.data_ptr() for x in (ray_o, ray_d, near, far, t))
    r.n_rays, r.n_samples = N, S
    return r, (ray_o, ray_d, near, far, t)


def render_rays(frame: Frame, ray_o, ray_d, near, far, S: int, mode: int = TH_RENDER_DENSE, want_raw: bool = False,
                want_mask: bool = False, t_vals=None) -> dict:
    """Fused a1-a11.  Returns ``rgb_map (N,3)``, ``acc_map (N)``, ``depth_map
    (N)`` and ``counters`` (points in radius, surviving rays, points evaluated)."""
    lib = _lib.load()
    r, keep = _rays(ray_o, ray_d, near, far, S, t_vals)
    N, dev = r.n_rays, keep[0].device
    out = {"rgb_map": torch.empty((N, 3), device=dev), "acc_map": torch.empty((N,), device=dev),
           "depth_map": torch.empty((N,), device=dev)}
    o = ThOut()
    o.rgb_map, o.acc_map, o.depth_map = (out[k].data_ptr() for k in ("rgb_map", "acc_map", "depth_map"))
    if want_raw:
        out["raw"] = torch.empty((N, S, 4), device=dev)
        o.raw = out["raw"].data_ptr()
    if want_mask and mode != TH_RENDER_DENSE:
        out["pts_mask"] = torch.empty((N, S), dtype=torch.uint8, device=dev)
        o.pts_mask = out["pts_mask"].data_ptr()
    counters = (C.c_int64 * 3)()
    o.counters_host = counters
    nbytes = lib.th_frame_workspace_bytes(C.byref(frame.c), N * S, 1)   # sized for the schedule the frame selects
    ws = _workspace(nbytes, dev)
    _lib.check(lib.th_render_rays(C.byref(frame.c), C.byref(r), C.byref(o), mode, _ptr(ws), ws.numel(), _stream()),
               "th_render_rays")
    out["counters"] = tuple(int(c) for c in counters)
    return out


def query_density(frame: Frame, pts):
    """a12: alpha_raw (P) (0 where culled) and the cull mask (P) uint8."""
    lib = _lib.load()
    pts = _f32(pts, "pts").view(-1, 3)
    P = pts.shape[0]
    alpha = torch.empty((P,), device=pts.device)
    mask = torch.empty((P,), dtype=torch.uint8, device=pts.device)
    nbytes = lib.th_frame_workspace_bytes(C.byref(frame.c), P, 1)
    ws = _workspace(nbytes, pts.device)
    _lib.check(lib.th_query_density(C.byref(frame.c), _ptr(pts), P, _ptr(alpha), _ptr(mask), _ptr(ws), ws.numel(),
                                    _stream()), "th_query_density")
    return alpha, mask


# ---- staged entry points (reference tensor layouts) --------------------------------------
def sample_points(ray_o, ray_d, near, far, S: int, t_vals=None):
    lib = _lib.load()
    r, keep = _rays(ray_o, ray_d, near, far, S, t_vals)
    pts = torch.empty((r.n_rays, S, 3), device=keep[0].device)
    z = torch.empty((r.n_rays, S), device=keep[0].device)
    _lib.check(lib.th_sample_points(C.byref(r), _ptr(pts), _ptr(z), _stream()), "th_sample_points")
    return pts, z


def cull_knn1(pts, verts, radius: float = 0.1):
    lib = _lib.load()
    pts, verts = _f32(pts, "pts").view(-1, 3), _f32(verts, "verts").view(-1, 3)
    P = pts.shape[0]
    d2 = torch.empty((P,), device=pts.device)
    idx = torch.empty((P,), dtype=torch.int64, device=pts.device)
    mask = torch.empty((P,), dtype=torch.uint8, device=pts.device)
    _lib.check(lib.th_cull_knn1(_ptr(pts), P, _ptr(verts), verts.shape[0], radius, _ptr(d2), _ptr(idx), _ptr(mask),
                                _stream()), "th_cull_knn1")
    return d2, idx, mask


def cull_grid(pts, verts, radius: float = 0.1):
    lib = _lib.load()
    pts, verts = _f32(pts, "pts").view(-1, 3), _f32(verts, "verts").view(-1, 3)
    P = pts.shape[0]
    mask = torch.empty((P,), dtype=torch.uint8, device=pts.device)
    ws = _workspace(lib.th_workspace_bytes(0, 1, verts.shape[0]), pts.device)
    _lib.check(lib.th_cull_grid(_ptr(pts), P, _ptr(verts), verts.shape[0], radius, _ptr(mask), _ptr(ws), ws.numel(),
                                _stream()), "th_cull_grid")
    return mask


def world2smpl(pts, Rh, Th):
    lib = _lib.load()
    pts = _f32(pts, "pts")
    Rh, Th = _f32(Rh, "Rh").view(3, 3), _f32(Th, "Th").view(3)
    out = torch.empty_like(pts)
    _lib.check(lib.th_world2smpl(_ptr(pts), pts.numel() // 3, _ptr(Rh), _ptr(Th), _ptr(out), _stream()),
               "th_world2smpl")
    return out


def view_embed(ray_d):
    lib = _lib.load()
    ray_d = _f32(ray_d, "ray_d").view(-1, 3)
    out = torch.empty((ray_d.shape[0], 27), device=ray_d.device)
    _lib.check(lib.th_view_embed(_ptr(ray_d), ray_d.shape[0], _ptr(out), _stream()), "th_view_embed")
    return out


def pixel_gather(frame: Frame, pts_world):
    lib = _lib.load()
    pts = _f32(pts_world, "pts").view(-1, 3)
    out = torch.empty((frame.V, 384, pts.shape[0]), device=pts.device)
    _lib.check(lib.th_pixel_gather(C.byref(frame.c), _ptr(pts), pts.shape[0], _ptr(out), _stream()),
               "th_pixel_gather")
    return out


def knn_dparf(frame: Frame, pts_smpl, token_grid=None):
    """a8 staged.  ``token_grid``: search the K nearest tokens through the token grid (True), by the scan over all
    tokens (False), or like the fused path does for culled rays / grid points (None: grid from 1024 tokens on)."""
    lib = _lib.load()
    pts = _f32(pts_smpl, "pts").view(-1, 3)
    P, K = pts.shape[0], frame.c.knn
    idx = torch.empty((P, K), dtype=torch.int64, device=pts.device)
    d2 = torch.empty((P, K), device=pts.device)
    rep = torch.empty((frame.V, 255, P), device=pts.device)
    if token_grid is None:
        token_grid = frame.c.n_tok >= 1024
    ws, nbytes = None, 0
    if token_grid:
        nbytes = lib.th_knn_workspace_bytes(frame.c.n_tok)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=pts.device)
    _lib.check(lib.th_knn_dparf(C.byref(frame.c), _ptr(pts), P, _ptr(idx), _ptr(d2), _ptr(rep), _ptr(ws), nbytes,
                                _stream()), "th_knn_dparf")
    return idx, d2, rep


def mlp_raw(frame: Frame, human_rep, pixel_feat, viewdir, pts_mask=None):
    lib = _lib.load()
    human_rep, pixel_feat = _f32(human_rep, "human_rep"), _f32(pixel_feat, "pixel_feat")
    viewdir = _f32(viewdir, "viewdir").view(-1, 27)
    P = viewdir.shape[0]
    assert human_rep.shape == (frame.V, 255, P) and pixel_feat.shape == (frame.V, 384, P)
    m = None if pts_mask is None else pts_mask.to(torch.uint8).contiguous().view(-1)
    raw = torch.empty((P, 4), device=viewdir.device)
    ws = _workspace(lib.th_workspace_bytes(P, frame.V, 0), viewdir.device)
    _lib.check(lib.th_mlp_raw(C.byref(frame.c), _ptr(human_rep), _ptr(pixel_feat), _ptr(viewdir), _ptr(m), P,
                              _ptr(raw), _ptr(ws), ws.numel(), _stream()), "th_mlp_raw")
    return raw


def integrate(raw, z_vals, ray_d, white_bkgd: bool = False):
    lib = _lib.load()
    raw, z_vals, ray_d = _f32(raw, "raw"), _f32(z_vals, "z_vals"), _f32(ray_d, "ray_d").view(-1, 3)
    N, S = z_vals.shape
    assert raw.shape == (N, S, 4)
    rgb = torch.empty((N, 3), device=raw.device)
    acc = torch.empty((N,), device=raw.device)
    depth = torch.empty((N,), device=raw.device)
    _lib.check(lib.th_integrate(_ptr(raw), _ptr(z_vals), _ptr(ray_d), N, S, 1 if white_bkgd else 0, _ptr(rgb),
                                _ptr(acc), _ptr(depth), _stream()), "th_integrate")
    return rgb, acc, depth


def nchw_to_nhwc(x):
    lib = _lib.load()
    x = _f32(x, "x")
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), device=x.device)
    _lib.check(lib.th_nchw_to_nhwc(_ptr(x), _ptr(out), n, c, h, w, _stream()), "th_nchw_to_nhwc")
    return out


def premap_features(x, weights: PackedWeights, out=None):
    """``TH_FLAG_PREMAPPED``: encoder maps (V,384,H,W) -> pre-mapped maps (V,H,W,512) with ``alpha_res_0`` /
    ``rgb_res_0`` / ``rgb_res_1`` already applied (``th_premap_features``, tcgen05 GEMM over the NCHW maps)."""
    lib = _lib.load()
    x = _f32(x, "x")
    n, c, h, w = x.shape
    assert c == 384 and n == weights.n_views
    if out is None:
        out = torch.empty((n, h, w, 512), device=x.device)
    assert out.shape == (n, h, w, 512) and out.is_contiguous() and out.dtype == torch.float32
    _lib.check(lib.th_premap_features(_ptr(x), weights.blob.data_ptr(), n, h, w, _ptr(out), _stream()),
               "th_premap_features")
    return out


# ---- the steps either side of the path -------------------------------------------
class ClusterIndex:
    """The k-means vertex clusters in CSR form (``cluster_start (n_tok+1)``, ``cluster_members (n_verts)`` int32):
    members of cluster c in the order of the reference's ``dict_voxel2pc_ind[c]``
    (if_clight_renderer.py:55; ascending vertex ids when only ``pc2voxel_ind`` is known)."""

    def __init__(self, pc2voxel_ind=None, dict_voxel2pc_ind=None, n_tok: int | None = None, device="cuda"):
        if dict_voxel2pc_ind is not None:
            keys = sorted(dict_voxel2pc_ind.keys())
            assert [int(k) for k in keys] == list(range(len(keys))), "cluster ids must be exactly arange(n)"
            lists = [np.asarray(dict_voxel2pc_ind[k], dtype=np.int64).reshape(-1) for k in keys]
        else:
            pc2 = np.asarray(pc2voxel_ind).reshape(-1).astype(np.int64)
            n = int(pc2.max()) + 1 if n_tok is None else int(n_tok)
            order = np.argsort(pc2, kind="stable")
            counts = np.bincount(pc2, minlength=n)
            lists = np.split(order, np.cumsum(counts)[:-1])
        assert all(len(l) > 0 for l in lists), "empty cluster"
        self.n_tok = len(lists)
        self.n_verts = int(sum(len(l) for l in lists))
        start = np.zeros(self.n_tok + 1, dtype=np.int32)
        start[1:] = np.cumsum([len(l) for l in lists])
        self.start_host, self.members_host = start, np.concatenate(lists).astype(np.int32)
        self._dev = {}
        self.to(device)

    def to(self, device):
        key = str(torch.device(device))
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.start_host).to(device), torch.from_numpy(self.members_host).to(device))
        self.start, self.members = self._dev[key]
        return self


def paint_group(holder_map, uv_scale, verts, cam_R, cam_T, cam_K, vizmap, clusters: ClusterIndex, want_painted=False):
    """8f-1: holder map (V,192,H,W) NCHW -> tokens (V,n_tok,192) (``th_paint_group``); ``vizmap`` (V,n_verts) bool
    or None."""
    lib = _lib.load()
    hm = _f32(holder_map, "holder_map")
    V, c, H, W = hm.shape
    assert c == 192
    verts = _f32(verts, "verts").view(-1, 3)
    nv = verts.shape[0]
    clusters.to(hm.device)
    assert clusters.n_verts == nv
    R, T, K = _f32(cam_R, "cam_R").view(V, 3, 3), _f32(cam_T, "cam_T").view(V, 3), _f32(cam_K, "cam_K").view(V, 3, 3)
    viz = None if vizmap is None else vizmap.to(torch.uint8).contiguous().view(V, nv)
    out = torch.empty((V, clusters.n_tok, 192), device=hm.device)
    painted = torch.empty((V, nv, 192), device=hm.device) if want_painted else None
    _lib.check(lib.th_paint_group(_ptr(hm), V, H, W, float(uv_scale[0]), float(uv_scale[1]), _ptr(verts), nv, _ptr(R),
                                  _ptr(T), _ptr(K), _ptr(viz), _ptr(clusters.start), _ptr(clusters.members),
                                  clusters.n_tok, _ptr(painted), _ptr(out), _stream()), "th_paint_group")
    return (out, painted) if want_painted else out


class EncoderTail:
    """The inputs of ``SpatialEncoder.forward``'s tail (encoder.py:133-146) for ``th_premap_from_latents`` /
    ``th_paint_group_latents``: the three backbone latents (V,64|64|128,lh,lw), the input images (V,3,H,W) and the
    ``upsample_color`` 1x1 convolution.  The C ABI takes the latents channel-last: tensors in
    ``torch.channels_last`` memory format (what a backbone run in that format returns) are passed as they are,
    NCHW-contiguous ones are converted here (22 MB per 512 x 512 view).  Keeps the tensors alive next to the C struct."""

    def __init__(self, latents, images, color_w, color_b):
        assert len(latents) == 3
        self.images = _f32(images, "images")
        self.color_w = _f32(color_w, "color_w").reshape(128, 3)
        self.color_b = _f32(color_b, "color_b").reshape(128)
        V, c, H, W = self.images.shape
        assert c == 3
        self.latents = []
        for l, ch in zip(latents, (64, 64, 128)):
            assert l.dim() == 4 and l.shape[0] == V and l.shape[1] == ch, "latents must be (V,64|64|128,lh,lw)"
            if not l.is_cuda:
                raise ValueError("latent: expected a CUDA tensor (the query path has no CPU implementation)")
            nhwc = l.float().permute(0, 2, 3, 1)          # a view; contiguous already for channels_last tensors
            self.latents.append(nhwc.contiguous())
        self.n_views, self.h, self.w = V, H, W
        self.c = _lib.ThEncoderTail()
        for i, l in enumerate(self.latents):
            self.c.latent[i] = l.data_ptr()
            self.c.lat_h[i], self.c.lat_w[i] = l.shape[1], l.shape[2]
        self.c.images, self.c.color_w, self.c.color_b = self.images.data_ptr(), self.color_w.data_ptr(), self.color_b.data_ptr()
        self.c.n_views, self.c.h, self.c.w = V, H, W

    @property
    def device(self):
        return self.images.device


def premap_from_latents(enc: EncoderTail, weights: PackedWeights, out=None):
    """8f-2: pre-mapped maps (V,H,W,512) straight from the encoder's latents (``th_premap_from_latents``): equals
    ``premap_features(pixel_feat_map)`` without the 1.2 GB ``pixel_feat_map`` ever being written."""
    lib = _lib.load()
    assert enc.n_views == weights.n_views
    if out is None:
        out = torch.empty((enc.n_views, enc.h, enc.w, 512), device=enc.device)
    assert out.shape == (enc.n_views, enc.h, enc.w, 512) and out.is_contiguous() and out.dtype == torch.float32
    nbytes = lib.th_premap_from_latents_workspace_bytes(C.byref(enc.c))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=enc.device)
    _lib.check(lib.th_premap_from_latents(C.byref(enc.c), weights.blob.data_ptr(), _ptr(out), _ptr(ws), nbytes,
                                          _stream()), "th_premap_from_latents")
    return out


def paint_group_latents(enc: EncoderTail, reduction_w, reduction_b, uv_scale, verts, cam_R, cam_T, cam_K, vizmap,
                        clusters: ClusterIndex):
    """8f-1 + 8f-2: tokens (V,n_tok,192) from the latents (``th_paint_group_latents``); ``reduction_w`` (192,384[,1,1]),
    ``reduction_b`` (192) = ``encoder.reduction_layer``.  No holder map is built."""
    lib = _lib.load()
    V, dev = enc.n_views, enc.device
    rw, rb = _f32(reduction_w, "reduction_w").reshape(192, 384), _f32(reduction_b, "reduction_b").reshape(192)
    verts = _f32(verts, "verts").view(-1, 3)
    nv = verts.shape[0]
    clusters.to(dev)
    assert clusters.n_verts == nv
    R, T, K = _f32(cam_R, "cam_R").view(V, 3, 3), _f32(cam_T, "cam_T").view(V, 3), _f32(cam_K, "cam_K").view(V, 3, 3)
    viz = None if vizmap is None else vizmap.to(torch.uint8).contiguous().view(V, nv)
    out = torch.empty((V, clusters.n_tok, 192), device=dev)
    nbytes = lib.th_paint_group_latents_workspace_bytes(V, nv, clusters.n_tok)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    _lib.check(lib.th_paint_group_latents(C.byref(enc.c), _ptr(rw), _ptr(rb), float(uv_scale[0]), float(uv_scale[1]),
                                          _ptr(verts), nv, _ptr(R), _ptr(T), _ptr(K), _ptr(viz), _ptr(clusters.start),
                                          _ptr(clusters.members), clusters.n_tok, _ptr(out), _ptr(ws), nbytes,
                                          _stream()), "th_paint_group_latents")
    return out


class PackedLinear:
    """An ``nn.Linear`` packed for ``th_linear`` (fp16 hi/lo operand images of the weight, host-side; the blob is
    1 KB-aligned on the device).  Re-pack when the parameters change (``key``: data pointers + versions)."""

    def __init__(self, weight, bias, device="cuda"):
        lib = _lib.load()
        w = weight.detach().to(torch.float32).cpu().contiguous()
        b = None if bias is None else bias.detach().to(torch.float32).cpu().contiguous()
        self.n_out, self.n_in = w.shape
        nbytes = lib.th_linear_packed_bytes(self.n_out, self.n_in)
        if nbytes == 0:
            raise ValueError(f"th_linear: unsupported shape ({self.n_out}, {self.n_in}): n_in % 64, n_out % 4")
        host = np.zeros(nbytes, dtype=np.uint8)
        _lib.check(lib.th_linear_pack(C.c_void_p(w.data_ptr()), None if b is None else C.c_void_p(b.data_ptr()),
                                      self.n_out, self.n_in, C.c_void_p(host.ctypes.data), nbytes), "th_linear_pack")
        buf = torch.empty((nbytes + 1024,), dtype=torch.uint8, device=device)
        shift = (-buf.data_ptr()) % 1024
        self._buf = buf
        self.blob = buf[shift:shift + nbytes]
        self.blob.copy_(torch.from_numpy(host))

    def __call__(self, x, relu: bool = False):
        lib = _lib.load()
        x = _f32(x, "x")
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        assert x2.shape[1] == self.n_in
        y = torch.empty((x2.shape[0], self.n_out), device=x.device)
        _lib.check(lib.th_linear(_ptr(x2), x2.shape[0], x2.stride(0), self.blob.data_ptr(), self.n_out, self.n_in, _ptr(y),
                                 self.n_out, 1 if relu else 0, _stream()), "th_linear")
        return y.reshape(*shape[:-1], self.n_out)


def marching_cubes(volume, iso: float):
    """8f-4: marching cubes on a device volume (X,Y,Z) fp32 (``th_marching_cubes``) -> (vertices (n,3) fp32 in index
    coordinates, triangles (m,3) int32), both on the device.  One counting call, one emitting call."""
    lib = _lib.load()
    vol = _f32(volume, "volume")
    assert vol.dim() == 3
    nx, ny, nz = vol.shape
    nbytes = lib.th_marching_cubes_workspace_bytes(nx, ny, nz)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=vol.device)
    counts = (C.c_int64 * 2)()
    _lib.check(lib.th_marching_cubes(_ptr(vol), nx, ny, nz, float(iso), None, 0, None, 0, counts, _ptr(ws), nbytes,
                                     _stream()), "th_marching_cubes")
    nv, nt = int(counts[0]), int(counts[1])
    verts = torch.empty((nv, 3), device=vol.device)
    tris = torch.empty((nt, 3), dtype=torch.int32, device=vol.device)
    if nv:
        _lib.check(lib.th_marching_cubes(_ptr(vol), nx, ny, nz, float(iso), _ptr(verts), nv, _ptr(tris) if nt else None,
                                         nt, counts, _ptr(ws), nbytes, _stream()), "th_marching_cubes")
    return verts, tris


def vit_attention(qkv, n_heads: int, scale: float):
    """8f-3: ``Attention.forward`` between ``self.qkv`` and ``self.proj`` (vision_transformer.py:267-275):
    qkv (B,N,3*H*64) fp32 -> (B,N,H*64), flash-style (``th_vit_attention``)."""
    lib = _lib.load()
    qkv = _f32(qkv, "qkv")
    B, N, C3 = qkv.shape
    hd = C3 // (3 * n_heads)
    assert C3 == 3 * n_heads * hd
    out = torch.empty((B, N, n_heads * hd), device=qkv.device)
    nbytes = lib.th_vit_attention_workspace_bytes(B, N, n_heads)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=qkv.device)
    _lib.check(lib.th_vit_attention(_ptr(qkv), B, N, n_heads, hd, float(scale), _ptr(out), _ptr(ws), nbytes, _stream()),
               "th_vit_attention")
    return out


def group_mean(x, clusters: ClusterIndex, outer_order=None):
    """``Renderer.voxelization`` of a per-vertex tensor (n_verts, ...), fp32 or fp64, bit-equal to torch-CPU's
    ``x[idx].mean(0)`` (``th_group_mean``).  ``outer_order`` defaults to torch's choice for the row width."""
    lib = _lib.load()
    assert x.is_cuda and x.dtype in (torch.float32, torch.float64)
    x = x.contiguous()
    n = x.shape[0]
    C_ = x.numel() // n
    clusters.to(x.device)
    assert clusters.n_verts == n
    f64 = x.dtype == torch.float64
    if outer_order is None:
        outer_order = C_ >= (16 if f64 else 32)
    out = torch.empty((clusters.n_tok,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    _lib.check(lib.th_group_mean(_ptr(x), 1 if f64 else 0, C_, _ptr(clusters.start), _ptr(clusters.members),
                                 clusters.n_tok, 1 if outer_order else 0, _ptr(out), _stream()), "th_group_mean")
    return out


def near_far(ray_o, ray_d, bounds) -> dict:
    """``get_near_far`` on given rays (``th_near_far``); returns near, far, mask_at_box and the clamped ray_d copy."""
    lib = _lib.load()
    ray_o = _f32(ray_o, "ray_o").view(-1, 3)
    ray_d = _f32(ray_d, "ray_d").view(-1, 3).clone()
    b = _f32(bounds, "bounds").view(2, 3)
    n = ray_o.shape[0]
    near, far = torch.empty((n,), device=ray_o.device), torch.empty((n,), device=ray_o.device)
    mask = torch.empty((n,), dtype=torch.uint8, device=ray_o.device)
    _lib.check(lib.th_near_far(_ptr(ray_o), _ptr(ray_d), n, _ptr(b), _ptr(near), _ptr(far), _ptr(mask), _stream()),
               "th_near_far")
    return {"near": near, "far": far, "mask_at_box": mask, "ray_d": ray_d}


def generate_rays(H: int, W: int, K, R, T, bounds=None, compact: bool = True) -> dict:
    """8f-4: ``get_rays`` + ``get_near_far`` + the ``[mask_at_box]`` selection for one target camera
    (``th_generate_rays``).  K (3,3), R (3,3), T (3,) or (3,1), bounds (2,3): CUDA tensors.  Returns dense
    ``ray_o, ray_d (H*W,3)`` and, with bounds, ``near, far, mask_at_box`` plus (``compact``) the rays inside the box
    in pixel order -- one device->host read of their count."""
    lib = _lib.load()
    K = _f32(K, "K").view(3, 3)
    dev = K.device
    Kinv = torch.linalg.inv(K).contiguous()             # np.linalg.inv(K) in float32 (if_nerf_data_utils.py:23)
    R, T = _f32(R, "R").view(3, 3), _f32(T, "T").view(3)
    n = H * W
    out = {"ray_o": torch.empty((n, 3), device=dev), "ray_d": torch.empty((n, 3), device=dev)}
    b = near = far = mask = oc = dc = nc = fc = cnt = ws = None
    if bounds is not None:
        b = _f32(bounds, "bounds").view(2, 3)
        near, far = torch.empty((n,), device=dev), torch.empty((n,), device=dev)
        mask = torch.empty((n,), dtype=torch.uint8, device=dev)
        if compact:
            oc, dc = torch.empty((n, 3), device=dev), torch.empty((n, 3), device=dev)
            nc, fc = torch.empty((n,), device=dev), torch.empty((n,), device=dev)
            cnt = torch.zeros((1,), dtype=torch.int64, device=dev)
            ws = torch.empty(lib.th_generate_rays_workspace_bytes(n), dtype=torch.uint8, device=dev)
    _lib.check(lib.th_generate_rays(H, W, _ptr(Kinv), _ptr(R), _ptr(T), _ptr(b), _ptr(out["ray_o"]), _ptr(out["ray_d"]),
                                    _ptr(near), _ptr(far), _ptr(mask), _ptr(oc), _ptr(dc), _ptr(nc), _ptr(fc), _ptr(cnt),
                                    _ptr(ws), 0 if ws is None else ws.numel(), _stream()), "th_generate_rays")
    if bounds is not None:
        out.update(near=near, far=far, mask_at_box=mask)
        if compact:
            m = int(cnt.item())
            out.update(ray_o_c=oc[:m], ray_d_c=dc[:m], near_c=nc[:m], far_c=fc[:m], count=m)
    return out
