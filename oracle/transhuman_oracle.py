"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TransHuman per-ray query path.

A restatement, in plain torch-CPU fp32 tensor ops, of the reference algorithm
for the hot path SURVEY.md section 8(a) lists (rows a1-a12).  It is the checker
the ``tests/`` parity tests, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs compare the CUDA path against.
It is never imported by ``transhuman_b200`` (the product path): that package
fails loudly when its CUDA library is missing instead of falling back here.

Parity pin: every function below is checked against the *genuine* reference
modules (imported in place from ``/root/reference`` under ``oracle/ref_shim.py``)
by ``tests/test_oracle_vs_reference.py`` when the reference tree is present,
and against the committed fixtures ``tests/golden/*.npz`` which
``oracle/make_golden.py`` generated from those reference modules.

The one primitive that is NOT in ``/root/reference`` is
``pytorch3d.ops.knn_points`` (facebookresearch/pytorch3d, un-vendored and
un-pinned: absent from ``requirements.txt``, ``README.md:63-64`` says "build
from source").  Parity at that boundary is therefore **unpinned**; this file
restates its published semantics (squared L2 distances, K smallest, sorted
ascending, int64 indices) with a fully specified rounding order -- see
``knn_points`` -- and the reference modules are run with this restatement
injected, so everything downstream of it *is* pinned to the reference's code.

Each function cites the reference file:line it follows (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

KNN_DIST_ALPHA = 0.5   # configs/train_or_eval.yaml:64
CULL_RADIUS = 0.1      # if_clight_renderer.py:442, if_mesh_renderer.py:55
CHUNK = 1024 * 32      # if_clight_renderer.py:575
TRAIN_BRANCH_MAX_RAYS = 2400   # if_clight_renderer.py:551


# --------------------------------------------------------------------------
# k-NN  (pytorch3d.ops.knn_points; call sites cross_transformer.py:170,
#        if_clight_renderer.py:440, if_mesh_renderer.py:53)
# --------------------------------------------------------------------------
def pairwise_d2(p: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """d2[i,j] = fl(fl(fl(dx*dx) + fl(dy*dy)) + fl(dz*dz)), dx = fl(p.x - q.x).
    Every product and sum individually rounded to fp32 (no FMA contraction):
    each line below is one elementwise torch op."""
    dx = p[:, None, 0] - q[None, :, 0]
    dy = p[:, None, 1] - q[None, :, 1]
    dz = p[:, None, 2] - q[None, :, 2]
    xx = dx * dx
    yy = dy * dy
    zz = dz * dz
    s = xx + yy
    return s + zz


def knn_points(p1: torch.Tensor, p2: torch.Tensor, K: int = 1, return_nn: bool = False,
               chunk: int = 8192):
    """Brute-force K nearest neighbours of every ``p1[b,i]`` in ``p2[b]``.

    Returns ``(dists (B,P,K) fp32 squared, idx (B,P,K) int64, nn or None)``
    sorted ascending by ``(d2, idx)`` -- ties go to the LOWER index.  This is
    the contract the CUDA kernels reproduce bit-exactly."""
    assert p1.dtype == torch.float32 and p2.dtype == torch.float32
    B, P, _ = p1.shape
    N = p2.shape[1]
    dists = torch.empty((B, P, K), dtype=torch.float32, device=p1.device)
    idx = torch.empty((B, P, K), dtype=torch.int64, device=p1.device)
    for b in range(B):
        for s in range(0, P, chunk):
            d2 = pairwise_d2(p1[b, s:s + chunk], p2[b])
            if K == 1:
                # torch.min returns the first minimal index on CPU
                v, i = torch.min(d2, dim=1, keepdim=True)
                # make the lower-index tie rule explicit
                first = (d2 == v).to(torch.int64).argmax(dim=1, keepdim=True)
                i = first
            else:
                # stable sort on d2 keeps equal keys in index order
                v, i = torch.sort(d2, dim=1, stable=True)
                v, i = v[:, :K], i[:, :K]
            dists[b, s:s + chunk] = v
            idx[b, s:s + chunk] = i
    nn = None
    if return_nn:
        nn = torch.gather(p2[:, None].expand(B, P, N, 3), 2,
                          idx[..., None].expand(B, P, K, 3)) if P * N < 1 << 24 else None
    return dists, idx, nn


# --------------------------------------------------------------------------
# a1  point sampler  (if_clight_renderer.py:271-287, inference branch)
# --------------------------------------------------------------------------
def t_vals_for(S: int) -> torch.Tensor:
    """``torch.linspace(0., 1., steps=S)`` evaluated on the CPU exactly as the
    reference does (273); the CUDA path takes this buffer as an input."""
    return torch.linspace(0., 1., steps=S)


def get_sampling_points(ray_o, ray_d, near, far, S: int):
    t_vals = t_vals_for(S).to(near)
    z_vals = near[..., None] * (1. - t_vals) + far[..., None] * t_vals      # 274
    pts = ray_o[:, :, None] + ray_d[:, :, None] * z_vals[..., None]         # 285
    return pts, z_vals


# --------------------------------------------------------------------------
# a3  world -> SMPL  (if_clight_renderer.py:289-295)
# --------------------------------------------------------------------------
def world2smpl(pts, Rh, Th):
    sh = pts.shape
    p = pts.view(sh[0], -1, sh[-1])
    p = p - Th
    p = torch.matmul(p, Rh)
    return p.view(*sh)


# --------------------------------------------------------------------------
# a4  view-direction embedding  (if_clight_renderer.py:525-526; embedder.py:4-53,
#     multires = cfg.view_res = 4: freqs 1,2,4,8; layout [v | sin f v | cos f v]...)
# --------------------------------------------------------------------------
def view_embed(ray_d):
    viewdir = ray_d / torch.norm(ray_d, dim=2, keepdim=True)
    freqs = 2. ** torch.linspace(0., 3., steps=4)
    out = [viewdir]
    for f in freqs:
        out.append(torch.sin(viewdir * f))
        out.append(torch.cos(viewdir * f))
    return torch.cat(out, -1)                                               # (1,N,27)


# --------------------------------------------------------------------------
# a8  positional encoding of the deformed offsets (vision_transformer.py:100-136,
#     num_freqs = cfg.KNN_FREQ = 10, freq_factor = pi)
# --------------------------------------------------------------------------
def pe_tables(num_freqs: int = 10):
    freqs = np.pi * 2.0 ** torch.arange(0, num_freqs)
    _freqs = torch.repeat_interleave(freqs, 2).view(1, -1, 1)
    _phases = torch.zeros(2 * num_freqs)
    _phases[1::2] = np.pi * 0.5
    return _freqs.to(torch.float32), _phases.view(1, -1, 1)


def positional_encoding(x, num_freqs: int = 10):
    _freqs, _phases = pe_tables(num_freqs)
    embed = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)
    embed = torch.sin(torch.addcmul(_phases, embed, _freqs))                # 132
    embed = embed.view(x.shape[0], -1)
    return torch.cat((x, embed), dim=-1)                                    # (n, 63)


# --------------------------------------------------------------------------
# token construction (if_clight_renderer.py:356-371 via 543-544): per-cluster
# mean over the vertices of each k-means cluster, in dict (= arange) order.
# --------------------------------------------------------------------------
def voxelization(pc2voxel: torch.Tensor, x: torch.Tensor, n_class: int) -> torch.Tensor:
    """Literal per-cluster ``x[pc_list].mean(0)`` loop (reference semantics,
    including its summation order inside ``mean``)."""
    out = []
    for c in range(n_class):
        pc_list = torch.nonzero(pc2voxel == c)[:, 0]
        out.append(x[pc_list].mean(0))
    return torch.stack(out)


def voxelization_lists(lists, x: torch.Tensor) -> torch.Tensor:
    """``Renderer.voxelization`` (356-371) with the cluster member lists as the reference holds them
    (``dict_voxel2pc_ind.values()``)."""
    return torch.stack([x[torch.as_tensor(l, dtype=torch.int64)].mean(0) for l in lists])


def paint_neural_human(smpl_vertice, input_R, input_T, input_K, holder_feat_map, image_shape, vizmap=None):
    """``Renderer.paint_neural_human`` (if_clight_renderer.py:95-184, t = 0, cfg.rasterize): vertices (6890,3) world,
    cameras (V,3,3)/(V,3,1)/(V,3,3), holder map (V,192,H,W), vizmap (V,6890) bool or None -> big_holder
    (V,6890,192): bilinear samples at the projected vertices, zero where invisible."""
    vertice_rot = torch.matmul(input_R[:, None], smpl_vertice[None].unsqueeze(-1))[..., 0]      # 121
    vertice = vertice_rot + input_T[:, None, :3, 0]                                             # 122
    vertice = torch.matmul(input_K[:, None], vertice.unsqueeze(-1))[..., 0]                     # 123
    uv = vertice[:, :, :2] / vertice[:, :, 2:]                                                  # 124
    latent = sample_from_feature_map(holder_feat_map, feat_scale_for(holder_feat_map), tuple(image_shape), uv)
    latent = latent.permute(0, 2, 1)                                                            # 167-171
    big = torch.zeros_like(latent)                                                              # 179
    if vizmap is None:
        return latent.clone()
    big[vizmap.bool()] = latent[vizmap.bool()]                                                  # 180
    return big


def can_body_grouping(lists, all_holders):
    """``Renderer.can_body_grouping`` (415-427): per view, per cluster mean of the painted vertices."""
    return torch.stack([voxelization_lists(lists, h) for h in all_holders])


# --------------------------------------------------------------------------
# 8f-4  rays of the target camera and their AABB near / far (numpy, like the reference's dataset code:
#       lib/utils/if_nerf/if_nerf_data_utils.py:11-30, 65-97 and the test split of sample_ray_grid, 190-199)
# --------------------------------------------------------------------------
def get_rays_np(H, W, K, R, T):
    rays_o = -np.dot(R.T, T).ravel()                                                            # 14
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    xy1 = np.stack([i, j, np.ones_like(i)], axis=2)
    pixel_camera = np.dot(xy1, np.linalg.inv(K).T)                                              # 25
    pixel_world = np.dot(pixel_camera - T.ravel(), R)                                           # 26
    rays_d = pixel_world - rays_o[None, None]
    rays_o = np.broadcast_to(rays_o, rays_d.shape)
    return rays_o, rays_d


def get_near_far_np(bounds, ray_o, ray_d):
    """-> near, far (float64, rays inside the box only), mask_at_box; CLAMPS ray_d in place like the reference (71)."""
    bounds = bounds + np.array([-0.01, 0.01])[:, None]                                          # 67
    nominator = bounds[None] - ray_o[:, None]
    ray_d[np.abs(ray_d) < 1e-5] = 1e-5                                                          # 71
    d_intersect = (nominator / ray_d[:, None]).reshape(-1, 6)
    p_intersect = d_intersect[..., None] * ray_d[:, None] + ray_o[:, None]
    min_x, min_y, min_z, max_x, max_y, max_z = bounds.ravel()
    eps = 1e-6
    p_mask_at_box = (p_intersect[..., 0] >= (min_x - eps)) * (p_intersect[..., 0] <= (max_x + eps)) * \
                    (p_intersect[..., 1] >= (min_y - eps)) * (p_intersect[..., 1] <= (max_y + eps)) * \
                    (p_intersect[..., 2] >= (min_z - eps)) * (p_intersect[..., 2] <= (max_z + eps))
    mask_at_box = p_mask_at_box.sum(-1) == 2                                                    # 86
    p_intervals = p_intersect[mask_at_box][p_mask_at_box[mask_at_box]].reshape(-1, 2, 3)
    ray_o = ray_o[mask_at_box]
    ray_d = ray_d[mask_at_box]
    norm_ray = np.linalg.norm(ray_d, axis=1)
    d0 = np.linalg.norm(p_intervals[:, 0] - ray_o, axis=1) / norm_ray
    d1 = np.linalg.norm(p_intervals[:, 1] - ray_o, axis=1) / norm_ray
    return np.minimum(d0, d1), np.maximum(d0, d1), mask_at_box


def test_split_rays(H, W, K, R, T, bounds):
    """The test split of ``sample_ray_grid`` (190-199) without the image: float32 rays, near / far, mask."""
    ray_o, ray_d = get_rays_np(H, W, K, R, T)
    ray_o = ray_o.reshape(-1, 3).astype(np.float32)
    ray_d = ray_d.reshape(-1, 3).astype(np.float32)
    near, far, mask_at_box = get_near_far_np(bounds, ray_o, ray_d)
    return {"ray_o_all": ray_o, "ray_d_all": ray_d, "near": near.astype(np.float32), "far": far.astype(np.float32),
            "mask_at_box": mask_at_box, "ray_o": ray_o[mask_at_box], "ray_d": ray_d[mask_at_box]}


# --------------------------------------------------------------------------
# a5  pixel-aligned feature gather (if_clight_renderer.py:186-208, 210-269)
# --------------------------------------------------------------------------
def sample_from_feature_map(feat_map, feat_scale, image_shape, uv):
    scale = feat_scale / image_shape                                        # 193
    scale = torch.tensor(scale).to(dtype=torch.float32)
    uv = uv * scale - 1.0                                                   # 197
    uv = uv.unsqueeze(2)
    samples = F.grid_sample(feat_map, uv, align_corners=True, mode="bilinear",
                            padding_mode="border")                          # 200-206
    return samples[:, :, :, 0]


def feat_scale_for(feat_map) -> np.ndarray:
    """``encoder.py:149-150``: scale = [W,H] / ([W,H]-1) * 2, numpy float64 (it
    is divided by ``image_shape`` in float64 and only then cast to float32,
    ``if_clight_renderer.py:193-195``)."""
    sc = np.array([feat_map.shape[-1], feat_map.shape[-2]])
    return sc / (sc - 1) * 2.0


def get_pixel_aligned_feature(xyz, input_R, input_T, input_K, pixel_feat_map, image_shape):
    """xyz (1,P,3) world -> (V,384,P)."""
    V = input_R.shape[0]
    x = xyz.unsqueeze(1).expand(-1, V, *xyz.shape[1:]).reshape(-1, *xyz.shape[1:])   # repeat_interleave 228
    xyz_rot = torch.matmul(input_R[:, None], x.unsqueeze(-1))[..., 0]       # 229
    x = xyz_rot + input_T[:, None, :3, 0]                                   # 230
    x = torch.matmul(input_K[:, None], x.unsqueeze(-1))[..., 0]             # 231
    uv = x[:, :, :2] / x[:, :, 2:]                                          # 232
    return sample_from_feature_map(pixel_feat_map, feat_scale_for(pixel_feat_map),
                                   tuple(image_shape), uv)


# --------------------------------------------------------------------------
# a8  DPaRF human representation (cross_transformer.py:151-205)
# --------------------------------------------------------------------------
def human_representation(pts_smpl, tok_xyz, blend_mtx, holder, K: int = 7, return_knn: bool = False):
    """pts_smpl (P,3), tok_xyz (N_c,3) fp32, blend_mtx (N_c,4,4) (fp64 ok),
    holder (V,N_c,192) -> human_rep (V,255,P)."""
    knn_dist, knn_idx, _ = knn_points(pts_smpl.unsqueeze(0), tok_xyz.unsqueeze(0), K=K)   # 170
    knn_dist = knn_dist[0].sqrt()                                           # 171
    knn_idx = knn_idx[0]
    w = F.softmax(-knn_dist / KNN_DIST_ALPHA, dim=1)                        # 151-156
    closest = tok_xyz[knn_idx]                                              # 183
    rel = pts_smpl.unsqueeze(1) - closest                                   # 184
    rot = blend_mtx[..., :3, :3].type(torch.float32)[knn_idx]               # 185-186
    deformed = torch.matmul(rel.unsqueeze(-2), rot).squeeze(-2)             # 187-188
    n, k = deformed.shape[:2]
    pe = positional_encoding(deformed.view(-1, 3)).view(n, k, -1)           # 192
    rep = []
    for _holder in holder:                                                  # 197-203
        f = torch.cat([_holder[knn_idx], pe], dim=-1)
        f = torch.sum(w.unsqueeze(-1) * f, dim=1)
        rep.append(f.t().unsqueeze(0))
    rep = torch.cat(rep, dim=0)
    if return_knn:
        return rep, knn_idx, knn_dist, w, deformed
    return rep


# --------------------------------------------------------------------------
# a9/a10  per-point network (cross_transformer.py:128-149, 273-353)
# --------------------------------------------------------------------------
def _conv(w: dict, name: str, x):
    """nn.Conv1d(k=1): x (B,Cin,P) -> (B,Cout,P)."""
    return F.conv1d(x, w[name + ".weight"][:, :, None], w[name + ".bias"])


def cross_attention(w, holder, pixel_feat):                                  # 128-149
    xp = pixel_feat.permute(2, 1, 0)
    key_embed = _conv(w, "spatial_key_value_0.key_embed", xp)
    value_embed = _conv(w, "spatial_key_value_0.value_embed", xp)
    xs = holder.permute(2, 1, 0)
    query_key = _conv(w, "spatial_key_value_1.key_embed", xs)
    query_value = _conv(w, "spatial_key_value_1.value_embed", xs)
    k_emb = key_embed.size(1)
    A = torch.bmm(key_embed.transpose(1, 2), query_key)
    A = A / math.sqrt(k_emb)
    A = F.softmax(A, dim=1)
    out = torch.bmm(value_embed, A)
    return query_value.permute(2, 1, 0) + out.permute(2, 1, 0)


def multiview_agg(w, human_rep, pixel_feat):                                # 313-322
    net_ske = F.relu(_conv(w, "fc_0", human_rep))
    net_pix = F.relu(_conv(w, "alpha_res_0", pixel_feat))
    net = cross_attention(w, net_ske, net_pix)
    net = F.relu(_conv(w, "fc_1", net))
    return F.relu(_conv(w, "fc_2", net))


def alpha_forward(w, inter_net, V):                                         # 324-328
    opa = inter_net.reshape(-1, V, *inter_net.shape[1:]).mean(dim=1)
    opa = F.relu(_conv(w, "fc_3", opa))
    return _conv(w, "alpha_fc", opa)


def rgb_forward(w, inter_net, pixel_feat, sincos_viewdir, V):               # 330-353
    features = _conv(w, "feature_fc", inter_net) + _conv(w, "rgb_res_0", pixel_feat)
    vd = sincos_viewdir.unsqueeze(1).expand(-1, V, *sincos_viewdir.shape[1:])
    vd = vd.reshape(-1, *sincos_viewdir.shape[1:]).transpose(1, 2)
    features = torch.cat((features, vd), dim=1)
    net = F.relu(_conv(w, "view_fc", features))
    net = net + _conv(w, "rgb_res_1", pixel_feat)
    net = net.reshape(-1, V, *net.shape[1:]).mean(dim=1)
    net = F.relu(_conv(w, "fc_4", net))
    return _conv(w, "rgb_fc", net)


def mlp_forward(w, human_rep, pixel_feat, sincos_viewdir, progressive: bool):
    """``MLP_forward_ori`` (280-289) / ``MLP_forward_ori_progressive`` (291-311).
    Returns raw (1,P,4) = (rgb x3, alpha)."""
    V = pixel_feat.shape[0]
    inter = multiview_agg(w, human_rep, pixel_feat)
    alpha = alpha_forward(w, inter, V)
    if not progressive:
        rgb = rgb_forward(w, inter, pixel_feat, sincos_viewdir, V)
    else:
        rgb = torch.zeros((alpha.shape[0], 3, alpha.shape[2]))
        m = (alpha > 0).flatten(0, 2)
        if m.sum() > 0:
            rgb[:, :, m] = rgb_forward(w, inter[..., m], pixel_feat[..., m], sincos_viewdir[:, m, :], V)
    return torch.cat((rgb, alpha), dim=1).transpose(1, 2)


def network_forward(w, pixel_feat, sincos_viewdir, pts_smpl, tok_xyz, blend_mtx, holder,
                    pts_mask=None, K: int = 7):
    """``Network.forward`` (cross_transformer.py:207-271) for B=1.
    pixel_feat (V,384,P), sincos_viewdir (1,P,27), pts_smpl (1,P,3),
    pts_mask (1,P) bool or None -> raw (1,P,4); masked-out points are 0."""
    if pts_mask is not None:
        raw_temp = torch.zeros((1, pts_smpl.shape[1], 4))
        if pts_mask.sum() == 0:
            return raw_temp
        pts_smpl = pts_smpl[pts_mask].unsqueeze(0)
        pixel_feat = pixel_feat[..., pts_mask[0]]
        sincos_viewdir = sincos_viewdir[:, pts_mask[0], :]
    rep = human_representation(pts_smpl[0], tok_xyz, blend_mtx, holder, K=K)
    raw = mlp_forward(w, rep, pixel_feat, sincos_viewdir, progressive=pts_mask is not None)
    if pts_mask is not None:
        raw_temp[pts_mask] = raw[0]
        raw = raw_temp
    return raw


# --------------------------------------------------------------------------
# a11  ray integration (nerf_net_utils.py:14-59; raw_noise_std = 0)
# --------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, white_bkgd: bool = False):
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.Tensor([1e10]).expand(dists[..., :1].shape).to(dists)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    weights = alpha * torch.cumprod(
        torch.cat([torch.ones((alpha.shape[0], 1)).to(alpha), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, acc_map, weights, depth_map


# --------------------------------------------------------------------------
# Frame plumbing: numpy synth frame -> torch CPU tensors + tokens
# --------------------------------------------------------------------------
def to_torch_frame(frame: dict) -> dict:
    t = {}
    for k, v in frame.items():
        if isinstance(v, np.ndarray):
            t[k] = torch.from_numpy(np.ascontiguousarray(v))
        elif isinstance(v, dict):
            t[k] = {kk: torch.from_numpy(np.ascontiguousarray(vv)) for kk, vv in v.items()}
        else:
            t[k] = v
    return t


def build_tokens(tf: dict):
    """Token xyz (N_c,3) fp32 and blend matrices (N_c,4,4) fp64
    (if_clight_renderer.py:543-544)."""
    n = tf["n_class"]
    pc2 = tf["pc2voxel_ind"].to(torch.int64)
    tok_xyz = voxelization(pc2, tf["tar_smpl_vertice_smplcoord"], n)
    tok_blend = voxelization(pc2, tf["blend_mtx"], n)
    return tok_xyz, tok_blend


def batchify(tf, xyz, pts_smpl, viewdir, tok_xyz, tok_blend, pts_mask=None, chunk: int = CHUNK,
             K: int = 7):
    """``batchify_rays`` (if_clight_renderer.py:607-656): chunk loop over points."""
    image_shape = tf["pixel_feat_map"].shape[-2:]
    out = []
    for i in range(0, pts_smpl.shape[1], chunk):
        pf = get_pixel_aligned_feature(xyz[:, i:i + chunk], tf["input_R"], tf["input_T"], tf["input_K"],
                                       tf["pixel_feat_map"], image_shape)
        out.append(network_forward(tf["weights"], pf, viewdir[:, i:i + chunk], pts_smpl[:, i:i + chunk],
                                   tok_xyz, tok_blend, tf["holder"],
                                   pts_mask[:, i:i + chunk] if pts_mask is not None else None, K=K))
    return torch.cat(out, 1)


def _render(tf, ray_o, ray_d, pts, z_vals, pts_mask=None, tokens=None, white_bkgd=False, K: int = 7,
            train_branch_max_rays: int = TRAIN_BRANCH_MAX_RAYS):
    """``Renderer._render`` (if_clight_renderer.py:500-605) from the point where
    tokens / feature maps exist (the encoder + ViT prologue is out of scope).

    Literal quirk (551-571): with at most 2400 rays the reference takes its
    un-chunked "train" branch, which calls the network WITHOUT ``pts_mask`` --
    so ``render_fast`` on a small ray set evaluates every sample of every
    surviving ray (and the full RGB branch).  ``train_branch_max_rays=0``
    forces the chunked/masked branch for testing at small sizes."""
    S = pts.shape[2]
    if ray_o.shape[1] <= train_branch_max_rays:
        pts_mask = None
    xyz = pts.clone().flatten(1, 2)
    pts_s = world2smpl(pts, tf["Rh"][None], tf["Th"][None])
    viewdir = view_embed(ray_d)
    viewdir = viewdir[:, :, None].repeat(1, 1, S, 1).contiguous().view(1, -1, 27)
    tok_xyz, tok_blend = tokens if tokens is not None else build_tokens(tf)
    raw = batchify(tf, xyz, pts_s.flatten(1, 2), viewdir, tok_xyz, tok_blend,
                   pts_mask.flatten(1, 2) if pts_mask is not None else None, K=K)
    raw = raw.reshape(-1, S, 4)
    rgb, acc, weights, depth = raw2outputs(raw, z_vals.view(-1, S), ray_d.view(-1, 3), white_bkgd)
    return {"rgb_map": rgb[None], "acc_map": acc[None], "depth_map": depth[None], "raw": raw}


def render(tf: dict, S: int, tokens=None, K: int = 7, white_bkgd: bool = False) -> dict:
    """``Renderer.render`` (486-498), dense mode (``pts_mask=None``)."""
    ray_o, ray_d = tf["ray_o"][None], tf["ray_d"][None]
    pts, z_vals = get_sampling_points(ray_o, ray_d, tf["near"][None], tf["far"][None], S)
    return _render(tf, ray_o, ray_d, pts, z_vals, tokens=tokens, K=K, white_bkgd=white_bkgd)


def cull_mask(pts_world, verts_world):
    """``render_fast`` cull (440-442): K=1 nearest vertex, sqrt(d2) < 0.1."""
    d2, _, _ = knn_points(pts_world, verts_world, K=1)
    return (d2.sqrt() < CULL_RADIUS).squeeze(-1)


def render_fast(tf: dict, S: int, tokens=None, K: int = 7,
                train_branch_max_rays: int = TRAIN_BRANCH_MAX_RAYS, white_bkgd: bool = False) -> dict:
    """``Renderer.render_fast`` (429-484), culled mode."""
    ray_o, ray_d = tf["ray_o"][None], tf["ray_d"][None]
    near, far = tf["near"][None], tf["far"][None]
    pts, z_vals = get_sampling_points(ray_o, ray_d, near, far, S)
    sh = pts.shape
    valid_pts = cull_mask(pts.flatten(1, 2), tf["tar_smpl_vertice"][None])
    valid_pix = valid_pts.view(1, *sh[1:3]).sum(-1) > 0
    valid_pix_pts = valid_pts.view(1, *sh[1:3])[valid_pix].unsqueeze(0)
    out = {"rgb_map": torch.zeros((1, sh[1], 3)), "acc_map": torch.zeros((1, sh[1])),
           "depth_map": torch.zeros((1, sh[1])), "valid_pts_mask": valid_pts.view(1, *sh[1:3]),
           "raw": torch.zeros((sh[1], S, 4))}
    if valid_pix.sum() == 0:
        return out
    r = _render(tf, ray_o[valid_pix].unsqueeze(0), ray_d[valid_pix].unsqueeze(0),
                pts[valid_pix].unsqueeze(0), z_vals[valid_pix].unsqueeze(0),
                pts_mask=valid_pix_pts, tokens=tokens, K=K, white_bkgd=white_bkgd,
                train_branch_max_rays=train_branch_max_rays)
    out["depth_map"][valid_pix] = r["depth_map"][0]
    out["rgb_map"][valid_pix] = r["rgb_map"][0]
    out["acc_map"][valid_pix] = r["acc_map"][0]
    out["raw"][valid_pix[0]] = r["raw"]
    return out


def query_density(tf: dict, pts_world, tokens=None, K: int = 7):
    """Grid query of ``if_mesh_renderer.Renderer.render`` (46-96): cull, zero
    view direction, chunked network forward; returns alpha_raw (P,) fp32 and
    the cull mask (P,)."""
    P = pts_world.shape[0]
    xyz = pts_world[None]
    mask = cull_mask(xyz, tf["tar_smpl_vertice"][None])
    pts_s = world2smpl(xyz, tf["Rh"][None], tf["Th"][None])
    viewdir = torch.zeros((1, P, 27))
    tok_xyz, tok_blend = tokens if tokens is not None else build_tokens(tf)
    raw = batchify(tf, xyz, pts_s, viewdir, tok_xyz, tok_blend, mask, K=K)
    return raw[0, :, 3], mask[0]
