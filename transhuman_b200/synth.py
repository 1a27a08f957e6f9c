"""Deterministic synthetic frames for the TransHuman query path (numpy only).

There is no network in the build/GPU boxes, so neither ZJU-MoCap nor the SMPL
model nor a checkpoint is available at run time.  This module generates, from
a seed, everything the hot path consumes:

* a 6890-vertex humanoid point set with SMPL-template extents (stands in for
  ``v_template``; the reference loads it at ``if_clight_renderer.py:43-48``),
* a vertex -> cluster map in the format of the reference's
  ``kmeans_dict_{N}.npy`` (``pc2voxel_ind``; ``if_clight_renderer.py:55``),
* per-vertex 4x4 float64 blend matrices (``lib/utils/SMPL.py:174``),
* transformer tokens ``holder (V,N_c,192)`` (the ViT output,
  ``if_clight_renderer.py:538``) and ``pixel_feat_map (V,384,H,W)`` (the
  encoder output, ``encoder.py:133-155``),
* input/target cameras and the ray bundle of ``get_rays``
  (``lib/utils/if_nerf/if_nerf_data_utils.py:11-30``),
* the 16 Conv1d(k=1) layers of the per-point network with the reference's
  state_dict names and default-init bounds (``cross_transformer.py:97-126``).

All randomness comes from ``numpy.random.default_rng`` (PCG64, stable across
numpy versions), so tests, goldens and the bench see identical inputs.
"""
from __future__ import annotations

import math

import numpy as np

N_VERTS = 6890
C_TOK = 192
C_PIX = 384

# (out, in) of every Conv1d(k=1) on the path, reference state_dict names
# (cross_transformer.py:97-126; SpatialKeyValue at 31-40).
LAYER_SHAPES = {
    "fc_0": (256, 255),
    "alpha_res_0": (256, 384),
    "spatial_key_value_0.key_embed": (128, 256),
    "spatial_key_value_0.value_embed": (256, 256),
    "spatial_key_value_1.key_embed": (128, 256),
    "spatial_key_value_1.value_embed": (256, 256),
    "fc_1": (256, 256),
    "fc_2": (256, 256),
    "fc_3": (256, 256),
    "alpha_fc": (1, 256),
    "feature_fc": (256, 256),
    "rgb_res_0": (256, 384),
    "view_fc": (128, 283),
    "rgb_res_1": (128, 384),
    "fc_4": (128, 128),
    "rgb_fc": (3, 128),
}


def make_body(seed: int = 0) -> np.ndarray:
    """6890 points on a T-pose humanoid made of capsules; float64 (6890,3).
    Extents follow the SMPL template (x +-0.87, y -1.16..0.56, z -0.12..0.17)."""
    rng = np.random.default_rng([seed, 101])
    # (p0, p1, radius, xscale, zscale, share of vertices)
    segs = [
        ((0.0, -0.42, 0.02), (0.0, 0.24, 0.02), 0.115, 1.35, 1.0, 0.26),    # torso
        ((0.0, 0.30, 0.03), (0.0, 0.46, 0.04), 0.095, 1.0, 1.15, 0.13),     # neck+head
        ((0.17, 0.22, 0.0), (0.83, 0.22, 0.0), 0.042, 1.0, 1.0, 0.12),      # left arm
        ((-0.17, 0.22, 0.0), (-0.83, 0.22, 0.0), 0.042, 1.0, 1.0, 0.12),    # right arm
        ((0.095, -0.45, 0.0), (0.115, -1.10, 0.0), 0.062, 1.0, 1.0, 0.185),  # left leg
        ((-0.095, -0.45, 0.0), (-0.115, -1.10, 0.0), 0.062, 1.0, 1.0, 0.185),  # right leg
    ]
    counts = [int(round(s[5] * N_VERTS)) for s in segs]
    counts[0] += N_VERTS - sum(counts)
    out = []
    for (p0, p1, r, xs, zs, _), n in zip(segs, counts):
        p0 = np.asarray(p0)
        p1 = np.asarray(p1)
        axis = p1 - p0
        length = np.linalg.norm(axis)
        axis = axis / length
        helper = np.array([0.0, 0.0, 1.0]) if abs(axis[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
        u = np.cross(axis, helper)
        u /= np.linalg.norm(u)
        v = np.cross(axis, u)
        # extend by the radius at both ends so capsule caps are covered
        t = rng.uniform(-r / length, 1.0 + r / length, size=n)
        ang = rng.uniform(0.0, 2.0 * math.pi, size=n)
        rad = np.where((t < 0) | (t > 1),
                       r * np.sqrt(np.clip(1.0 - (np.minimum(np.abs(t), np.abs(t - 1)) * length / r) ** 2, 0, 1)),
                       r)
        tc = t
        pts = p0[None] + tc[:, None] * (length * axis)[None] + \
            rad[:, None] * (np.cos(ang)[:, None] * u[None] + np.sin(ang)[:, None] * v[None])
        centre = p0[None] + tc[:, None] * (length * axis)[None]
        off = pts - centre
        off[:, 0] *= xs
        off[:, 2] *= zs
        out.append(centre + off)
    verts = np.concatenate(out, 0)
    perm = np.random.default_rng([seed, 102]).permutation(N_VERTS)  # SMPL order is not spatial
    return np.ascontiguousarray(verts[perm])


def cluster_body(verts: np.ndarray, n_class: int) -> np.ndarray:
    """Vertex -> cluster map, int32 (6890,), every cluster non-empty and the
    cluster ids exactly ``arange(n_class)`` (as in the reference's k-means
    dictionaries).  Farthest-point seeds + nearest-seed assignment."""
    v = np.asarray(verts, dtype=np.float64)
    n = v.shape[0]
    assert 1 <= n_class <= n
    seeds = np.empty(n_class, dtype=np.int64)
    seeds[0] = int(np.argmin(v[:, 1]))
    d = np.full(n, np.inf)
    for i in range(1, n_class):
        d = np.minimum(d, ((v - v[seeds[i - 1]]) ** 2).sum(1))
        seeds[i] = int(np.argmax(d))
    # nearest seed (chunked to bound memory at n_class = 6000)
    assign = np.empty(n, dtype=np.int32)
    sv = v[seeds]
    for s in range(0, n, 512):
        dd = ((v[s:s + 512, None, :] - sv[None]) ** 2).sum(-1)
        assign[s:s + 512] = dd.argmin(1)
    assign[seeds] = np.arange(n_class, dtype=np.int32)
    return assign


def segment_mean(x: np.ndarray, pc2voxel: np.ndarray, n_class: int) -> np.ndarray:
    """Per-cluster mean over the vertex axis -- the vectorised form of the
    reference's ``voxelization`` loop (``if_clight_renderer.py:356-371``)."""
    x = np.asarray(x)
    flat = x.reshape(x.shape[0], -1)
    acc = np.zeros((n_class, flat.shape[1]), dtype=flat.dtype)
    np.add.at(acc, pc2voxel, flat)
    cnt = np.bincount(pc2voxel, minlength=n_class).astype(flat.dtype)
    return (acc / cnt[:, None]).reshape((n_class,) + x.shape[1:])


def _rodrigues(rvec: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(rvec, axis=-1, keepdims=True)
    k = rvec / np.maximum(th, 1e-12)
    K = np.zeros(rvec.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s = np.sin(th)[..., None]
    c = np.cos(th)[..., None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def look_at_camera(azimuth: float, radius: float, y_off: float, centre, size: int):
    """OpenCV-style camera (x right, y down, z forward) on a ring around the
    body.  Returns K (3,3), R (3,3), T (3,1) float32 with x_cam = R x + T."""
    centre = np.asarray(centre, dtype=np.float64)
    c = centre + np.array([radius * math.sin(azimuth), y_off, radius * math.cos(azimuth)])
    f = centre - c
    f /= np.linalg.norm(f)
    down = np.array([0.0, -1.0, 0.0])
    xc = np.cross(down, f)
    xc /= np.linalg.norm(xc)
    yc = np.cross(f, xc)
    R = np.stack([xc, yc, f], 0)
    T = -R @ c
    K = np.array([[size, 0, size / 2], [0, size, size / 2], [0, 0, 1]], dtype=np.float64)
    return K.astype(np.float32), R.astype(np.float32), T.reshape(3, 1).astype(np.float32)


def get_rays(H: int, W: int, K, R, T):
    """Ray bundle of ``get_rays`` (``if_nerf_data_utils.py:11-30``), float32."""
    K = np.asarray(K, np.float32)
    R = np.asarray(R, np.float32)
    T = np.asarray(T, np.float32)
    rays_o = -np.dot(R.T, T).ravel()
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    xy1 = np.stack([i, j, np.ones_like(i)], axis=2)
    pixel_camera = np.dot(xy1, np.linalg.inv(K).T)
    pixel_world = np.dot(pixel_camera - T.ravel(), R)
    rays_d = pixel_world - rays_o[None, None]
    rays_o = np.broadcast_to(rays_o, rays_d.shape)
    return (np.ascontiguousarray(rays_o.reshape(-1, 3), dtype=np.float32),
            np.ascontiguousarray(rays_d.reshape(-1, 3), dtype=np.float32))


def make_weights(seed: int = 0, alpha_bias_shift: float = 0.0, gain: float = math.sqrt(6.0),
                 alpha_gain: float = 30.0) -> dict:
    """The 16 Conv1d layers, reference names -> ``{name.weight (out,in),
    name.bias (out,)}`` float32.  U(-b, b) with b = gain/sqrt(fan_in): gain 1 is
    the reference's default ``nn.Conv1d`` init, under which the ReLU stack
    shrinks the signal to a near-constant output (alpha_raw = -0.025 +- 0.0025,
    an all-transparent image); the default gain sqrt(6) (He scaling) keeps unit
    variance through the stack so every layer matters to the result, and
    ``alpha_gain`` scales the density head so that rays saturate over a few
    samples like a trained model's do."""
    rng = np.random.default_rng([seed, 303])
    w = {}
    for name, (o, i) in LAYER_SHAPES.items():
        b = 1.0 / math.sqrt(i)
        w[name + ".weight"] = (rng.uniform(-b, b, size=(o, i)) * gain).astype(np.float32)
        w[name + ".bias"] = rng.uniform(-b, b, size=(o,)).astype(np.float32)
    w["alpha_fc.weight"] = (w["alpha_fc.weight"] * np.float32(alpha_gain)).astype(np.float32)
    w["alpha_fc.bias"] = (w["alpha_fc.bias"] * np.float32(alpha_gain) + np.float32(alpha_bias_shift)).astype(np.float32)
    return w


def make_frame(H: int = 64, W: int = 64, n_class: int = 300, V: int = 3, feat_hw: int | None = None,
               seed: int = 0, near: float = 1.5, far: float = 3.5, target_azimuth: float = 1.0,
               rotate_rh: bool = False, posed: bool = False, with_feature_maps: bool = True,
               alpha_bias_shift: float = 0.0) -> dict:
    """One synthetic frame.  ``feat_hw`` is the (square) size of the input
    views / feature maps (defaults to ``H``).  Returned arrays are numpy,
    float32 unless noted; names follow the reference batch dict
    (``lib/datasets/light_stage/can_smpl.py:537-594``)."""
    assert H == W, "all reference configs are square (SURVEY 3.5-6)"
    feat_hw = feat_hw or H
    rng = np.random.default_rng([seed, 202])
    body = make_body(seed)
    if posed:
        body = body + rng.normal(0.0, 0.01, size=body.shape)
    pc2voxel = cluster_body(body, n_class)
    if rotate_rh:
        Rh = _rodrigues(np.array([0.3, -0.8, 0.2]))
        Th = np.array([[0.05, -0.02, 0.1]])
    else:
        Rh = np.eye(3)
        Th = np.zeros((1, 3))
    verts_smpl = body.astype(np.float32)
    # world = smpl @ Rh^-1 + Th  (inverse of world2smpl, if_clight_renderer.py:289-295)
    verts_world = (verts_smpl.astype(np.float64) @ np.linalg.inv(Rh) + Th).astype(np.float32)
    blend = np.zeros((N_VERTS, 4, 4), dtype=np.float64)
    blend[:, :3, :3] = _rodrigues(rng.normal(0.0, 0.3, size=(N_VERTS, 3)))
    blend[:, :3, 3] = rng.normal(0.0, 0.05, size=(N_VERTS, 3))
    blend[:, 3, 3] = 1.0
    centre = np.array([0.0, -0.3, 0.0]) @ np.linalg.inv(Rh) + Th[0]
    cams = [look_at_camera(2.0 * math.pi * v / max(V, 1), 2.5, 0.3, centre, feat_hw) for v in range(V)]
    Kt, Rt, Tt = look_at_camera(target_azimuth, 2.5, 0.3, centre, H)
    ray_o, ray_d = get_rays(H, W, Kt, Rt, Tt)
    frame = {
        "H": H, "W": W, "V": V, "n_class": n_class, "feat_hw": feat_hw,
        "tar_smpl_vertice": verts_world,
        "tar_smpl_vertice_smplcoord": verts_smpl,
        "Rh": Rh.astype(np.float32), "Th": Th.astype(np.float32),
        "blend_mtx": blend,
        "pc2voxel_ind": pc2voxel,
        "input_K": np.stack([c[0] for c in cams]),
        "input_R": np.stack([c[1] for c in cams]),
        "input_T": np.stack([c[2] for c in cams]),
        "target_K": Kt, "target_R": Rt, "target_T": Tt,
        "ray_o": ray_o, "ray_d": ray_d,
        "near": np.full((H * W,), near, dtype=np.float32),
        "far": np.full((H * W,), far, dtype=np.float32),
        "holder": rng.standard_normal((V, n_class, C_TOK), dtype=np.float32),
        "weights": make_weights(seed, alpha_bias_shift),
    }
    if with_feature_maps:
        frng = np.random.default_rng([seed, 404])
        frame["pixel_feat_map"] = frng.standard_normal((V, C_PIX, feat_hw, feat_hw), dtype=np.float32)
    return frame


def make_grid_points(frame: dict, res: int, pad: float = 0.05) -> np.ndarray:
    """Dense voxel grid over the body AABB (+-pad), (res,res,res,3) float32 --
    the ``batch['pts']`` of the mesh path (``can_smpl_mesh.py:78-85``)."""
    v = frame["tar_smpl_vertice"]
    lo = v.min(0) - pad
    hi = v.max(0) + pad
    axes = [np.linspace(lo[a], hi[a], res, dtype=np.float32) for a in range(3)]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), -1)
    return np.ascontiguousarray(g, dtype=np.float32)
