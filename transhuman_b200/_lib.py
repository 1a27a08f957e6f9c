"""ctypes binding of ``libtranshuman_b200.so`` (the C ABI of
``include/transhuman_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (``nvcc`` for
sm_100a).  There is no CPU fallback: if the library is missing, loading raises,
and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TH_LIB_PATH") or os.path.join(HERE, "libtranshuman_b200.so")  # TH_LIB_PATH: A/B builds

TH_MAX_VIEWS = 4
TH_MAX_KNN = 16
TH_FLAG_WHITE_BKGD = 1
TH_FLAG_SIMT_MLP = 2
TH_FLAG_LAYERWISE = 4
TH_FLAG_PREMAPPED = 8  # experimental, see include/transhuman_b200.h
TH_RENDER_DENSE, TH_RENDER_MASKED, TH_RENDER_FAST = 0, 1, 2
TH_TRAIN_BRANCH_MAX_RAYS = 2400

_fp = C.c_void_p  # device pointers travel as integers


class ThFrame(C.Structure):
    _fields_ = [
        ("tok_feat", _fp), ("tok_xyz", _fp), ("tok_rot", _fp), ("verts", _fp), ("feat", _fp),
        ("cam_R", _fp), ("cam_T", _fp), ("cam_K", _fp), ("Rh", _fp), ("Th", _fp), ("weights", _fp),
        ("n_views", C.c_int32), ("n_tok", C.c_int32), ("n_verts", C.c_int32),
        ("feat_h", C.c_int32), ("feat_w", C.c_int32), ("knn", C.c_int32),
        ("uv_scale_x", C.c_float), ("uv_scale_y", C.c_float),
        ("knn_dist_alpha", C.c_float), ("cull_radius", C.c_float), ("flags", C.c_uint32),
    ]


class ThRays(C.Structure):
    _fields_ = [
        ("ray_o", _fp), ("ray_d", _fp), ("near_", _fp), ("far_", _fp), ("t_vals", _fp),
        ("n_rays", C.c_int64), ("n_samples", C.c_int32),
    ]


class ThOut(C.Structure):
    _fields_ = [
        ("rgb_map", _fp), ("acc_map", _fp), ("depth_map", _fp), ("raw", _fp), ("pts_mask", _fp),
        ("counters_host", C.POINTER(C.c_int64)),
    ]


WEIGHT_FIELDS = [
    ("fc_0", "fc_0"), ("alpha_res_0", "alpha_res_0"),
    ("skv0_key", "spatial_key_value_0.key_embed"), ("skv0_value", "spatial_key_value_0.value_embed"),
    ("skv1_key", "spatial_key_value_1.key_embed"), ("skv1_value", "spatial_key_value_1.value_embed"),
    ("fc_1", "fc_1"), ("fc_2", "fc_2"), ("fc_3", "fc_3"), ("alpha_fc", "alpha_fc"),
    ("feature_fc", "feature_fc"), ("rgb_res_0", "rgb_res_0"), ("view_fc", "view_fc"),
    ("rgb_res_1", "rgb_res_1"), ("fc_4", "fc_4"), ("rgb_fc", "rgb_fc"),
]


class ThEncoderTail(C.Structure):
    _fields_ = [
        ("latent", _fp * 3), ("lat_h", C.c_int32 * 3), ("lat_w", C.c_int32 * 3),
        ("images", _fp), ("color_w", _fp), ("color_b", _fp),
        ("n_views", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
    ]


class ThWeightsF32(C.Structure):
    _fields_ = [(f"{c}_{s}", _fp) for c, _ in WEIGHT_FIELDS for s in ("w", "b")]


# name -> (restype, argtypes); must list every symbol include/transhuman_b200.h declares
SIGNATURES = {
    "th_version": (C.c_char_p, []),
    "th_last_error": (C.c_char_p, []),
    "th_launch_count": (C.c_int64, [C.c_int32]),
    "th_profile_start": (C.c_int, []),
    "th_profile_stop": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "th_packed_weights_bytes": (C.c_size_t, [C.c_int32]),
    "th_pack_weights": (C.c_int, [C.POINTER(ThWeightsF32), C.c_int32, _fp, C.c_size_t]),
    "th_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "th_frame_workspace_bytes": (C.c_size_t, [C.POINTER(ThFrame), C.c_int64, C.c_int32]),
    "th_render_rays": (C.c_int, [C.POINTER(ThFrame), C.POINTER(ThRays), C.POINTER(ThOut), C.c_int32, _fp,
                                 C.c_size_t, _fp]),
    "th_query_density": (C.c_int, [C.POINTER(ThFrame), _fp, C.c_int64, _fp, _fp, _fp, C.c_size_t, _fp]),
    "th_sample_points": (C.c_int, [C.POINTER(ThRays), _fp, _fp, _fp]),
    "th_cull_knn1": (C.c_int, [_fp, C.c_int64, _fp, C.c_int32, C.c_float, _fp, _fp, _fp, _fp]),
    "th_cull_grid": (C.c_int, [_fp, C.c_int64, _fp, C.c_int32, C.c_float, _fp, _fp, C.c_size_t, _fp]),
    "th_world2smpl": (C.c_int, [_fp, C.c_int64, _fp, _fp, _fp, _fp]),
    "th_view_embed": (C.c_int, [_fp, C.c_int64, _fp, _fp]),
    "th_pixel_gather": (C.c_int, [C.POINTER(ThFrame), _fp, C.c_int64, _fp, _fp]),
    "th_knn_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "th_knn_dparf": (C.c_int, [C.POINTER(ThFrame), _fp, C.c_int64, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "th_mlp_raw": (C.c_int, [C.POINTER(ThFrame), _fp, _fp, _fp, _fp, C.c_int64, _fp, _fp, C.c_size_t, _fp]),
    "th_integrate": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int32, C.c_int32, _fp, _fp, _fp, _fp]),
    "th_nchw_to_nhwc": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp]),
    "th_premap_features": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, C.c_int32, _fp, _fp]),
    "th_paint_group": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _fp, C.c_int32, _fp, _fp, _fp,
                                 _fp, _fp, _fp, C.c_int32, _fp, _fp, _fp]),
    "th_premap_from_latents_workspace_bytes": (C.c_size_t, [C.POINTER(ThEncoderTail)]),
    "th_premap_from_latents": (C.c_int, [C.POINTER(ThEncoderTail), _fp, _fp, _fp, C.c_size_t, _fp]),
    "th_paint_group_latents_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "th_paint_group_latents": (C.c_int, [C.POINTER(ThEncoderTail), _fp, _fp, C.c_float, C.c_float, _fp, C.c_int32, _fp,
                                         _fp, _fp, _fp, _fp, _fp, C.c_int32, _fp, _fp, C.c_size_t, _fp]),
    "th_linear_packed_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "th_linear_pack": (C.c_int, [_fp, _fp, C.c_int32, C.c_int32, _fp, C.c_size_t]),
    "th_linear": (C.c_int, [_fp, C.c_int64, C.c_int32, _fp, C.c_int32, C.c_int32, _fp, C.c_int32, C.c_int32, _fp]),
    "th_marching_cubes_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "th_marching_cubes": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_float, _fp, C.c_int64, _fp, C.c_int64,
                                    C.POINTER(C.c_int64), _fp, C.c_size_t, _fp]),
    "th_vit_attention_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "th_vit_attention": (C.c_int, [_fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, _fp, _fp, C.c_size_t, _fp]),
    "th_group_mean": (C.c_int, [_fp, C.c_int32, C.c_int32, _fp, _fp, C.c_int32, C.c_int32, _fp, _fp]),
    "th_near_far": (C.c_int, [_fp, _fp, C.c_int64, _fp, _fp, _fp, _fp, _fp]),
    "th_generate_rays_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "th_generate_rays": (C.c_int, [C.c_int32, C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                   _fp, _fp, C.c_size_t, _fp]),
    "th_debug_chain_program": (C.c_int64, [_fp, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _fp, C.c_int64]),
}

_lib = None


class TransHumanLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TransHumanLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the query path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().th_last_error().decode("utf-8", "replace")
        raise TransHumanLibraryError(f"{what} failed ({rc}): {msg}")
