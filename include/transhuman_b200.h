/*
 * transhuman_b200 -- C ABI of the B200-native TransHuman query path.
 *
 * Drop-in boundary for the per-ray volumetric query path of
 * pansanity666/TransHuman (sample -> cull -> k-NN/DPaRF -> pixel-aligned
 * gather -> per-point MLP -> alpha compositing).  The reference has no FFI:
 * its boundary is the Python plugin surface `Renderer(net).render /
 * render_fast` selected by the YAML keys `renderer_module` / `renderer_path`
 * (lib/networks/renderer/make_renderer.py:4-8).  `transhuman_b200/renderer.py`
 * implements that surface and calls the entry points below through ctypes;
 * INTEGRATION.md shows the binding.  Each entry point cites the reference
 * code it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless its name
 *    ends in `_host`; fp32 unless stated; all tensors contiguous;
 *  - return 0 on success, a negative TH_E* code otherwise (never throws, never
 *    exits); `th_last_error()` returns a thread-local message;
 *  - asynchronous with respect to the host on `stream` unless stated; the
 *    library allocates nothing persistent: the caller passes a workspace of
 *    at least `th_workspace_bytes(...)` bytes (256-byte aligned);
 *  - forward only (training keeps the reference's torch path).
 */
#ifndef TRANSHUMAN_B200_H
#define TRANSHUMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TH_OK 0
#define TH_EINVAL (-1)     /* bad argument (null pointer, unsupported shape) */
#define TH_EWORKSPACE (-2) /* workspace too small */
#define TH_ECUDA (-3)      /* a CUDA call / launch failed */
#define TH_EUNSUPPORTED (-4)

#define TH_C_TOK 192 /* token feature width, cfg.embed_size (vit_tiny)            */
#define TH_C_PIX 384 /* pixel-aligned feature width, encoder.py:133-146           */
#define TH_C_REP 255 /* 192 + 63, cross_transformer.py:106                        */
#define TH_C_VIEW 27 /* view-direction embedding, embedder.py (view_res = 4)      */
#define TH_MAX_VIEWS 4
#define TH_TRAIN_BRANCH_MAX_RAYS 2400 /* if_clight_renderer.py:551 */

/* `culled` argument of th_render_rays */
#define TH_RENDER_DENSE 0   /* Renderer.render: every sample evaluated             */
#define TH_RENDER_MASKED 1  /* render_fast, chunked branch: only points within the
                               cull radius are evaluated, progressive RGB          */
#define TH_RENDER_FAST 2    /* render_fast literally: as MASKED when more than
                               TH_TRAIN_BRANCH_MAX_RAYS rays survive the cull, else
                               every sample of the surviving rays (the reference's
                               un-chunked branch drops pts_mask)                    */
#define TH_MAX_KNN 16

/* Flags for ThFrame.flags / th_render_rays */
#define TH_FLAG_WHITE_BKGD 1u /* cfg.white_bkgd, nerf_net_utils.py:56-57 */
#define TH_FLAG_SIMT_MLP 2u   /* force the fp32 CUDA-core GEMM path (debug / parity) */
#define TH_FLAG_LAYERWISE 4u  /* tcgen05 GEMMs one layer per launch instead of the layer-chained kernel (cross-check) */
/* `feat` holds the PRE-MAPPED maps (V, H, W, 512) written by th_premap_features -- alpha_res_0,
 * rgb_res_0 and rgb_res_1 (cross_transformer.py:315, 333, 343; linear maps that the reference applies
 * to the bilinear blend of the feature maps) applied to the maps once per frame instead of to every
 * blended sample: blend and map commute exactly in real arithmetic.  This is the path the Renderer
 * plugin uses (23 % fewer tensor MACs per point, no alpha_res_0 jobs).  Layer-chained tensor-core
 * schedule only (V <= 3); without the flag `feat` is the plain channel-last map (V, H, W, 384). */
#define TH_FLAG_PREMAPPED 8u

/* Per-frame state: the outputs of the (out-of-scope, torch) prologue that the
 * query path consumes.  Built once per frame by the Python Renderer. */
typedef struct ThFrame {
  /* tokens: ViT output `holder_completed` (if_clight_renderer.py:538) and the
   * DPaRF parameters `obs_smpl_smplcoord`, `blend_mtx[..., :3, :3].float()`
   * (if_clight_renderer.py:541-544, cross_transformer.py:185) */
  const float* tok_feat; /* (V, n_tok, 192)                               */
  const float* tok_xyz;  /* (n_tok, 3)   SMPL coordinates                 */
  const float* tok_rot;  /* (n_tok, 3, 3) row-major                       */
  /* cull set: batch['tar_smpl_vertice'] (if_clight_renderer.py:440)      */
  const float* verts;    /* (n_verts, 3) world coordinates                */
  /* encoder output `pixel_feat_map` (if_clight_renderer.py:399), stored
   * channel-last: (V, H, W, 384).  th_nchw_to_nhwc converts.             */
  const float* feat;
  /* input cameras batch['input_R'|'input_T'|'input_K'][0]
   * (if_clight_renderer.py:215-221)                                      */
  const float* cam_R;    /* (V, 3, 3) */
  const float* cam_T;    /* (V, 3)    */
  const float* cam_K;    /* (V, 3, 3) */
  /* batch['Rh'], batch['Th'] (if_clight_renderer.py:289-295, 520)        */
  const float* Rh;       /* (3, 3)    */
  const float* Th;       /* (3)       */
  const void* weights;   /* blob written by th_pack_weights, uploaded by the caller */
  int32_t n_views;       /* V: 1..TH_MAX_VIEWS                            */
  int32_t n_tok;         /* cfg.num_class                                 */
  int32_t n_verts;       /* 6890 for SMPL                                 */
  int32_t feat_h, feat_w;
  int32_t knn;           /* cfg.KNN (7), <= TH_MAX_KNN                    */
  /* uv -> [-1,1]: `feat_scale / image_shape` evaluated in float64 on the host
   * then cast (if_clight_renderer.py:193-197; x uses index 0, y index 1) */
  float uv_scale_x, uv_scale_y;
  float knn_dist_alpha;  /* cfg.KNN_DIST_ALPHA = 0.5, cross_transformer.py:154 */
  float cull_radius;     /* 0.1, if_clight_renderer.py:442                */
  uint32_t flags;
} ThFrame;

/* Ray bundle: batch['ray_o','ray_d','near','far'] (if_clight_renderer.py:431-434)
 * and the S-float `torch.linspace(0,1,S)` table the reference evaluates on the
 * CPU (if_clight_renderer.py:273) -- taken as an input so z_vals are bit-exact. */
typedef struct ThRays {
  const float* ray_o;  /* (N, 3) */
  const float* ray_d;  /* (N, 3) */
  const float* near_;  /* (N)    */
  const float* far_;   /* (N)    */
  const float* t_vals; /* (S)    */
  int64_t n_rays;
  int32_t n_samples;
} ThRays;

/* Outputs of Renderer._render (if_clight_renderer.py:599): caller-allocated. */
typedef struct ThOut {
  float* rgb_map;   /* (N, 3) */
  float* acc_map;   /* (N)    */
  float* depth_map; /* (N)    */
  /* optional debug taps (may be NULL) */
  float* raw;       /* (N, S, 4): (rgb raw x3, alpha raw); 0 where culled */
  uint8_t* pts_mask; /* (N, S): cull result (1 = within cull_radius)      */
  int64_t* counters_host; /* HOST, optional: [0] = points within the cull radius,
                             [1] = surviving rays, [2] = points evaluated        */
} ThOut;

/* The 16 Conv1d(k=1) layers of the per-point network, HOST pointers, reference
 * state_dict names (cross_transformer.py:97-126); weight (out, in) row-major. */
typedef struct ThWeightsF32 {
  const float *fc_0_w, *fc_0_b;               /* (256,255) */
  const float *alpha_res_0_w, *alpha_res_0_b; /* (256,384) */
  const float *skv0_key_w, *skv0_key_b;       /* spatial_key_value_0.key_embed   (128,256) */
  const float *skv0_value_w, *skv0_value_b;   /* spatial_key_value_0.value_embed (256,256) */
  const float *skv1_key_w, *skv1_key_b;       /* spatial_key_value_1.key_embed   (128,256) */
  const float *skv1_value_w, *skv1_value_b;   /* spatial_key_value_1.value_embed (256,256) */
  const float *fc_1_w, *fc_1_b;               /* (256,256) */
  const float *fc_2_w, *fc_2_b;               /* (256,256) */
  const float *fc_3_w, *fc_3_b;               /* (256,256) */
  const float *alpha_fc_w, *alpha_fc_b;       /* (1,256)   */
  const float *feature_fc_w, *feature_fc_b;   /* (256,256) */
  const float *rgb_res_0_w, *rgb_res_0_b;     /* (256,384) */
  const float *view_fc_w, *view_fc_b;         /* (128,283) */
  const float *rgb_res_1_w, *rgb_res_1_b;     /* (128,384) */
  const float *fc_4_w, *fc_4_b;               /* (128,128) */
  const float *rgb_fc_w, *rgb_fc_b;           /* (3,128)   */
} ThWeightsF32;

/* ---- library ---------------------------------------------------------- */
const char* th_version(void);
const char* th_last_error(void);

/* ---- weights (host side, no CUDA call): replaces load_network's role of
 * turning a state_dict into what the kernels read (net_utils.py:361-392) ---- */
size_t th_packed_weights_bytes(int32_t n_views);
int th_pack_weights(const ThWeightsF32* w_host, int32_t n_views, void* packed_host, size_t bytes);

/* ---- fused path --------------------------------------------------------- */
/* Workspace needed by th_render_rays / th_query_density for up to `n_points`
 * sample points in flight (n_rays * n_samples for rays). */
size_t th_workspace_bytes(int64_t n_points, int32_t n_views, int32_t n_verts);
/* The same for the schedule `frame->flags` select (th_render_rays / th_query_density accept either size): with
 * TH_FLAG_PREMAPPED -- the layer-chained kernel, the default path -- the chunk block holds only the feature
 * kernel's operand images (8.25 KB per point of a 284,160-point chunk) and the kernel's fixed per-CTA scratch
 * (138 MB on 148 SMs): 2.6 GB per stream at configs[1] instead of the 7.8 GB th_workspace_bytes reserves for the
 * layer-at-a-time schedule.  with_cull = 0 leaves the cull grid of frame->n_verts vertices out (dense rays).
 * Sized for the current device's SM count (148 when no device is present). */
size_t th_frame_workspace_bytes(const ThFrame* frame, int64_t n_points, int32_t with_cull);

/* Renderer.render (dense: every sample evaluated, pts_mask=None;
 * if_clight_renderer.py:486-498 -> 500-605) when culled == 0, and
 * Renderer.render_fast (K=1 cull at cull_radius, masked points give raw = 0,
 * culled rays give 0; if_clight_renderer.py:429-484 with the chunk loop
 * 607-656, Network.forward cross_transformer.py:207-271 and raw2outputs
 * nerf_net_utils.py:14-59) when culled != 0 (TH_RENDER_*).  Rows a1-a11 of
 * SURVEY 8(a).  Synchronises the stream once when culled != 0 (to read the
 * survivor count, which sizes the launches over the compacted point list and
 * decides the reference's <= 2400-ray branch; measured cost: the launches of
 * a culled 512x512x64 frame add up to within 85 us = 1 % of the frame time).
 * Dense rays whose sample count divides 128 are composited inside the chain
 * kernel's last epilogue (raw2outputs fused: `out->raw` is written only when
 * it is not NULL and no separate integration kernel runs); other sample
 * counts and the culled modes write raw to the workspace and composite with
 * th_integrate's kernel -- the same two device functions, the same bits. */
int th_render_rays(const ThFrame* frame, const ThRays* rays, ThOut* out, int32_t culled,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Mesh renderer's grid query (if_mesh_renderer.py:46-96): cull, zero view
 * direction, alpha_raw only.  pts (P,3) world; alpha_raw (P) gets raw[...,-1]
 * (0 where culled); mask (P) optional.  Row a12. */
int th_query_density(const ThFrame* frame, const float* pts, int64_t n_points, float* alpha_raw,
                     uint8_t* mask, void* workspace, size_t workspace_bytes, void* stream);

/* ---- staged entry points (parity / debugging; same arithmetic as the fused
 * path, reference tensor layouts) ------------------------------------------ */
/* a1: Renderer.get_sampling_points (if_clight_renderer.py:271-287, no jitter):
 * pts (N,S,3), z_vals (N,S). */
int th_sample_points(const ThRays* rays, float* pts, float* z_vals, void* stream);
/* a2: knn_points(pts, verts, K=1) + sqrt + `< radius` (if_clight_renderer.py:440-442,
 * if_mesh_renderer.py:53-55).  Brute force: d2 (P) and idx (P) of the nearest
 * vertex (lower index on ties), mask (P).  Any output may be NULL. */
int th_cull_knn1(const float* pts, int64_t n_points, const float* verts, int32_t n_verts, float radius,
                 float* d2, int64_t* idx, uint8_t* mask, void* stream);
/* a2 (fast): the same mask from a uniform-grid search (exactly equal to the
 * brute-force mask); needs a workspace of th_workspace_bytes(0, 1, n_verts). */
int th_cull_grid(const float* pts, int64_t n_points, const float* verts, int32_t n_verts, float radius,
                 uint8_t* mask, void* workspace, size_t workspace_bytes, void* stream);
/* a3: Renderer.world2smpl (if_clight_renderer.py:289-295). */
int th_world2smpl(const float* pts, int64_t n_points, const float* Rh, const float* Th, float* out,
                  void* stream);
/* a4: view-direction embedding (if_clight_renderer.py:525-526, embedder.py): (N,27). */
int th_view_embed(const float* ray_d, int64_t n_rays, float* out, void* stream);
/* a5: get_pixel_aligned_feature + sample_from_feature_map
 * (if_clight_renderer.py:186-269): pts (P,3) world -> pixel_feat (V,384,P). */
int th_pixel_gather(const ThFrame* frame, const float* pts, int64_t n_points, float* pixel_feat,
                    void* stream);
/* a8: get_human_representation (cross_transformer.py:158-205): pts (P,3) SMPL
 * coordinates -> knn_idx (P,K) int64, knn_d2 (P,K) squared, human_rep (V,255,P).
 * Outputs may be NULL.  workspace (DEVICE, 256-byte aligned, >= th_knn_workspace_bytes(n_tok)) is optional: with it
 * the K nearest tokens are searched through a uniform grid over the tokens (shells of cells until the K-th distance
 * is covered; exactly the same (d2, index)-ordered result as the scan over all tokens, which remains the fall-back
 * for points far from every token) -- what the fused path does for culled rays and grid points from 1024 tokens on;
 * NULL = scan all tokens. */
size_t th_knn_workspace_bytes(int32_t n_tok);
int th_knn_dparf(const ThFrame* frame, const float* pts_smpl, int64_t n_points, int64_t* knn_idx,
                 float* knn_d2, float* human_rep, void* workspace, size_t workspace_bytes, void* stream);
/* a9+a10: the per-point network from its reference inputs
 * (cross_transformer.py:273-353): human_rep (V,255,P), pixel_feat (V,384,P),
 * viewdir (P,27), pts_mask (P) or NULL -> raw (P,4).  With a mask, masked-out
 * points get raw = 0 and rgb is 0 where alpha_raw <= 0 (the progressive
 * variant, 291-311).  workspace >= th_workspace_bytes(P, V, 0). */
int th_mlp_raw(const ThFrame* frame, const float* human_rep, const float* pixel_feat, const float* viewdir,
               const uint8_t* pts_mask, int64_t n_points, float* raw, void* workspace,
               size_t workspace_bytes, void* stream);
/* a11: raw2outputs (nerf_net_utils.py:14-59; raw_noise_std = 0). */
int th_integrate(const float* raw, const float* z_vals, const float* ray_d, int64_t n_rays,
                 int32_t n_samples, int32_t white_bkgd, float* rgb_map, float* acc_map, float* depth_map,
                 void* stream);
/* layout helper: (V,C,H,W) -> (V,H,W,C). */
int th_nchw_to_nhwc(const float* src, float* dst, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);
/* With TH_FLAG_PREMAPPED: encoder output `pixel_feat_map` (V,384,H,W) NCHW (encoder.py:133-146) ->
 * pre-mapped maps (V,H,W,512) channel-last =
 * [alpha_res_0 F + b | view_fc[:, :256] rgb_res_0 F | fc_4 rgb_res_1 F / V], weights from the packed
 * blob (device pointer, th_pack_weights).  Two tcgen05 GEMM launches per view reading the NCHW map as a
 * channel-major operand (no transpose pass).  Replaces th_nchw_to_nhwc for such a frame. */
int th_premap_features(const float* feat_nchw, const void* packed_weights, int32_t n_views, int32_t h, int32_t w,
                       float* out, void* stream);

/* ---- the steps either side of the path --------------------------------------- */
/* Token prologue, one kernel: paint_neural_human + can_body_grouping (if_clight_renderer.py:95-184, 415-427 with
 * voxelization 356-371): project the SMPL vertices `verts` (n_verts,3; batch['input_smpl_vertice']) into every
 * input view, bilinearly sample the encoder's `holder_feat_map` (V,192,H,W) NCHW at uv * (uv_scale) - 1
 * (align_corners, border padding), zero the vertices that `vizmap` (V,n_verts; may be NULL = all visible) marks
 * invisible, and mean-pool each k-means cluster -> tokens (V,n_tok,192).  Clusters in CSR form: the members of
 * cluster c are cluster_members[cluster_start[c] .. cluster_start[c+1]) in the order of the reference's
 * dict_voxel2pc_ind[c]; the mean uses torch-CPU's summation order (csrc/prologue.cu).  `painted` (V,n_verts,192)
 * optionally receives the per-vertex features (= big_holder). */
int th_paint_group(const float* holder_map, int32_t n_views, int32_t h, int32_t w, float uv_scale_x, float uv_scale_y,
                   const float* verts, int32_t n_verts, const float* cam_R, const float* cam_T, const float* cam_K,
                   const uint8_t* vizmap, const int32_t* cluster_start, const int32_t* cluster_members, int32_t n_tok,
                   float* painted, float* tokens, void* stream);
/* ---- the encoder's tail without its full-resolution maps -------------------- */
/* What SpatialEncoder.forward does AFTER the ResNet backbone (encoder.py:133-146): the three latents are bilinearly
 * upsampled (align_corners=True) to the image size, concatenated with upsample_color(images) (a 1x1 convolution
 * 3 -> 128) into pixel_feat_map (V,384,H,W), and reduction_layer (1x1, 384 -> 192) makes holder_feat_map.  The two
 * entry points below take the LATENTS and evaluate that tail per pixel / per vertex on the fly, so neither
 * pixel_feat_map (1.2 GB at 3 x 512 x 512), nor its channel-last transpose, nor holder_feat_map (0.6 GB) is ever
 * written.  All pointers DEVICE fp32, contiguous; images NCHW, latents channel-last. */
typedef struct ThEncoderTail {
  const float* latent[3]; /* CHANNEL-LAST (V,lat_h[0],lat_w[0],64), (V,lat_h[1],lat_w[1],64), (V,lat_h[2],lat_w[2],128):
                           * conv1/bn/relu, layer1, layer2 outputs (encoder.py:113-125, num_layers = 3); what a
                           * torch backbone run in torch.channels_last memory format produces in place         */
  int32_t lat_h[3], lat_w[3];
  const float* images;    /* (V,3,H,W) the encoder's input                                                  */
  const float* color_w;   /* upsample_color.weight (128,3) */
  const float* color_b;   /* upsample_color.bias (128)     */
  int32_t n_views, h, w;
} ThEncoderTail;
/* = th_premap_features(pixel_feat_map of these latents): pre-mapped maps (V,H,W,512) channel-last.  A 1x1
 * convolution commutes with bilinear upsampling, so W_pre is applied to the LOW-RESOLUTION latents (tcgen05 GEMMs
 * over 86,016 instead of 262,144 rows per 512 x 512 view, K = 64 / 64 / 128) and one kernel upsamples the three
 * results, adds the folded colour convolution and writes the channel-last map.  workspace (DEVICE, 256-byte
 * aligned) >= th_premap_from_latents_workspace_bytes(enc) holds the low-resolution maps of one view. */
size_t th_premap_from_latents_workspace_bytes(const ThEncoderTail* enc);
int th_premap_from_latents(const ThEncoderTail* enc, const void* packed_weights, float* out, void* workspace,
                           size_t workspace_bytes, void* stream);
/* = th_paint_group(holder_feat_map of these latents, ...) without `painted`: reduction_w (192,384), reduction_b (192)
 * = reduction_layer.  The reduction and the cluster mean are both linear, so the 384 pixel-feature channels are
 * interpolated at each visible member vertex, summed per cluster and reduced once per (cluster, view): tokens equal
 * the reference's to rounding (a few 1e-7 relative; not bit-equal -- th_paint_group is).  workspace (DEVICE,
 * 256-byte aligned) >= th_paint_group_latents_workspace_bytes(n_views, n_verts, n_tok): the per-vertex 384-channel
 * samples and the per-cluster sums. */
size_t th_paint_group_latents_workspace_bytes(int32_t n_views, int32_t n_verts, int32_t n_tok);
int th_paint_group_latents(const ThEncoderTail* enc, const float* reduction_w, const float* reduction_b,
                           float uv_scale_x, float uv_scale_y, const float* verts, int32_t n_verts, const float* cam_R,
                           const float* cam_T, const float* cam_K, const uint8_t* vizmap, const int32_t* cluster_start,
                           const int32_t* cluster_members, int32_t n_tok, float* tokens, void* workspace,
                           size_t workspace_bytes, void* stream);
/* ---- linear layers of the token transformer on the tensor cores ---------------- */
/* y = x W^T + b for an nn.Linear (vision_transformer.py: qkv 269, proj 276, Mlp.fc1 / fc2 246-250) with the same
 * fp16 hi/lo three-product scheme as the per-point network (fp32-equivalent: ~1e-6 relative), through the tcgen05
 * GEMM: th_linear_pack (HOST) turns weight (n_out, n_in) + bias (n_out, or NULL) into the operand images the kernel
 * copies into shared memory (n_in a multiple of 64; n_out any multiple of 4 -- it is cut into chunks of 256 / 128
 * output columns, the last one zero padded); th_linear runs on the DEVICE copy of that blob: x (m, n_in) fp32 rows with
 * leading dimension ldx, y (m, n_out) with ldy (both multiples of 4, pointers 16-byte aligned).  No workspace, no
 * synchronisation: capturable in a CUDA graph. */
size_t th_linear_packed_bytes(int32_t n_out, int32_t n_in);
int th_linear_pack(const float* weight_host, const float* bias_host, int32_t n_out, int32_t n_in, void* packed_host,
                   size_t bytes);
int th_linear(const float* x, int64_t m, int32_t ldx, const void* packed_dev, int32_t n_out, int32_t n_in, float* y,
              int32_t ldy, int32_t relu, void* stream);

/* ---- mesh extraction ---------------------------------------------------------- */
/* The step after the density-grid query: if_mesh_renderer.py:98-104 pads the cube by 10 voxels on the CPU and calls
 * the third-party `mcubes.marching_cubes(cube, cfg.mesh_th)`.  th_marching_cubes runs marching cubes on a DEVICE
 * volume (nx,ny,nz) fp32 (dim 0 = x slowest ... dim 2 = z fastest, like the numpy cube): a corner is inside when its
 * value is > iso; one vertex per cut lattice edge at a + (iso - v_a) / (v_b - v_a) in index coordinates (fp32; the
 * caller applies voxel_size and the lower bound as lines 106-109 do); triangles (vertex ids) from a 256-case table.
 * PyMCubes is not available to this repository (absent from the reference tree and unpinned), so its table is not
 * reproduced: the table is derived (tools/gen_mc_table.py) -- surfaces differ from mcubes' only in how ambiguous
 * faces are joined and polygons fanned.  Indexed, duplicate-free, deterministic order (vertices by lattice edge,
 * triangles by cube).  counts_host[0..1] (HOST) receive the vertex and triangle counts -- the call synchronises the
 * stream for that; vertices / triangles may be NULL (count only) and are filled up to max_vertices / max_triangles
 * (TH_EWORKSPACE if a non-NULL buffer was too small).  workspace (DEVICE, 256-byte aligned) >=
 * th_marching_cubes_workspace_bytes(nx, ny, nz) (12 bytes per voxel). */
size_t th_marching_cubes_workspace_bytes(int32_t nx, int32_t ny, int32_t nz);
int th_marching_cubes(const float* volume, int32_t nx, int32_t ny, int32_t nz, float iso, float* vertices,
                      int64_t max_vertices, int32_t* triangles, int64_t max_triangles, int64_t* counts_host,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- token transformer attention ---------------------------------------------- */
/* Attention.forward between its two Linear layers (vision_transformer.py:267-275) for vit_tiny (3 heads of 64):
 * qkv (B,N,3,H,64) fp32 = the output of `self.qkv` viewed as in line 269; out (B,N,H*64) fp32 = the input of
 * `self.proj`: softmax(q k^T * scale, dim=-1) v per head, flash-style -- the (B,H,N,N) attention matrix (1.3 GB per
 * layer at 6000 tokens) is never formed.  fp32-equivalent arithmetic (fp16 hi/lo operand split, three tensor-core
 * products, fp32 accumulation and softmax).  head_dim must be 64.  workspace (DEVICE, 256-byte aligned) >=
 * th_vit_attention_workspace_bytes(B, N, H). */
size_t th_vit_attention_workspace_bytes(int32_t batch, int32_t n_tokens, int32_t n_heads);
int th_vit_attention(const float* qkv, int32_t batch, int32_t n_tokens, int32_t n_heads, int32_t head_dim, float scale,
                     float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Renderer.voxelization (if_clight_renderer.py:356-371) of a per-vertex quantity x (n_verts, C), fp32 (is_f64 = 0)
 * or fp64 (blend_mtx, 543-544): out (n_tok, C) of the same type, bit-equal to torch-CPU's `x[idx].mean(0)`:
 * `outer_order` = 1 for wide rows (fp32 C >= 32, fp64 C >= 16: rows added sequentially with a 16-row cascade),
 * 0 for narrow rows such as (n,3) coordinates (four interleaved partial sums). */
int th_group_mean(const void* x, int32_t is_f64, int32_t n_cols, const int32_t* cluster_start,
                  const int32_t* cluster_members, int32_t n_tok, int32_t outer_order, void* out, void* stream);
/* Target-view rays + AABB near/far + compaction (get_rays, get_near_far and the `[mask_at_box]` selection of
 * sample_ray_grid's test split, if_nerf_data_utils.py:11-30, 65-97, 190-199).  K_inv (3,3) = inverse intrinsics,
 * R (3,3), T (3), bounds (2,3) or NULL (rays only), all DEVICE fp32.  Dense outputs ray_o, ray_d (H*W,3), near,
 * far (H*W), mask_at_box (H*W); compacted outputs (optional, all four or none) hold the rays with mask_at_box set,
 * in pixel order, and *count_dev (DEVICE int64) their number.  near/far are evaluated in float64 like the
 * reference.  workspace >= th_generate_rays_workspace_bytes(H*W). */
size_t th_generate_rays_workspace_bytes(int64_t n_pixels);
/* get_near_far alone (if_nerf_data_utils.py:65-97) on caller-supplied rays: float64 arithmetic on the float32 rays,
 * bit-equal to the reference.  ray_d is clamped IN PLACE where |d| < 1e-5, like the reference does (71). */
int th_near_far(const float* ray_o, float* ray_d, int64_t n_rays, const float* bounds, float* near_, float* far_,
                uint8_t* mask_at_box, void* stream);
int th_generate_rays(int32_t h, int32_t w, const float* K_inv, const float* R, const float* T, const float* bounds,
                     float* ray_o, float* ray_d, float* near_, float* far_, uint8_t* mask_at_box, float* ray_o_c,
                     float* ray_d_c, float* near_c, float* far_c, int64_t* count_dev, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Optional device-time profile: between start and stop every kernel launch is
 * bracketed by CUDA events on its stream; stop synchronises the device and
 * returns, per category, the summed elapsed milliseconds and launch counts.
 * Categories: 0 cull, 1 features (sampler + k-NN/DPaRF + pixel gather),
 * 2 GEMM layers, 3 point-wise (attention mix, heads), 4 integration,
 * 5 pre-map GEMM (th_premap_features), 6 token prologue (paint / group / ray kernels). */
#define TH_PROF_NCAT 7
int th_profile_start(void);
int th_profile_stop(double* ms_per_category_host, int64_t* launches_per_category_host, int32_t n);

/* Host-only test hook (needs no device): the job program the layer-chained kernel would run for a chunk
 * of n_points, as a table of int64 (layout documented at the definition in csrc/mlp_chain.cu), so that a
 * CPU test can interpret it against the oracle.  packed_host = the HOST copy of the th_pack_weights blob.
 * Returns the number of words written, or a negative TH_E* code. */
int64_t th_debug_chain_program(const void* packed_host, int32_t n_views, int64_t n_points, int32_t alpha_only,
                               int32_t premapped, int64_t* table, int64_t capacity);

/* Number of kernels this library launched on the calling thread since the last
 * reset (bench.py reports it as gpu_launches). */
int64_t th_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* TRANSHUMAN_B200_H */
