// Internal declarations shared by geometry.cu, mlp_simt.cu, mlp_tc.cu and api.cu.
#pragma once
#include <stddef.h>

#include "common.cuh"

namespace th {

struct CullGrid;
// Frame parameters as the kernels see them (by value).
struct FrameDev {
  const float *tok_feat, *tok_xyz, *tok_rot, *feat, *cam_R, *cam_T, *cam_K, *Rh, *Th;
  int V, n_tok, H, W, K;
  float sx, sy, knn_alpha;
  // uniform grid over the token positions (launch_token_grid) for the exact K-NN of points near the body; nullptr =
  // scan all tokens from shared memory
  const CullGrid* tok_grid;
};

// Output description of k_features.  Element (v, p, c) of the DPaRF
// representation goes to rep[v*rep_sv + p*rep_sp + c*rep_sc]; same for the
// pixel-aligned features.  GEMM layout: sv = P*ld, sp = ld, sc = 1.  Reference
// layout (V,C,P): sv = C*P, sp = 1, sc = P.
struct FeatOut {
  float* rep;
  int64_t rep_sv, rep_sp, rep_sc;
  float* pix;
  int64_t pix_sv, pix_sp, pix_sc;
  float* pix_mean;   // (P, 384) mean over views, GEMM layout only (or nullptr)
  float* vd;         // (P, 32) view-direction embedding, zero padded (or nullptr)
  int64_t* knn_idx;  // (P, K) or nullptr
  float* knn_d2;     // (P, K) or nullptr
  // GEMM-layout outputs can instead be written as fp16 hi/lo tile images (tensor-core
  // path): rep (V*Pp,256), pix (V*Pp,384), pix_mean (Pp,384), vd (Pp,64: 27 + zeros).
  unsigned char *rep_img, *pix_img, *pixm_img, *vd_img;
  int64_t img_view_rows;  // Pp
  int do_rep, do_pix, do_vd;
  int rep_pad;       // write channel 255 = 0 (GEMM layout)
  int pts_are_smpl;  // explicit points are already in SMPL coordinates (staged a8)
};

// Uniform grid over the cull vertices (device memory, built per frame).
struct CullGrid {
  static constexpr int MAX_DIM = 64;
  float ox, oy, oz, inv_h;
  int nx, ny, nz, ncell;
  int* cell_start;   // (ncell + 1)
  int* cursor;       // (ncell) build scratch; after the build: 1 where the 3 x 3 x 3 block around the cell holds a vertex
  float4* sorted;    // (n_verts) vertices grouped by cell
  int covers;        // the grid box spans every vertex + one cell (no dimension was clamped to MAX_DIM)
  float d2_max;      // largest float d2 with sqrt_rn(d2) < radius: `sqrt(d2) < radius` <=> `d2 <= d2_max` (sqrt_rn is monotonic)
  float prune2;      // (radius + 1e-4)^2: a cell row farther than this from the point cannot hold a vertex within radius
  float h;           // cell size
  float4* rowbox;    // (ncell, 2): bounding box (min, max) of the vertices in cells x-1 .. x+1 of the cell's row --
                     // the range one step of the 27-cell scan reads; an empty range has min = +inf, max = -inf
};
inline size_t cull_grid_bytes(int n_verts) {
  size_t cells = (size_t)CullGrid::MAX_DIM * CullGrid::MAX_DIM * CullGrid::MAX_DIM;
  return align_up(sizeof(CullGrid), 256) + align_up((cells + 1) * 4, 256) + align_up(cells * 4, 256) +
         align_up((size_t)n_verts * 16, 256) + align_up(cells * 32, 256);
}

// ---- geometry.cu -------------------------------------------------------------
int launch_sample_points(const PointSource& src, int64_t n_points, float* pts, float* z, cudaStream_t st);
int launch_cull_brute(const PointSource& src, int64_t n_points, const float* verts, int n_verts, float radius,
                      float* d2, int64_t* idx, uint8_t* mask, cudaStream_t st);
int launch_grid_build(const float* verts, int n_verts, float radius, void* grid_mem, cudaStream_t st);
// token grid for the K-NN (geometry.cu: knn_grid): the same structure over tok_xyz, cell size from the token density;
// grid_mem >= cull_grid_bytes(n_tok).  Worth it from TOKEN_GRID_MIN tokens on (1500- and 6000-token configs).
constexpr int TOKEN_GRID_MIN = 1024;
int launch_token_grid(const float* tok_xyz, int n_tok, void* grid_mem, cudaStream_t st);
// cand (optional, >= n_points int32 of scratch; needs counters): two-pass form -- classify, then scan the candidates
// with full warps; counters[3] = candidate count
int launch_cull_grid(const PointSource& src, int64_t n_points, const void* grid_mem, float radius, uint8_t* mask,
                     int32_t* ids, uint8_t* ray_any, unsigned long long* counters, cudaStream_t st,
                     int32_t* cand = nullptr);
int launch_scan_counts(int32_t* counts, int n, unsigned long long* total, cudaStream_t st);
// mesh.cu: marching cubes on a device volume (th_marching_cubes)
size_t marching_cubes_workspace_bytes(int nx, int ny, int nz);
int launch_marching_cubes(const float* vol, int nx, int ny, int nz, float iso, float* verts, int64_t max_verts,
                          int32_t* tris, int64_t max_tris, unsigned long long* counts_dev, void* workspace,
                          cudaStream_t st);
int launch_count_nonzero(const uint8_t* flags, int64_t n, unsigned long long* out, cudaStream_t st);
int launch_expand_rays(const uint8_t* ray_any, int64_t n_points, int S, uint8_t* mask, int32_t* ids,
                       unsigned long long* counter, cudaStream_t st);
int launch_world2smpl(const float* pts, int64_t n, const float* Rh, const float* Th, float* out, cudaStream_t st);
int launch_view_embed(const float* ray_d, int64_t n_rays, float* out, cudaStream_t st);
int launch_features(const FrameDev& fr, const PointSource& src, int64_t n_points, const FeatOut& out,
                    cudaStream_t st, bool premapped = false);
int launch_integrate(const float* raw, const uint8_t* mask, const uint8_t* ray_alive, const PointSource& src,
                     const float* z_vals, const float* ray_d, int64_t n_rays, int S, int white_bkgd, float* rgb,
                     float* acc, float* depth, cudaStream_t st);
int launch_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, cudaStream_t st);

// ---- prologue.cu -------------------------------------------------
int launch_paint_group(const float* map, int V, int H, int W, float sx, float sy, const float* verts, const float* cam_R,
                       const float* cam_T, const float* cam_K, const uint8_t* viz, int n_verts, const int32_t* start,
                       const int32_t* members, int n_tok, float* painted, float* out, cudaStream_t st);
// vit_attn.cu: flash-style self-attention of the token transformer (th_vit_attention), head_dim 64
size_t vit_attention_workspace_bytes(int B, int N, int H);
int launch_vit_attention(const float* qkv, int B, int N, int H, float scale, float* out, void* workspace,
                         cudaStream_t st);
int launch_group_mean(const void* x, int is_f64, int C, const int32_t* start, const int32_t* members, int n_tok, int outer,
                      void* out, cudaStream_t st);
int launch_near_far(const float* ray_o, float* ray_d, int64_t n, const float* bounds, float* near_, float* far_,
                    uint8_t* mask, cudaStream_t st);
size_t generate_rays_workspace_bytes(int64_t n);
int launch_generate_rays(int H, int W, const float* Kinv, const float* R, const float* T, const float* bounds,
                         float* ray_o, float* ray_d, float* near_, float* far_, uint8_t* mask, float* o_c, float* d_c,
                         float* n_c, float* f_c, int64_t* count, void* workspace, cudaStream_t st);

// ---- packed weights (api.cu writes, mlp_*.cu read) --------------------------------
// Offsets (in floats) into the fp32 section of the blob.  Folded matrices are
// built in float64 by th_pack_weights:
//   W_v    = [value_embed_1 | value_embed_0]            b_v = b_v1 + b_v0
//   W_fc3m = [fc_3/V | ... | fc_3/V]   (mean over views folded into K)
//   W_f    = [feature_fc | rgb_res_0]                   b_f = b_ff + b_r0
//   W_view = [view_fc[:, :256] | view_fc[:, 256:283] | 0(37)]   (K = 320)
//   W_t    = [fc_4/V | ... | fc_4/V | fc_4 @ rgb_res_1] b_t = fc_4 @ b_r1 + b_4
// Tensor-core schedule only (no non-linearity sits between the folded layers, so
// the products are exact in real arithmetic; the fp32 CUDA-core schedule keeps the
// unfolded W_v / fc_1 and W_f / W_view and so cross-checks the folds):
//   W_fc1f = [fc_1 @ value_embed_1 | fc_1 @ value_embed_0]      b = fc_1 @ b_v + b_fc1
//   W_gvf  = [V1 @ feature_fc | V1 @ rgb_res_0 | view_fc[:, 256:283] | 0(37)]  (K = 704)
//            with V1 = view_fc[:, :256];  b = V1 @ b_f + b_view
struct PackedHeader {
  uint32_t magic;      // 'THW5'
  int32_t n_views;
  uint64_t total_bytes;
  // fp32 matrices, row-major (N, K) with K contiguous; offsets in bytes from blob start
  uint64_t fc0_w, fc0_b;        // (256,256)  K = 255 + 1 zero
  uint64_t ar0_w, ar0_b;        // (256,384)
  uint64_t k0_w, k0_b;          // (128,256)
  uint64_t k1_w, k1_b;          // (128,256)
  uint64_t v_w, v_b;            // (256,512)
  uint64_t fc1_w, fc1_b;        // (256,256)
  uint64_t fc2_w, fc2_b;        // (256,256)
  uint64_t fc3m_w, fc3m_b;      // (256,256*V)
  uint64_t afc_w, afc_b;        // (256), (1)
  uint64_t f_w, f_b;            // (256,640)
  uint64_t view_w, view_b;      // (128,320)
  uint64_t t_w, t_b;            // (128,128*V+384)
  uint64_t rgb_w, rgb_b;        // (3,128), (3)
  uint64_t fc1f_w, fc1f_b;      // (256,512)
  uint64_t gvf_w, gvf_b;        // (128,704)
  // fp16 hi/lo tile images for the tensor-core path (see th_pack_weights)
  uint64_t h_fc0, h_ar0, h_k0, h_k1, h_v, h_fc1, h_fc2, h_fc3m, h_f, h_view, h_t, h_fc1f, h_gvf;
  // Pre-mapped feature maps (packed, not yet consumed by a kernel):
  // the three 1x1 convolutions that read the bilinear blend of the feature maps commute with the
  // blend, so they can be applied once per frame to the (V,H,W,384) maps instead of once per point:
  //   W_pre = [alpha_res_0 ; V1 @ rgb_res_0 ; fc_4 @ rgb_res_1 / V]  (512,384), b_pre = [b_ar0 ; 0 ; 0]
  //           (the other two biases already sit in b_gvf and b_t; bilinear weights sum to 1)
  //   W_gvfp = [V1 @ feature_fc | I_128 | view_fc[:, 256:283] | 0(37)]  (K = 448): the blended rows
  //            256..383 of the pre-mapped map enter view_fc' through an identity block
  //   W_tp   = [fc_4/V | ... | fc_4/V | I_128]  (K = 128 V + 128): the view sum of rows 384..511
  //   W_xid  = I_256: copies the blended, rectified rows 0..255 (= X_v) into the chain kernel's scratch
  uint64_t pre_w, pre_b;        // (512,384), (512)
  uint64_t gvfp_w, gvfp_b;      // (128,448)
  uint64_t tp_w, tp_b;          // (128,128*V+128)
  uint64_t xid_w, xid_b;        // (256,256), zero bias
  uint64_t h_gvfp, h_tp, h_xid;
  // second half of W_pre as a GEMM matrix of its own ([V1 @ rgb_res_0 ; fc_4 @ rgb_res_1 / V], zero bias): the
  // pre-map GEMM (th_premap_features) runs as two N = 256 tcgen05 launches, h_ar0 / ar0_b and h_preb / preb_b
  uint64_t preb_w, preb_b;      // (256,384), (256)
  uint64_t h_preb;
  // Per-matrix power-of-two scale of the fp16 hi/lo images: the image holds W * 2^e with e chosen so that
  // max|W| 2^e lies in (2^13, 2^14]; every epilogue multiplies its accumulator by 2^-e (exact) before the bias.
  // The fp16 lo plane is subnormal below 6.1e-5 (absolute step 6e-8): unscaled, a weight of 1e-3 would keep
  // 15 instead of 22 significant bits and weights below 6e-5 almost none of the lo half.
  uint64_t img_off[24];
  float img_inv_scale[24];
  int32_t n_img, pad_;
  // fc_1' = [fc_1 @ v1 | fc_1 @ v0] cut into its S part and its X part and into two halves of 128 output rows
  // each (N = 128 jobs of the TMEM-side attention mix, mlp_chain.cu): fc1s{h} = W_fc1f[128h:128h+128, 0:256]
  // with bias b_fc1f[128h:], fc1x{h} = W_fc1f[128h:128h+128, 256:512] with zero bias
  uint64_t fc1s0_w, fc1s0_b, fc1s1_w, fc1s1_b, fc1x0_w, fc1x0_b, fc1x1_w, fc1x1_b;
  uint64_t h_fc1s0, h_fc1s1, h_fc1x0, h_fc1x1;
};
// DEVICE address of the 2^-e of the weight image at byte offset `off` (nullptr = unknown image, scale 1).  The scale
// is per blob (it depends on the weights), so kernels read it from the blob itself; the host only needs the INDEX,
// which -- like every offset in this header -- is a pure function of the view count (safe to cache across blobs).
inline const float* img_inv_scale_ptr(const unsigned char* weights_dev, const PackedHeader& h, uint64_t off) {
  for (int i = 0; i < h.n_img && i < 24; ++i)
    if (h.img_off[i] == off)
      return reinterpret_cast<const float*>(weights_dev + offsetof(PackedHeader, img_inv_scale)) + i;
  return nullptr;
}
constexpr uint32_t PACK_MAGIC = 0x35574854u;

// Byte offset of the hi element (row, col) of a (rows, C) activation in tile-image
// format; the lo element sits 16384 bytes further.
__host__ __device__ inline size_t img_offset(int64_t row, int col, int C) {
  const int64_t tile = row >> 7;
  const int r = (int)(row & 127), kb = col >> 6, kk = col & 63;
  return ((size_t)(tile * (C >> 6) + kb) << 15) + (size_t)r * 128 + (size_t)(((kk >> 3) ^ (r & 7)) << 4) +
         (size_t)(kk & 7) * 2;
}

// ---- per-point network on GEMM-layout activations (mlp_simt.cu / mlp_tc.cu) ----
// Activation buffers of one chunk (fp32, row-major, rows = view-major (v*Pp + p)
// with Pp = P rounded up to 256 so that every view starts on a 2-CTA super-tile;
// on the tensor-core path S, NET, N1, INTER, F, G hold tile images instead).
inline int64_t pad_points(int64_t P) { return (P + 255) / 256 * 256; }
struct MlpBuffers {
  float *rep;      // (V*P, 256)
  float *pix;      // (V*P, 384)
  float *pix_mean; // (P, 384)
  float *vd;       // (P, 32)
  float *s, *x;    // (V*P, 256) each
  float *kp, *ks;  // (V*P, 128) each
  float *xt;       // (V*P, 256)
  float *net;      // (V*P, 256)
};
size_t mlp_buffer_floats_per_point(int V);
// chain_scratch > 0: compact carve for the layer-chained schedule on pre-mapped maps (mlp_simt.cu)
void mlp_carve(float* base, int64_t P, int V, MlpBuffers* b, size_t chain_scratch = 0);
size_t mlp_compact_bytes(int64_t P, int V, size_t chain_scratch);

// Runs fc_0 ... rgb_fc for P points whose inputs sit in `b`; writes raw rows
// (rgb x3, alpha) to raw[dst_ids ? dst_ids[first + i] : first + i].  alpha_only
// skips the colour branch (a12).  zero_rgb_if_transparent reproduces the
// progressive variant's output (rgb = 0 where alpha_raw <= 0).
// Fused compositing (chain schedule, dense rays): raw2outputs (nerf_net_utils.py:14-59) inside the fc_4' epilogue.
// A 128-row tile of the chain kernel holds whole rays when the sample count divides 128 and the chunk starts on a ray
// boundary, so a ray's alpha, transmittance and weighted sums never leave the CTA: no raw tensor, no k_integrate.
struct CompositeArgs {
  float *rgb_map, *acc_map, *depth_map;  // rgb_map == nullptr: off (raw is written, k_integrate composites)
  const float *near_, *far_, *t_vals, *ray_d;
  int32_t S, white_bkgd;
};
struct MlpRun {
  const unsigned char* weights;  // device blob
  int64_t P;
  int V;
  const int32_t* dst_ids;
  int64_t first;
  float* raw;         // (.., 4) or nullptr
  float* alpha_out;   // (..) or nullptr (density query)
  int alpha_only;
  int zero_rgb_if_transparent;
  int use_tensor_cores;
  int inputs_are_images;  // rep / pix / pix_mean / vd buffers hold tile images (fused path)
  // experimental (TH_FLAG_PREMAPPED, chain schedule only): the pix block holds the images
  // [X (V*Pp,256) | P2 (V*Pp,128)] and pix_mean the image R (Pp,128) -- see k_features PRE
  int premapped;
  CompositeArgs cmp;  // chain schedule only; requires dst_ids == nullptr, first % S == 0, P % S == 0, 128 % S == 0
  // test hook (th_debug_chain_program): mlp_forward_chain copies its job program here and returns
  // before touching the device
  void* program_dump;
};
// Pp = pad_points(P): view stride of every buffer in `b`.
int mlp_forward(const MlpRun& run, const MlpBuffers& b, const PackedHeader& hdr_host, cudaStream_t st);

// mlp_chain.cu: the same network as ONE launch per chunk (layer-chained persistent kernel;
// inputs must be the feature kernel's tile images, V <= 3).  `scratch` is 1 KB aligned and
// holds chain_scratch_bytes(P, V, sm count) bytes; `alpha` (Pp floats) may be nullptr.
bool chain_supported(int V);
size_t chain_scratch_bytes(int64_t P, int V, int num_sms);
int mlp_forward_chain(const MlpRun& run, const MlpBuffers& b, const PackedHeader& hdr_host, unsigned char* scratch,
                      float* alpha, cudaStream_t st);

// mlp_simt.cu: C[M,N] = act(sum_seg A_seg[M,K_seg] W[:, koff:koff+K_seg]^T + bias)
// One K-segment of the A operand.  Either fp32 rows (`ptr`, `ld`), converted to
// fp16 hi/lo on the fly, or -- tensor-core path only -- an activation already in
// operand TILE-IMAGE format (`img`): per 128-row tile and 64-wide k-block a
// 32 KB block [hi 128x128B | lo 128x128B], K-major, 128-byte swizzled, tiles
// row-major over (row tile, k-block); `K` must then be a multiple of 64.
struct GemmSeg {
  const float* ptr;
  int ld;       // row stride in floats
  int K;        // multiple of 16
  int64_t row_mod;  // rows wrap modulo this (0 = no wrap)
  const unsigned char* img;
  int64_t img_tile_mod;  // image segments: row tiles wrap modulo this (0 = no wrap)
  // fp32 segments, tensor-core path: non-zero = the operand is stored K-MAJOR-OUTER ("channel-major"):
  // element (row m, column k) sits at ptr[k * col_stride + m] -- an NCHW feature map read as a
  // (pixels, channels) matrix without a transpose pass (th_premap_features)
  int64_t col_stride;
};
// SpatialEncoder.forward after the backbone (encoder.py:133-146): pixel_feat_map = [up(latent_0) 64 | up(latent_1) 64 |
// up(latent_2) 128 | upsample_color(image) 128], `up` = bilinear, align_corners=True (PyTorch's CUDA formula), evaluated
// where it is consumed: th_premap_from_latents applies W_pre to the LOW-RESOLUTION latents (1x1 convolutions commute with
// bilinear upsampling) and upsamples the results, th_paint_group_latents interpolates per vertex.  Neither the
// (V,384,H,W) map, nor a transpose of it, nor the (V,192,H,W) holder map ever exists.  Pointers of ONE view.
struct EncTail {
  const float* lat[3];  // (lh, lw, 64 | 64 | 128) latents of the view, CHANNEL-LAST
  int lh[3], lw[3];
  const float* img;     // (3, H, W)
  const float* wc;      // upsample_color.weight (128, 3)
  const float* bc;      // upsample_color.bias (128)
  int H, W;
};
// value of the upsampled plane at full-resolution texel (y, x): src = dst * (in - 1) / (out - 1), i0 = (int)src,
// l1 = src - i0, val = h0 (w0 v00 + w1 v01) + h1 (w0 v10 + w1 v11) (UpSampleBilinear2d.cu, align_corners=True)
struct UpTap {
  int o00, o01, o10, o11;
  float h0, h1, w0, w1;
};
__device__ __forceinline__ UpTap up_tap(int lh, int lw, int H, int W, int y, int x) {
  const float rh = H > 1 ? (float)(lh - 1) / (float)(H - 1) : 0.f;
  const float rw = W > 1 ? (float)(lw - 1) / (float)(W - 1) : 0.f;
  const float sy = rh * (float)y, sx = rw * (float)x;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < lh - 1 ? 1 : 0), x1 = x0 + (x0 < lw - 1 ? 1 : 0);
  UpTap t;
  t.h1 = sy - (float)y0, t.h0 = 1.f - t.h1, t.w1 = sx - (float)x0, t.w0 = 1.f - t.w1;
  t.o00 = y0 * lw + x0, t.o01 = y0 * lw + x1, t.o10 = y1 * lw + x0, t.o11 = y1 * lw + x1;
  return t;
}
// paint + group straight from the latents (th_paint_group_latents): holder_feat_map = reduction_layer(pixel_feat_map)
// is linear, and so is the cluster mean, so the 384 pixel-feature channels are interpolated at every visible vertex
// (-> painted (V, n_verts, 384), scratch), summed per cluster, and the 192 x 384 reduction is applied ONCE per
// (cluster, view).  enc = V views (device array); red_w (192,384), red_b (192) = reduction_layer; scratch >=
// paint_latents_scratch_bytes.
size_t paint_latents_scratch_bytes(int V, int n_verts, int n_tok);
int launch_paint_group_latents(const EncTail* enc_dev, int V, const float* red_w, const float* red_b, float sx, float sy,
                               const float* verts, const float* cam_R, const float* cam_T, const float* cam_K,
                               const uint8_t* viz, int n_verts, const int32_t* start, const int32_t* members, int n_tok,
                               void* scratch, float* out, cudaStream_t st);
struct GemmArgs {
  GemmSeg seg[TH_MAX_VIEWS + 1];
  int nseg;
  const float* W;  // (N, Ktot) row-major
  int ldw;
  const float* bias;
  float* C;            // fp32 output (M, ldc) or nullptr
  int ldc;
  unsigned char* C_img;  // tile-image output (tensor-core path) or nullptr
  int64_t M;
  int N;  // multiple of 128
  int relu;
  int n_store;             // fp32 C output: only columns < n_store are written (0 = all N): the padded last chunk of a
                           // linear layer whose width is not a multiple of 128
  const float* acc_scale;  // tensor-core path: DEVICE pointer to the 2^-e of the weight image (img_inv_scale_ptr);
                           // the accumulator is multiplied by it before the bias; nullptr = 1
};
int launch_gemm_simt(const GemmArgs& a, cudaStream_t st);
int launch_gemm_tc(const GemmArgs& a, const void* w_hi_lo, cudaStream_t st, int prof_cat = PROF_GEMM);
// mlp_tc.cu: pre-mapped feature maps.  dst[v][hw][n] = b[n] + sum_k W_pre[n][k] src[v][k][hw]: (V,384,H,W) NCHW in (the
// encoder's layout, encoder.py:133-146), (V,H,W,512) channel-last out -- the layout change of th_nchw_to_nhwc and the
// three 1x1 convolutions that read the blended maps, as tcgen05 GEMMs over the maps (2 launches of N = 256 per view).
int launch_premap(const float* src_nchw, const unsigned char* weights, const PackedHeader& hdr, float* dst, int n_views,
                  int h, int w, cudaStream_t st);
// the same maps straight from the encoder's latents (prologue.cu): enc[v] describes view v; `scratch`
// holds the low-resolution pre-mapped maps of ONE view, premap_latents_scratch_bytes(enc[0])
size_t premap_latents_scratch_bytes(const int* lh, const int* lw);
int launch_premap_latents(const EncTail* enc_host, const unsigned char* weights, const PackedHeader& hdr, float* dst,
                          int n_views, void* scratch, cudaStream_t st);
int launch_pack_inputs(const float* human_rep, const float* pixel_feat, const float* viewdir, int64_t P, int V,
                       const MlpBuffers& b, cudaStream_t st);

}  // namespace th
