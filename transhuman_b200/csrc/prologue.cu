// The steps either side of the query path, one kernel each instead of the reference's
// Python loops / numpy:
//   * token prologue: project the 6890 SMPL vertices into every input view, bilinearly sample the 192-channel
//     holder map, mask by visibility, and mean-pool per k-means cluster (paint_neural_human +
//     can_body_grouping + voxelization, if_clight_renderer.py:95-184, 356-371, 415-427), plus the per-cluster
//     means of the token coordinates / blend matrices (541-547);
//   * ray generation for a target camera and the AABB near / far test with ray compaction
//     (if_nerf_data_utils.py:11-30, 65-97, 190-199).
// Summation orders of the cluster means are those of torch's CPU `x[idx].mean(0)` as measured in the build
// container (AVX2 kernels of aten/native/cpu/SumKernel.cpp), so that tokens are bit-equal to the reference's
// CPU `voxelization` (the goldens' tok_xyz / tok_rot are checked bitwise):
//   "outer" (>= 4 vectors of columns: fp32 C >= 32, fp64 C >= 16): rows added one by one into a level-0 accumulator
//           that is flushed into level 1 every 16 rows (level 2 every 256, ...), levels summed at the end;
//   "row"   (narrow rows, e.g. fp32 (n,3)): four interleaved partial sums over the first 4*floor(n/4) rows (each with
//           the same cascade every 16 of ITS rows), the tail rows added to partial 0, then p0 + p1 + p2 + p3.
// Both divide by n (true division).
#include <stdlib.h>

#include "kernels.cuh"

namespace th {

// ---- cascade accumulator of SumKernel.cpp's multi_row_sum (level step 16 for n < 2^20) ----
template <typename T>
struct Cascade {
  T acc[4];
  int64_t i;
  __device__ Cascade() : i(0) { acc[0] = acc[1] = acc[2] = acc[3] = T(0); }
  __device__ void add(T x) {
    acc[0] += x;
    ++i;
    if ((i & 15) == 0) {
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        acc[j] += acc[j - 1];
        acc[j - 1] = T(0);
        if ((i & ((int64_t)15 << (4 * j))) != 0) break;
      }
    }
  }
  // rows that do not fill a whole level-0 block are added without a flush (the reference's tail loop)
  __device__ T total() const { return ((acc[0] + acc[1]) + acc[2]) + acc[3]; }
};
// The reference flushes only after COMPLETE blocks of 16: `add` above flushes exactly when i hits a multiple of 16,
// which is the same thing (a partial last block never reaches a multiple of 16).

template <typename T>
__device__ __forceinline__ T mean_outer(const T* __restrict__ x, int64_t ld, const int32_t* __restrict__ members, int n) {
  Cascade<T> c;
  for (int r = 0; r < n; ++r) c.add(x[(int64_t)members[r] * ld]);
  return c.total() / (T)n;
}

template <typename T>
__device__ __forceinline__ T mean_row(const T* __restrict__ x, int64_t ld, const int32_t* __restrict__ members, int n) {
  const int n4 = n / 4;
  Cascade<T> c[4];
  for (int r = 0; r < n4; ++r) {
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k].add(x[(int64_t)members[4 * r + k] * ld]);
  }
  T p[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = c[k].total();
  for (int r = 4 * n4; r < n; ++r) p[0] += x[(int64_t)members[r] * ld];
  p[0] += p[1];
  p[0] += p[2];
  p[0] += p[3];
  return p[0] / (T)n;
}

// out[c][col] = mean over the members of cluster c of x[member][col]; one thread per (cluster, column)
template <typename T>
__global__ void k_group_mean(const T* __restrict__ x, int C, const int32_t* __restrict__ start,
                             const int32_t* __restrict__ members, int n_tok, int outer, T* __restrict__ out) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= (int64_t)n_tok * C) return;
  const int c = (int)(g / C), col = (int)(g - (int64_t)c * C);
  const int b = start[c], n = start[c + 1] - b;
  out[g] = outer ? mean_outer(x + col, C, members + b, n) : mean_row(x + col, C, members + b, n);
}

// projection of a vertex into a view and its bilinear taps: separately rounded products, left to right, like the
// batched matmuls of the reference's CPU path; align_corners=True, border padding (ATen grid_sample)
struct VertTap {
  int x0i, y0i, x1i, y1i;
  float wx, wy, ex, ey;
};
__device__ __forceinline__ VertTap vertex_tap(const float* __restrict__ verts, int vi, const float* __restrict__ R,
                                              const float* __restrict__ T, const float* __restrict__ Km, float sx,
                                              float sy, int H, int W) {
  const float px = verts[vi * 3], py = verts[vi * 3 + 1], pz = verts[vi * 3 + 2];
  float xc[3], xk[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    xc[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[i * 3], px), __fmul_rn(R[i * 3 + 1], py)),
                                __fmul_rn(R[i * 3 + 2], pz)),
                      T[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    xk[i] = __fadd_rn(__fadd_rn(__fmul_rn(Km[i * 3], xc[0]), __fmul_rn(Km[i * 3 + 1], xc[1])),
                      __fmul_rn(Km[i * 3 + 2], xc[2]));
  const float u = __fdiv_rn(xk[0], xk[2]), w_ = __fdiv_rn(xk[1], xk[2]);
  const float gx = __fsub_rn(__fmul_rn(u, sx), 1.0f), gy = __fsub_rn(__fmul_rn(w_, sy), 1.0f);
  float ix = __fmul_rn(__fadd_rn(gx, 1.0f), 0.5f * (float)(W - 1));
  float iy = __fmul_rn(__fadd_rn(gy, 1.0f), 0.5f * (float)(H - 1));
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float x0 = floorf(ix), y0 = floorf(iy);
  VertTap t;
  t.wx = __fsub_rn(ix, x0), t.wy = __fsub_rn(iy, y0);
  t.ex = __fsub_rn(1.0f, t.wx), t.ey = __fsub_rn(1.0f, t.wy);
  t.x0i = (int)x0, t.y0i = (int)y0;
  t.x1i = min(t.x0i + 1, W - 1), t.y1i = min(t.y0i + 1, H - 1);
  return t;
}
// FMA chain over (nw, ne, sw, se)
__device__ __forceinline__ float tap_blend(const VertTap& t, float a, float bq, float cq, float d) {
  return __fmaf_rn(d, __fmul_rn(t.wy, t.wx),
                   __fmaf_rn(cq, __fmul_rn(t.wy, t.ex), __fmaf_rn(bq, __fmul_rn(t.ey, t.wx), __fmul_rn(a, __fmul_rn(t.ey, t.ex)))));
}

// ---------------------------------------------------------------------------
// paint + group: grid (n_tok, V), 192 threads = channels.  For every member vertex of the cluster: project into
// view v (separately rounded products, left to right, like the batched matmuls of the reference's CPU path),
// bilinear taps with align_corners=True / border padding (ATen grid_sample), FMA chain over (nw, ne, sw, se),
// zero where the vertex is not visible in the view, cluster mean in the "outer" order.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TH_C_TOK) k_paint_group(const float* __restrict__ map, int V, int H, int W, float sx,
                                                          float sy, const float* __restrict__ verts,
                                                          const float* __restrict__ cam_R, const float* __restrict__ cam_T,
                                                          const float* __restrict__ cam_K, const uint8_t* __restrict__ viz,
                                                          int n_verts, const int32_t* __restrict__ start,
                                                          const int32_t* __restrict__ members, int n_tok,
                                                          float* __restrict__ painted, float* __restrict__ out) {
  const int c = blockIdx.x, v = blockIdx.y, ch = threadIdx.x;
  const int b = start[c], n = start[c + 1] - b;
  const float* R = cam_R + v * 9;
  const float* T = cam_T + v * 3;
  const float* Km = cam_K + v * 9;
  const float* plane = map + ((int64_t)v * TH_C_TOK + ch) * H * W;
  Cascade<float> acc;
  for (int r = 0; r < n; ++r) {
    const int vi = members[b + r];
    float val = 0.f;
    if (!viz || viz[(int64_t)v * n_verts + vi]) {
      const VertTap t = vertex_tap(verts, vi, R, T, Km, sx, sy, H, W);
      const float a = __ldg(plane + t.y0i * W + t.x0i), bq = __ldg(plane + t.y0i * W + t.x1i);
      const float cq = __ldg(plane + t.y1i * W + t.x0i), d = __ldg(plane + t.y1i * W + t.x1i);
      val = tap_blend(t, a, bq, cq, d);
    }
    if (painted) painted[((int64_t)v * n_verts + vi) * TH_C_TOK + ch] = val;
    acc.add(val);
  }
  out[((int64_t)v * n_tok + c) * TH_C_TOK + ch] = acc.total() / (float)n;
}

// ---------------------------------------------------------------------------
// paint + group straight from the encoder's latents (kernels.cuh: EncTail; latents channel-last), three kernels.
//  k_paint_latents: one warp per (vertex, view).  Per visible vertex the four full-resolution texels of every channel
//    are evaluated from the latents (bilinear upsampling: 4 coalesced row loads per texel and level; colour
//    convolution from the image) and blended like grid_sample -> painted (V, n_verts, 384), zero rows if invisible.
//    Lane owns channels {2l, 2l+1} of levels 0 and 1, {4l..4l+3} of level 2 and of the colour block.
//  k_cluster_sum_latents: grid (n_tok, V), 96 threads x float4: sum of the member rows, visible count.
//  k_reduce_rows_latents: reduction_layer as a small GEMM, 24 (cluster, view) rows per block in shared memory,
//    thread = (output channel, 12 rows): token = (W_r sum + n_visible b_r) / n.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float2 lerp4(float2 a, float2 b, float2 c, float2 d, const UpTap& t) {
  return make_float2(t.h0 * (t.w0 * a.x + t.w1 * b.x) + t.h1 * (t.w0 * c.x + t.w1 * d.x),
                     t.h0 * (t.w0 * a.y + t.w1 * b.y) + t.h1 * (t.w0 * c.y + t.w1 * d.y));
}
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float4 c, float4 d, const UpTap& t) {
  return make_float4(t.h0 * (t.w0 * a.x + t.w1 * b.x) + t.h1 * (t.w0 * c.x + t.w1 * d.x),
                     t.h0 * (t.w0 * a.y + t.w1 * b.y) + t.h1 * (t.w0 * c.y + t.w1 * d.y),
                     t.h0 * (t.w0 * a.z + t.w1 * b.z) + t.h1 * (t.w0 * c.z + t.w1 * d.z),
                     t.h0 * (t.w0 * a.w + t.w1 * b.w) + t.h1 * (t.w0 * c.w + t.w1 * d.w));
}
constexpr int PL_WARPS = 8;
__global__ void __launch_bounds__(PL_WARPS * 32) k_paint_latents(const EncTail* __restrict__ encs, float sx, float sy,
                                                                 const float* __restrict__ verts,
                                                                 const float* __restrict__ cam_R,
                                                                 const float* __restrict__ cam_T,
                                                                 const float* __restrict__ cam_K,
                                                                 const uint8_t* __restrict__ viz, int n_verts,
                                                                 float* __restrict__ painted) {
  const int v = blockIdx.y, lane = threadIdx.x & 31;
  const int vi = blockIdx.x * PL_WARPS + (threadIdx.x >> 5);
  if (vi >= n_verts) return;
  float* dst = painted + ((int64_t)v * n_verts + vi) * TH_C_PIX;
  float2 o0 = make_float2(0.f, 0.f), o1 = o0;
  float4 o2 = make_float4(0.f, 0.f, 0.f, 0.f), o3 = o2;
  if (!viz || viz[(int64_t)v * n_verts + vi]) {
    const EncTail e = encs[v];
    const int H = e.H, W = e.W;
    const VertTap t = vertex_tap(verts, vi, cam_R + v * 9, cam_T + v * 3, cam_K + v * 9, sx, sy, H, W);
    float2 q0[4], q1[4];
    float4 q2[4], q3[4];
    float4 wr, wg, wb, wbias;  // colour convolution of the lane's 4 channels
    {
      const float* w = e.wc + lane * 12;
      wr = make_float4(w[0], w[3], w[6], w[9]), wg = make_float4(w[1], w[4], w[7], w[10]);
      wb = make_float4(w[2], w[5], w[8], w[11]);
      wbias = __ldg(reinterpret_cast<const float4*>(e.bc) + lane);
    }
    const int64_t hw = (int64_t)H * W;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = (k & 2) ? t.y1i : t.y0i, xx = (k & 1) ? t.x1i : t.x0i;
      {
        const UpTap u = up_tap(e.lh[0], e.lw[0], H, W, yy, xx);
        const float2* p = reinterpret_cast<const float2*>(e.lat[0]) + lane;
        q0[k] = lerp4(__ldg(p + u.o00 * 32), __ldg(p + u.o01 * 32), __ldg(p + u.o10 * 32), __ldg(p + u.o11 * 32), u);
      }
      {
        const UpTap u = up_tap(e.lh[1], e.lw[1], H, W, yy, xx);
        const float2* p = reinterpret_cast<const float2*>(e.lat[1]) + lane;
        q1[k] = lerp4(__ldg(p + u.o00 * 32), __ldg(p + u.o01 * 32), __ldg(p + u.o10 * 32), __ldg(p + u.o11 * 32), u);
      }
      {
        const UpTap u = up_tap(e.lh[2], e.lw[2], H, W, yy, xx);
        const float4* p = reinterpret_cast<const float4*>(e.lat[2]) + lane;
        q2[k] = lerp4(__ldg(p + u.o00 * 32), __ldg(p + u.o01 * 32), __ldg(p + u.o10 * 32), __ldg(p + u.o11 * 32), u);
      }
      const int64_t o = (int64_t)yy * W + xx;
      const float r = __ldg(e.img + o), g = __ldg(e.img + hw + o), b = __ldg(e.img + 2 * hw + o);
      q3[k] = make_float4(wr.x * r + wg.x * g + wb.x * b + wbias.x, wr.y * r + wg.y * g + wb.y * b + wbias.y,
                          wr.z * r + wg.z * g + wb.z * b + wbias.z, wr.w * r + wg.w * g + wb.w * b + wbias.w);
    }
    o0 = make_float2(tap_blend(t, q0[0].x, q0[1].x, q0[2].x, q0[3].x), tap_blend(t, q0[0].y, q0[1].y, q0[2].y, q0[3].y));
    o1 = make_float2(tap_blend(t, q1[0].x, q1[1].x, q1[2].x, q1[3].x), tap_blend(t, q1[0].y, q1[1].y, q1[2].y, q1[3].y));
    o2 = make_float4(tap_blend(t, q2[0].x, q2[1].x, q2[2].x, q2[3].x), tap_blend(t, q2[0].y, q2[1].y, q2[2].y, q2[3].y),
                     tap_blend(t, q2[0].z, q2[1].z, q2[2].z, q2[3].z), tap_blend(t, q2[0].w, q2[1].w, q2[2].w, q2[3].w));
    o3 = make_float4(tap_blend(t, q3[0].x, q3[1].x, q3[2].x, q3[3].x), tap_blend(t, q3[0].y, q3[1].y, q3[2].y, q3[3].y),
                     tap_blend(t, q3[0].z, q3[1].z, q3[2].z, q3[3].z), tap_blend(t, q3[0].w, q3[1].w, q3[2].w, q3[3].w));
  }
  reinterpret_cast<float2*>(dst)[lane] = o0;
  reinterpret_cast<float2*>(dst + 64)[lane] = o1;
  reinterpret_cast<float4*>(dst + 128)[lane] = o2;
  reinterpret_cast<float4*>(dst + 256)[lane] = o3;
}

// sums (V * n_tok, 384 + 2): [sum of the member rows | n_visible | n]
constexpr int SUM_LD = TH_C_PIX + 4;
__global__ void __launch_bounds__(TH_C_PIX / 4) k_cluster_sum_latents(const float* __restrict__ painted,
                                                                      const uint8_t* __restrict__ viz, int n_verts,
                                                                      const int32_t* __restrict__ start,
                                                                      const int32_t* __restrict__ members, int n_tok,
                                                                      float* __restrict__ sums) {
  const int c = blockIdx.x, v = blockIdx.y, q = threadIdx.x;
  const int b = start[c], n = start[c + 1] - b;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int n_vis = 0;
#pragma unroll 4
  for (int m = 0; m < n; ++m) {
    const int vi = members[b + m];
    const float4 x = __ldg(reinterpret_cast<const float4*>(painted + ((int64_t)v * n_verts + vi) * TH_C_PIX) + q);
    acc.x += x.x, acc.y += x.y, acc.z += x.z, acc.w += x.w;
    if (q == 0) n_vis += (!viz || viz[(int64_t)v * n_verts + vi]) ? 1 : 0;
  }
  float* row = sums + ((int64_t)v * n_tok + c) * SUM_LD;
  reinterpret_cast<float4*>(row)[q] = acc;
  if (q == 0) row[TH_C_PIX] = (float)n_vis, row[TH_C_PIX + 1] = (float)n;
}

constexpr int RR_ROWS = 24;
__global__ void __launch_bounds__(2 * TH_C_TOK) k_reduce_rows_latents(const float* __restrict__ sums, int n_rows,
                                                                      const float* __restrict__ red_w,
                                                                      const float* __restrict__ red_b,
                                                                      float* __restrict__ out) {
  __shared__ float4 s_x[RR_ROWS][SUM_LD / 4];
  const int row0 = blockIdx.x * RR_ROWS, tid = threadIdx.x;
  for (int i = tid; i < RR_ROWS * (SUM_LD / 4); i += 2 * TH_C_TOK) {
    const int r = i / (SUM_LD / 4), k = i - r * (SUM_LD / 4);
    s_x[r][k] = row0 + r < n_rows ? __ldg(reinterpret_cast<const float4*>(sums + (int64_t)(row0 + r) * SUM_LD) + k)
                                  : make_float4(0.f, 0.f, 0.f, 1.f);
  }
  __syncthreads();
  const int oc = tid % TH_C_TOK, half = tid / TH_C_TOK;
  float acc[RR_ROWS / 2];
#pragma unroll
  for (int r = 0; r < RR_ROWS / 2; ++r) acc[r] = 0.f;
  const float4* w = reinterpret_cast<const float4*>(red_w + (int64_t)oc * TH_C_PIX);
#pragma unroll 2
  for (int k = 0; k < TH_C_PIX / 4; ++k) {
    const float4 wk = __ldg(w + k);
#pragma unroll
    for (int r = 0; r < RR_ROWS / 2; ++r) {
      const float4 x = s_x[half * (RR_ROWS / 2) + r][k];
      acc[r] = fmaf(wk.w, x.w, fmaf(wk.z, x.z, fmaf(wk.y, x.y, fmaf(wk.x, x.x, acc[r]))));
    }
  }
  const float bias = __ldg(red_b + oc);
#pragma unroll
  for (int r = 0; r < RR_ROWS / 2; ++r) {
    const int row = row0 + half * (RR_ROWS / 2) + r;
    const float4 tail = s_x[half * (RR_ROWS / 2) + r][TH_C_PIX / 4];  // (n_visible, n, -, -)
    if (row < n_rows) out[(int64_t)row * TH_C_TOK + oc] = (acc[r] + tail.x * bias) / tail.y;
  }
}

// ---------------------------------------------------------------------------
// Pre-mapped maps from the encoder's latents (th_premap_from_latents).  With pixel_feat_map = [up(l_0) | up(l_1) |
// up(l_2) | Wc img + bc] (encoder.py:133-146) and W_pre = [W_0 | W_1 | W_2 | W_3] cut the same way,
//   W_pre pixel_feat_map + b = up(W_0 l_0) + up(W_1 l_1) + up(W_2 l_2) + (W_3 Wc) img + (W_3 bc + b):
// a 1x1 convolution commutes with bilinear upsampling (whose weights sum to 1).  The three products are tcgen05
// GEMMs over the LOW-RESOLUTION latents (86,016 instead of 262,144 rows per 512 x 512 view, K = 64 / 64 / 128 instead
// of 384: 6 % of the MACs of the full-resolution pre-map GEMM); k_premap_combine upsamples and adds them.
// ---------------------------------------------------------------------------
// fold[n] = (W_3 Wc)[n][0..2], (W_3 bc + b)[n]; one block per output channel n, 128 threads = colour channels
__global__ void __launch_bounds__(128) k_premap_fold(const float* __restrict__ w_pre, const float* __restrict__ b_a,
                                                     const float* __restrict__ b_b, const float* __restrict__ wc,
                                                     const float* __restrict__ bc, float4* __restrict__ fold) {
  __shared__ float4 s_part[4];
  const int n = blockIdx.x, c = threadIdx.x;
  const float w = w_pre[(int64_t)n * TH_C_PIX + 256 + c];
  float4 p = make_float4(w * wc[c * 3], w * wc[c * 3 + 1], w * wc[c * 3 + 2], w * bc[c]);
  p.x = warp_sum(p.x), p.y = warp_sum(p.y), p.z = warp_sum(p.z), p.w = warp_sum(p.w);
  if ((c & 31) == 0) s_part[c >> 5] = p;
  __syncthreads();
  if (c == 0) {
    float4 t = s_part[0];
#pragma unroll
    for (int i = 1; i < 4; ++i) t.x += s_part[i].x, t.y += s_part[i].y, t.z += s_part[i].z, t.w += s_part[i].w;
    t.w += n < 256 ? b_a[n] : b_b[n - 256];
    fold[n] = t;
  }
}

// out (H*W, 512) channel-last.  Block = cb_w x CB_H output pixels, 4 warps per output row (one per 128-channel
// group, so that the 2 KB of a pixel are written at about the same time), lane = one float4 of channels.
// Phase 1: the warp brings every low-resolution column its row segment touches into shared memory, vertically
// blended -- all its loads (two coalesced 512-byte rows per column) in flight at once, a column loaded once per row
// segment instead of once per pixel (about 2 loads per output pixel instead of 12).  Phase 2: left to right, per
// level two shared-memory reads and one FMA per channel.  (The vertical-then-horizontal blend differs from PyTorch's
// h0 (w0 a + w1 b) + h1 (w0 c + w1 d) by rounding only.)
struct CombineArgs {
  const float* P[3];  // (lh*lw, 512)
  int lh[3], lw[3];
  int ncol[3];        // columns of shared memory per level and warp (>= the columns a row segment can touch)
  const float* img;   // (3, H, W)
  const float4* fold; // (512)
  float* out;
  int H, W, cb_w;
};
constexpr int CB_H = 2;
__global__ void __launch_bounds__(CB_H * 128) k_premap_combine(const CombineArgs a) {
  extern __shared__ float4 s_cols[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x_base = blockIdx.x * a.cb_w, y = blockIdx.y * CB_H + (warp >> 2), cg = warp & 3;
  if (y >= a.H) return;
  const int ch = cg * 128 + lane * 4;
  const int x_end = min(x_base + a.cb_w, a.W);
  float4* mine = s_cols + (size_t)warp * (a.ncol[0] + a.ncol[1] + a.ncol[2]) * 32 + lane;
  float rw[3];
  int cmin[3], lwm1[3];
  float4* lvl_base[3];
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const int lh = a.lh[l], lw = a.lw[l];
    lwm1[l] = lw - 1;
    const float rh = a.H > 1 ? (float)(lh - 1) / (float)(a.H - 1) : 0.f;
    const float sy = rh * (float)y;
    const int y0 = (int)sy, y1 = y0 + (y0 < lh - 1 ? 1 : 0);
    const float g1 = sy - (float)y0, g0 = 1.f - g1;
    rw[l] = a.W > 1 ? (float)(lw - 1) / (float)(a.W - 1) : 0.f;
    cmin[l] = (int)(rw[l] * (float)x_base);
    const int cl = (int)(rw[l] * (float)(x_end - 1));
    const int cmax = cl + (cl < lw - 1 ? 1 : 0);
    lvl_base[l] = mine + (l == 0 ? 0 : (l == 1 ? a.ncol[0] : a.ncol[0] + a.ncol[1])) * 32;
    const float* r0 = a.P[l] + ((int64_t)y0 * lw + cmin[l]) * 512 + ch;
    const float* r1 = a.P[l] + ((int64_t)y1 * lw + cmin[l]) * 512 + ch;
    const int n = min(cmax - cmin[l] + 1, a.ncol[l]);
#pragma unroll 4
    for (int c = 0; c < n; ++c) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(r0 + (int64_t)c * 512));
      const float4 w = __ldg(reinterpret_cast<const float4*>(r1 + (int64_t)c * 512));
      lvl_base[l][c * 32] = make_float4(g0 * u.x + g1 * w.x, g0 * u.y + g1 * w.y, g0 * u.z + g1 * w.z, g0 * u.w + g1 * w.w);
    }
  }
  float4 f[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) f[j] = __ldg(a.fold + ch + j);
  __syncwarp();
  const int64_t hw = (int64_t)a.H * a.W;
  const float* im = a.img + (int64_t)y * a.W + x_base;
  float4* o = reinterpret_cast<float4*>(a.out + ((int64_t)y * a.W + x_base) * 512 + ch);
  for (int x = x_base; x < x_end; ++x, ++im, o += 128) {
    const float r = __ldg(im), g = __ldg(im + hw), b = __ldg(im + 2 * hw);
    float4 acc;
    acc.x = fmaf(f[0].x, r, fmaf(f[0].y, g, fmaf(f[0].z, b, f[0].w)));
    acc.y = fmaf(f[1].x, r, fmaf(f[1].y, g, fmaf(f[1].z, b, f[1].w)));
    acc.z = fmaf(f[2].x, r, fmaf(f[2].y, g, fmaf(f[2].z, b, f[2].w)));
    acc.w = fmaf(f[3].x, r, fmaf(f[3].y, g, fmaf(f[3].z, b, f[3].w)));
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      const float sxf = rw[l] * (float)x;
      const int c0 = (int)sxf;
      const float w1 = sxf - (float)c0;
      const float4* p = lvl_base[l] + (c0 - cmin[l]) * 32;
      const float4 t0 = p[0];
      const float4 t1 = p[c0 < lwm1[l] ? 32 : 0];
      acc.x += fmaf(w1, t1.x - t0.x, t0.x);
      acc.y += fmaf(w1, t1.y - t0.y, t0.y);
      acc.z += fmaf(w1, t1.z - t0.z, t0.z);
      acc.w += fmaf(w1, t1.w - t0.w, t0.w);
    }
    __stcs(o, acc);
  }
}

// ---------------------------------------------------------------------------
// rays of a pinhole camera (get_rays, if_nerf_data_utils.py:11-30): o = -R^T T; d = ((x, y, 1) Kinv^T - T) R - o
// ---------------------------------------------------------------------------
__global__ void k_gen_rays(int H, int W, const float* __restrict__ Kinv, const float* __restrict__ R,
                           const float* __restrict__ T, float* __restrict__ ray_o, float* __restrict__ ray_d) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= (int64_t)H * W) return;
  const float x = (float)(g % W), y = (float)(g / W);
  float o[3], pc[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[i] = -(R[0 * 3 + i] * T[0] + R[1 * 3 + i] * T[1] + R[2 * 3 + i] * T[2]);
    pc[i] = (x * Kinv[i * 3] + y * Kinv[i * 3 + 1] + Kinv[i * 3 + 2]) - T[i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float pw = pc[0] * R[0 * 3 + i] + pc[1] * R[1 * 3 + i] + pc[2] * R[2 * 3 + i];
    ray_o[g * 3 + i] = o[i];
    ray_d[g * 3 + i] = pw - o[i];
  }
}

// get_near_far (if_nerf_data_utils.py:65-97): the reference evaluates it in float64 on the float32 rays (bounds
// are promoted by `+ np.array([-0.01, 0.01])`), clamps |d| < 1e-5 to 1e-5 IN PLACE (the returned ray_d carries the
// clamp), keeps rays with exactly two face intersections inside the box (+- 1e-6), near / far = the two distances.
__global__ void k_near_far(const float* __restrict__ ray_o, float* __restrict__ ray_d, int64_t n,
                           const float* __restrict__ bounds, float* __restrict__ near_, float* __restrict__ far_,
                           uint8_t* __restrict__ mask) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n) return;
  double bd[2][3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    bd[0][a] = (double)bounds[a] + (-0.01);
    bd[1][a] = (double)bounds[3 + a] + 0.01;
  }
  float df[3];
  double o[3], d[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    df[a] = ray_d[g * 3 + a];
    if (fabsf(df[a]) < 1e-5f) {
      df[a] = 1e-5f;
      ray_d[g * 3 + a] = df[a];
    }
    o[a] = (double)ray_o[g * 3 + a];
    d[a] = (double)df[a];
  }
  const double eps = 1e-6;
  int hits = 0;
  double dist[2] = {0.0, 0.0};
  // numpy evaluates every product and sum separately (no FMA): explicit _rn intrinsics keep nvcc from contracting.
  // face order of `(nominator / ray_d[:, None]).reshape(-1, 6)`: (min x, min y, min z, max x, max y, max z)
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const int side = f / 3, a = f % 3;
    const double t = __ddiv_rn(__dsub_rn(bd[side][a], o[a]), d[a]);
    const double p0 = __dadd_rn(__dmul_rn(t, d[0]), o[0]), p1 = __dadd_rn(__dmul_rn(t, d[1]), o[1]),
                 p2 = __dadd_rn(__dmul_rn(t, d[2]), o[2]);
    const bool in = p0 >= bd[0][0] - eps && p0 <= bd[1][0] + eps && p1 >= bd[0][1] - eps && p1 <= bd[1][1] + eps &&
                    p2 >= bd[0][2] - eps && p2 <= bd[1][2] + eps;
    if (in) {
      if (hits < 2) {
        const double q0 = __dsub_rn(p0, o[0]), q1 = __dsub_rn(p1, o[1]), q2 = __dsub_rn(p2, o[2]);
        dist[hits] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(q0, q0), __dmul_rn(q1, q1)), __dmul_rn(q2, q2)));
      }
      ++hits;
    }
  }
  const bool ok = hits == 2;
  mask[g] = ok ? 1 : 0;
  if (ok) {
    // `np.linalg.norm(ray_d, axis=1)` runs on the float32 rays: a float32 norm (95), promoted only by the division
    const float ndf = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(df[0], df[0]), __fmul_rn(df[1], df[1])), __fmul_rn(df[2], df[2])));
    const double d0 = __ddiv_rn(dist[0], (double)ndf), d1 = __ddiv_rn(dist[1], (double)ndf);
    near_[g] = (float)fmin(d0, d1);
    far_[g] = (float)fmax(d0, d1);
  } else {
    near_[g] = 0.f;
    far_[g] = 0.f;
  }
}

// ordered compaction of the rays that hit the box (`ray_o[mask_at_box]`): block counts -> scan -> scatter
__global__ void __launch_bounds__(256) k_mask_block_counts(const uint8_t* __restrict__ mask, int64_t n,
                                                           int32_t* __restrict__ counts) {
  const int64_t g = blockIdx.x * 256LL + threadIdx.x;
  const int c = __syncthreads_count(g < n && mask[g]);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) k_scan_counts(int32_t* __restrict__ counts, int nblocks, int64_t* __restrict__ total) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? counts[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    const int incl = s[threadIdx.x], c0 = carry;
    if (i < nblocks) counts[i] = c0 + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_compact_rays(const uint8_t* __restrict__ mask, int64_t n,
                                                      const int32_t* __restrict__ offsets, const float* __restrict__ ray_o,
                                                      const float* __restrict__ ray_d, const float* __restrict__ near_,
                                                      const float* __restrict__ far_, float* __restrict__ o_out,
                                                      float* __restrict__ d_out, float* __restrict__ n_out,
                                                      float* __restrict__ f_out) {
  __shared__ int wsum[8];
  const int64_t g = blockIdx.x * 256LL + threadIdx.x;
  const bool hit = g < n && mask[g];
  const unsigned bal = __ballot_sync(0xffffffffu, hit);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wsum[warp] = __popc(bal);
  __syncthreads();
  int before = offsets[blockIdx.x];
  for (int w = 0; w < warp; ++w) before += wsum[w];
  if (hit) {
    const int64_t p = before + __popc(bal & ((1u << lane) - 1));
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      o_out[p * 3 + a] = ray_o[g * 3 + a];
      d_out[p * 3 + a] = ray_d[g * 3 + a];
    }
    n_out[p] = near_[g];
    f_out[p] = far_[g];
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
int launch_paint_group(const float* map, int V, int H, int W, float sx, float sy, const float* verts, const float* cam_R,
                       const float* cam_T, const float* cam_K, const uint8_t* viz, int n_verts, const int32_t* start,
                       const int32_t* members, int n_tok, float* painted, float* out, cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  k_paint_group<<<dim3((unsigned)n_tok, (unsigned)V), TH_C_TOK, 0, st>>>(map, V, H, W, sx, sy, verts, cam_R, cam_T, cam_K,
                                                                        viz, n_verts, start, members, n_tok, painted, out);
  TH_LAUNCHED();
  return TH_OK;
}

size_t paint_latents_scratch_bytes(int V, int n_verts, int n_tok) {
  return align_up((size_t)V * n_verts * TH_C_PIX * sizeof(float), 256) +
         align_up((size_t)V * n_tok * SUM_LD * sizeof(float), 256);
}
int launch_paint_group_latents(const EncTail* enc_dev, int V, const float* red_w, const float* red_b, float sx, float sy,
                               const float* verts, const float* cam_R, const float* cam_T, const float* cam_K,
                               const uint8_t* viz, int n_verts, const int32_t* start, const int32_t* members, int n_tok,
                               void* scratch, float* out, cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  float* painted = static_cast<float*>(scratch);
  float* sums = reinterpret_cast<float*>(static_cast<unsigned char*>(scratch) +
                                         align_up((size_t)V * n_verts * TH_C_PIX * sizeof(float), 256));
  k_paint_latents<<<dim3((unsigned)cdiv(n_verts, PL_WARPS), (unsigned)V), PL_WARPS * 32, 0, st>>>(
      enc_dev, sx, sy, verts, cam_R, cam_T, cam_K, viz, n_verts, painted);
  TH_LAUNCHED();
  k_cluster_sum_latents<<<dim3((unsigned)n_tok, (unsigned)V), TH_C_PIX / 4, 0, st>>>(painted, viz, n_verts, start,
                                                                                   members, n_tok, sums);
  TH_LAUNCHED();
  const int n_rows = V * n_tok;
  k_reduce_rows_latents<<<(unsigned)cdiv(n_rows, RR_ROWS), 2 * TH_C_TOK, 0, st>>>(sums, n_rows, red_w, red_b, out);
  TH_LAUNCHED();
  return TH_OK;
}

static size_t premap_latents_p_floats(const int* lh, const int* lw, size_t* off) {
  size_t n = 0;
  for (int i = 0; i < 3; ++i) {
    if (off) off[i] = n;
    n += (size_t)lh[i] * lw[i] * 512;
  }
  return n;
}
size_t premap_latents_scratch_bytes(const int* lh, const int* lw) {
  return align_up(premap_latents_p_floats(lh, lw, nullptr) * sizeof(float), 256) + 512 * sizeof(float4);
}

int launch_premap_latents(const EncTail* enc_host, const unsigned char* weights, const PackedHeader& hdr, float* dst,
                          int n_views, void* scratch, cudaStream_t st) {
  const EncTail& e0 = enc_host[0];
  size_t off[3];
  const size_t pf = premap_latents_p_floats(e0.lh, e0.lw, off);
  float* P = static_cast<float*>(scratch);
  float4* fold = reinterpret_cast<float4*>(static_cast<unsigned char*>(scratch) + align_up(pf * sizeof(float), 256));
  {
    ProfScope prof_(PROF_PREMAP, st);
    k_premap_fold<<<512, 128, 0, st>>>(reinterpret_cast<const float*>(weights + hdr.pre_w),
                                       reinterpret_cast<const float*>(weights + hdr.ar0_b),
                                       reinterpret_cast<const float*>(weights + hdr.preb_b), e0.wc, e0.bc, fold);
    TH_LAUNCHED();
  }
  static const int kb0[3] = {0, 1, 2}, kch[3] = {64, 64, 128};
  for (int v = 0; v < n_views; ++v) {
    const EncTail& e = enc_host[v];
    for (int lvl = 0; lvl < 3; ++lvl)
      for (int half = 0; half < 2; ++half) {
        GemmArgs g{};
        g.nseg = 1;
        g.seg[0].ptr = e.lat[lvl];  // (lh*lw, C) rows: the latents are channel-last
        g.seg[0].K = kch[lvl];
        g.seg[0].ld = kch[lvl];
        g.bias = nullptr;  // both biases travel through `fold`
        g.C = P + off[lvl] + half * 256;
        g.ldc = 512;
        g.M = (int64_t)e.lh[lvl] * e.lw[lvl];
        g.N = 256;
        g.relu = 0;
        const uint64_t img = half ? hdr.h_preb : hdr.h_ar0;
        g.acc_scale = img_inv_scale_ptr(weights, hdr, img);
        // k-blocks kb0.. of the (256, 384) weight image: 2 planes x 256 rows x 128 bytes each
        int rc = launch_gemm_tc(g, weights + img + (size_t)kb0[lvl] * (2 * 256 * 128), st, PROF_PREMAP);
        if (rc) return rc;
      }
    ProfScope prof_(PROF_PREMAP, st);
    CombineArgs a;
    for (int i = 0; i < 3; ++i) a.P[i] = P + off[i], a.lh[i] = e.lh[i], a.lw[i] = e.lw[i];
    a.img = e.img, a.fold = fold, a.out = dst + (int64_t)v * e.H * e.W * 512, a.H = e.H, a.W = e.W;
    // widest row segment whose columns fit the shared memory of a block (x-ratio <= 1 per level: at most
    // ceil(cb_w * ratio) + 2 columns)
    size_t smem = 0;
    const size_t smem_cap = 96 * 1024;  // two blocks per SM (measured: 40 / 56 / 96 / 170 KB -> 1.53 / 1.42 / 1.39 / 2.07 ms)
    for (a.cb_w = 32; a.cb_w >= 1; a.cb_w >>= 1) {
      int cols = 0;
      for (int i = 0; i < 3; ++i) {
        const double ratio = e.W > 1 ? (double)(e.lw[i] - 1) / (double)(e.W - 1) : 0.0;
        int n = (int)(a.cb_w * ratio) + 3;
        a.ncol[i] = n < e.lw[i] ? n : e.lw[i];
        cols += a.ncol[i];
      }
      smem = (size_t)cols * 512 * (CB_H * 4);
      if (smem <= smem_cap) break;
    }
    if (a.cb_w < 1) {
      set_error("premap_from_latents: latents wider than the image are not supported");
      return TH_EINVAL;
    }
    TH_CUDA(cudaFuncSetAttribute(k_premap_combine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_premap_combine<<<dim3((unsigned)cdiv(e.W, a.cb_w), (unsigned)cdiv(e.H, CB_H)), CB_H * 128, smem, st>>>(a);
    TH_LAUNCHED();
  }
  return TH_OK;
}

int launch_group_mean(const void* x, int is_f64, int C, const int32_t* start, const int32_t* members, int n_tok, int outer,
                      void* out, cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  const unsigned grid = (unsigned)cdiv((int64_t)n_tok * C, 128);
  if (is_f64)
    k_group_mean<double><<<grid, 128, 0, st>>>(static_cast<const double*>(x), C, start, members, n_tok, outer,
                                               static_cast<double*>(out));
  else
    k_group_mean<float><<<grid, 128, 0, st>>>(static_cast<const float*>(x), C, start, members, n_tok, outer,
                                              static_cast<float*>(out));
  TH_LAUNCHED();
  return TH_OK;
}

int launch_near_far(const float* ray_o, float* ray_d, int64_t n, const float* bounds, float* near_, float* far_,
                    uint8_t* mask, cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  if (n <= 0) return TH_OK;
  k_near_far<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(ray_o, ray_d, n, bounds, near_, far_, mask);
  TH_LAUNCHED();
  return TH_OK;
}

size_t generate_rays_workspace_bytes(int64_t n) { return align_up((size_t)(cdiv(n, 256) + 1) * 4, 256) + 256; }

int launch_generate_rays(int H, int W, const float* Kinv, const float* R, const float* T, const float* bounds,
                         float* ray_o, float* ray_d, float* near_, float* far_, uint8_t* mask, float* o_c, float* d_c,
                         float* n_c, float* f_c, int64_t* count, void* workspace, cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  const int64_t n = (int64_t)H * W;
  k_gen_rays<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(H, W, Kinv, R, T, ray_o, ray_d);
  TH_LAUNCHED();
  if (!bounds) return TH_OK;
  k_near_far<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(ray_o, ray_d, n, bounds, near_, far_, mask);
  TH_LAUNCHED();
  if (!o_c) return TH_OK;
  int32_t* counts = static_cast<int32_t*>(workspace);
  const int nblocks = (int)cdiv(n, 256);
  k_mask_block_counts<<<nblocks, 256, 0, st>>>(mask, n, counts);
  TH_LAUNCHED();
  k_scan_counts<<<1, 1024, 0, st>>>(counts, nblocks, count);
  TH_LAUNCHED();
  k_compact_rays<<<nblocks, 256, 0, st>>>(mask, n, counts, ray_o, ray_d, near_, far_, o_c, d_c, n_c, f_c);
  TH_LAUNCHED();
  return TH_OK;
}

}  // namespace th
