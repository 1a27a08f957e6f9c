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
      const float wx = __fsub_rn(ix, x0), wy = __fsub_rn(iy, y0);
      const float ex = __fsub_rn(1.0f, wx), ey = __fsub_rn(1.0f, wy);
      const int x0i = (int)x0, y0i = (int)y0;
      const int x1i = min(x0i + 1, W - 1), y1i = min(y0i + 1, H - 1);
      const float a = __ldg(plane + y0i * W + x0i), bq = __ldg(plane + y0i * W + x1i);
      const float cq = __ldg(plane + y1i * W + x0i), d = __ldg(plane + y1i * W + x1i);
      val = __fmaf_rn(d, __fmul_rn(wy, wx),
                      __fmaf_rn(cq, __fmul_rn(wy, ex), __fmaf_rn(bq, __fmul_rn(ey, wx), __fmul_rn(a, __fmul_rn(ey, ex)))));
    }
    if (painted) painted[((int64_t)v * n_verts + vi) * TH_C_TOK + ch] = val;
    acc.add(val);
  }
  out[((int64_t)v * n_tok + c) * TH_C_TOK + ch] = acc.total() / (float)n;
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
