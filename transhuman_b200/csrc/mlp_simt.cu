// Per-point network (rows a9 + a10): the fp32 CUDA-core GEMM, the cross-view
// attention mix, the alpha / rgb heads and the layer schedule shared with the
// tensor-core GEMM (mlp_tc.cu).
//
// Schedule for P points, V views (rows of per-view activations are view-major,
// r = v*P + p).  Reference: cross_transformer.py:128-149, 273-353.
//   S   = relu(rep  fc_0^T)                         (V*P,256) <- K 256
//   X   = relu(pix  alpha_res_0^T)                  (V*P,256) <- K 384
//   KP  = X  key_embed_0^T ;  KS = S key_embed_1^T  (V*P,128) <- K 256
//   A[i,j] = softmax_i(KP_i . KS_j / sqrt(128)) ;  XT_j = sum_i A[i,j] X_i
//   NET = [S | XT] [value_embed_1 | value_embed_0]^T            <- K 512
//         (= query_value + value_embed(x) @ A because sum_i A[i,j] = 1)
//   N1  = relu(NET fc_1^T) ; INTER = relu(N1 fc_2^T)
//   O   = relu([INTER_0|..|INTER_{V-1}] [fc_3/V ...]^T)   (P,256) <- K 256 V
//   alpha = O . alpha_fc
//   F   = [INTER | pix] [feature_fc | rgb_res_0]^T              <- K 640
//   G   = relu([F | viewdir] view_fc^T)                  (V*P,128) <- K 288
//   T   = relu([G_0|..|G_{V-1}| mean_v pix] [fc_4/V ...| fc_4 rgb_res_1]^T)
//   rgb = T rgb_fc^T
#include <cuda_fp16.h>

#include <initializer_list>

#include "kernels.cuh"

namespace th {

// ---------------------------------------------------------------------------
// fp32 GEMM, 128x128x16 tiles, 256 threads, 8x8 micro-tiles, register
// prefetch + double-buffered shared memory.  A is given as up to 4 K-segments
// (concatenation along K without materialising it).
// ---------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 128, GBK = 16, GLD = GBM + 4;

__global__ void __launch_bounds__(256, 2) k_gemm_simt(GemmArgs a) {
  __shared__ __align__(16) float As[2][GBK][GLD];
  __shared__ __align__(16) float Bs[2][GBK][GLD];
  const int tid = threadIdx.x;
  const int64_t m0 = blockIdx.x * (int64_t)GBM;
  const int n0 = blockIdx.y * GBN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader coordinates: rows lr, lr+64; k offset lk
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  int total_tiles = 0;
  for (int s = 0; s < a.nseg; ++s) total_tiles += a.seg[s].K / GBK;

  float4 ra[2], rb[2];
  int seg = 0, kin = 0, kw = 0;  // current segment, k offset inside it, k offset in W
  auto fetch = [&]() {
    const GemmSeg& sg = a.seg[seg];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t m = m0 + lr + 64 * h;
      if (m < a.M) {
        int64_t row = sg.row_mod ? m % sg.row_mod : m;
        ra[h] = *reinterpret_cast<const float4*>(sg.ptr + row * sg.ld + kin + lk);
      } else {
        ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      int n = n0 + lr + 64 * h;
      rb[h] = *reinterpret_cast<const float4*>(a.W + (int64_t)n * a.ldw + kw + lk);
    }
    kin += GBK;
    kw += GBK;
    if (kin >= sg.K) {
      kin = 0;
      ++seg;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lr + 64 * h;
      As[buf][lk + 0][r] = ra[h].x;
      As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z;
      As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x;
      Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z;
      Bs[buf][lk + 3][r] = rb[h].w;
    }
  };
  fetch();
  stash(0);
  __syncthreads();
  for (int t = 0; t < total_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < total_tiles) fetch();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < total_tiles) stash(buf ^ 1);
    __syncthreads();
  }
  // epilogue: bias, optional ReLU, float4 stores
  float bias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias[j] = a.bias ? a.bias[n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4)] : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= a.M) continue;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = acc[i][j] + bias[j];
      if (a.relu) v[j] = fmaxf(v[j], 0.f);
    }
    float* crow = a.C + m * a.ldc + n0;
    *reinterpret_cast<float4*>(crow + tx * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(crow + 64 + tx * 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

int launch_gemm_simt(const GemmArgs& a, cudaStream_t st) {
  ProfScope prof_(PROF_GEMM, st);
  if (a.M <= 0) return TH_OK;
  if (a.N % GBN != 0) {
    set_error("gemm_simt: N=%d not a multiple of %d", a.N, GBN);
    return TH_EINVAL;
  }
  for (int s = 0; s < a.nseg; ++s)
    if (a.seg[s].K % GBK != 0 || a.seg[s].ld % 4 != 0) {
      set_error("gemm_simt: segment %d K=%d ld=%d unsupported", s, a.seg[s].K, a.seg[s].ld);
      return TH_EINVAL;
    }
  dim3 grid((unsigned)cdiv(a.M, GBM), a.N / GBN);
  k_gemm_simt<<<grid, 256, 0, st>>>(a);
  TH_LAUNCHED();
  return TH_OK;
}

// ---------------------------------------------------------------------------
// cross-view attention mix (cross_transformer.py:128-149): one warp per point.
// ---------------------------------------------------------------------------
// IMG: X is read from, and XT written to, fp16 hi/lo tile images (tensor-core path).
template <bool IMG>
__global__ void __launch_bounds__(256) k_attn_mix(const float* __restrict__ kp, const float* __restrict__ ks,
                                                  const float* __restrict__ x, float* __restrict__ xt, int64_t P,
                                                  int64_t Pp, int V) {
  const int lane = threadIdx.x & 31;
  const int64_t p = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (p >= P) return;
  float4 kpv[TH_MAX_VIEWS], ksv[TH_MAX_VIEWS];
#pragma unroll
  for (int v = 0; v < TH_MAX_VIEWS; ++v)
    if (v < V) {
      kpv[v] = *reinterpret_cast<const float4*>(kp + (v * Pp + p) * 128 + lane * 4);
      ksv[v] = *reinterpret_cast<const float4*>(ks + (v * Pp + p) * 128 + lane * 4);
    }
  float A[TH_MAX_VIEWS][TH_MAX_VIEWS];
#pragma unroll
  for (int i = 0; i < TH_MAX_VIEWS; ++i)
#pragma unroll
    for (int j = 0; j < TH_MAX_VIEWS; ++j)
      if (i < V && j < V) {
        float d = kpv[i].x * ksv[j].x + kpv[i].y * ksv[j].y + kpv[i].z * ksv[j].z + kpv[i].w * ksv[j].w;
        A[i][j] = __fdiv_rn(warp_sum(d), 11.313708498984761f);  // / sqrt(128)
      }
#pragma unroll
  for (int j = 0; j < TH_MAX_VIEWS; ++j)
    if (j < V) {  // softmax over i (dim=1 of (P, V_i, V_j))
      float m = -3.4e38f, sum = 0.f;
#pragma unroll
      for (int i = 0; i < TH_MAX_VIEWS; ++i)
        if (i < V) m = fmaxf(m, A[i][j]);
#pragma unroll
      for (int i = 0; i < TH_MAX_VIEWS; ++i)
        if (i < V) {
          A[i][j] = expf(A[i][j] - m);
          sum += A[i][j];
        }
#pragma unroll
      for (int i = 0; i < TH_MAX_VIEWS; ++i)
        if (i < V) A[i][j] = __fdiv_rn(A[i][j], sum);
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = (lane + 32 * h) * 4;
    float4 xv[TH_MAX_VIEWS];
#pragma unroll
    for (int i = 0; i < TH_MAX_VIEWS; ++i)
      if (i < V) {
        if (IMG) {
          const unsigned char* q = reinterpret_cast<const unsigned char*>(x) + img_offset(i * Pp + p, c, 256);
          const uint2 h = *reinterpret_cast<const uint2*>(q), l = *reinterpret_cast<const uint2*>(q + 16384);
          const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)),
                       h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y)),
                       l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)),
                       l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
          xv[i] = make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
        } else {
          xv[i] = *reinterpret_cast<const float4*>(x + (i * Pp + p) * 256 + c);
        }
      }
#pragma unroll
    for (int j = 0; j < TH_MAX_VIEWS; ++j)
      if (j < V) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < TH_MAX_VIEWS; ++i)
          if (i < V) {
            o.x = fmaf(A[i][j], xv[i].x, o.x);
            o.y = fmaf(A[i][j], xv[i].y, o.y);
            o.z = fmaf(A[i][j], xv[i].z, o.z);
            o.w = fmaf(A[i][j], xv[i].w, o.w);
          }
        if (IMG) {
          uint32_t ha, la, hb, lb;
          split_hl2(o.x, o.y, ha, la);
          split_hl2(o.z, o.w, hb, lb);
          unsigned char* q = reinterpret_cast<unsigned char*>(xt) + img_offset(j * Pp + p, c, 256);
          *reinterpret_cast<uint2*>(q) = make_uint2(ha, hb);
          *reinterpret_cast<uint2*>(q + 16384) = make_uint2(la, lb);
        } else {
          *reinterpret_cast<float4*>(xt + (j * Pp + p) * 256 + c) = o;
        }
      }
  }
}

// ---------------------------------------------------------------------------
// heads: alpha = O . alpha_fc + b ; rgb = T rgb_fc^T + b.  One warp per point.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_alpha_head(const float* __restrict__ o, const float* __restrict__ w,
                                                    const float* __restrict__ b, float* __restrict__ alpha,
                                                    int64_t P) {
  const int lane = threadIdx.x & 31;
  const int64_t p = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (p >= P) return;
  float acc = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int c = (lane + 32 * h) * 4;
    float4 a = *reinterpret_cast<const float4*>(o + p * 256 + c);
    float4 ww = *reinterpret_cast<const float4*>(w + c);
    acc += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) alpha[p] = acc + b[0];
}

__global__ void __launch_bounds__(256) k_write_alpha(const float* __restrict__ alpha, const int32_t* __restrict__ ids,
                                                    int64_t first, int64_t P, float* __restrict__ alpha_out,
                                                    float* __restrict__ raw) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= P) return;
  int64_t dst = ids ? (int64_t)ids[first + p] : first + p;
  if (alpha_out) alpha_out[dst] = alpha[p];
  if (raw) reinterpret_cast<float4*>(raw)[dst] = make_float4(0.f, 0.f, 0.f, alpha[p]);
}

__global__ void __launch_bounds__(256) k_rgb_head(const float* __restrict__ t, const float* __restrict__ w,
                                                  const float* __restrict__ b, const float* __restrict__ alpha,
                                                  const int32_t* __restrict__ ids, int64_t first, int64_t P,
                                                  int zero_if_transparent, float* __restrict__ raw) {
  const int lane = threadIdx.x & 31;
  const int64_t p = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (p >= P) return;
  float4 a = *reinterpret_cast<const float4*>(t + p * 128 + lane * 4);
  float out[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float4 ww = *reinterpret_cast<const float4*>(w + k * 128 + lane * 4);
    out[k] = warp_sum(a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w) + b[k];
  }
  if (lane == 0) {
    float al = alpha[p];
    if (zero_if_transparent && !(al > 0.f)) out[0] = out[1] = out[2] = 0.f;
    int64_t dst = ids ? (int64_t)ids[first + p] : first + p;
    reinterpret_cast<float4*>(raw)[dst] = make_float4(out[0], out[1], out[2], al);
  }
}

// ---------------------------------------------------------------------------
// staged a9/a10 input packing: reference layouts (V,255,P), (V,384,P), (P,27)
// -> GEMM layout rows.
// ---------------------------------------------------------------------------
__global__ void k_pack_cmajor(const float* __restrict__ src, int C, int64_t P, int64_t Pp, float* __restrict__ dst,
                              int ld) {
  // src (V, C, P) -> dst (V*Pp, ld), zero padded; 32x32 tile transpose
  __shared__ float tile[32][33];
  const int64_t v = blockIdx.z;
  const int64_t p0 = blockIdx.x * 32LL;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    int64_t p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < P) ? src[(v * C + c) * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t p = p0 + i;
    int c = c0 + threadIdx.x;
    if (p < P && c < ld) dst[(v * Pp + p) * ld + c] = tile[threadIdx.x][i];
  }
}

__global__ void k_pack_misc(const float* __restrict__ viewdir, const float* __restrict__ pix, int64_t P, int64_t Pp,
                            int V, float* __restrict__ vd, float* __restrict__ pix_mean) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < P * VD_LD) {
    int64_t p = i / VD_LD;
    int c = (int)(i % VD_LD);
    vd[i] = c < TH_C_VIEW ? viewdir[p * TH_C_VIEW + c] : 0.f;
  }
  if (i < P * PIX_LD) {
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += pix[(int64_t)v * Pp * PIX_LD + i];
    pix_mean[i] = __fdiv_rn(s, (float)V);
  }
}

int launch_pack_inputs(const float* human_rep, const float* pixel_feat, const float* viewdir, int64_t P, int V,
                       const MlpBuffers& b, cudaStream_t st) {
  if (P <= 0) return TH_OK;
  dim3 block(32, 8);
  const int64_t Pp = pad_points(P);
  k_pack_cmajor<<<dim3((unsigned)cdiv(P, 32), REP_LD / 32, V), block, 0, st>>>(human_rep, TH_C_REP, P, Pp, b.rep,
                                                                             REP_LD);
  TH_LAUNCHED();
  k_pack_cmajor<<<dim3((unsigned)cdiv(P, 32), PIX_LD / 32, V), block, 0, st>>>(pixel_feat, TH_C_PIX, P, Pp, b.pix,
                                                                             PIX_LD);
  TH_LAUNCHED();
  k_pack_misc<<<(unsigned)cdiv(P * PIX_LD, 256), 256, 0, st>>>(viewdir, b.pix, P, Pp, V, b.vd, b.pix_mean);
  TH_LAUNCHED();
  return TH_OK;
}

// ---------------------------------------------------------------------------
// buffers + schedule
// ---------------------------------------------------------------------------
size_t mlp_buffer_floats_per_point(int V) {
  // rep, pix, s, x, xt, net: per view; kp, ks: per view; per point: pix_mean, vd, o, t, alpha
  return (size_t)V * (REP_LD + PIX_LD + 256 * 4 + 128 * 2) + PIX_LD + 2 * VD_LD + 256 + 128 + 4;
}

// All buffers are sized and strided for Pp = pad_points(P) rows per view.
// chain_scratch > 0 = COMPACT carve for the layer-chained schedule on pre-mapped maps (the default path): that
// schedule reads rep, the pix block [X | P2], R (128 wide, in pix_mean) and vd, and keeps everything between its
// layers in a per-CTA scratch of a FIXED size (chain_scratch_bytes) -- none of the per-point activation buffers of
// the layer-at-a-time schedule (s .. ks, 15 KB per point) exist.  8.25 KB per point + the scratch.
size_t mlp_compact_bytes(int64_t P, int V, size_t chain_scratch) {
  return (size_t)pad_points(P) * ((size_t)V * (REP_LD + PIX_LD) + 128 + 2 * VD_LD) * 4 + ((chain_scratch + 1023) & ~size_t(1023));
}
void mlp_carve(float* base, int64_t P, int V, MlpBuffers* b, size_t chain_scratch) {
  const int64_t Pp = pad_points(P);
  // tile images need 1 KB alignment; every buffer size below is a multiple of 1 KB because Pp % 256 == 0
  float* p = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(base) + 1023) & ~uintptr_t(1023));
  auto take = [&](size_t per_point) {
    float* r = p;
    p += (size_t)Pp * per_point;
    return r;
  };
  b->rep = take((size_t)V * REP_LD);
  b->pix = take((size_t)V * PIX_LD);
  if (chain_scratch) {
    b->s = p;  // the chain kernel's scratch (1 KB aligned: everything before it is a multiple of 1 KB)
    p += ((chain_scratch + 1023) & ~size_t(1023)) / 4;
    b->x = b->xt = b->net = b->kp = b->ks = nullptr;
    b->pix_mean = take(128);
    b->vd = take(2 * VD_LD);
    return;
  }
  b->s = take((size_t)V * 256);
  b->x = take((size_t)V * 256);
  b->xt = take((size_t)V * 256);
  b->net = take((size_t)V * 256);
  b->kp = take((size_t)V * 128);
  b->ks = take((size_t)V * 128);
  b->pix_mean = take(PIX_LD);
  b->vd = take(2 * VD_LD);  // 32 fp32 channels, or a 64-wide fp16 hi/lo image (same 256 B / point)
  // o (256), t (128), alpha (4, padded) follow: see mlp_forward
}

static const float* wf(const MlpRun& run, uint64_t off) { return reinterpret_cast<const float*>(run.weights + off); }

static GemmSeg seg_f32(const float* ptr, int ld, int K, int64_t row_mod = 0) {
  GemmSeg g{};
  g.ptr = ptr;
  g.ld = ld;
  g.K = K;
  g.row_mod = row_mod;
  return g;
}
static GemmSeg seg_img(const float* buf, int K) {
  GemmSeg g{};
  g.K = K;
  g.img = reinterpret_cast<const unsigned char*>(buf);
  return g;
}

// Heads shared by both schedules.
static int run_heads(const MlpRun& run, const float* o, const float* t, float* alpha, const PackedHeader& h,
                     bool rgb, cudaStream_t st) {
  const int64_t P = run.P;
  if (!rgb) {
    {
      ProfScope prof_(PROF_POINTWISE, st);
      k_alpha_head<<<(unsigned)cdiv(P, 8), 256, 0, st>>>(o, wf(run, h.afc_w), wf(run, h.afc_b), alpha, P);
    }
    TH_LAUNCHED();
    if (run.alpha_only) {
      ProfScope prof_(PROF_POINTWISE, st);
      k_write_alpha<<<(unsigned)cdiv(P, 256), 256, 0, st>>>(alpha, run.dst_ids, run.first, P, run.alpha_out, run.raw);
      TH_LAUNCHED();
    }
    return TH_OK;
  }
  {
    ProfScope prof_(PROF_POINTWISE, st);
    k_rgb_head<<<(unsigned)cdiv(P, 8), 256, 0, st>>>(t, wf(run, h.rgb_w), wf(run, h.rgb_b), alpha, run.dst_ids,
                                                     run.first, P, run.zero_rgb_if_transparent, run.raw);
  }
  TH_LAUNCHED();
  return TH_OK;
}

// fp32 schedule (CUDA-core GEMM): every activation is an fp32 row-major buffer.
static int mlp_forward_simt(const MlpRun& run, const MlpBuffers& b, const PackedHeader& h, cudaStream_t st) {
  const int64_t P = run.P, Pp = pad_points(P);
  const int V = run.V;
  const int64_t R = (int64_t)V * Pp;
  float* o = b.vd + (size_t)Pp * 2 * VD_LD;
  float* t = o + (size_t)Pp * 256;
  float* alpha = t + (size_t)Pp * 128;
  auto gemm = [&](std::initializer_list<GemmSeg> segs, const float* W, int ldw, const float* bias, float* C, int N,
                  int64_t M, int relu) -> int {
    GemmArgs g{};
    for (const GemmSeg& sg : segs) g.seg[g.nseg++] = sg;
    g.W = W;
    g.ldw = ldw;
    g.bias = bias;
    g.C = C;
    g.ldc = N;
    g.M = M;
    g.N = N;
    g.relu = relu;
    return launch_gemm_simt(g, st);
  };
  int rc;
  if ((rc = gemm({seg_f32(b.rep, REP_LD, 256)}, wf(run, h.fc0_w), 256, wf(run, h.fc0_b), b.s, 256, R, 1))) return rc;
  if ((rc = gemm({seg_f32(b.pix, PIX_LD, 384)}, wf(run, h.ar0_w), 384, wf(run, h.ar0_b), b.x, 256, R, 1))) return rc;
  if ((rc = gemm({seg_f32(b.x, 256, 256)}, wf(run, h.k0_w), 256, wf(run, h.k0_b), b.kp, 128, R, 0))) return rc;
  if ((rc = gemm({seg_f32(b.s, 256, 256)}, wf(run, h.k1_w), 256, wf(run, h.k1_b), b.ks, 128, R, 0))) return rc;
  {
    ProfScope prof_(PROF_POINTWISE, st);
    k_attn_mix<false><<<(unsigned)cdiv(P, 8), 256, 0, st>>>(b.kp, b.ks, b.x, b.xt, P, Pp, V);
  }
  TH_LAUNCHED();
  if ((rc = gemm({seg_f32(b.s, 256, 256), seg_f32(b.xt, 256, 256)}, wf(run, h.v_w), 512, wf(run, h.v_b), b.net, 256, R,
                 0)))
    return rc;
  float* n1 = b.x;      // X is dead after the mix
  float* inter = b.xt;  // XT is dead after NET
  if ((rc = gemm({seg_f32(b.net, 256, 256)}, wf(run, h.fc1_w), 256, wf(run, h.fc1_b), n1, 256, R, 1))) return rc;
  if ((rc = gemm({seg_f32(n1, 256, 256)}, wf(run, h.fc2_w), 256, wf(run, h.fc2_b), inter, 256, R, 1))) return rc;
  {  // O = relu(mean_v(INTER) fc_3^T): mean folded into K
    GemmArgs g{};
    for (int v = 0; v < V; ++v) g.seg[v] = seg_f32(inter + (size_t)v * Pp * 256, 256, 256);
    g.nseg = V;
    g.W = wf(run, h.fc3m_w);
    g.ldw = 256 * V;
    g.bias = wf(run, h.fc3m_b);
    g.C = o;
    g.ldc = 256;
    g.M = Pp;
    g.N = 256;
    g.relu = 1;
    if ((rc = launch_gemm_simt(g, st))) return rc;
  }
  if ((rc = run_heads(run, o, nullptr, alpha, h, false, st))) return rc;
  if (run.alpha_only) return TH_OK;
  float* f = b.s;      // S is dead after NET
  float* gbuf = b.kp;  // keys are dead after the mix
  if ((rc = gemm({seg_f32(inter, 256, 256), seg_f32(b.pix, PIX_LD, 384)}, wf(run, h.f_w), 640, wf(run, h.f_b), f, 256,
                 R, 0)))
    return rc;
  if ((rc = gemm({seg_f32(f, 256, 256), seg_f32(b.vd, VD_LD, VD_LD, Pp)}, wf(run, h.view_w), 320, wf(run, h.view_b),
                 gbuf, 128, R, 1)))
    return rc;
  {  // T = relu([G_0|..|G_{V-1}| mean pix] W_t^T)
    GemmArgs g{};
    for (int v = 0; v < V; ++v) g.seg[v] = seg_f32(gbuf + (size_t)v * Pp * 128, 128, 128);
    g.seg[V] = seg_f32(b.pix_mean, PIX_LD, 384);
    g.nseg = V + 1;
    g.W = wf(run, h.t_w);
    g.ldw = 128 * V + 384;
    g.bias = wf(run, h.t_b);
    g.C = t;
    g.ldc = 128;
    g.M = Pp;
    g.N = 128;
    g.relu = 1;
    if ((rc = launch_gemm_simt(g, st))) return rc;
  }
  return run_heads(run, o, t, alpha, h, true, st);
}

// Tensor-core schedule: the GEMM -> GEMM activations (S, NET, N1, INTER, F, G) are
// written by the producing epilogue as fp16 hi/lo tile images and consumed by the
// next layer with one 32 KB bulk copy per k-block; inputs produced by other
// kernels (rep, pix, X for the attention, XT, viewdir, mean pix) and the outputs
// read by the heads / attention (X, KP, KS, O, T) stay fp32 rows.
static int mlp_forward_tc(const MlpRun& run, const MlpBuffers& b, const PackedHeader& h, cudaStream_t st) {
  const int64_t P = run.P, Pp = pad_points(P);
  const int V = run.V;
  const int64_t R = (int64_t)V * Pp;
  float* o = b.vd + (size_t)Pp * 2 * VD_LD;
  float* t = o + (size_t)Pp * 256;
  float* alpha = t + (size_t)Pp * 128;
  auto gemm = [&](std::initializer_list<GemmSeg> segs, uint64_t w_img, const float* bias, float* C, bool c_is_img, int N,
                  int64_t M, int relu) -> int {
    GemmArgs g{};
    for (const GemmSeg& sg : segs) g.seg[g.nseg++] = sg;
    g.bias = bias;
    if (c_is_img)
      g.C_img = reinterpret_cast<unsigned char*>(C);
    else
      g.C = C;
    g.ldc = N;
    g.M = M;
    g.N = N;
    g.relu = relu;
    g.acc_scale = img_inv_scale_ptr(run.weights, h, w_img);
    return launch_gemm_tc(g, run.weights + w_img, st);
  };
  // image of a (V*Pp, C) activation: the slice of view v starts (v*Pp/128) row tiles in
  auto view_img = [&](const float* buf, int C, int v) {
    return seg_img(buf + (size_t)v * Pp * C, C);
  };
  const bool in_img = run.inputs_are_images != 0;
  auto in_seg = [&](const float* buf, int ld, int K) { return in_img ? seg_img(buf, K) : seg_f32(buf, ld, K); };
  int rc;
  if ((rc = gemm({in_seg(b.rep, REP_LD, 256)}, h.h_fc0, wf(run, h.fc0_b), b.s, true, 256, R, 1))) return rc;
  if ((rc = gemm({in_seg(b.pix, PIX_LD, 384)}, h.h_ar0, wf(run, h.ar0_b), b.x, true, 256, R, 1))) return rc;
  if ((rc = gemm({seg_img(b.x, 256)}, h.h_k0, wf(run, h.k0_b), b.kp, false, 128, R, 0))) return rc;
  if ((rc = gemm({seg_img(b.s, 256)}, h.h_k1, wf(run, h.k1_b), b.ks, false, 128, R, 0))) return rc;
  {
    ProfScope prof_(PROF_POINTWISE, st);
    k_attn_mix<true><<<(unsigned)cdiv(P, 8), 256, 0, st>>>(b.kp, b.ks, b.x, b.xt, P, Pp, V);
  }
  TH_LAUNCHED();
  // value embeds folded into fc_1 (no non-linearity between them): N1 = relu([S | XT] W_fc1f^T)
  float* n1 = b.x;      // X is dead after the mix
  float* inter = b.xt;  // XT is dead after N1
  if ((rc = gemm({seg_img(b.s, 256), seg_img(b.xt, 256)}, h.h_fc1f, wf(run, h.fc1f_b), n1, true, 256, R, 1))) return rc;
  if ((rc = gemm({seg_img(n1, 256)}, h.h_fc2, wf(run, h.fc2_b), inter, true, 256, R, 1))) return rc;
  {
    GemmArgs g{};
    for (int v = 0; v < V; ++v) g.seg[v] = view_img(inter, 256, v);
    g.nseg = V;
    g.bias = wf(run, h.fc3m_b);
    g.C = o;
    g.ldc = 256;
    g.M = Pp;
    g.N = 256;
    g.relu = 1;
    g.acc_scale = img_inv_scale_ptr(run.weights, h, h.h_fc3m);
    if ((rc = launch_gemm_tc(g, run.weights + h.h_fc3m, st))) return rc;
  }
  if ((rc = run_heads(run, o, nullptr, alpha, h, false, st))) return rc;
  if (run.alpha_only) return TH_OK;
  float* gbuf = b.kp;  // keys are dead after the mix
  {
    // feature_fc + rgb_res_0 folded into view_fc: G = relu([INTER | pix | viewdir] W_gvf^T)
    GemmSeg vd = seg_f32(b.vd, VD_LD, VD_LD, Pp);
    if (in_img) {
      vd = seg_img(b.vd, 64);
      vd.img_tile_mod = Pp / 128;
    }
    if ((rc = gemm({seg_img(inter, 256), in_seg(b.pix, PIX_LD, 384), vd}, h.h_gvf, wf(run, h.gvf_b), gbuf, true, 128, R,
                   1)))
      return rc;
  }
  {
    GemmArgs g{};
    for (int v = 0; v < V; ++v) g.seg[v] = view_img(gbuf, 128, v);
    g.seg[V] = in_seg(b.pix_mean, PIX_LD, 384);
    g.nseg = V + 1;
    g.bias = wf(run, h.t_b);
    g.C = t;
    g.ldc = 128;
    g.M = Pp;
    g.N = 128;
    g.relu = 1;
    g.acc_scale = img_inv_scale_ptr(run.weights, h, h.h_t);
    if ((rc = launch_gemm_tc(g, run.weights + h.h_t, st))) return rc;
  }
  return run_heads(run, o, t, alpha, h, true, st);
}

int mlp_forward(const MlpRun& run, const MlpBuffers& b, const PackedHeader& h, cudaStream_t st) {
  if (run.P <= 0) return TH_OK;
  return run.use_tensor_cores ? mlp_forward_tc(run, b, h, st) : mlp_forward_simt(run, b, h, st);
}

}  // namespace th
