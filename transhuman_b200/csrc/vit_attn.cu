// Self-attention of the token transformer: vision_transformer.py:257-278 (`Attention.forward` between
// its two Linear layers) for vit_tiny (head_dim 64), flash-style -- softmax(q k^T * scale) v without the (B,H,N,N)
// attention matrix (1.3 GB per layer at 6000 tokens in the reference).
//
// Every operand is split x = hi + lo into two fp16 values (22 significant bits; the same split as the per-point
// network, common.cuh) and every product is three tensor-core products hi*hi + lo*hi + hi*lo with fp32 accumulation;
// the softmax runs in fp32 in base 2 (q carries scale * log2 e).
//
//                      k_attn_images : qkv (B,N,3,H,64) fp32 (the output of `self.qkv`, 269) -> per (view, head) and
//                      128-query / 64-key tile the shared-memory OPERAND IMAGES (K-major rows of 128 bytes, 128-byte
//                      swizzle; zero padded past N): Q [hi | lo] 32 KB, and per key tile K [hi | lo] and V^T [hi | lo]
//                      = 32 KB, so a pipeline stage is ONE bulk copy.
//                      k_attn_tc : one CTA per TWO 128-query tiles x (view, head); warp 8 = loader (cp.async.bulk
//                      into two 32 KB stages), warp 9 = MMA issuer (tcgen05.mma.cta_group::1, M = 128, N = 64,
//                      K = 16: S_g(j) = Q_g K_j^T and O_g(j) = P_g(j) V_j into TMEM), warps 0-3 / 4-7 = the softmax
//                      groups of the two query tiles, running out of phase so that each group's latency chain (TMEM
//                      load, exp2, operand store, proxy fence, barrier) is covered by the tensor work of the other:
//                      thread = query row (its TMEM lane), row max / exp2 / sum without any shuffle, P written as an
//                      fp16 hi/lo A-operand image into shared memory, O accumulated in registers with the running
//                      rescale.
// (A first version on the legacy tensor path -- mma.sync.m16n8k16, 8 warps x 16 rows, P split in registers -- was
// 1.7x slower at 6000 tokens and, because that path's accumulation is less accurate, 10x further from the float64
// result; it was removed.)
// Result (B,N,H*64) fp32 = the input of `self.proj` (275-276); within ~1e-6 of the fp32 evaluation.
#include <cuda_fp16.h>

#include <stdlib.h>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace th {
namespace attn {

constexpr int D = 64;        // head dim
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ===========================================================================
// tcgen05 schedule
// ===========================================================================
constexpr int TQ = 128;                 // queries per CTA = MMA M
constexpr int TK = 64;                  // keys per tile = MMA N of S, K of P V
constexpr int IMG_Q = 2 * TQ * 128;     // [hi | lo] 32 KB
constexpr int IMG_KV = 4 * TK * 128;    // [K hi | K lo | V^T hi | V^T lo] 32 KB
constexpr int IMG_P = 2 * TQ * 128;     // [hi | lo] 32 KB per buffer
constexpr int TC_CTRL = 256;
constexpr int TC_SMEM = 1024 + 2 * IMG_Q + 2 * IMG_KV + 2 * IMG_P + TC_CTRL;
constexpr int TC_THREADS = 10 * 32;   // 2 x 4 softmax warps (one group per 128-query tile) + loader + MMA issuer

__host__ __device__ inline size_t images_bytes(int B, int N, int H) {
  const size_t nq = 2 * ((size_t)(N + 2 * TQ - 1) / (2 * TQ)), nk = (size_t)(N + TK - 1) / TK;  // query tiles in pairs
  return (size_t)B * H * (nq * IMG_Q + nk * IMG_KV);
}

// One thread per 16-byte chunk (8 fp16) of the hi plane and of the lo plane.  Image element (row r, chunk c) lives at
// r * 128 + ((c ^ (r & 7)) << 4).  Q and K rows are tokens (chunk = 8 consecutive d); V^T rows are d (chunk = 8
// consecutive keys of the tile).
__global__ void __launch_bounds__(256) k_attn_images(const float* __restrict__ qkv, int B, int N, int H, float qscale,
                                                     unsigned char* __restrict__ img) {
  const int nq = 2 * ((N + 2 * TQ - 1) / (2 * TQ)), nk = (N + TK - 1) / TK;
  const int64_t q_chunks = (int64_t)nq * TQ * 8, k_chunks = (int64_t)nk * TK * 8;  // per (b, h); V^T has k_chunks too
  const int64_t per_bh = q_chunks + 2 * k_chunks;
  const int64_t total = (int64_t)B * H * per_bh;
  const int64_t row_floats = (int64_t)3 * H * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int bh = (int)(i / per_bh);
    int64_t r = i - (int64_t)bh * per_bh;
    const int b = bh / H, h = bh - b * H;
    unsigned char* base = img + (size_t)bh * ((size_t)nq * IMG_Q + (size_t)nk * IMG_KV);
    float x[8];
    unsigned char* dst;  // hi chunk; lo chunk at + plane
    int plane;
    if (r < q_chunks) {
      const int c = (int)(r & 7), tok = (int)(r >> 3), tile = tok / TQ, row = tok - tile * TQ;
      const float* src = qkv + ((int64_t)b * N + tok) * row_floats + h * D + c * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = tok < N ? __ldg(src + e) * qscale : 0.f;
      dst = base + (size_t)tile * IMG_Q + row * 128 + ((c ^ (row & 7)) << 4);
      plane = TQ * 128;
    } else if (r < q_chunks + k_chunks) {
      r -= q_chunks;
      const int c = (int)(r & 7), tok = (int)(r >> 3), tile = tok / TK, row = tok - tile * TK;
      const float* src = qkv + ((int64_t)b * N + tok) * row_floats + (H + h) * D + c * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = tok < N ? __ldg(src + e) : 0.f;
      dst = base + (size_t)nq * IMG_Q + (size_t)tile * IMG_KV + row * 128 + ((c ^ (row & 7)) << 4);
      plane = TK * 128;
    } else {
      r -= q_chunks + k_chunks;
      // d fastest so that a warp reads 32 consecutive floats of a token row
      const int d = (int)(r % D);
      const int64_t t = r / D;
      const int c = (int)(t & 7), tile = (int)(t >> 3);
      const int key0 = tile * TK + c * 8;
      const float* src = qkv + ((int64_t)b * N + key0) * row_floats + (2 * H + h) * D + d;
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = key0 + e < N ? __ldg(src + e * row_floats) : 0.f;
      dst = base + (size_t)nq * IMG_Q + (size_t)tile * IMG_KV + 2 * TK * 128 + d * 128 + ((c ^ (d & 7)) << 4);
      plane = TK * 128;
    }
    uint4 hi, lo;
    split_hl2(x[0], x[1], hi.x, lo.x);
    split_hl2(x[2], x[3], hi.y, lo.y);
    split_hl2(x[4], x[5], hi.z, lo.z);
    split_hl2(x[6], x[7], hi.w, lo.w);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + plane) = lo;
  }
}

// One CTA = TWO 128-query tiles (g = 0, 1) of one (view, head) against all key tiles.  The two softmax groups run
// out of phase: while group g turns S_g(j) into P_g(j), the tensor core computes O_{1-g}(j) = P_{1-g}(j) V_j and
// S_{1-g}(j+1) -- every latency of one group's chain (TMEM load, proxy fence, barrier round trip) is covered by the
// other group's work.
// G = query tiles per CTA: 2 when that still fills the machine, else 1 (one softmax group; twice the CTAs).
template <int G>
__global__ void __launch_bounds__(TC_THREADS, 1) k_attn_tc(const unsigned char* __restrict__ img, int B, int N, int H,
                                                           float* __restrict__ out) {
  using namespace th::tc;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t s_q = base, s_kv = base + 2 * IMG_Q, s_p = s_kv + 2 * IMG_KV, ctrl = s_p + 2 * IMG_P;
  unsigned char* p_ptr = base_ptr + 2 * IMG_Q + 2 * IMG_KV;
  const uint32_t bar_q = ctrl, bar_kvf = ctrl + 8, bar_kve = ctrl + 24, bar_sf = ctrl + 40, bar_pf = ctrl + 56,
                 bar_of = ctrl + 72;  // kvf / kve per stage; sf / pf / of per query tile g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + 2 * IMG_Q + 2 * IMG_KV + 2 * IMG_P + 96);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, qp = blockIdx.x;  // qp = group of G query tiles
  const int nq = 2 * ((N + 2 * TQ - 1) / (2 * TQ)), nk = (N + TK - 1) / TK;
  const unsigned char* img_bh = img + (size_t)bh * ((size_t)nq * IMG_Q + (size_t)nk * IMG_KV);
  const unsigned char* img_q = img_bh + (size_t)(G * qp) * IMG_Q;
  const unsigned char* img_kv = img_bh + (size_t)nq * IMG_Q;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_kvf + 8 * s, 1);
      mbar_init(bar_kve + 8 * s, 1);
      mbar_init(bar_sf + 8 * s, 1);
      mbar_init(bar_pf + 8 * s, 128);
      mbar_init(bar_of + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_o = tmem_base + 2 * TK;  // S_g at columns 64 g, O_g at 128 + 64 g

  if (warp == 8) {
    // ===================== loader =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, G * IMG_Q);
      for (int g = 0; g < G; ++g) bulk_g2s(s_q + g * IMG_Q, img_q + (size_t)g * IMG_Q, IMG_Q, bar_q);
      for (int j = 0; j < nk; ++j) {
        const int s = j & 1;
        mbar_wait(bar_kve + 8 * s, ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar_kvf + 8 * s, IMG_KV);
        bulk_g2s(s_kv + s * IMG_KV, img_kv + (size_t)j * IMG_KV, IMG_KV, bar_kvf + 8 * s);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(TK);  // M = 128, N = 64, fp16 x fp16 -> fp32, both operands K-major
      auto issue_s = [&](int g, int j) {      // S_g(j) = Q_g K_j^T; the caller has waited for stage j & 1
        const int s = j & 1;
        const uint64_t d_qhi = umma_desc(s_q + g * IMG_Q), d_qlo = umma_desc(s_q + g * IMG_Q + TQ * 128);
        const uint64_t d_khi = umma_desc(s_kv + s * IMG_KV), d_klo = umma_desc(s_kv + s * IMG_KV + TK * 128);
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 32) >> 4);
          umma_f16(t_s + g * TK, d_qhi + adv, d_khi + adv, idesc, ks ? 1u : 0u);
          umma_f16(t_s + g * TK, d_qlo + adv, d_khi + adv, idesc, 1u);
          umma_f16(t_s + g * TK, d_qhi + adv, d_klo + adv, idesc, 1u);
        }
        umma_commit(bar_sf + 8 * g);
      };
      mbar_wait(bar_q, 0);
      mbar_wait(bar_kvf, 0);
      tc_fence_after();
      for (int g = 0; g < G; ++g) issue_s(g, 0);
      for (int j = 0; j < nk; ++j) {
        const int s = j & 1;
        for (int g = 0; g < G; ++g) {
          mbar_wait(bar_pf + 8 * g, j & 1);   // P_g(j) written; S_g(j) and O_g(j-1) read
          tc_fence_after();
          const uint64_t d_phi = umma_desc(s_p + g * IMG_P), d_plo = umma_desc(s_p + g * IMG_P + TQ * 128);
          const uint64_t d_vhi = umma_desc(s_kv + s * IMG_KV + 2 * TK * 128),
                         d_vlo = umma_desc(s_kv + s * IMG_KV + 3 * TK * 128);
#pragma unroll
          for (int ks = 0; ks < TK / 16; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
            umma_f16(t_o + g * D, d_phi + adv, d_vhi + adv, idesc, ks ? 1u : 0u);
            umma_f16(t_o + g * D, d_plo + adv, d_vhi + adv, idesc, 1u);
            umma_f16(t_o + g * D, d_phi + adv, d_vlo + adv, idesc, 1u);
          }
          umma_commit(bar_of + 8 * g);
          if (g == G - 1) umma_commit(bar_kve + 8 * s);  // stage j & 1 (K_j, V_j) no longer read
          if (j + 1 < nk) {
            if (g == 0) {
              mbar_wait(bar_kvf + 8 * (s ^ 1), ((j + 1) >> 1) & 1);
              tc_fence_after();
            }
            issue_s(g, j + 1);
          }
        }
      }
    }
  } else if ((warp >> 2) < G) {
    // ===================== softmax + output accumulation: group g = query tile, thread = query row ============
    const int g = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t ts = t_s + lane_addr + g * TK, to = t_o + lane_addr + g * D;
    unsigned char* prow = p_ptr + g * IMG_P + row * 128;
    float o[D];
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(bar_sf + 8 * g, j & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(ts, v0);
      tmem_ld32(ts + 32, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int n_valid = N - j * TK;  // keys of this tile that exist (>= 1)
      if (n_valid < TK) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c >= n_valid) v0[c] = 0xff800000u;  // -inf
          if (c + 32 >= n_valid) v1[c] = 0xff800000u;
        }
      }
      float mx = m;
#pragma unroll
      for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[c]), __uint_as_float(v1[c])));
      const float alpha = ex2(m - mx);  // first tile: ex2(-inf) = 0
      m = mx;
      // O += O_g(j-1) (same scale as the running sum before this tile), then rescale to the new maximum.  Its
      // product was issued while the other group ran; waiting for it also frees P_g for this tile.
      if (j > 0) {
        mbar_wait(bar_of + 8 * g, (j - 1) & 1);
        tc_fence_after();
        uint32_t w0[32], w1[32];
        tmem_ld32(to, w0);
        tmem_ld32(to + 32, w1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          o[c] = (o[c] + __uint_as_float(w0[c])) * alpha;
          o[c + 32] = (o[c + 32] + __uint_as_float(w1[c])) * alpha;
        }
      }
      // P_g(j) = exp2(S - m) as an fp16 hi/lo A-operand image (row = 128 bytes = the tile's 64 keys)
      float sum = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float p[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t raw = c8 < 4 ? v0[c8 * 8 + e] : v1[(c8 - 4) * 8 + e];
          p[e] = ex2(__uint_as_float(raw) - mx);
          sum += p[e];
        }
        uint4 hi, lo;
        split_hl2(p[0], p[1], hi.x, lo.x);
        split_hl2(p[2], p[3], hi.y, lo.y);
        split_hl2(p[4], p[5], hi.z, lo.z);
        split_hl2(p[6], p[7], hi.w, lo.w);
        const int off = (c8 ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(prow + off) = hi;
        *reinterpret_cast<uint4*>(prow + TQ * 128 + off) = lo;
      }
      l = l * alpha + sum;
      fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
      tc_fence_before();     // the TMEM loads of S_g(j) / O_g(j-1) are complete before the issuer overwrites them
      mbar_arrive(bar_pf + 8 * g);
    }
    {
      mbar_wait(bar_of + 8 * g, (nk - 1) & 1);
      tc_fence_after();
      uint32_t w0[32], w1[32];
      tmem_ld32(to, w0);
      tmem_ld32(to + 32, w1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const float inv = 1.f / l;
      const int b = bh / H, h = bh - b * H;
      const int q = (G * qp + g) * TQ + row;
      if (q < N) {
        float4* dst = reinterpret_cast<float4*>(out + ((int64_t)b * N + q) * (H * D) + h * D);
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          dst[c >> 2] = make_float4((o[c] + __uint_as_float(w0[c])) * inv, (o[c + 1] + __uint_as_float(w0[c + 1])) * inv,
                                    (o[c + 2] + __uint_as_float(w0[c + 2])) * inv,
                                    (o[c + 3] + __uint_as_float(w0[c + 3])) * inv);
          dst[8 + (c >> 2)] =
              make_float4((o[c + 32] + __uint_as_float(w1[c])) * inv, (o[c + 33] + __uint_as_float(w1[c + 1])) * inv,
                          (o[c + 34] + __uint_as_float(w1[c + 2])) * inv, (o[c + 35] + __uint_as_float(w1[c + 3])) * inv);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace attn

size_t vit_attention_workspace_bytes(int B, int N, int H) {
  return align_up(attn::images_bytes(B, N, H), 256);
}

int launch_vit_attention(const float* qkv, int B, int N, int H, float scale, float* out, void* workspace,
                         cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  const float qscale = scale * 1.4426950408889634f;
  unsigned char* img = static_cast<unsigned char*>(workspace);
  const int64_t chunks = (int64_t)attn::images_bytes(B, N, H) / 32;  // one thread per hi + lo chunk pair
  const unsigned grid = (unsigned)(cdiv(chunks, 256) < 148 * 16 ? cdiv(chunks, 256) : 148 * 16);
  attn::k_attn_images<<<grid, 256, 0, st>>>(qkv, B, N, H, qscale, img);
  TH_LAUNCHED();
  int num_sms = 0;
  if (device_sm_count(&num_sms)) return TH_ECUDA;
  if (cdiv(N, 2 * attn::TQ) * B * H >= num_sms) {
    TH_CUDA(cudaFuncSetAttribute(attn::k_attn_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::TC_SMEM));
    attn::k_attn_tc<2><<<dim3((unsigned)cdiv(N, 2 * attn::TQ), (unsigned)(B * H)), attn::TC_THREADS, attn::TC_SMEM, st>>>(
        img, B, N, H, out);
  } else {
    TH_CUDA(cudaFuncSetAttribute(attn::k_attn_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::TC_SMEM));
    attn::k_attn_tc<1><<<dim3((unsigned)cdiv(N, attn::TQ), (unsigned)(B * H)), attn::TC_THREADS, attn::TC_SMEM, st>>>(
        img, B, N, H, out);
  }
  TH_LAUNCHED();
  return TH_OK;
}

}  // namespace th
