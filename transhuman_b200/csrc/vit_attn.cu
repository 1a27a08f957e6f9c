// Self-attention of the token transformer: vision_transformer.py:257-278 (`Attention.forward` between
// its two Linear layers) for vit_tiny (head_dim 64), flash-style -- softmax(q k^T * scale) v without the (B,H,N,N)
// attention matrix (1.3 GB per layer at 6000 tokens in the reference).
//
//   k_attn_split : qkv (B,N,3,H,64) fp32 (the output of `self.qkv`, 269) -> fp16 hi/lo planes [B,H,N,64] of
//                  q * scale * log2(e), k and v (x = hi + lo, 22 significant bits; the same operand split as the
//                  per-point network, common.cuh)
//   k_attn_fwd   : block = 128 queries (8 warps x 16 rows), key tiles of 64 double-buffered through cp.async;
//                  S = q k^T and O += P v as three fp16 tensor-core products each (hi*hi + lo*hi + hi*lo, fp32
//                  accumulate, mma.sync.m16n8k16), online softmax in fp32 registers in base 2, P split hi/lo in
//                  registers (the accumulator fragment of S is the A fragment of the next product).
// Result (B,N,H*64) fp32 = the input of `self.proj` (275-276); within ~1e-6 of the fp32 evaluation.
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace th {
namespace attn {

constexpr int D = 64;        // head dim
constexpr int BQ = 128;      // queries per block
constexpr int BK = 64;       // keys per tile
constexpr int NWARP = BQ / 16;
constexpr int PLANE = BK * D * 2;              // bytes of one fp16 tile plane (8 KB)
constexpr int STAGE = 4 * PLANE;               // K hi, K lo, V hi, V lo
constexpr int SMEM = 2 * STAGE;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool ok) {
  const int n = ok ? 16 : 0;  // zero-fill past the last key
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// planes: [which (q,k,v)][hi/lo][B*H][N][64] fp16
__global__ void __launch_bounds__(256) k_attn_split(const float* __restrict__ qkv, int B, int N, int H, float qscale,
                                                    __half* __restrict__ planes) {
  // one thread per (b, n, which, h, pair of channels); qkv row = [which][h][64]
  const int64_t total = (int64_t)B * N * 3 * H * (D / 2);
  const int64_t plane_elems = (int64_t)B * H * N * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c2 = (int)(i % (D / 2));
    int64_t r = i / (D / 2);
    const int h = (int)(r % H);
    r /= H;
    const int which = (int)(r % 3);
    r /= 3;
    const int n = (int)(r % N), b = (int)(r / N);
    float2 x = *reinterpret_cast<const float2*>(qkv + 2 * i);
    if (which == 0) x.x *= qscale, x.y *= qscale;
    uint32_t hi, lo;
    split_hl2(x.x, x.y, hi, lo);
    const int64_t o = (((int64_t)b * H + h) * N + n) * D + 2 * c2;
    *reinterpret_cast<uint32_t*>(planes + (int64_t)(which * 2) * plane_elems + o) = hi;
    *reinterpret_cast<uint32_t*>(planes + (int64_t)(which * 2 + 1) * plane_elems + o) = lo;
  }
}

// tile element (row, 16-byte chunk c of the 128-byte row) lives at row * 128 + ((c ^ (row & 7)) << 4)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(NWARP * 32, 1) k_attn_fwd(const __half* __restrict__ planes, int B, int N, int H,
                                                            float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y;  // b * H + h
  const int q0 = blockIdx.x * BQ;
  const int64_t plane_elems = (int64_t)B * H * N * D;
  const __half* q_hi = planes + (int64_t)bh * N * D;
  const __half* q_lo = q_hi + plane_elems;
  const __half* kv[4] = {q_hi + 2 * plane_elems, q_hi + 3 * plane_elems, q_hi + 4 * plane_elems,
                         q_hi + 5 * plane_elems};  // K hi, K lo, V hi, V lo
  const uint32_t sbase = smem_addr(smem);

  auto load_tile = [&](int stage, int k0) {
    // 4 planes x 64 rows x 8 chunks of 16 bytes = 2048 chunks, 256 threads
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * (NWARP * 32);
      const int p = idx >> 9, row = (idx >> 3) & 63, c = idx & 7;
      const int key = k0 + row;
      const bool ok = key < N;
      const __half* src = kv[p] + ((int64_t)(ok ? key : 0) * D + c * 8);
      cp_async16(sbase + stage * STAGE + p * PLANE + tile_off(row, c), src, ok);
    }
    cp_async_commit();
  };

  const int n_tiles = (N + BK - 1) / BK;
  load_tile(0, 0);

  // Q fragments of this warp's 16 rows: 4 k-steps (d chunks of 16) x {hi, lo}
  uint32_t qa[2][4][4];
  {
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const __half* qp = p ? q_lo : q_hi;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int c = kk * 16 + 2 * t;
        qa[p][kk][0] = r0 < N ? *reinterpret_cast<const uint32_t*>(qp + (int64_t)r0 * D + c) : 0u;
        qa[p][kk][1] = r1 < N ? *reinterpret_cast<const uint32_t*>(qp + (int64_t)r1 * D + c) : 0u;
        qa[p][kk][2] = r0 < N ? *reinterpret_cast<const uint32_t*>(qp + (int64_t)r0 * D + c + 8) : 0u;
        qa[p][kk][3] = r1 < N ? *reinterpret_cast<const uint32_t*>(qp + (int64_t)r1 * D + c + 8) : 0u;
      }
    }
  }

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  // ldmatrix lane -> (row within 16, chunk within 2) of a 16 x 16 block: lanes 0-7 rows 0-7 chunk 0, 8-15 rows 8-15
  // chunk 0, 16-23 rows 0-7 chunk 1, 24-31 rows 8-15 chunk 1
  const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8, lm_chunk = lane >> 4;

  for (int it = 0; it < n_tiles; ++it) {
    const int stage = it & 1;
    if (it + 1 < n_tiles) {
      load_tile(stage ^ 1, (it + 1) * BK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t s_khi = sbase + stage * STAGE, s_klo = s_khi + PLANE, s_vhi = s_khi + 2 * PLANE,
                   s_vlo = s_khi + 3 * PLANE;

    // ---- S = q k^T (16 x 64 per warp) -------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of key n-tiles: keys 16 jp .. 16 jp + 15
        // non-transposed ldmatrix over K rows (keys) x d chunk: matrices (keys lo8, d lo8), (keys hi8, d lo8),
        // (keys lo8, d hi8), (keys hi8, d hi8) -> b0 of tile 2jp, b0 of tile 2jp+1, b1 of 2jp, b1 of 2jp+1
        const uint32_t off = tile_off(jp * 16 + lm_row, kk * 2 + lm_chunk);
        uint32_t h0, h1, h2, h3, e0, e1, e2, e3;
        ldsm_x4(s_khi + off, h0, h1, h2, h3);
        ldsm_x4(s_klo + off, e0, e1, e2, e3);
        mma16816(s[2 * jp], qa[0][kk], h0, h2);
        mma16816(s[2 * jp + 1], qa[0][kk], h1, h3);
        mma16816(s[2 * jp], qa[1][kk], h0, h2);
        mma16816(s[2 * jp + 1], qa[1][kk], h1, h3);
        mma16816(s[2 * jp], qa[0][kk], e0, e2);
        mma16816(s[2 * jp + 1], qa[0][kk], e1, e3);
      }
    }
    // ---- mask the keys past N (last tile) ------------------------------------------------
    const int k0 = it * BK;
    if (k0 + BK > N) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = k0 + j * 8 + 2 * t;
        if (key >= N) s[j][0] = s[j][2] = -INFINITY;
        if (key + 1 >= N) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    // ---- online softmax (base 2; q carries scale * log2 e) -------------------------------
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float a0 = ex2(m0 - mx0), a1 = ex2(m1 - mx1);  // first tile: ex2(-inf) = 0
    m0 = mx0, m1 = mx1;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[2][4][4];  // [hi/lo][key chunk of 16][a0..a3]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = ex2(s[j][0] - mx0), p1 = ex2(s[j][1] - mx0), p2 = ex2(s[j][2] - mx1), p3 = ex2(s[j][3] - mx1);
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      uint32_t hi01, lo01, hi23, lo23;
      split_hl2(p0, p1, hi01, lo01);
      split_hl2(p2, p3, hi23, lo23);
      const int kk = j >> 1, half = j & 1;
      pa[0][kk][half * 2] = hi01, pa[0][kk][half * 2 + 1] = hi23;
      pa[1][kk][half * 2] = lo01, pa[1][kk][half * 2 + 1] = lo23;
    }
    l0 = l0 * a0 + sum0;
    l1 = l1 * a1 + sum1;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] *= a0, o[j][1] *= a0, o[j][2] *= a1, o[j][3] *= a1;
    // ---- O += P v ------------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {       // keys 16 kk .. 16 kk + 15
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {     // d n-tiles 2jp, 2jp+1
        // transposed ldmatrix over V rows (keys) x d chunk: matrices (keys lo8, d chunk 2jp), (keys hi8, 2jp),
        // (keys lo8, 2jp+1), (keys hi8, 2jp+1) -> b0, b1 of tile 2jp, b0, b1 of tile 2jp+1
        const uint32_t off = tile_off(kk * 16 + lm_row, jp * 2 + lm_chunk);
        uint32_t h0, h1, h2, h3, e0, e1, e2, e3;
        ldsm_x4_t(s_vhi + off, h0, h1, h2, h3);
        ldsm_x4_t(s_vlo + off, e0, e1, e2, e3);
        mma16816(o[2 * jp], pa[0][kk], h0, h1);
        mma16816(o[2 * jp + 1], pa[0][kk], h2, h3);
        mma16816(o[2 * jp], pa[1][kk], h0, h1);
        mma16816(o[2 * jp + 1], pa[1][kk], h2, h3);
        mma16816(o[2 * jp], pa[0][kk], e0, e1);
        mma16816(o[2 * jp + 1], pa[0][kk], e2, e3);
      }
    }
    __syncthreads();  // the next iteration's prefetch overwrites this stage's sibling; all warps must be done
  }
  // ---- normalise and store (B, N, H*64) --------------------------------------------------
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int b = bh / H, h = bh - b * H;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = h * D + j * 8 + 2 * t;
    if (r0 < N) *reinterpret_cast<float2*>(out + ((int64_t)b * N + r0) * (H * D) + c) = make_float2(o[j][0] * i0, o[j][1] * i0);
    if (r1 < N) *reinterpret_cast<float2*>(out + ((int64_t)b * N + r1) * (H * D) + c) = make_float2(o[j][2] * i1, o[j][3] * i1);
  }
}

}  // namespace attn

size_t vit_attention_workspace_bytes(int B, int N, int H) {
  return align_up((size_t)6 * B * H * N * attn::D * sizeof(__half), 256);
}

int launch_vit_attention(const float* qkv, int B, int N, int H, float scale, float* out, void* workspace,
                         cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  __half* planes = static_cast<__half*>(workspace);
  const int64_t total = (int64_t)B * N * 3 * H * (attn::D / 2);
  const unsigned grid = (unsigned)(cdiv(total, 256) < 148 * 16 ? cdiv(total, 256) : 148 * 16);
  attn::k_attn_split<<<grid, 256, 0, st>>>(qkv, B, N, H, scale * 1.4426950408889634f, planes);
  TH_LAUNCHED();
  TH_CUDA(cudaFuncSetAttribute(attn::k_attn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM));
  attn::k_attn_fwd<<<dim3((unsigned)cdiv(N, attn::BQ), (unsigned)(B * H)), attn::NWARP * 32, attn::SMEM, st>>>(planes, B, N, H,
                                                                                                        out);
  TH_LAUNCHED();
  return TH_OK;
}

}  // namespace th
