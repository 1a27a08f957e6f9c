// Shared helpers for the transhuman_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/transhuman_b200.h"

namespace th {

// ---- error plumbing (thread-local message, negative return codes) ----------
void set_error(const char* fmt, ...);
int64_t& launch_counter();

#define TH_CHECK_ARG(cond, msg)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      th::set_error("%s: %s", __func__, msg);           \
      return TH_EINVAL;                                 \
    }                                                   \
  } while (0)

#define TH_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      th::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return TH_ECUDA;                                                             \
    }                                                                              \
  } while (0)

// count + check a kernel launch
#define TH_LAUNCHED()                                                      \
  do {                                                                     \
    ++th::launch_counter();                                                \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      th::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return TH_ECUDA;                                                     \
    }                                                                      \
  } while (0)

// ---- optional per-category device timing (th_profile_start/stop) -------------
enum ProfCat { PROF_CULL = 0, PROF_FEATURES = 1, PROF_GEMM = 2, PROF_POINTWISE = 3, PROF_INTEGRATE = 4, PROF_PREMAP = 5, PROF_PROLOGUE = 6, PROF_NCAT = 7 };
void prof_begin(int cat, cudaStream_t st);
void prof_end(int cat, cudaStream_t st);
struct ProfScope {
  int cat;
  cudaStream_t st;
  ProfScope(int c, cudaStream_t s) : cat(c), st(s) { prof_begin(cat, st); }
  ~ProfScope() { prof_end(cat, st); }
};

// SM count of the CURRENT device (cached per device id: a process may render on several devices)
int device_sm_count(int* out);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- GEMM-layout activation widths (per (view, point) row) ------------------
constexpr int REP_LD = 256;   // human_rep row: 192 token + 63 PE + 1 zero pad
constexpr int PIX_LD = 384;
constexpr int VD_LD = 32;     // 27 view-direction channels + 5 zero pad
constexpr int TILE_PTS = 128; // points per feature-kernel CTA

// Where the points of a launch come from: either a ray bundle (sampler fused,
// row a1) or an explicit point array (grid query, row a12), optionally through
// an index list produced by the cull compaction.
struct PointSource {
  const float* ray_o;
  const float* ray_d;
  const float* near_;
  const float* far_;
  const float* t_vals;
  const float* pts;        // explicit points (world), or nullptr
  const int32_t* ids;      // optional: global point ids to process, or nullptr
  int64_t first;           // first list position (or first global id when ids == nullptr)
  int32_t n_samples;
};

// ---- device math with the reference's rounding order ---------
// z = near*(1-t) + far*t ; p = o + d*z  -- every op individually rounded
// (if_clight_renderer.py:274,285 are separate elementwise torch ops).
__device__ __forceinline__ float sample_z(float near_, float far_, float t) {
  float a = __fsub_rn(1.0f, t);
  return __fadd_rn(__fmul_rn(near_, a), __fmul_rn(far_, t));
}
__device__ __forceinline__ float3 sample_point(const float* o, const float* d, float z) {
  return make_float3(__fadd_rn(o[0], __fmul_rn(d[0], z)), __fadd_rn(o[1], __fmul_rn(d[1], z)),
                     __fadd_rn(o[2], __fmul_rn(d[2], z)));
}
// squared distance, no FMA contraction: the k-NN contract (oracle pairwise_d2)
__device__ __forceinline__ float dist2(float px, float py, float pz, float qx, float qy, float qz) {
  float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// (p - Th) @ Rh: torch.matmul (1,P,3)x(1,3,3) is an sgemm whose k-loop is an
// FMA chain on the reference's CPU path (measured; backend-defined in general).
__device__ __forceinline__ float3 world2smpl(float3 p, const float* Rh, const float* Th) {
  float x = __fsub_rn(p.x, Th[0]), y = __fsub_rn(p.y, Th[1]), z = __fsub_rn(p.z, Th[2]);
  float3 r;
  r.x = __fmaf_rn(z, Rh[6], __fmaf_rn(y, Rh[3], __fmul_rn(x, Rh[0])));
  r.y = __fmaf_rn(z, Rh[7], __fmaf_rn(y, Rh[4], __fmul_rn(x, Rh[1])));
  r.z = __fmaf_rn(z, Rh[8], __fmaf_rn(y, Rh[5], __fmul_rn(x, Rh[2])));
  return r;
}
// torch.norm over 3 components: sqrt(fma(z,z,fma(y,y,x*x))) on the CPU path.
__device__ __forceinline__ float norm3(float x, float y, float z) {
  return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}

// ---- fp32 -> fp16 hi/lo operand split of the tensor-core path ------------------------------
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits while |x| < 65504.  Both conversions
// SATURATE (cvt.satfinite): a plain conversion turns |x| > 65504 into hi = inf, lo = -inf and the three-product
// GEMM into inf - inf = NaN, where the fp32 reference just carries a large number.  Saturated, the pair
// represents x exactly-to-fp16 up to 2 x 65504 and clamps (finite, sign-correct) beyond.  Residuals below
// the fp16 subnormal step (6e-8) flush to zero: an absolute operand error of <= 3e-8.  NaN stays NaN.
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float lo_half, float hi_half) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}
__device__ __forceinline__ float2 f16x2_to_f32(uint32_t h) {
  float2 f;
  asm("{\n\t.reg .f16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}"
      : "=f"(f.x), "=f"(f.y)
      : "r"(h));
  return f;
}
// two adjacent values -> packed hi pair and packed lo pair
__device__ __forceinline__ void split_hl2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(x, y);
  const float2 hf = f16x2_to_f32(hi);
  lo = cvt_f16x2_sat(x - hf.x, y - hf.y);
}

// ---- raw2outputs (nerf_net_utils.py:14-59) in two parts, shared by k_integrate and by the chain kernel's fused
// compositing so that both produce the same bits: the per-sample terms (independent of the other samples) and the
// sequential step along the ray (transmittance as a running product, sums in sample order).
struct SampleTerm {
  float alpha, r, g, b;
};
// dist = (z_{s+1} - z_s) or 1e10 for the last sample, already multiplied by |ray_d|
__device__ __forceinline__ SampleTerm composite_sample(float4 raw, float dist) {
  SampleTerm t;
  const float sigma = fmaxf(raw.w, 0.f);
  t.alpha = __fsub_rn(1.0f, expf(-__fmul_rn(sigma, dist)));
  t.r = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw.x)));
  t.g = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw.y)));
  t.b = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-raw.z)));
  return t;
}
struct RayAcc {
  float T, r, g, b, acc, depth;
};
__device__ __forceinline__ RayAcc ray_acc_init() { return RayAcc{1.0f, 0.f, 0.f, 0.f, 0.f, 0.f}; }
__device__ __forceinline__ void composite_step(RayAcc& a, const SampleTerm& t, float z) {
  const float w = __fmul_rn(t.alpha, a.T);
  a.r = __fmaf_rn(w, t.r, a.r);
  a.g = __fmaf_rn(w, t.g, a.g);
  a.b = __fmaf_rn(w, t.b, a.b);
  a.depth = __fmaf_rn(w, z, a.depth);
  a.acc = __fadd_rn(a.acc, w);
  a.T = __fmul_rn(a.T, __fadd_rn(__fsub_rn(1.0f, t.alpha), 1e-10f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace th
