// C-ABI entry points (include/transhuman_b200.h): argument checking, host-side
// weight packing, workspace planning and the kernel schedule of the fused path.
#include <cuda_fp16.h>
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "kernels.cuh"

namespace th {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t& launch_counter() { return g_launches; }

int device_sm_count(int* out) {
  static int cache[64] = {0};  // benign race: every writer stores the same value
  int dev = 0;
  TH_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !cache[dev]) {
    int n = 0;
    TH_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) cache[dev] = n;
    *out = n;
    return TH_OK;
  }
  *out = cache[dev];
  return TH_OK;
}

// ---- profiler: CUDA events around every launch of a category, summed at stop ----
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev[PROF_NCAT];  // start/stop pairs
  std::vector<cudaEvent_t> pool;
  int launches[PROF_NCAT] = {0};
};
static thread_local ProfState g_prof;
static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof.pool.empty()) {
    e = g_prof.pool.back();
    g_prof.pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
void prof_begin(int cat, cudaStream_t st) {
  if (!g_prof.on) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, st);
  g_prof.ev[cat].push_back(e);
}
void prof_end(int cat, cudaStream_t st) {
  if (!g_prof.on) return;
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, st);
  g_prof.ev[cat].push_back(e);
  ++g_prof.launches[cat];
}

// Points per chunk.  284160 = 148 SMs x 5 resident feature-kernel CTAs x 128 points x 3 waves
// = 74 clusters x 15 units of 256 points for the chain kernel: both kernels run whole waves
// (262144 left the feature kernel's third wave 77 % full).  Workspace ~ 26 KiB / point.
// TH_CHUNK_PTS overrides (multiple of 256).
static int64_t chunk_pts() {
  static int64_t v = 0;
  if (!v) {
    const char* e = getenv("TH_CHUNK_PTS");
    v = e ? atoll(e) : 284160;
    if (v < 256) v = 256;
    v = v / 256 * 256;
  }
  return v;
}

// fp16 hi/lo tile image of an (n_rows, K) fp32 matrix (rows beyond n_rows up to N are zero) scaled by `up`, in the
// layout k_gemm_tc2 / k_chain copy into shared memory: per 64-wide k-block and per half of the N rows (one half per
// CTA of a cta_group::2 pair) [hi | lo], each N/2 rows of 128 bytes, K-major, 16-byte chunk ^= row & 7.
static void pack_weight_image(const float* src, int N, int K, int n_rows, float up, unsigned char* img) {
  auto sat = [](float f) { return f > 65504.f ? 65504.f : (f < -65504.f ? -65504.f : f); };  // like the device split
  for (int kb = 0; kb < K / 64; ++kb)
    for (int n = 0; n < N; ++n)
      for (int kk = 0; kk < 64; ++kk) {
        const float x = n < n_rows ? src[(size_t)n * K + kb * 64 + kk] * up : 0.f;
        const __half a = __float2half_rn(sat(x));
        const __half b = __float2half_rn(sat(x - __half2float(a)));
        const size_t off = (size_t)n * 128 + (size_t)(((kk >> 3) ^ (n & 7)) << 4) + (size_t)(kk & 7) * 2;
        const int half_rows = N / 2, r = n / half_rows;
        unsigned char* tile = img + (size_t)kb * (2 * N * 128) + (size_t)r * (N * 128);  // this half: [hi | lo]
        memcpy(tile + off - (size_t)r * half_rows * 128, &a, 2);
        memcpy(tile + (size_t)half_rows * 128 + off - (size_t)r * half_rows * 128, &b, 2);
      }
}
static int weight_scale_exp(const float* src, size_t n) {  // max|w| 2^e in (2^13, 2^14]
  float wmax = 0.f;
  for (size_t i = 0; i < n; ++i) wmax = fmaxf(wmax, fabsf(src[i]));
  int e = 0;
  if (wmax > 0.f && isfinite(wmax)) {
    int ex = 0;
    frexpf(wmax, &ex);  // wmax = f * 2^ex, f in [0.5, 1)
    e = 14 - ex;
    if (e > 60) e = 60;
    if (e < -60) e = -60;
  }
  return e;
}

// ---------------------------------------------------------------------------
// weight packing (host)
// ---------------------------------------------------------------------------
struct MatSpec {
  uint64_t PackedHeader::*w;
  uint64_t PackedHeader::*b;
  uint64_t PackedHeader::*h;  // fp16 hi/lo planes (nullptr for the heads)
  int N, K;
};

static std::vector<MatSpec> mat_specs(int V) {
  typedef PackedHeader H;
  return {
      {&H::fc0_w, &H::fc0_b, &H::h_fc0, 256, 256},
      {&H::ar0_w, &H::ar0_b, &H::h_ar0, 256, 384},
      {&H::k0_w, &H::k0_b, &H::h_k0, 128, 256},
      {&H::k1_w, &H::k1_b, &H::h_k1, 128, 256},
      {&H::v_w, &H::v_b, &H::h_v, 256, 512},
      {&H::fc1_w, &H::fc1_b, &H::h_fc1, 256, 256},
      {&H::fc2_w, &H::fc2_b, &H::h_fc2, 256, 256},
      {&H::fc3m_w, &H::fc3m_b, &H::h_fc3m, 256, 256 * V},
      {&H::afc_w, &H::afc_b, nullptr, 1, 256},
      {&H::f_w, &H::f_b, &H::h_f, 256, 640},
      {&H::view_w, &H::view_b, &H::h_view, 128, 320},
      {&H::t_w, &H::t_b, &H::h_t, 128, 128 * V + 384},
      {&H::rgb_w, &H::rgb_b, nullptr, 3, 128},
      {&H::fc1f_w, &H::fc1f_b, &H::h_fc1f, 256, 512},
      {&H::gvf_w, &H::gvf_b, &H::h_gvf, 128, 704},
      {&H::pre_w, &H::pre_b, nullptr, 512, 384},
      {&H::gvfp_w, &H::gvfp_b, &H::h_gvfp, 128, 448},
      {&H::tp_w, &H::tp_b, &H::h_tp, 128, 128 * V + 128},
      {&H::xid_w, &H::xid_b, &H::h_xid, 256, 256},
      {&H::preb_w, &H::preb_b, &H::h_preb, 256, 384},
      {&H::fc1s0_w, &H::fc1s0_b, &H::h_fc1s0, 128, 256},
      {&H::fc1s1_w, &H::fc1s1_b, &H::h_fc1s1, 128, 256},
      {&H::fc1x0_w, &H::fc1x0_b, &H::h_fc1x0, 128, 256},
      {&H::fc1x1_w, &H::fc1x1_b, &H::h_fc1x1, 128, 256},
  };
}

static void layout_header(int V, PackedHeader* h) {
  memset(h, 0, sizeof(*h));
  h->magic = PACK_MAGIC;
  h->n_views = V;
  uint64_t off = align_up(sizeof(PackedHeader), 256);
  for (const MatSpec& m : mat_specs(V)) {
    h->*(m.w) = off;
    off += align_up((size_t)m.N * m.K * 4, 256);
    h->*(m.b) = off;
    off += align_up((size_t)m.N * 4, 256);
  }
  for (const MatSpec& m : mat_specs(V)) {
    if (!m.h) continue;
    h->*(m.h) = off;
    off += align_up((size_t)m.N * m.K * 2 * 2, 1024);
  }
  h->total_bytes = off;
}

// zero the raw rows of masked-out points (Network.forward scatter into zeros,
// cross_transformer.py:229-233, 267-269)
__global__ void k_apply_mask(float4* raw, const uint8_t* __restrict__ mask, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && !mask[i]) raw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
int launch_apply_mask(float* raw, const uint8_t* mask, int64_t n, cudaStream_t st) {
  k_apply_mask<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(reinterpret_cast<float4*>(raw), mask, n);
  TH_LAUNCHED();
  return TH_OK;
}

}  // namespace th

using namespace th;

extern "C" {

const char* th_version(void) { return "transhuman_b200 0.1 (sm_100a)"; }
const char* th_last_error(void) { return g_err; }
int64_t th_launch_count(int32_t reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int th_profile_start(void) {
  for (int c = 0; c < PROF_NCAT; ++c) {
    for (cudaEvent_t e : g_prof.ev[c]) g_prof.pool.push_back(e);
    g_prof.ev[c].clear();
    g_prof.launches[c] = 0;
  }
  g_prof.on = true;
  return TH_OK;
}

int th_profile_stop(double* ms_per_category, int64_t* launches_per_category, int32_t n) {
  g_prof.on = false;
  TH_CUDA(cudaDeviceSynchronize());
  for (int c = 0; c < PROF_NCAT; ++c) {
    double total = 0.0;
    for (size_t i = 0; i + 1 < g_prof.ev[c].size(); i += 2) {
      float ms = 0.f;
      TH_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[c][i], g_prof.ev[c][i + 1]));
      total += ms;
    }
    if (c < n) {
      if (ms_per_category) ms_per_category[c] = total;
      if (launches_per_category) launches_per_category[c] = g_prof.launches[c];
    }
  }
  return TH_OK;
}

size_t th_packed_weights_bytes(int32_t n_views) {
  if (n_views < 1 || n_views > TH_MAX_VIEWS) return 0;
  PackedHeader h;
  layout_header(n_views, &h);
  return (size_t)h.total_bytes;
}

int th_pack_weights(const ThWeightsF32* w, int32_t V, void* packed_host, size_t bytes) {
  TH_CHECK_ARG(w && packed_host, "null pointer");
  TH_CHECK_ARG(V >= 1 && V <= TH_MAX_VIEWS, "n_views out of range");
  const float* const* fields = reinterpret_cast<const float* const*>(w);
  for (size_t i = 0; i < sizeof(ThWeightsF32) / sizeof(float*); ++i) TH_CHECK_ARG(fields[i], "null weight pointer");
  PackedHeader h;
  layout_header(V, &h);
  if (bytes < h.total_bytes) {
    set_error("th_pack_weights: need %llu bytes, got %zu", (unsigned long long)h.total_bytes, bytes);
    return TH_EWORKSPACE;
  }
  unsigned char* blob = static_cast<unsigned char*>(packed_host);
  memset(blob, 0, h.total_bytes);
  memcpy(blob, &h, sizeof(h));
  auto W = [&](uint64_t off) { return reinterpret_cast<float*>(blob + off); };
  auto copy_cols = [](float* dst, int ldd, int col0, const float* src, int N, int K, double scale) {
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) dst[(size_t)n * ldd + col0 + k] = (float)((double)src[(size_t)n * K + k] * scale);
  };
  // plain layers
  copy_cols(W(h.fc0_w), 256, 0, w->fc_0_w, 256, 255, 1.0);
  memcpy(W(h.fc0_b), w->fc_0_b, 256 * 4);
  copy_cols(W(h.ar0_w), 384, 0, w->alpha_res_0_w, 256, 384, 1.0);
  memcpy(W(h.ar0_b), w->alpha_res_0_b, 256 * 4);
  copy_cols(W(h.k0_w), 256, 0, w->skv0_key_w, 128, 256, 1.0);
  memcpy(W(h.k0_b), w->skv0_key_b, 128 * 4);
  copy_cols(W(h.k1_w), 256, 0, w->skv1_key_w, 128, 256, 1.0);
  memcpy(W(h.k1_b), w->skv1_key_b, 128 * 4);
  // W_v = [value_embed_1 | value_embed_0] ; b = b1 + b0
  copy_cols(W(h.v_w), 512, 0, w->skv1_value_w, 256, 256, 1.0);
  copy_cols(W(h.v_w), 512, 256, w->skv0_value_w, 256, 256, 1.0);
  for (int n = 0; n < 256; ++n) W(h.v_b)[n] = (float)((double)w->skv1_value_b[n] + (double)w->skv0_value_b[n]);
  copy_cols(W(h.fc1_w), 256, 0, w->fc_1_w, 256, 256, 1.0);
  memcpy(W(h.fc1_b), w->fc_1_b, 256 * 4);
  copy_cols(W(h.fc2_w), 256, 0, w->fc_2_w, 256, 256, 1.0);
  memcpy(W(h.fc2_b), w->fc_2_b, 256 * 4);
  // mean over views folded into K
  for (int v = 0; v < V; ++v) copy_cols(W(h.fc3m_w), 256 * V, 256 * v, w->fc_3_w, 256, 256, 1.0 / V);
  memcpy(W(h.fc3m_b), w->fc_3_b, 256 * 4);
  memcpy(W(h.afc_w), w->alpha_fc_w, 256 * 4);
  W(h.afc_b)[0] = w->alpha_fc_b[0];
  // W_f = [feature_fc | rgb_res_0]
  copy_cols(W(h.f_w), 640, 0, w->feature_fc_w, 256, 256, 1.0);
  copy_cols(W(h.f_w), 640, 256, w->rgb_res_0_w, 256, 384, 1.0);
  for (int n = 0; n < 256; ++n) W(h.f_b)[n] = (float)((double)w->feature_fc_b[n] + (double)w->rgb_res_0_b[n]);
  // W_view: 283 -> 320 columns (view direction padded to 32, then to a whole 64-wide k-block)
  copy_cols(W(h.view_w), 320, 0, w->view_fc_w, 128, 283, 1.0);
  memcpy(W(h.view_b), w->view_fc_b, 128 * 4);
  // W_t = [fc_4/V x V | fc_4 @ rgb_res_1] ; b_t = fc_4 @ b_r1 + b_4
  {
    const int ld = 128 * V + 384;
    for (int v = 0; v < V; ++v) copy_cols(W(h.t_w), ld, 128 * v, w->fc_4_w, 128, 128, 1.0 / V);
    for (int n = 0; n < 128; ++n) {
      for (int k = 0; k < 384; ++k) {
        double s = 0.0;
        for (int j = 0; j < 128; ++j) s += (double)w->fc_4_w[n * 128 + j] * (double)w->rgb_res_1_w[j * 384 + k];
        W(h.t_w)[(size_t)n * ld + 128 * V + k] = (float)s;
      }
      double s = w->fc_4_b[n];
      for (int j = 0; j < 128; ++j) s += (double)w->fc_4_w[n * 128 + j] * (double)w->rgb_res_1_b[j];
      W(h.t_b)[n] = (float)s;
    }
  }
  // layers folded across a missing non-linearity (tensor-core schedule), products in float64
  {
    // N1 = relu(fc_1 (v1 S + v0 XT + b_v) + b_1): W_fc1f = [fc_1 @ v1 | fc_1 @ v0]
    for (int n = 0; n < 256; ++n) {
      for (int k = 0; k < 256; ++k) {
        double s1 = 0.0, s0 = 0.0;
        for (int j = 0; j < 256; ++j) {
          const double f = (double)w->fc_1_w[n * 256 + j];
          s1 += f * (double)w->skv1_value_w[j * 256 + k];
          s0 += f * (double)w->skv0_value_w[j * 256 + k];
        }
        W(h.fc1f_w)[(size_t)n * 512 + k] = (float)s1;
        W(h.fc1f_w)[(size_t)n * 512 + 256 + k] = (float)s0;
      }
      double s = w->fc_1_b[n];
      for (int j = 0; j < 256; ++j)
        s += (double)w->fc_1_w[n * 256 + j] * ((double)w->skv1_value_b[j] + (double)w->skv0_value_b[j]);
      W(h.fc1f_b)[n] = (float)s;
    }
    // G = relu(V1 (feature_fc INTER + rgb_res_0 pix + b_f) + V2 viewdir + b_view), view_fc = [V1 | V2]
    for (int n = 0; n < 128; ++n) {
      const float* v1 = w->view_fc_w + (size_t)n * 283;
      for (int k = 0; k < 256; ++k) {
        double s = 0.0;
        for (int j = 0; j < 256; ++j) s += (double)v1[j] * (double)w->feature_fc_w[j * 256 + k];
        W(h.gvf_w)[(size_t)n * 704 + k] = (float)s;
      }
      for (int k = 0; k < 384; ++k) {
        double s = 0.0;
        for (int j = 0; j < 256; ++j) s += (double)v1[j] * (double)w->rgb_res_0_w[j * 384 + k];
        W(h.gvf_w)[(size_t)n * 704 + 256 + k] = (float)s;
      }
      for (int k = 0; k < 27; ++k) W(h.gvf_w)[(size_t)n * 704 + 640 + k] = v1[256 + k];
      double s = w->view_fc_b[n];
      for (int j = 0; j < 256; ++j) s += (double)v1[j] * ((double)w->feature_fc_b[j] + (double)w->rgb_res_0_b[j]);
      W(h.gvf_b)[n] = (float)s;
    }
  }
  memcpy(W(h.rgb_w), w->rgb_fc_w, 3 * 128 * 4);
  memcpy(W(h.rgb_b), w->rgb_fc_b, 3 * 4);
  // pre-mapped feature maps (kernels.cuh): built from the folded matrices above
  {
    const int ldt = 128 * V + 384, ldp = 128 * V + 128;
    memcpy(W(h.pre_w), W(h.ar0_w), (size_t)256 * 384 * 4);
    memcpy(W(h.pre_b), W(h.ar0_b), 256 * 4);  // rows 256..511 keep a zero bias
    for (int n = 0; n < 128; ++n) {
      for (int k = 0; k < 384; ++k) {
        W(h.pre_w)[(size_t)(256 + n) * 384 + k] = W(h.gvf_w)[(size_t)n * 704 + 256 + k];
        W(h.pre_w)[(size_t)(384 + n) * 384 + k] = (float)((double)W(h.t_w)[(size_t)n * ldt + 128 * V + k] / V);
      }
      for (int k = 0; k < 256; ++k) W(h.gvfp_w)[(size_t)n * 448 + k] = W(h.gvf_w)[(size_t)n * 704 + k];
      W(h.gvfp_w)[(size_t)n * 448 + 256 + n] = 1.0f;
      for (int k = 0; k < 64; ++k) W(h.gvfp_w)[(size_t)n * 448 + 384 + k] = W(h.gvf_w)[(size_t)n * 704 + 640 + k];
      W(h.gvfp_b)[n] = W(h.gvf_b)[n];
      for (int k = 0; k < 128 * V; ++k) W(h.tp_w)[(size_t)n * ldp + k] = W(h.t_w)[(size_t)n * ldt + k];
      W(h.tp_w)[(size_t)n * ldp + 128 * V + n] = 1.0f;
      W(h.tp_b)[n] = W(h.t_b)[n];
    }
    for (int n = 0; n < 256; ++n) W(h.xid_w)[(size_t)n * 256 + n] = 1.0f;
    // rows 256..511 of W_pre as a matrix of their own (the second launch of the pre-map GEMM; zero bias)
    memcpy(W(h.preb_w), W(h.pre_w) + (size_t)256 * 384, (size_t)256 * 384 * 4);
    // fc_1' cut into S / X parts and halves of 128 output rows (PackedHeader)
    float* const fs[2] = {W(h.fc1s0_w), W(h.fc1s1_w)};
    float* const fx[2] = {W(h.fc1x0_w), W(h.fc1x1_w)};
    float* const fb[2] = {W(h.fc1s0_b), W(h.fc1s1_b)};
    for (int hh = 0; hh < 2; ++hh)
      for (int n = 0; n < 128; ++n) {
        memcpy(fs[hh] + (size_t)n * 256, W(h.fc1f_w) + (size_t)(128 * hh + n) * 512, 256 * 4);
        memcpy(fx[hh] + (size_t)n * 256, W(h.fc1f_w) + (size_t)(128 * hh + n) * 512 + 256, 256 * 4);
        fb[hh][n] = W(h.fc1f_b)[128 * hh + n];
      }
  }
  // fp16 hi/lo split of the GEMM matrices, stored as shared-memory tile images for
  // the tensor-core path: per 64-wide k-block and per half of the N rows (one half per CTA of a
  // cta_group::2 pair) [hi image | lo image], each N/2 rows of 128 bytes, K-major with the 128-byte
  // swizzle (16-byte chunk ^= row & 7) -- a CTA's whole share of a k-block is ONE contiguous copy.
  for (const MatSpec& m : mat_specs(V)) {
    if (!m.h) continue;
    const float* src = W(h.*(m.w));
    unsigned char* img = blob + h.*(m.h);
    // power-of-two scale (PackedHeader::img_inv_scale): max|w| 2^e in (2^13, 2^14]
    const int e = weight_scale_exp(src, (size_t)m.N * m.K);
    const float up = ldexpf(1.0f, e);
    PackedHeader* hb = reinterpret_cast<PackedHeader*>(blob);
    if (hb->n_img < 24) {
      hb->img_off[hb->n_img] = h.*(m.h);
      hb->img_inv_scale[hb->n_img] = ldexpf(1.0f, -e);
      ++hb->n_img;
    }
    pack_weight_image(src, m.N, m.K, m.N, up, img);
  }
  return TH_OK;
}

// ---------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------
struct Workspace {
  unsigned char* grid;
  unsigned long long* counters;  // 8 slots
  uint8_t* ray_any;
  uint8_t* mask;
  int32_t* ids;
  float* raw;
  float* chunk;
  int64_t chunk_pts;
  unsigned char* tok_grid;  // token grid of the K-NN (culled rays / density grid with >= TOKEN_GRID_MIN tokens)
  int compact_sms;          // > 0: compact chunk block of the layer-chained schedule on pre-mapped maps (mlp_carve), scratch sized for this SM count
};

// The schedule a frame's flags select keeps either the per-point activation buffers of the layer-at-a-time
// schedule (26.9 KB per point of the chunk) or -- pre-mapped maps, hence the layer-chained kernel: the default
// path -- only the feature kernel's operand images and a fixed per-CTA scratch (8.25 KB per point + 138 MB).
static int compact_sms_for(const ThFrame* f) {
  if (!f || !(f->flags & TH_FLAG_PREMAPPED) || (f->flags & (TH_FLAG_SIMT_MLP | TH_FLAG_LAYERWISE)) ||
      !chain_supported(f->n_views))
    return 0;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n < 2) {
    cudaGetLastError();  // no device (host-side sizing): size for a B200
    n = 148;
  }
  return n;
}

constexpr int TOKEN_GRID_MAX = 65536;  // tokens the reserved grid block can hold (th_workspace_bytes has no token count)
// TH_TOKEN_GRID=<n>: use the token grid from n tokens on (default TOKEN_GRID_MIN; a huge n = never) -- measurement knob
static int token_grid_min() {
  static const int v = getenv("TH_TOKEN_GRID") ? atoi(getenv("TH_TOKEN_GRID")) : TOKEN_GRID_MIN;
  return v;
}
static size_t ws_plan(int64_t n_points, int V, int n_verts, unsigned char* base, Workspace* ws, int compact_sms = 0) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  int64_t np = n_points > 0 ? n_points : 0;
  unsigned char* grid = take(n_verts > 0 ? cull_grid_bytes(n_verts) : 0);
  unsigned char* counters = take(64);
  unsigned char* ray_any = take((size_t)np);
  unsigned char* mask = take((size_t)np);
  unsigned char* ids = take((size_t)np * 4);
  unsigned char* raw = take((size_t)np * 16);
  int64_t cp = np < chunk_pts() ? np : chunk_pts();
  unsigned char* chunk = take((compact_sms > 0 ? mlp_compact_bytes(cp, V, chain_scratch_bytes(cp, V, compact_sms))
                                               : (size_t)pad_points(cp) * mlp_buffer_floats_per_point(V) * 4) + 1024);
  unsigned char* tok_grid = take(n_verts > 0 ? cull_grid_bytes(TOKEN_GRID_MAX) : 0);
  if (ws) {
    ws->tok_grid = n_verts > 0 ? tok_grid : nullptr;
    ws->grid = grid;
    ws->counters = reinterpret_cast<unsigned long long*>(counters);
    ws->ray_any = ray_any;
    ws->mask = mask;
    ws->ids = reinterpret_cast<int32_t*>(ids);
    ws->raw = reinterpret_cast<float*>(raw);
    ws->chunk = reinterpret_cast<float*>(chunk);
    ws->chunk_pts = cp;
    ws->compact_sms = compact_sms;
  }
  return off;
}

size_t th_workspace_bytes(int64_t n_points, int32_t n_views, int32_t n_verts) {
  if (n_views < 1) n_views = 1;
  return ws_plan(n_points, n_views, n_verts, nullptr, nullptr);
}
size_t th_frame_workspace_bytes(const ThFrame* f, int64_t n_points, int32_t with_cull) {
  if (!f) return 0;
  const int V = f->n_views < 1 ? 1 : f->n_views;
  return ws_plan(n_points, V, with_cull ? f->n_verts : 0, nullptr, nullptr, compact_sms_for(f));
}

static int frame_dev(const ThFrame* f, FrameDev* d, bool need_tokens, bool need_feat) {
  TH_CHECK_ARG(f, "null frame");
  TH_CHECK_ARG(f->n_views >= 1 && f->n_views <= TH_MAX_VIEWS, "n_views out of range");
  if (need_tokens) {
    TH_CHECK_ARG(f->tok_feat && f->tok_xyz && f->tok_rot, "null token pointer");
    TH_CHECK_ARG(f->knn >= 1 && f->knn <= TH_MAX_KNN && f->knn <= f->n_tok, "knn out of range");
    TH_CHECK_ARG(f->knn_dist_alpha > 0.f, "knn_dist_alpha must be positive");
  }
  if (need_feat) {
    TH_CHECK_ARG(f->feat && f->cam_R && f->cam_T && f->cam_K, "null feature-map / camera pointer");
    TH_CHECK_ARG(f->feat_h >= 1 && f->feat_w >= 1, "bad feature-map size");
  }
  d->tok_feat = f->tok_feat;
  d->tok_xyz = f->tok_xyz;
  d->tok_rot = f->tok_rot;
  d->feat = f->feat;
  d->cam_R = f->cam_R;
  d->cam_T = f->cam_T;
  d->cam_K = f->cam_K;
  d->Rh = f->Rh;
  d->Th = f->Th;
  d->V = f->n_views;
  d->n_tok = f->n_tok;
  d->H = f->feat_h;
  d->W = f->feat_w;
  d->K = f->knn;
  d->sx = f->uv_scale_x;
  d->sy = f->uv_scale_y;
  d->knn_alpha = f->knn_dist_alpha;
  d->tok_grid = nullptr;
  return TH_OK;
}

static int read_header(const ThFrame* f, PackedHeader* h, cudaStream_t st) {
  TH_CHECK_ARG(f->weights, "null weights blob");
  TH_CUDA(cudaMemcpyAsync(h, f->weights, sizeof(PackedHeader), cudaMemcpyDeviceToHost, st));
  TH_CUDA(cudaStreamSynchronize(st));
  if (h->magic != PACK_MAGIC || h->n_views != f->n_views) {
    set_error("weights blob: bad magic or packed for %d views (frame has %d)", h->n_views, f->n_views);
    return TH_EINVAL;
  }
  return TH_OK;
}

// host copy of the header, cached per blob pointer (the header never changes
// for a given upload; avoids a device round trip per call)
static int cached_header(const ThFrame* f, PackedHeader* h, cudaStream_t st) {
  static thread_local const void* key = nullptr;
  static thread_local PackedHeader cache;
  if (key != f->weights || cache.n_views != f->n_views) {
    int rc = read_header(f, &cache, st);
    if (rc) {
      key = nullptr;
      return rc;
    }
    key = f->weights;
  }
  *h = cache;
  return TH_OK;
}

// feature kernel + MLP over a list of points, in chunks
// the layer-chained kernel runs this frame's network (mlp_chain.cu) -- otherwise the layer-at-a-time schedules
static bool uses_chain(const ThFrame* f) {
  const char* e = getenv("TH_CHAIN");
  const int use_chain = e ? atoi(e) : 1;
  const bool premapped = (f->flags & TH_FLAG_PREMAPPED) != 0;
  return !(f->flags & (TH_FLAG_SIMT_MLP | TH_FLAG_LAYERWISE)) && (use_chain || premapped) && chain_supported(f->n_views);
}

static int run_points(const ThFrame* f, const FrameDev& fr, const PackedHeader& hdr, PointSource src,
                      const int32_t* ids, int64_t n_list, const Workspace& ws, float* raw, float* alpha_out,
                      int alpha_only, int zero_rgb, cudaStream_t st, const CompositeArgs* cmp = nullptr) {
  const int V = fr.V;
  for (int64_t first = 0; first < n_list; first += ws.chunk_pts) {
    int64_t P = n_list - first < ws.chunk_pts ? n_list - first : ws.chunk_pts;
    MlpBuffers b;
    mlp_carve(ws.chunk, P, V, &b, ws.compact_sms > 0 ? chain_scratch_bytes(P, V, ws.compact_sms) : 0);
    const int64_t Pp = pad_points(P);
    FeatOut fo{};
    fo.rep = b.rep;
    fo.rep_sv = Pp * REP_LD;
    fo.rep_sp = REP_LD;
    fo.rep_sc = 1;
    fo.pix = b.pix;
    fo.pix_sv = Pp * PIX_LD;
    fo.pix_sp = PIX_LD;
    fo.pix_sc = 1;
    fo.pix_mean = alpha_only ? nullptr : b.pix_mean;
    fo.vd = b.vd;
    fo.do_rep = 1;
    fo.do_pix = 1;
    fo.do_vd = alpha_only ? 0 : 1;
    fo.rep_pad = 1;
    const int use_tc = (f->flags & TH_FLAG_SIMT_MLP) ? 0 : 1;
    const bool premapped = (f->flags & TH_FLAG_PREMAPPED) != 0;
    if (premapped && !(use_tc && !(f->flags & TH_FLAG_LAYERWISE) && chain_supported(V))) {
      set_error("TH_FLAG_PREMAPPED needs the layer-chained tensor-core schedule (V <= 3, no SIMT/LAYERWISE flag)");
      return TH_EUNSUPPORTED;
    }
    if (use_tc) {  // the feature kernel writes the GEMM operands directly as fp16 hi/lo tile images
      fo.rep_img = reinterpret_cast<unsigned char*>(b.rep);
      fo.pix_img = reinterpret_cast<unsigned char*>(b.pix);
      fo.pixm_img = alpha_only ? nullptr : reinterpret_cast<unsigned char*>(b.pix_mean);
      fo.vd_img = reinterpret_cast<unsigned char*>(b.vd);
      fo.img_view_rows = Pp;
      fo.pix_mean = nullptr;
      if (premapped)  // pix block = [X (V*Pp,256) | P2 (V*Pp,128)], pix_mean block = R (Pp,128): see k_features PRE
        fo.pix = alpha_only ? nullptr : b.pix + (size_t)V * Pp * 256;
    }
    src.ids = ids;
    src.first = first;
    int rc = launch_features(fr, src, P, fo, st, premapped);
    if (rc) return rc;
    MlpRun run{};
    run.weights = static_cast<const unsigned char*>(f->weights);
    run.P = P;
    run.V = V;
    run.dst_ids = ids;
    run.first = first;
    run.raw = raw;
    run.alpha_out = alpha_out;
    run.alpha_only = alpha_only;
    run.zero_rgb_if_transparent = zero_rgb;
    run.use_tensor_cores = use_tc;
    run.inputs_are_images = use_tc;
    run.premapped = premapped ? 1 : 0;
    if (cmp) run.cmp = *cmp;
    // one layer-chained launch per chunk (mlp_chain.cu) unless TH_CHAIN=0 asks for the
    // layer-at-a-time schedule; its scratch is the (otherwise unused) S/X/XT/NET/KP/KS block
    int num_sms = 0;
    if (device_sm_count(&num_sms)) return TH_ECUDA;
    const size_t scratch_room = ws.compact_sms > 0 ? chain_scratch_bytes(P, V, ws.compact_sms)  // sized for it
                                                   : (size_t)Pp * V * (256 * 4 + 128 * 2) * 4;  // b.s ... b.ks are contiguous
    if (uses_chain(f) && chain_scratch_bytes(P, V, num_sms) <= scratch_room) {
      rc = mlp_forward_chain(run, b, hdr, reinterpret_cast<unsigned char*>(b.s), nullptr, st);
    } else if (cmp) {
      set_error("fused compositing needs the layer-chained schedule");
      rc = TH_EUNSUPPORTED;
    } else if (premapped) {
      set_error("TH_FLAG_PREMAPPED: the layer-chained schedule is not available for this chunk");
      rc = TH_EUNSUPPORTED;
    } else {
      rc = mlp_forward(run, b, hdr, st);
    }
    if (rc) return rc;
  }
  return TH_OK;
}

int th_render_rays(const ThFrame* f, const ThRays* r, ThOut* o, int32_t culled, void* workspace,
                   size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(f && r && o, "null argument");
  TH_CHECK_ARG(r->n_rays >= 0 && r->n_samples >= 1, "bad ray counts");
  TH_CHECK_ARG(culled >= 0 && culled <= 2, "bad render mode");
  if (r->n_rays == 0) {  // empty bundle: nothing to read or write
    if (o->counters_host) o->counters_host[0] = o->counters_host[1] = o->counters_host[2] = 0;
    return TH_OK;
  }
  TH_CHECK_ARG(r->ray_o && r->ray_d && r->near_ && r->far_ && r->t_vals, "null ray pointer");
  TH_CHECK_ARG(o->rgb_map && o->acc_map && o->depth_map, "null output pointer");
  const int64_t N = r->n_rays;
  const int S = r->n_samples;
  const int64_t NP = N * S;
  TH_CHECK_ARG(NP < (1LL << 31), "more than 2^31 sample points in one call");
  if (o->counters_host) o->counters_host[0] = o->counters_host[1] = o->counters_host[2] = 0;
  if (N == 0) return TH_OK;
  FrameDev fr;
  int rc = frame_dev(f, &fr, true, true);
  if (rc) return rc;
  TH_CHECK_ARG(f->Rh && f->Th, "null Rh/Th");
  if (culled) TH_CHECK_ARG(f->verts && f->n_verts > 0, "culled mode needs the cull vertices");
  Workspace ws;
  size_t need = ws_plan(NP, fr.V, culled ? f->n_verts : 0, static_cast<unsigned char*>(workspace), &ws, compact_sms_for(f));
  if (!workspace || workspace_bytes < need) {
    set_error("th_render_rays: workspace %zu < %zu bytes", workspace_bytes, need);
    return TH_EWORKSPACE;
  }
  PackedHeader hdr;
  if ((rc = cached_header(f, &hdr, st))) return rc;

  PointSource src{};
  src.ray_o = r->ray_o;
  src.ray_d = r->ray_d;
  src.near_ = r->near_;
  src.far_ = r->far_;
  src.t_vals = r->t_vals;
  src.n_samples = S;
  float* raw = o->raw ? o->raw : ws.raw;
  const uint8_t* mask = nullptr;
  const int white = (f->flags & TH_FLAG_WHITE_BKGD) ? 1 : 0;

  if (!culled) {
    // Dense rays on the chain schedule with S | 128: the rows of a 128-point tile are whole rays, so the chain kernel
    // composites them in its fc_4' epilogue (CompositeArgs) -- no raw tensor unless the caller asks for it, no
    // k_integrate.  TH_FUSE_INTEGRATE=0 keeps the two-kernel form (tests compare the two bit for bit).
    const char* fuse_env = getenv("TH_FUSE_INTEGRATE");
    const bool fuse = uses_chain(f) && S <= 128 && 128 % S == 0 && ws.chunk_pts % S == 0 && !(fuse_env && !atoi(fuse_env));
    CompositeArgs cmp{};
    if (fuse) {
      cmp.rgb_map = o->rgb_map;
      cmp.acc_map = o->acc_map;
      cmp.depth_map = o->depth_map;
      cmp.near_ = r->near_;
      cmp.far_ = r->far_;
      cmp.t_vals = r->t_vals;
      cmp.ray_d = r->ray_d;
      cmp.S = S;
      cmp.white_bkgd = white;
    }
    if ((rc = run_points(f, fr, hdr, src, nullptr, NP, ws, fuse ? o->raw : raw, nullptr, 0, 0, st, fuse ? &cmp : nullptr)))
      return rc;
    if (o->counters_host) {
      o->counters_host[0] = NP;
      o->counters_host[1] = N;
      o->counters_host[2] = NP;
    }
    if (fuse) return TH_OK;
  } else {
    uint8_t* m = o->pts_mask ? o->pts_mask : ws.mask;
    TH_CUDA(cudaMemsetAsync(ws.counters, 0, 64, st));
    TH_CUDA(cudaMemsetAsync(ws.ray_any, 0, (size_t)N, st));
    if ((rc = launch_grid_build(f->verts, f->n_verts, f->cull_radius, ws.grid, st))) return rc;
    // candidate list of the two-pass cull: the workspace's raw block is idle until the network runs
    if ((rc = launch_cull_grid(src, NP, ws.grid, f->cull_radius, m, ws.ids, ws.ray_any, ws.counters, st,
                               reinterpret_cast<int32_t*>(ws.raw))))
      return rc;
    if ((rc = launch_count_nonzero(ws.ray_any, N, ws.counters + 1, st))) return rc;
    unsigned long long cnt[3] = {0, 0, 0};
    TH_CUDA(cudaMemcpyAsync(cnt, ws.counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    TH_CUDA(cudaStreamSynchronize(st));
    int64_t n_eval = (int64_t)cnt[0];
    int zero_rgb = 1;
    const uint8_t* integ_mask = m;
    if (culled == TH_RENDER_FAST && (int64_t)cnt[1] <= TH_TRAIN_BRANCH_MAX_RAYS && cnt[1] > 0) {
      // reference quirk: <= 2400 surviving rays -> un-chunked branch without pts_mask
      if ((rc = launch_expand_rays(ws.ray_any, NP, S, ws.mask, ws.ids, ws.counters + 2, st))) return rc;
      TH_CUDA(cudaMemcpyAsync(cnt + 2, ws.counters + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      TH_CUDA(cudaStreamSynchronize(st));
      n_eval = (int64_t)cnt[2];
      zero_rgb = 0;
      integ_mask = ws.mask;
    }
    if (o->counters_host) {
      o->counters_host[0] = (int64_t)cnt[0];
      o->counters_host[1] = (int64_t)cnt[1];
      o->counters_host[2] = n_eval;
    }
    if (o->raw) TH_CUDA(cudaMemsetAsync(o->raw, 0, (size_t)NP * 16, st));
    // every evaluated point lies within the cull radius of the body: its K nearest tokens come from the token grid
    if (ws.tok_grid && fr.n_tok >= token_grid_min() && fr.n_tok <= TOKEN_GRID_MAX && n_eval > 0) {
      if ((rc = launch_token_grid(fr.tok_xyz, fr.n_tok, ws.tok_grid, st))) return rc;
      fr.tok_grid = reinterpret_cast<const CullGrid*>(ws.tok_grid);
    }
    if ((rc = run_points(f, fr, hdr, src, ws.ids, n_eval, ws, raw, nullptr, 0, zero_rgb, st))) return rc;
    mask = integ_mask;
  }
  return launch_integrate(raw, mask, culled ? ws.ray_any : nullptr, src, nullptr, r->ray_d, N, S, white, o->rgb_map,
                          o->acc_map, o->depth_map, st);
}

int th_query_density(const ThFrame* f, const float* pts, int64_t n_points, float* alpha_raw, uint8_t* mask_out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(f && pts && alpha_raw, "null argument");
  TH_CHECK_ARG(n_points >= 0 && n_points < (1LL << 31), "bad point count");
  if (n_points == 0) return TH_OK;
  FrameDev fr;
  int rc = frame_dev(f, &fr, true, true);
  if (rc) return rc;
  TH_CHECK_ARG(f->Rh && f->Th && f->verts && f->n_verts > 0, "null Rh/Th/verts");
  Workspace ws;
  size_t need = ws_plan(n_points, fr.V, f->n_verts, static_cast<unsigned char*>(workspace), &ws, compact_sms_for(f));
  if (!workspace || workspace_bytes < need) {
    set_error("th_query_density: workspace %zu < %zu bytes", workspace_bytes, need);
    return TH_EWORKSPACE;
  }
  PackedHeader hdr;
  if ((rc = cached_header(f, &hdr, st))) return rc;
  PointSource src{};
  src.pts = pts;
  src.n_samples = 1;
  uint8_t* m = mask_out ? mask_out : ws.mask;
  TH_CUDA(cudaMemsetAsync(ws.counters, 0, 64, st));
  TH_CUDA(cudaMemsetAsync(alpha_raw, 0, (size_t)n_points * 4, st));
  if ((rc = launch_grid_build(f->verts, f->n_verts, f->cull_radius, ws.grid, st))) return rc;
  if ((rc = launch_cull_grid(src, n_points, ws.grid, f->cull_radius, m, ws.ids, nullptr, ws.counters, st,
                             reinterpret_cast<int32_t*>(ws.raw))))
    return rc;
  unsigned long long cnt = 0;
  TH_CUDA(cudaMemcpyAsync(&cnt, ws.counters, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  TH_CUDA(cudaStreamSynchronize(st));
  if (ws.tok_grid && fr.n_tok >= token_grid_min() && fr.n_tok <= TOKEN_GRID_MAX && cnt > 0) {
    if ((rc = launch_token_grid(fr.tok_xyz, fr.n_tok, ws.tok_grid, st))) return rc;
    fr.tok_grid = reinterpret_cast<const CullGrid*>(ws.tok_grid);
  }
  return run_points(f, fr, hdr, src, ws.ids, (int64_t)cnt, ws, nullptr, alpha_raw, 1, 0, st);
}

// ---------------------------------------------------------------------------
// staged entry points
// ---------------------------------------------------------------------------
int th_sample_points(const ThRays* r, float* pts, float* z_vals, void* stream) {
  TH_CHECK_ARG(r && r->ray_o && r->ray_d && r->near_ && r->far_ && r->t_vals, "null ray pointer");
  TH_CHECK_ARG(r->n_rays >= 0 && r->n_samples >= 1, "bad ray counts");
  PointSource src{};
  src.ray_o = r->ray_o;
  src.ray_d = r->ray_d;
  src.near_ = r->near_;
  src.far_ = r->far_;
  src.t_vals = r->t_vals;
  src.n_samples = r->n_samples;
  return launch_sample_points(src, r->n_rays * r->n_samples, pts, z_vals, static_cast<cudaStream_t>(stream));
}

int th_cull_knn1(const float* pts, int64_t n_points, const float* verts, int32_t n_verts, float radius, float* d2,
                 int64_t* idx, uint8_t* mask, void* stream) {
  TH_CHECK_ARG(pts && verts, "null pointer");
  TH_CHECK_ARG(n_points >= 0 && n_verts >= 1, "bad counts");
  PointSource src{};
  src.pts = pts;
  src.n_samples = 1;
  return launch_cull_brute(src, n_points, verts, n_verts, radius, d2, idx, mask, static_cast<cudaStream_t>(stream));
}

int th_cull_grid(const float* pts, int64_t n_points, const float* verts, int32_t n_verts, float radius,
                 uint8_t* mask, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(pts && verts && mask, "null pointer");
  TH_CHECK_ARG(n_points >= 0 && n_verts >= 1, "bad counts");
  if (!workspace || workspace_bytes < cull_grid_bytes(n_verts)) {
    set_error("th_cull_grid: workspace %zu < %zu bytes", workspace_bytes, cull_grid_bytes(n_verts));
    return TH_EWORKSPACE;
  }
  PointSource src{};
  src.pts = pts;
  src.n_samples = 1;
  int rc = launch_grid_build(verts, n_verts, radius, workspace, st);
  if (rc) return rc;
  return launch_cull_grid(src, n_points, workspace, radius, mask, nullptr, nullptr, nullptr, st);
}

int th_world2smpl(const float* pts, int64_t n_points, const float* Rh, const float* Th, float* out, void* stream) {
  TH_CHECK_ARG(pts && Rh && Th && out, "null pointer");
  return launch_world2smpl(pts, n_points, Rh, Th, out, static_cast<cudaStream_t>(stream));
}

int th_view_embed(const float* ray_d, int64_t n_rays, float* out, void* stream) {
  TH_CHECK_ARG(ray_d && out, "null pointer");
  return launch_view_embed(ray_d, n_rays, out, static_cast<cudaStream_t>(stream));
}

int th_pixel_gather(const ThFrame* f, const float* pts, int64_t n_points, float* pixel_feat, void* stream) {
  TH_CHECK_ARG(pts && pixel_feat, "null pointer");
  FrameDev fr;
  int rc = frame_dev(f, &fr, false, true);
  if (rc) return rc;
  fr.n_tok = 0;
  PointSource src{};
  src.pts = pts;
  src.n_samples = 1;
  FeatOut fo{};
  fo.pix = pixel_feat;
  fo.pix_sv = (int64_t)TH_C_PIX * n_points;
  fo.pix_sp = 1;
  fo.pix_sc = n_points;
  fo.do_pix = 1;
  return launch_features(fr, src, n_points, fo, static_cast<cudaStream_t>(stream));
}

size_t th_knn_workspace_bytes(int32_t n_tok) { return cull_grid_bytes(n_tok > 0 ? n_tok : 1); }

int th_knn_dparf(const ThFrame* f, const float* pts_smpl, int64_t n_points, int64_t* knn_idx, float* knn_d2,
                 float* human_rep, void* workspace, size_t workspace_bytes, void* stream) {
  TH_CHECK_ARG(pts_smpl, "null pointer");
  FrameDev fr;
  int rc = frame_dev(f, &fr, true, false);
  if (rc) return rc;
  // with a workspace the K-NN goes through the token grid (what the fused path does for culled rays / grid points)
  if (workspace && fr.n_tok <= TOKEN_GRID_MAX) {
    TH_CHECK_ARG(workspace_bytes >= cull_grid_bytes(fr.n_tok) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                 "token-grid workspace too small or misaligned");
    if ((rc = launch_token_grid(fr.tok_xyz, fr.n_tok, workspace, static_cast<cudaStream_t>(stream)))) return rc;
    fr.tok_grid = static_cast<const CullGrid*>(workspace);
  }
  PointSource src{};
  src.pts = pts_smpl;
  src.n_samples = 1;
  FeatOut fo{};
  fo.rep = human_rep;
  fo.rep_sv = (int64_t)TH_C_REP * n_points;
  fo.rep_sp = 1;
  fo.rep_sc = n_points;
  fo.knn_idx = knn_idx;
  fo.knn_d2 = knn_d2;
  fo.do_rep = 1;
  fo.pts_are_smpl = 1;
  return launch_features(fr, src, n_points, fo, static_cast<cudaStream_t>(stream));
}

int th_mlp_raw(const ThFrame* f, const float* human_rep, const float* pixel_feat, const float* viewdir,
               const uint8_t* pts_mask, int64_t n_points, float* raw, void* workspace, size_t workspace_bytes,
               void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(f && human_rep && pixel_feat && viewdir && raw, "null pointer");
  TH_CHECK_ARG(f->n_views >= 1 && f->n_views <= TH_MAX_VIEWS, "n_views out of range");
  if (n_points <= 0) return TH_OK;
  const int V = f->n_views;
  size_t need = align_up((size_t)pad_points(n_points) * mlp_buffer_floats_per_point(V) * 4 + 1024, 256);
  if (!workspace || workspace_bytes < need) {
    set_error("th_mlp_raw: workspace %zu < %zu bytes", workspace_bytes, need);
    return TH_EWORKSPACE;
  }
  PackedHeader hdr;
  int rc = cached_header(f, &hdr, st);
  if (rc) return rc;
  MlpBuffers b;
  mlp_carve(static_cast<float*>(workspace), n_points, V, &b);
  if ((rc = launch_pack_inputs(human_rep, pixel_feat, viewdir, n_points, V, b, st))) return rc;
  MlpRun run{};
  run.weights = static_cast<const unsigned char*>(f->weights);
  run.P = n_points;
  run.V = V;
  run.raw = raw;
  run.zero_rgb_if_transparent = pts_mask ? 1 : 0;
  run.use_tensor_cores = (f->flags & TH_FLAG_SIMT_MLP) ? 0 : 1;
  if ((rc = mlp_forward(run, b, hdr, st))) return rc;
  if (pts_mask) {
    if ((rc = launch_apply_mask(raw, pts_mask, n_points, st))) return rc;
  }
  return TH_OK;
}

int th_integrate(const float* raw, const float* z_vals, const float* ray_d, int64_t n_rays, int32_t n_samples,
                 int32_t white_bkgd, float* rgb_map, float* acc_map, float* depth_map, void* stream) {
  TH_CHECK_ARG(raw && z_vals && ray_d && rgb_map && acc_map && depth_map, "null pointer");
  TH_CHECK_ARG(n_rays >= 0 && n_samples >= 1, "bad counts");
  PointSource src{};
  return launch_integrate(raw, nullptr, nullptr, src, z_vals, ray_d, n_rays, n_samples, white_bkgd, rgb_map, acc_map,
                          depth_map, static_cast<cudaStream_t>(stream));
}

int th_premap_features(const float* feat_nchw, const void* packed_weights, int32_t n_views, int32_t h, int32_t w,
                       float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(feat_nchw && packed_weights && out, "null pointer");
  TH_CHECK_ARG(n_views >= 1 && n_views <= TH_MAX_VIEWS && h >= 1 && w >= 1, "bad argument");
  PackedHeader hdr;
  TH_CUDA(cudaMemcpyAsync(&hdr, packed_weights, sizeof(PackedHeader), cudaMemcpyDeviceToHost, st));
  TH_CUDA(cudaStreamSynchronize(st));
  if (hdr.magic != PACK_MAGIC) {
    set_error("th_premap_features: bad weights blob");
    return TH_EINVAL;
  }
  return launch_premap(feat_nchw, static_cast<const unsigned char*>(packed_weights), hdr, out, n_views, h, w, st);
}

int th_paint_group(const float* holder_map, int32_t n_views, int32_t h, int32_t w, float uv_scale_x, float uv_scale_y,
                   const float* verts, int32_t n_verts, const float* cam_R, const float* cam_T, const float* cam_K,
                   const uint8_t* vizmap, const int32_t* cluster_start, const int32_t* cluster_members, int32_t n_tok,
                   float* painted, float* tokens, void* stream) {
  TH_CHECK_ARG(holder_map && verts && cam_R && cam_T && cam_K && cluster_start && cluster_members && tokens,
               "null pointer");
  TH_CHECK_ARG(n_views >= 1 && n_views <= 65535 && h >= 1 && w >= 1 && n_verts >= 1 && n_tok >= 1, "bad sizes");
  return launch_paint_group(holder_map, n_views, h, w, uv_scale_x, uv_scale_y, verts, cam_R, cam_T, cam_K, vizmap,
                            n_verts, cluster_start, cluster_members, n_tok, painted, tokens,
                            static_cast<cudaStream_t>(stream));
}

// one EncTail per view from the C-ABI struct
static int enc_views(const ThEncoderTail* enc, EncTail* ev) {
  TH_CHECK_ARG(enc && enc->images && enc->color_w && enc->color_b, "null pointer");
  TH_CHECK_ARG(enc->n_views >= 1 && enc->n_views <= TH_MAX_VIEWS && enc->h >= 1 && enc->w >= 1, "bad sizes");
  static const int chans[3] = {64, 64, 128};
  for (int i = 0; i < 3; ++i)
    TH_CHECK_ARG(enc->latent[i] && enc->lat_h[i] >= 1 && enc->lat_w[i] >= 1, "bad latent");
  for (int v = 0; v < enc->n_views; ++v) {
    EncTail& e = ev[v];
    for (int i = 0; i < 3; ++i) {
      e.lat[i] = enc->latent[i] + (int64_t)v * chans[i] * enc->lat_h[i] * enc->lat_w[i];
      e.lh[i] = enc->lat_h[i];
      e.lw[i] = enc->lat_w[i];
    }
    e.img = enc->images + (int64_t)v * 3 * enc->h * enc->w;
    e.wc = enc->color_w;
    e.bc = enc->color_b;
    e.H = enc->h;
    e.W = enc->w;
  }
  return TH_OK;
}

size_t th_premap_from_latents_workspace_bytes(const ThEncoderTail* enc) {
  if (!enc) return 0;
  int lh[3], lw[3];
  for (int i = 0; i < 3; ++i) lh[i] = enc->lat_h[i] > 0 ? enc->lat_h[i] : 1, lw[i] = enc->lat_w[i] > 0 ? enc->lat_w[i] : 1;
  return premap_latents_scratch_bytes(lh, lw);
}

int th_premap_from_latents(const ThEncoderTail* enc, const void* packed_weights, float* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(packed_weights && out && workspace, "null pointer");
  EncTail ev[TH_MAX_VIEWS];
  int rc = enc_views(enc, ev);
  if (rc) return rc;
  TH_CHECK_ARG(workspace_bytes >= th_premap_from_latents_workspace_bytes(enc), "workspace too small");
  TH_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  PackedHeader hdr;
  TH_CUDA(cudaMemcpyAsync(&hdr, packed_weights, sizeof(PackedHeader), cudaMemcpyDeviceToHost, st));
  TH_CUDA(cudaStreamSynchronize(st));
  if (hdr.magic != PACK_MAGIC) {
    set_error("th_premap_from_latents: bad weights blob");
    return TH_EINVAL;
  }
  return launch_premap_latents(ev, static_cast<const unsigned char*>(packed_weights), hdr, out, enc->n_views, workspace,
                               st);
}

size_t th_paint_group_latents_workspace_bytes(int32_t n_views, int32_t n_verts, int32_t n_tok) {
  const int v = n_views > 0 ? n_views : 1;
  return align_up((size_t)v * sizeof(EncTail), 256) +
         paint_latents_scratch_bytes(v, n_verts > 0 ? n_verts : 1, n_tok > 0 ? n_tok : 1);
}

int th_paint_group_latents(const ThEncoderTail* enc, const float* reduction_w, const float* reduction_b,
                           float uv_scale_x, float uv_scale_y, const float* verts, int32_t n_verts, const float* cam_R,
                           const float* cam_T, const float* cam_K, const uint8_t* vizmap, const int32_t* cluster_start,
                           const int32_t* cluster_members, int32_t n_tok, float* tokens, void* workspace,
                           size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(reduction_w && reduction_b && verts && cam_R && cam_T && cam_K && cluster_start && cluster_members &&
                   tokens && workspace,
               "null pointer");
  TH_CHECK_ARG(n_verts >= 1 && n_tok >= 1, "bad sizes");
  EncTail ev[TH_MAX_VIEWS];
  int rc = enc_views(enc, ev);
  if (rc) return rc;
  TH_CHECK_ARG(workspace_bytes >= th_paint_group_latents_workspace_bytes(enc->n_views, n_verts, n_tok),
               "workspace too small");
  TH_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  void* scratch = static_cast<unsigned char*>(workspace) + align_up((size_t)enc->n_views * sizeof(EncTail), 256);
  // the per-view descriptors travel through the caller's workspace (stream-ordered copy from a pageable buffer:
  // cudaMemcpyAsync stages it before returning, so `ev` may die with this frame)
  TH_CUDA(cudaMemcpyAsync(workspace, ev, (size_t)enc->n_views * sizeof(EncTail), cudaMemcpyHostToDevice, st));
  return launch_paint_group_latents(static_cast<const EncTail*>(workspace), enc->n_views, reduction_w, reduction_b,
                                    uv_scale_x, uv_scale_y, verts, cam_R, cam_T, cam_K, vizmap, n_verts, cluster_start,
                                    cluster_members, n_tok, scratch, tokens, st);
}

// ---- generic linear layer through the tcgen05 GEMM --------------------------------------------
struct LinearPlan {
  int n_chunks;
  int chunk_n[16], chunk_valid[16], col0[16];
  size_t img_off[16], bias_off, scale_off, total;
};
static bool linear_plan(int n_out, int n_in, LinearPlan* p) {
  if (n_out < 4 || n_out % 4 || n_in < 64 || n_in % 64 || n_out > 16 * 256) return false;
  p->n_chunks = 0;
  size_t off = 256;  // [0]: magic, [16]: inv_scale
  p->scale_off = 16;
  for (int c = 0; c < n_out;) {
    const int rem = n_out - c;
    const int n = rem > 128 ? 256 : 128, valid = rem < n ? rem : n;
    const int i = p->n_chunks++;
    p->chunk_n[i] = n, p->chunk_valid[i] = valid, p->col0[i] = c;
    p->img_off[i] = off;
    off += align_up((size_t)n * n_in * 2 * 2, 1024);
    c += valid;
  }
  p->bias_off = off;
  off += align_up((size_t)p->n_chunks * 256 * 4, 256);
  p->total = off;
  return true;
}
constexpr uint32_t LINEAR_MAGIC = 0x314E4C54u;  // 'TLN1'

size_t th_linear_packed_bytes(int32_t n_out, int32_t n_in) {
  LinearPlan p;
  return linear_plan(n_out, n_in, &p) ? p.total : 0;
}

int th_linear_pack(const float* weight_host, const float* bias_host, int32_t n_out, int32_t n_in, void* packed_host,
                   size_t bytes) {
  TH_CHECK_ARG(weight_host && packed_host, "null pointer");
  LinearPlan p;
  TH_CHECK_ARG(linear_plan(n_out, n_in, &p), "n_in must be a multiple of 64, n_out a multiple of 4 (<= 4096)");
  TH_CHECK_ARG(bytes >= p.total, "buffer too small");
  unsigned char* blob = static_cast<unsigned char*>(packed_host);
  memset(blob, 0, p.total);
  const int e = weight_scale_exp(weight_host, (size_t)n_out * n_in);
  const float up = ldexpf(1.0f, e), inv = ldexpf(1.0f, -e);
  memcpy(blob, &LINEAR_MAGIC, 4);
  memcpy(blob + p.scale_off, &inv, 4);
  for (int i = 0; i < p.n_chunks; ++i) {
    pack_weight_image(weight_host + (size_t)p.col0[i] * n_in, p.chunk_n[i], n_in, p.chunk_valid[i], up,
                      blob + p.img_off[i]);
    if (bias_host) memcpy(blob + p.bias_off + (size_t)i * 256 * 4, bias_host + p.col0[i], (size_t)p.chunk_valid[i] * 4);
  }
  return TH_OK;
}

int th_linear(const float* x, int64_t m, int32_t ldx, const void* packed_dev, int32_t n_out, int32_t n_in, float* y,
              int32_t ldy, int32_t relu, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(x && packed_dev && y, "null pointer");
  LinearPlan p;
  TH_CHECK_ARG(linear_plan(n_out, n_in, &p), "n_in must be a multiple of 64, n_out a multiple of 4 (<= 4096)");
  TH_CHECK_ARG(m >= 0 && ldx >= n_in && ldx % 4 == 0 && ldy >= n_out && ldy % 4 == 0, "bad leading dimension");
  TH_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(packed_dev) & 1023) == 0,
               "misaligned pointer (x, y: 16 bytes; packed blob: 1024 bytes)");
  if (m == 0) return TH_OK;
  const unsigned char* blob = static_cast<const unsigned char*>(packed_dev);
  for (int i = 0; i < p.n_chunks; ++i) {
    GemmArgs g{};
    g.nseg = 1;
    g.seg[0].ptr = x;
    g.seg[0].K = n_in;
    g.seg[0].ld = ldx;
    g.bias = reinterpret_cast<const float*>(blob + p.bias_off + (size_t)i * 256 * 4);
    g.C = y + p.col0[i];
    g.ldc = ldy;
    g.M = m;
    g.N = p.chunk_n[i];
    g.n_store = p.chunk_valid[i];
    g.relu = relu ? 1 : 0;
    g.acc_scale = reinterpret_cast<const float*>(blob + p.scale_off);
    int rc = launch_gemm_tc(g, blob + p.img_off[i], st, PROF_PROLOGUE);
    if (rc) return rc;
  }
  return TH_OK;
}

size_t th_marching_cubes_workspace_bytes(int32_t nx, int32_t ny, int32_t nz) {
  if (nx < 1 || ny < 1 || nz < 1) return 256;
  return marching_cubes_workspace_bytes(nx, ny, nz);
}

int th_marching_cubes(const float* volume, int32_t nx, int32_t ny, int32_t nz, float iso, float* vertices,
                      int64_t max_vertices, int32_t* triangles, int64_t max_triangles, int64_t* counts_host,
                      void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  TH_CHECK_ARG(volume && counts_host && workspace, "null pointer");
  TH_CHECK_ARG(nx >= 1 && ny >= 1 && nz >= 1 && (int64_t)nx * ny * nz < (1LL << 31) / 3, "bad volume size");
  TH_CHECK_ARG(max_vertices >= 0 && max_triangles >= 0, "bad capacity");
  TH_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  if (workspace_bytes < th_marching_cubes_workspace_bytes(nx, ny, nz)) {
    set_error("th_marching_cubes: workspace %zu < %zu bytes", workspace_bytes, th_marching_cubes_workspace_bytes(nx, ny, nz));
    return TH_EWORKSPACE;
  }
  unsigned long long* counts_dev =
      reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + workspace_bytes - 256);
  int rc = launch_marching_cubes(volume, nx, ny, nz, iso, vertices, vertices ? max_vertices : 0, triangles,
                                 triangles ? max_triangles : 0, counts_dev, workspace, st);
  if (rc) return rc;
  unsigned long long cnt[2] = {0, 0};
  TH_CUDA(cudaMemcpyAsync(cnt, counts_dev, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  TH_CUDA(cudaStreamSynchronize(st));
  counts_host[0] = (int64_t)cnt[0];
  counts_host[1] = (int64_t)cnt[1];
  if ((vertices && (int64_t)cnt[0] > max_vertices) || (triangles && (int64_t)cnt[1] > max_triangles)) {
    set_error("th_marching_cubes: %llu vertices / %llu triangles exceed the buffers (%lld / %lld)", cnt[0], cnt[1],
              (long long)max_vertices, (long long)max_triangles);
    return TH_EWORKSPACE;
  }
  return TH_OK;
}

size_t th_vit_attention_workspace_bytes(int32_t batch, int32_t n_tokens, int32_t n_heads) {
  if (batch < 1 || n_tokens < 1 || n_heads < 1) return 256;
  return vit_attention_workspace_bytes(batch, n_tokens, n_heads);
}

int th_vit_attention(const float* qkv, int32_t batch, int32_t n_tokens, int32_t n_heads, int32_t head_dim, float scale,
                     float* out, void* workspace, size_t workspace_bytes, void* stream) {
  TH_CHECK_ARG(qkv && out && workspace, "null pointer");
  TH_CHECK_ARG(batch >= 1 && n_tokens >= 1 && n_heads >= 1 && (int64_t)batch * n_heads <= 65535, "bad sizes");
  TH_CHECK_ARG(head_dim == 64, "head_dim must be 64 (vit_tiny)");
  TH_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "misaligned pointer");
  TH_CHECK_ARG(workspace_bytes >= th_vit_attention_workspace_bytes(batch, n_tokens, n_heads), "workspace too small");
  return launch_vit_attention(qkv, batch, n_tokens, n_heads, scale, out, workspace, static_cast<cudaStream_t>(stream));
}

int th_group_mean(const void* x, int32_t is_f64, int32_t n_cols, const int32_t* cluster_start,
                  const int32_t* cluster_members, int32_t n_tok, int32_t outer_order, void* out, void* stream) {
  TH_CHECK_ARG(x && cluster_start && cluster_members && out, "null pointer");
  TH_CHECK_ARG(n_cols >= 1 && n_tok >= 1, "bad sizes");
  return launch_group_mean(x, is_f64, n_cols, cluster_start, cluster_members, n_tok, outer_order, out,
                           static_cast<cudaStream_t>(stream));
}

int th_near_far(const float* ray_o, float* ray_d, int64_t n_rays, const float* bounds, float* near_, float* far_,
                uint8_t* mask_at_box, void* stream) {
  TH_CHECK_ARG(ray_o && ray_d && bounds && near_ && far_ && mask_at_box && n_rays >= 0, "bad argument");
  return launch_near_far(ray_o, ray_d, n_rays, bounds, near_, far_, mask_at_box, static_cast<cudaStream_t>(stream));
}

size_t th_generate_rays_workspace_bytes(int64_t n_pixels) {
  return n_pixels > 0 ? generate_rays_workspace_bytes(n_pixels) : 256;
}

int th_generate_rays(int32_t h, int32_t w, const float* K_inv, const float* R, const float* T, const float* bounds,
                     float* ray_o, float* ray_d, float* near_, float* far_, uint8_t* mask_at_box, float* ray_o_c,
                     float* ray_d_c, float* near_c, float* far_c, int64_t* count_dev, void* workspace,
                     size_t workspace_bytes, void* stream) {
  TH_CHECK_ARG(h >= 1 && w >= 1 && K_inv && R && T && ray_o && ray_d, "bad argument");
  if (bounds) TH_CHECK_ARG(near_ && far_ && mask_at_box, "bounds given: near / far / mask_at_box outputs are required");
  const bool compact = ray_o_c || ray_d_c || near_c || far_c;
  if (compact) {
    TH_CHECK_ARG(bounds && ray_o_c && ray_d_c && near_c && far_c && count_dev, "compacted outputs: all four + count, with bounds");
    if (!workspace || workspace_bytes < generate_rays_workspace_bytes((int64_t)h * w)) {
      set_error("th_generate_rays: workspace %zu < %zu bytes", workspace_bytes,
                generate_rays_workspace_bytes((int64_t)h * w));
      return TH_EWORKSPACE;
    }
  }
  return launch_generate_rays(h, w, K_inv, R, T, bounds, ray_o, ray_d, near_, far_, mask_at_box, compact ? ray_o_c : nullptr,
                              ray_d_c, near_c, far_c, count_dev, workspace, static_cast<cudaStream_t>(stream));
}

int th_nchw_to_nhwc(const float* src, float* dst, int32_t n, int32_t c, int32_t h, int32_t w, void* stream) {
  TH_CHECK_ARG(src && dst && n >= 1 && c >= 1 && h >= 1 && w >= 1, "bad argument");
  return launch_nchw_to_nhwc(src, dst, n, c, h, w, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

