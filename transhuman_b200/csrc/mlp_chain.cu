// Layer-chained tcgen05 kernel for the per-point network (rows a9/a10), sm_100a.
//
// mlp_tc.cu runs the network one layer per launch: every activation makes a round
// trip through HBM (55 KB per point) and the layers sit on the HBM roof.  This kernel
// runs a whole PROGRAM of layers per launch.  The unit of work is 256 sample points
// (128 per CTA of a 2-CTA cluster, all V views of them); a cluster walks its unit
// through every job of the program -- one job = one layer applied to one view's
// 128-row tile -- before it moves to the next unit, so that
//   * activations between layers live in a small per-CTA scratch (768 KB) that is
//     written by the epilogue warps (32-byte stores straight into the tile image) and
//     read back by the loader's bulk copies while it is still in L2 -- no second launch,
//     HBM sees the chunk inputs plus whatever part of the 115 MB of scratch L2 evicts;
//   * the cross-view attention needs no kernel of its own: the key embeds stay in
//     TMEM, the 3x3 scores and their softmax are computed by the epilogue threads
//     (one thread = one point), and the otherwise idle warps 0-5 mix X in place;
//   * the alpha / rgb heads are dot products inside the fc_3 / fc_4 epilogues.
// Roles (16 warps = 4 per scheduler, 128 registers each): warps 0-5 mix, 6 loader
// (2-D TMA copies of tile images; both CTAs' bytes complete on the leader's stage
// barrier, cta_group::2, so there is no relay between the CTAs), 7 MMA issuer (leader
// CTA only), 8-15 epilogue (two per TMEM lane quadrant).
// All synchronisation is CTA- or cluster-local (same rows stay on the same CTA):
//   stage full/empty mbarriers (loader <-> MMA), tfull mbarriers (MMA -> epilogue),
//   and monotonic shared-memory counters for epilogue -> MMA (TMEM reuse, one counter
//   per CTA of the pair), epilogue -> loader (a stored tile may be loaded: CTA-scope
//   release by the writers, gpu-scope + proxy fence by the loader after its acquire),
//   scores -> mix -> loader (one counter per k-block).  Every counter has one
//   spinning reader; TH_CHAIN_DBG=16|32|64|128 delays the roles at random and the
//   tests demand bit-identical output under it.
// The MMA scheme (fp16 hi/lo split, 3 products, cta_group::2, M = 256) and the tile
// image format are those of mlp_tc.cu.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace th {
namespace chain {
using namespace tc;

// 16 warps = 4 per scheduler, so every thread may use 128 registers:
// warps 0-5 mix, 6 loader, 7 MMA issuer / relay, 8-15 epilogue (2 per TMEM lane quadrant)
constexpr int NUM_THREADS = 512;
constexpr int MIX_THREADS = 192, LOADER_WARP = 6, MMA_WARP = 7, EPI_WARP0 = 8;
constexpr int EPI_WARPS = 8;
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = 65536;  // A hi/lo (32 KB) + this CTA's half of B hi/lo (<= 32 KB)
constexpr int MAX_SEG = 5, MAX_JOBS = 32;
constexpr uint32_t TILE_IMG = 2 * A_TILE_BYTES;       // one 128-row x 64-column hi/lo k-block
constexpr uint32_t SCR_ACT = 4 * TILE_IMG;            // a 256-wide activation tile: 128 KB
constexpr int CHAIN_MAX_V = 3;                        // V key embeds + one more must fit 512 TMEM columns
// Per-CTA scratch for V views: slot A (S -> N1 -> INTER) and slot B (X -> XT -> G), V tiles each.
// (deferred-tail program: slot set A twice -- it alternates with the unit's parity -- and slot B at half size, G only)
__host__ __device__ inline uint32_t scratch_stride(int V) { return 2u * (uint32_t)V * SCR_ACT + (uint32_t)V * (SCR_ACT / 2); }

// EPI_MIX: the attention mix on the ACCUMULATOR side (pre-mapped program): out = relu(Z + b + sum_i A[i][j] Y_i) with
// Z = this job's accumulator (the S part of fc_1', view j) and Y_i = kept accumulators of the X part for every view i
enum { EPI_IMG = 0, EPI_KEEP = 2, EPI_SCORES = 3, EPI_ALPHA = 4, EPI_RGB = 5, EPI_MIX = 6 };

struct Seg {
  const unsigned char* img;  // chunk-level tile image, or nullptr = this CTA's scratch
  int64_t tile_off;          // chunk images: row tile of point tile 0 (view * Pp / 128)
  uint32_t scratch_off;
  int32_t kbs;
  int32_t dep;      // job (same unit) whose stored output this segment reads, -1 = chunk input
  int32_t dep_mix;  // 1 = written by the mix warps, released per k-block
  int32_t keep;     // chunk input that a later job of the unit reads again: keep it in L2 (default: evict first)
  int32_t aset;     // scratch segment in slot set A, which alternates with the unit's parity (deferred-tail program)
};
struct Job {
  Seg seg[MAX_SEG];
  const unsigned char* wimg;
  const float* bias;
  const float* bias2;       // EPI_SCORES: bias of the kept key embed
  uint32_t out_off;         // EPI_IMG: destination inside the CTA's scratch
  int32_t nseg, nkb, N, relu, epi, tmem_col, wait_back, view;
  const float* acc_scale;   // DEVICE pointer to the 2^-e of this job's weight image (img_inv_scale_ptr): acc * scale + bias
  const float* acc_scale2;  // EPI_SCORES: the scale of the kept key embeds' weight image; EPI_MIX: of the Y jobs'
  int32_t reader;           // host only: last job whose epilogue reads this accumulator (-1 = its own epilogue)
  // Software pipelining across units: a job with shift = -1 belongs to the unit of the PREVIOUS iteration (the tail
  // of the network -- fc_3, view_fc', fc_4' and the heads -- is issued into the next unit's attention-mix window,
  // where the tensor pipe would otherwise wait for the mix warps).  Iteration `it` of a cluster runs the shift-0
  // jobs of its unit number `it` and the shift -1 jobs of unit number `it - 1`; one extra iteration drains the tail.
  int32_t shift;
  int32_t out_aset;         // out_off lies in slot set A (alternates with the unit's parity)
};
struct Program {
  // TMA descriptors: tile images as rows of 64 fp16 (128 B).  tm_a: box 256 rows (one 32 KB hi|lo
  // k-block of an A tile) over the chunk workspace; tm_b256 / tm_b128: box N rows (one CTA's [hi | lo]
  // share of a weight k-block) over the packed weight blob.
  CUtensorMap tm_a, tm_b256, tm_b128;
  const unsigned char* a_base;  // first byte tm_a covers
  const unsigned char* w_base;  // first byte tm_b* cover
  Job job[MAX_JOBS];
  unsigned char* scratch;
  const float *afc_w, *afc_b, *rgb_w, *rgb_b;
  float* alpha;  // (Pp) chunk-local alpha_raw
  float* raw;
  float* alpha_out;
  const int32_t* dst_ids;
  int64_t first, P;
  // pre-mapped inputs without the X copy jobs: the mix works on the X tiles of the chunk image in place
  // (x_img = first X tile of view 0, x_view_stride = bytes between the views); nullptr = slot B of the scratch
  unsigned char* x_img;
  int64_t x_view_stride;
  int32_t njobs, num_units, V, has_mix, zero_rgb, alpha_only, kp_col;
  int32_t deferred_tail;  // 1 = the program has shift -1 jobs (and two alternating A slot sets)
  int32_t ks_col[TH_MAX_VIEWS];
  CompositeArgs cmp;  // fused compositing in the EPI_RGB epilogue (rgb_map == nullptr: off)
  int32_t dbg;  // TH_CHAIN_DBG: timing experiments only, 1 = skip the mix, 2 = skip the store fences (results wrong); 16/32/64/128 = random delays in the loader / MMA / epilogue / mix role, 512 = writer-side fences as well (results valid)
  unsigned long long* stats;  // TH_CHAIN_STATS=1: per-CTA wait-time counters (cycles), else nullptr
};

// ---- shared-memory counters (monotonic; one writer side, one spinning reader) ----
__device__ __forceinline__ uint32_t ld_acquire_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// for opportunistic reads that a fence.acq_rel follows (relaxed load + fence = acquire)
__device__ __forceinline__ uint32_t ld_relaxed_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// `nap` nanoseconds between polls: a spinning single thread otherwise competes for issue slots with
// the epilogue warp that shares its scheduler.
// Polls with RELAXED loads and acquires once at the end: an acquire load at cluster scope makes ptxas emit an L1
// invalidation (CCTL.IVALL) per poll -- 34 M of them per launch in the round-1 kernel (ncu source page).
__device__ __forceinline__ void wait_counter(uint32_t addr, uint32_t target, int what, unsigned nap = 32) {
  if ((int32_t)(ld_acquire_u32(addr) - target) >= 0) return;
  const long long t0 = clock64();
  while (true) {
    if ((int32_t)(ld_relaxed_u32(addr) - target) >= 0) {
      if ((int32_t)(ld_acquire_u32(addr) - target) >= 0) break;  // the acquire that orders the data behind the count
    }
    __nanosleep(nap);
    if (clock64() - t0 > WAIT_TIMEOUT_CYCLES) {
      printf("k_chain: counter wait timed out (block %d thread %d what %d target %u have %u)\n", blockIdx.x,
             threadIdx.x, what, target, ld_acquire_u32(addr));
      __trap();
    }
  }
}
// TH_CHAIN_DBG & (16|32|64|128): random delays per role, to shake out ordering bugs (results stay valid)
__device__ __forceinline__ void jitter(int dbg, int role_bit = 16) {
  if (dbg & role_bit) __nanosleep((unsigned)((clock64() * 2654435761ull) >> 13) & 2047u);
}
__device__ __forceinline__ void add_release_local(uint32_t addr) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void add_release_remote(uint32_t addr, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "red.release.cluster.shared::cluster.add.u32 [ra], 1;\n\t}" ::"r"(addr),
      "r"(cta)
      : "memory");
}
// "Accumulator drained" from the peer CTA to the leader's MMA issuer.  It orders nothing but the epilogue's TMEM
// reads, and those have COMPLETED when it is sent (tcgen05.wait::ld returned: the values sit in registers), so the
// signal needs no release: a release at cluster scope is cumulative over the warp's global stores of the tile it has
// just written and made every epilogue warp of the peer CTA wait for their acknowledgement once per job (ncu: 3 % of
// the kernel's stall samples were `membar` stalls on this instruction, ~10 % of the peer's epilogue time, and the pair
// is gated by its slower CTA).  The stored tile itself is published separately (cnt_job, release).
__device__ __forceinline__ void add_relaxed_remote(uint32_t addr, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "red.relaxed.cluster.shared::cluster.add.u32 [ra], 1;\n\t}" ::"r"(addr),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// scratch accesses of the mix warps: never allocate in L1 (the tiles are rewritten by the async
// proxy every unit, L1 is not coherent with it), keep them in L2 (evict last)
__device__ __forceinline__ uint4 ldcg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(L2_EVICT_LAST));
  return v;
}
__device__ __forceinline__ void st16_keep(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "l"(L2_EVICT_LAST)
               : "memory");
}
// 16-byte asynchronous copy global -> shared (generic proxy), L1 bypassed, kept in L2
__device__ __forceinline__ void cp_async16_keep(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(L2_EVICT_LAST)
               : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ldcg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// 32-byte store of 16 fp16 (two 16-byte chunks of one 32-byte sector) that stays in L2
__device__ __forceinline__ void st32_keep(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p),
               "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "l"(L2_EVICT_LAST)
               : "memory");
}
// Packed fp32x2 arithmetic (FADD2 / FFMA2 on sm_100): same rounding as the scalar forms, half the
// issue slots -- the epilogue and the mix are bound by instruction issue, not by memory.
__device__ __forceinline__ unsigned long long pk2(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 up2(unsigned long long r) {
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return up2(d);
}
__device__ __forceinline__ void split2(float2 v, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(v.x, v.y);  // saturating: see common.cuh
  const float2 r = sub2(v, __half22float2(*reinterpret_cast<const __half2*>(&hi)));
  lo = cvt_f16x2_sat(r.x, r.y);
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  split2(make_float2(x, y), hi, lo);
}
__device__ __forceinline__ float2 join2(uint32_t hi, uint32_t lo) {
  return add2(__half22float2(*reinterpret_cast<const __half2*>(&hi)),
              __half22float2(*reinterpret_cast<const __half2*>(&lo)));
}

constexpr int STATS_PER_CTA = 32 + 3 * MAX_JOBS;  // role totals + per-job MMA waits
constexpr int ATAB_LD = 9;  // attention table row stride (floats): V x V <= 9 entries per point
// stages | bias (2 x 256) | attention table | partial scores | control (256 B) | mix staging (96 B per mix thread)
constexpr uint32_t MIX_RING_OFF = NSTAGE * STAGE_BYTES + 2 * 256 * 4 + 128 * ATAB_LD * 4 + 128 * 4 * 4 + 256;
// ... | per-sample compositing terms (alpha, r, g, b, z per row of the tile)
constexpr uint32_t COMP_OFF = MIX_RING_OFF + MIX_THREADS * 96;
// ... | second attention table: the table alternates with the unit's parity, so that the score epilogues of unit u + 1
// may run while the mix warps still read the table of unit u (deep deferral, TH_CHAIN_DEFER=2)
constexpr uint32_t ATAB1_OFF = COMP_OFF + 128 * 5 * 4;
constexpr size_t SMEM_BYTES = (size_t)ATAB1_OFF + 128 * ATAB_LD * 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    k_chain(const __grid_constant__ Program pg) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  float* s_bias = reinterpret_cast<float*>(smem_raw + NSTAGE * STAGE_BYTES);  // 2 x 256
  float* s_atab = s_bias + 512;                                               // 128 x ATAB_LD
  float* s_part = s_atab + 128 * ATAB_LD;                                     // 128 x 4 partial scores
  float* s_atab1 = reinterpret_cast<float*>(smem_raw + ATAB1_OFF);            // the table of odd units
  unsigned char* ctrl_ptr = reinterpret_cast<unsigned char*>(s_part + 128 * 4);
  const uint32_t ctrl = base + NSTAGE * STAGE_BYTES + 2048 + 128 * ATAB_LD * 4 + 128 * 4 * 4;
  const uint32_t bar_full = ctrl, bar_empty = ctrl + 24, bar_pfull = ctrl + 48, bar_tfull = ctrl + 72;
  // counters (u32): epilogue-done of this CTA / of the peer, scores, mix per k-block [4], stored tile per job
  const uint32_t cnt_epi = ctrl + 96, cnt_epi_peer = ctrl + 100, cnt_scores = ctrl + 104, cnt_mix = ctrl + 112,
                 cnt_job = ctrl + 128;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl_ptr + 88);
  uint32_t* counters = reinterpret_cast<uint32_t*>(ctrl_ptr + 96);  // see cnt_* below

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int njobs = pg.njobs, V = pg.V;
  const uint32_t flip_on = (njobs & 1) ? 256u : 0u;  // odd programs alternate TMEM halves from unit to unit
  unsigned char* scratch = pg.scratch + (size_t)blockIdx.x * scratch_stride(V);
  // units of this cluster, and iterations of its job loop (one more when the tail is deferred)
  const int n_units_mine = cluster_id < pg.num_units ? (pg.num_units - cluster_id + nclusters - 1) / nclusters : 0;
  const int n_iter = n_units_mine + ((pg.deferred_tail && n_units_mine > 0) ? 1 : 0);
  const uint32_t aset_bytes = (uint32_t)V * SCR_ACT;  // distance between the two A slot sets

  if (tid == 0) {
    if (base & 1023u) {
      printf("k_chain: dynamic shared memory is not 1 KB aligned (%u)\n", base);
      __trap();
    }
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tfull + 8, 1);
    for (int i = 0; i < 8 + MAX_JOBS; ++i) counters[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // optional wait-time accounting (one slot per question; cycles summed over the launch)
  unsigned long long* stats = pg.stats ? pg.stats + (size_t)blockIdx.x * STATS_PER_CTA : nullptr;
  long long tw[6] = {0, 0, 0, 0, 0, 0};
#define TH_TIMED(slot, stmt)            \
  do {                                  \
    if (stats) {                        \
      const long long t_ = clock64();   \
      stmt;                             \
      tw[slot] += clock64() - t_;       \
    } else {                            \
      stmt;                             \
    }                                   \
  } while (0)
  const long long t_begin = clock64();

  if (warp < LOADER_WARP) {
    // ===================== mix warps: XT_j = sum_i A[i][j] X_i, in place =====================
    // (cross_transformer.py:144-146).  Same tile-image layout in and out, so the work is
    // elementwise over 16-byte chunks: position = (row, physical chunk); a thread owns 4 rows.
    if (pg.has_mix) {
      unsigned char* xbase = scratch + (size_t)V * SCR_ACT;  // slot B: X_v at v * SCR_ACT
      size_t xstride = SCR_ACT;
      int it = 0;
      for (int u = cluster_id; u < pg.num_units; u += nclusters, ++it) {
        if (pg.x_img) {  // this CTA's 128-row tile of every view's X image (4 k-blocks per tile)
          xbase = pg.x_img + (size_t)(2 * (int64_t)u + rank) * 4 * TILE_IMG;
          xstride = (size_t)pg.x_view_stride;
        }
        TH_TIMED(0, if (lane == 0) wait_counter(cnt_scores, 4u * (uint32_t)(it + 1), 1, 256); __syncwarp());
        const long long t_mix = clock64();
        fence_proxy_async_all();
        if (pg.dbg & 1) {  // timing experiment: no mix at all (results wrong)
          __syncwarp();
          if (lane == 0)
            for (int kb = 0; kb < 4; ++kb) add_release_local(cnt_mix + 4 * kb);
          if (stats) tw[1] += clock64() - t_mix;
          continue;
        }
        jitter(pg.dbg, 128);
        // One stream of 4096 positions per unit (4 k-blocks x 128 rows x 8 chunks of 16 bytes; thread t takes
        // t, t + 192, ...), software pipelined TWO positions deep with no drain at the k-block boundaries:
        // while position p is mixed, p + 192 already sits in registers and the 2V 16-byte loads of p + 384 are
        // in flight as cp.async copies into this thread's private 96-byte slot of shared memory (the mix is bound
        // by the latency of its loads and by its ALU work, ncu: 46 % of its samples wait for the first use of a
        // loaded value with one position in flight).  A warp's 32 positions never straddle a k-block (1024
        // positions), so the warp publishes k-block kb when it has stored its last position of it.
        const uint32_t ring = base + MIX_RING_OFF + (uint32_t)tid * 96u;
        auto src_off = [&](int pos) {
          return (uint32_t)(pos >> 10) * TILE_IMG + (uint32_t)((pos & 1023) >> 3) * 128u + (uint32_t)(pos & 7) * 16u;
        };
        auto stage = [&](int pos) {  // global -> shared, asynchronous, L1 bypassed, kept in L2
          const uint32_t off = src_off(pos);
#pragma unroll
          for (int i = 0; i < CHAIN_MAX_V; ++i)
            if (i < V) {
              cp_async16_keep(ring + 32u * i, xbase + (size_t)i * xstride + off);
              cp_async16_keep(ring + 32u * i + 16u, xbase + (size_t)i * xstride + off + A_TILE_BYTES);
            }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto unstage = [&](uint4 (&h)[CHAIN_MAX_V], uint4 (&l)[CHAIN_MAX_V]) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
          for (int i = 0; i < CHAIN_MAX_V; ++i)
            if (i < V) {
              h[i] = lds16(ring + 32u * i);
              l[i] = lds16(ring + 32u * i + 16u);
            }
        };
        uint4 ch[CHAIN_MAX_V], cl[CHAIN_MAX_V], nh[CHAIN_MAX_V], nl[CHAIN_MAX_V];
        stage(tid);
        unstage(ch, cl);
        if (tid + MIX_THREADS < 4096) stage(tid + MIX_THREADS);
#pragma unroll 1
        for (int pos = tid; pos < 4096; pos += MIX_THREADS) {
          const bool more = pos + MIX_THREADS < 4096;
          if (more) {
            unstage(nh, nl);                                              // p + 192: arrived during the last position
            if (pos + 2 * MIX_THREADS < 4096) stage(pos + 2 * MIX_THREADS);  // p + 384: in flight during this one
          }
          const int row = (pos & 1023) >> 3;
          const uint32_t off = src_off(pos);
          float2 x[CHAIN_MAX_V][4];
#pragma unroll
          for (int i = 0; i < CHAIN_MAX_V; ++i)
            if (i < V) {
              x[i][0] = join2(ch[i].x, cl[i].x);
              x[i][1] = join2(ch[i].y, cl[i].y);
              x[i][2] = join2(ch[i].z, cl[i].z);
              x[i][3] = join2(ch[i].w, cl[i].w);
            }
          const float* A = ((it & 1) ? s_atab1 : s_atab) + row * ATAB_LD;
#pragma unroll
          for (int j = 0; j < CHAIN_MAX_V; ++j)
            if (j < V) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 o = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < CHAIN_MAX_V; ++i)
                  if (i < V) {
                    const float a = A[i * V + j];
                    o = fma2(make_float2(a, a), x[i][e], o);
                  }
                split2(o, hi[e], lo[e]);
              }
              unsigned char* dst = xbase + (size_t)j * xstride + off;
              st16_keep(dst, make_uint4(hi[0], hi[1], hi[2], hi[3]));
              st16_keep(dst + A_TILE_BYTES, make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
          // last position of this warp in its k-block -> publish the k-block (one counter per k-block)
          if (!more || ((pos + MIX_THREADS) >> 10) != (pos >> 10)) {
            if (pg.dbg & 512) {
              __threadfence();
              fence_proxy_async_all();
            }
            __syncwarp();
            if (lane == 0) add_release_local(cnt_mix + 4 * (pos >> 10));
          }
          if (more) {
#pragma unroll
            for (int i = 0; i < CHAIN_MAX_V; ++i) {
              ch[i] = nh[i];
              cl[i] = nl[i];
            }
          }
        }
        if (stats) tw[1] += clock64() - t_mix;
      }
    }
  } else if (warp == LOADER_WARP) {
    // ===================== loader =====================
    // One thread; its per-k-block path is kept short (incremental pointers and stage index, two bulk
    // copies: the A tile image and this CTA's contiguous [hi | lo] share of the weight k-block)
    // because at ~6 cycles per dependent instruction a long path makes the loader the bottleneck.
    if (lane == 0) {
      uint32_t s = 0, ph = 0;  // stage index and its phase bit
      for (int it = 0; it < n_iter; ++it) {
        // What this thread's fences already cover in this iteration: bit j = the tile job j stored, bit kb = the
        // mix of k-block kb.  A fence executed after a counter was seen complete orders those stores before
        // every later copy, so each dependency costs one wait + fence per unit, not one per reader -- a
        // fence is ~1-2 kcycles here and sits between "tile stored" and the dependent job's first copy.
        uint32_t dep_seen = 0, mix_seen = 0;
        for (int j = 0; j < njobs; ++j) {
          const Job& jb = pg.job[j];
          const int ui = it + jb.shift;  // which of this cluster's units the job works on
          if (ui < 0 || ui >= n_units_mine) continue;
          const int u = cluster_id + ui * nclusters;
          const int64_t ptile = 2 * (int64_t)u + rank;
          // job D (shift s_D) has run it + s_D + 1 times once it is done for the unit the shift-s_D jobs of this
          // iteration work on; a consumer needs its producers' outputs for ITS unit: it + shift + 1 executions
          // (a deferred consumer of a shift-0 producer needs one execution less than the producer's own class)
          const uint32_t done_target = (uint32_t)EPI_WARPS * (uint32_t)(it + jb.shift + 1);
          const uint32_t mix_target = (uint32_t)(MIX_THREADS / 32) * (uint32_t)(it + jb.shift + 1);  // the mix of ITS unit
          const uint32_t aset_off = (ui & 1) ? aset_bytes : 0u;
          const uint32_t b_bytes = (uint32_t)jb.N * 128u;  // N/2 rows x 128 B x (hi, lo)
          const unsigned char* w = jb.wimg + (size_t)rank * b_bytes;
          const uint32_t w_step = 2u * b_bytes, tx = TILE_IMG + b_bytes;
          const CUtensorMap* b_map = jb.N == 256 ? &pg.tm_b256 : &pg.tm_b128;
          const int nseg = jb.nseg;
          for (int sgi = 0; sgi < nseg; ++sgi) {
            const Seg& sg = jb.seg[sgi];
            const int kbs = sg.kbs, dep_mix = sg.dep_mix;
            const bool from_chunk = sg.img != nullptr;
            const unsigned char* src = from_chunk ? sg.img + (size_t)(sg.tile_off + ptile) * kbs * TILE_IMG
                                                  : scratch + sg.scratch_off + (sg.aset ? aset_off : 0u);
            // chunk inputs stream through L2 once (evict first); scratch tiles and weights are the
            // working set that should stay resident (evict last)
            const uint64_t pol = (from_chunk && !sg.keep) ? L2_EVICT_FIRST : L2_EVICT_LAST;
            if (sg.dep >= 0 && !((dep_seen >> sg.dep) & 1u)) {
              TH_TIMED(0, wait_counter(cnt_job + 4 * sg.dep, done_target, 2));
              // whatever else is stored by now rides on the same fence
              for (int j2 = 0; j2 < njobs; ++j2)  // (each producer against the count of ITS class: conservative)
                if ((int32_t)(ld_relaxed_u32(cnt_job + 4 * j2) -
                              (uint32_t)EPI_WARPS * (uint32_t)(it + pg.job[j2].shift + 1)) >= 0)
                  dep_seen |= 1u << j2;
              // cumulative: covers the epilogue warps' stores observed through the counters
              TH_TIMED(4, __threadfence(); fence_proxy_async_all());
            }
            for (int kk = 0; kk < kbs; ++kk) {
              if (dep_mix && !((mix_seen >> kk) & 1u)) {
                // one counter per k-block: the mix warps are not in lockstep
                TH_TIMED(1, wait_counter(cnt_mix + 4 * kk, mix_target, 3));
                for (int k2 = kk; k2 < 4; ++k2)
                  if ((int32_t)(ld_relaxed_u32(cnt_mix + 4 * k2) - mix_target) >= 0) mix_seen |= 1u << k2;
                TH_TIMED(4, __threadfence(); fence_proxy_async_all());
              }
              jitter(pg.dbg, 16);
              TH_TIMED(2, mbar_wait(bar_empty + 8 * s, ph ^ 1));
              const long long t_issue = stats ? clock64() : 0;
              // both CTAs' bytes are counted by the leader's barrier: the leader arms it for the pair
              const uint32_t sa = base + s * STAGE_BYTES, bar = (bar_full + 8 * s) & PEER_BIT_MASK;
              if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 2 * tx);
              tma2d_pair(sa, &pg.tm_a, 0, (int)((src - pg.a_base) >> 7), bar, pol);
              tma2d_pair(sa + TILE_IMG, b_map, 0, (int)((w - pg.w_base) >> 7), bar, L2_EVICT_LAST);
              if (stats) tw[3] += clock64() - t_issue;
              src += TILE_IMG;
              w += w_step;
              if (++s == NSTAGE) {
                s = 0;
                ph ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      if (rank == 0) {
        // ===================== MMA issuer (leader CTA) =====================
        uint32_t s = 0, ph = 0, G = 0;
        for (int it = 0; it < n_iter; ++it) {
          const uint32_t flip = (it & 1) ? flip_on : 0u;
          for (int j = 0; j < njobs; ++j) {
            const Job& jb = pg.job[j];
            if (it + jb.shift < 0 || it + jb.shift >= n_units_mine) continue;  // (G counts executed jobs)
            jitter(pg.dbg, 32);
            const long long w0 = tw[0], w1 = tw[1], w2 = tw[2];
            if ((int32_t)(G - jb.wait_back) >= 0) {
              // one counter per CTA: warps of a CTA stay within one job of each other (named barrier),
              // the two CTAs of the pair do not
              TH_TIMED(0, wait_counter(cnt_epi, (uint32_t)EPI_WARPS * (G - jb.wait_back + 1), 4, 20);
                       wait_counter(cnt_epi_peer, (uint32_t)EPI_WARPS * (G - jb.wait_back + 1), 5, 20));
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (((uint32_t)jb.tmem_col + flip) & 511u);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(jb.N >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
            const uint32_t half_b = (uint32_t)(jb.N / 2) * 128u;
            const int nkb = jb.nkb;
            for (int kb = 0; kb < nkb; ++kb) {
              TH_TIMED(1, mbar_wait(bar_full + 8 * s, ph));  // both CTAs' halves of the stage
              tc_fence_after();
              const uint32_t sa = base + s * STAGE_BYTES;
              const uint64_t d_ahi = umma_desc(sa), d_alo = umma_desc(sa + A_TILE_BYTES);
              const uint64_t d_bhi = umma_desc(sa + TILE_IMG), d_blo = umma_desc(sa + TILE_IMG + half_b);
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                const uint64_t adv = (uint64_t)((ks * 32) >> 4);
                umma_f16_2cta(d_tmem, d_ahi + adv, d_bhi + adv, idesc, (kb | ks) ? 1u : 0u);
                umma_f16_2cta(d_tmem, d_alo + adv, d_bhi + adv, idesc, 1u);
                umma_f16_2cta(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
              }
              umma_commit_2cta(bar_empty + 8 * s);
              if (++s == NSTAGE) {
                s = 0;
                ph ^= 1;
              }
            }
            umma_commit_2cta(bar_tfull + 8 * (G & 1));
            if (stats) {  // per-job split of the three waits (leader CTA only)
              stats[32 + 3 * j + 0] += (unsigned long long)(tw[0] - w0);
              stats[32 + 3 * j + 1] += (unsigned long long)(tw[1] - w1);
              stats[32 + 3 * j + 2] += (unsigned long long)(tw[2] - w2);
            }
            ++G;
          }
        }
      } else {
        // (peer CTA: nothing to do -- its tensor copies complete on the leader's barriers)
      }
    }
  } else {
    // ===================== epilogue =====================
    // Two warps per TMEM lane quadrant: group g takes half of the accumulator's columns of an image
    // job; the per-point epilogues (scores, heads) are group 0's.
    const int q = warp & 3;          // TMEM lane quadrant
    const int grp = (warp - EPI_WARP0) >> 2;
    const int et = q * 32 + lane;    // row inside the 128-row tile
    float alpha_reg = 0.f;
    uint32_t G = 0;
    for (int it = 0; it < n_iter; ++it) {
      const uint32_t flip = (it & 1) ? flip_on : 0u;
      for (int j = 0; j < njobs; ++j) {
        const Job& jb = pg.job[j];
        const int ui = it + jb.shift;
        if (ui < 0 || ui >= n_units_mine) continue;
        const int64_t ptile = 2 * (int64_t)(cluster_id + ui * nclusters) + rank;
        const int64_t pt = ptile * BM + et;  // chunk-local point of this thread (per-point jobs)
        const uint32_t out_aset_off = (jb.out_aset && (ui & 1)) ? aset_bytes : 0u;
        float* atab = ((ui & 1) ? s_atab1 : s_atab) + et * ATAB_LD;  // this unit's attention table
        const int N = jb.N, epi = jb.epi;
        float* bias_s = s_bias + (G & 1) * 256;
        // this job's bias -> shared memory (double buffered by job parity; the named barrier of EVERY
        // job keeps a warp at most one job ahead of the slowest reader of the other buffer)
        if (epi != EPI_KEEP) {
          for (int c = grp * 128 + et; c < N; c += 256) bias_s[c] = jb.bias ? __ldg(jb.bias + c) : 0.f;
          if (epi == EPI_SCORES && grp == 1) bias_s[128 + et] = __ldg(jb.bias2 + et);
        }
        TH_TIMED(0, epi_bar());
        TH_TIMED(1, mbar_wait(bar_tfull + 8 * (G & 1), (G >> 1) & 1));
        const long long t_work = clock64();
        jitter(pg.dbg, 64);
        tc_fence_after();
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t taddr = lane_addr + (((uint32_t)jb.tmem_col + flip) & 511u);
        if (epi == EPI_IMG || epi == EPI_MIX) {
          // bias / ReLU / fp16 hi-lo split straight from the accumulator to the scratch tile image:
          // a thread owns one row; 16 columns = one 32-byte sector of the hi plane and one of the lo
          // plane (the 128-byte swizzle permutes 16-byte chunks inside a sector pair-wise).
          unsigned char* out = scratch + jb.out_off + out_aset_off + (size_t)et * 128;
          const int ncol = N >> 1, cbeg = grp * ncol;
          const bool swap = (et & 1) != 0;
          const bool relu = jb.relu != 0;
          const float acc_s = __ldg(jb.acc_scale);
          // EPI_MIX adds the bias to an already scaled and mixed value: scale 1
          const float2 sc2 = epi == EPI_MIX ? make_float2(1.f, 1.f) : make_float2(acc_s, acc_s);
          // 32 accumulator columns -> 2 x (hi sector, lo sector)
          auto emit = [&](const uint32_t (&v)[32], int c0) {
            unsigned char* kb_out = out + (size_t)(c0 >> 6) * TILE_IMG;
#pragma unroll
            for (int half = 0; half < 2; ++half) {  // 16 columns each
              uint4 hi[2], lo[2];
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0 + half * 16 + cc * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + half * 16 + cc * 8 + 4);
                const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                      make_float2(b1.z, b1.w)};
                float2 x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  x[e] = fma2(make_float2(__uint_as_float(v[half * 16 + cc * 8 + 2 * e]),
                                          __uint_as_float(v[half * 16 + cc * 8 + 2 * e + 1])),
                              sc2, bb[e]);
                  if (relu) {
                    x[e].x = fmaxf(x[e].x, 0.f);
                    x[e].y = fmaxf(x[e].y, 0.f);
                  }
                }
                split2(x[0], hi[cc].x, lo[cc].x);
                split2(x[1], hi[cc].y, lo[cc].y);
                split2(x[2], hi[cc].z, lo[cc].z);
                split2(x[3], hi[cc].w, lo[cc].w);
              }
              // logical chunks (2m, 2m+1) -> physical (2m ^ x, (2m+1) ^ x), x = row & 7: same sector,
              // halves swapped when x is odd
              const int chunk = ((c0 & 63) >> 3) + half * 2;
              unsigned char* dst = kb_out + (((chunk ^ (et & 7)) & ~1) << 4);
              st32_keep(dst, swap ? hi[1] : hi[0], swap ? hi[0] : hi[1]);
              st32_keep(dst + A_TILE_BYTES, swap ? lo[1] : lo[0], swap ? lo[0] : lo[1]);
            }
          };
          uint32_t va[32], vb[32];
          if (epi == EPI_IMG) {
            // software pipelined: the next 32 columns are in flight from TMEM while these are converted
            tmem_ld32(taddr + cbeg, va);
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + ncol; c0 += 64) {
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              tmem_ld32(taddr + c0 + 32, vb);  // ncol is a multiple of 64
              emit(va, c0);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              if (c0 + 64 < cbeg + ncol) tmem_ld32(taddr + c0 + 64, va);
              emit(vb, c0 + 32);
            }
          } else {
            // N1_j = relu(Z s_z + b + sum_i (A[i][j] s_y) Y_i): Z = this job's accumulator, Y_i = the kept
            // accumulators of the X part (columns ks_col[i]); the attention table was finished by the last
            // score epilogue (ordered by the named barrier of the jobs in between).  fp32 throughout: the mix
            // is never re-quantised to fp16 operands, and the tensor pipe never waits for it.
            const int jv = jb.view;
            const float s_y = __ldg(jb.acc_scale2);
            float am[CHAIN_MAX_V];
#pragma unroll
            for (int i = 0; i < CHAIN_MAX_V; ++i) am[i] = i < V ? atab[i * V + jv] * s_y : 0.f;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + ncol; c0 += 32) {
              tmem_ld32(taddr + c0, va);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              tmem_ld32(lane_addr + (((uint32_t)pg.ks_col[0] + flip) & 511u) + c0, vb);
#pragma unroll
              for (int e = 0; e < 32; ++e) va[e] = __float_as_uint(__uint_as_float(va[e]) * acc_s);
#pragma unroll
              for (int i = 0; i < CHAIN_MAX_V; ++i)
                if (i < V) {
                  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                  const float a_i = am[i];
#pragma unroll
                  for (int e = 0; e < 32; ++e) va[e] = __float_as_uint(fmaf(a_i, __uint_as_float(vb[e]), __uint_as_float(va[e])));
                  if (i + 1 < V) tmem_ld32(lane_addr + (((uint32_t)pg.ks_col[i + 1] + flip) & 511u) + c0, vb);
                }
              emit(va, c0);
            }
          }
          // Publication: a CTA-scope release of a counter, nothing else.  The loader, having acquired
          // it, executes the gpu-scope fence (cumulative over the stores it has thereby observed) and the
          // proxy fence before its bulk copy reads the tile -- one fence per dependency on an otherwise
          // idle thread instead of one per warp and job on the epilogue's critical path.
          if (pg.dbg & 512) {
            TH_TIMED(2, __threadfence(); fence_proxy_async_all());
          }
          __syncwarp();
          if (lane == 0) add_release_local(cnt_job + 4 * j);
        } else if (epi == EPI_SCORES) {
          // A[i][j] = (KP_i + b0) . (KS_j + b1) / sqrt(128) for i = this job's view; the key embeds of
          // every view j sit in TMEM (EPI_KEEP jobs).  One thread = one point, no shuffles; the two
          // warps of a lane quadrant take 64 of the 128 key channels each and group 1 hands its
          // partial sums over through shared memory.
          float sc[TH_MAX_VIEWS];
          const float s_kp = __ldg(jb.acc_scale), s_ks = __ldg(jb.acc_scale2);
#pragma unroll
          for (int jv = 0; jv < TH_MAX_VIEWS; ++jv) sc[jv] = 0.f;
#pragma unroll 1
          for (int c0 = 64 * grp; c0 < 64 * grp + 64; c0 += 32) {
            uint32_t kp[32];
            tmem_ld32(taddr + c0, kp);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 bq = *reinterpret_cast<const float4*>(bias_s + c0 + e);
              kp[e + 0] = __float_as_uint(fmaf(__uint_as_float(kp[e + 0]), s_kp, bq.x));
              kp[e + 1] = __float_as_uint(fmaf(__uint_as_float(kp[e + 1]), s_kp, bq.y));
              kp[e + 2] = __float_as_uint(fmaf(__uint_as_float(kp[e + 2]), s_kp, bq.z));
              kp[e + 3] = __float_as_uint(fmaf(__uint_as_float(kp[e + 3]), s_kp, bq.w));
            }
#pragma unroll
            for (int jv = 0; jv < TH_MAX_VIEWS; ++jv)
              if (jv < V) {
                uint32_t ks[32];
                tmem_ld32(lane_addr + (((uint32_t)pg.ks_col[jv] + flip) & 511u) + c0, ks);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                  const float4 bq = *reinterpret_cast<const float4*>(bias_s + 128 + c0 + e);
                  sc[jv] = fmaf(__uint_as_float(kp[e + 0]), fmaf(__uint_as_float(ks[e + 0]), s_ks, bq.x), sc[jv]);
                  sc[jv] = fmaf(__uint_as_float(kp[e + 1]), fmaf(__uint_as_float(ks[e + 1]), s_ks, bq.y), sc[jv]);
                  sc[jv] = fmaf(__uint_as_float(kp[e + 2]), fmaf(__uint_as_float(ks[e + 2]), s_ks, bq.z), sc[jv]);
                  sc[jv] = fmaf(__uint_as_float(kp[e + 3]), fmaf(__uint_as_float(ks[e + 3]), s_ks, bq.w), sc[jv]);
                }
              }
          }
          if (grp == 1) {
#pragma unroll
            for (int jv = 0; jv < TH_MAX_VIEWS; ++jv) s_part[et * 4 + jv] = sc[jv];
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
          if (grp == 0) {
#pragma unroll
          for (int jv = 0; jv < TH_MAX_VIEWS; ++jv) sc[jv] += s_part[et * 4 + jv];
          const int i = jb.view;
#pragma unroll
          for (int jv = 0; jv < TH_MAX_VIEWS; ++jv)
            if (jv < V) atab[i * V + jv] = __fdiv_rn(sc[jv], 11.313708498984761f);
          if (i == V - 1) {
            // softmax over i for every j (dim=1 of (P, V_i, V_j), cross_transformer.py:144)
            for (int jv = 0; jv < V; ++jv) {
              float a[TH_MAX_VIEWS], m = -3.4e38f, sum = 0.f;
              for (int ii = 0; ii < V; ++ii) {
                a[ii] = atab[ii * V + jv];
                m = fmaxf(m, a[ii]);
              }
              for (int ii = 0; ii < V; ++ii) {
                a[ii] = expf(a[ii] - m);
                sum += a[ii];
              }
              for (int ii = 0; ii < V; ++ii) atab[ii * V + jv] = __fdiv_rn(a[ii], sum);
            }
            __syncwarp();
            if (lane == 0) add_release_local(cnt_scores);
          }
          }  // grp == 0
        } else if (grp != 0) {
          // the per-point head epilogues below are group 0's
        } else if (epi == EPI_ALPHA) {
          // alpha = relu(O) . alpha_fc + b (cross_transformer.py:324-328): O never leaves the SM
          float acc = 0.f;
          const float acc_s = __ldg(jb.acc_scale);
#pragma unroll 1
          for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 32; ++e)
              acc = fmaf(fmaxf(fmaf(__uint_as_float(v[e]), acc_s, bias_s[c0 + e]), 0.f), __ldg(pg.afc_w + c0 + e), acc);
          }
          alpha_reg = acc + __ldg(pg.afc_b);
          if (pt < pg.P) {
            if (pg.alpha) pg.alpha[pt] = alpha_reg;
            if (pg.alpha_only) {
              const int64_t dst = pg.dst_ids ? (int64_t)pg.dst_ids[pg.first + pt] : pg.first + pt;
              if (pg.alpha_out) pg.alpha_out[dst] = alpha_reg;
              if (pg.raw) reinterpret_cast<float4*>(pg.raw)[dst] = make_float4(0.f, 0.f, 0.f, alpha_reg);
            }
          }
        } else if (epi == EPI_RGB) {
          // rgb = relu(T) rgb_fc^T + b (cross_transformer.py:349-351); raw = (rgb, alpha)
          float o0 = 0.f, o1 = 0.f, o2 = 0.f;
          const float acc_s = __ldg(jb.acc_scale);
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float t = fmaxf(fmaf(__uint_as_float(v[e]), acc_s, bias_s[c0 + e]), 0.f);
              o0 = fmaf(t, __ldg(pg.rgb_w + c0 + e), o0);
              o1 = fmaf(t, __ldg(pg.rgb_w + 128 + c0 + e), o1);
              o2 = fmaf(t, __ldg(pg.rgb_w + 256 + c0 + e), o2);
            }
          }
          const bool live = pt < pg.P;
          if (live) {
            o0 += __ldg(pg.rgb_b);
            o1 += __ldg(pg.rgb_b + 1);
            o2 += __ldg(pg.rgb_b + 2);
            if (pg.zero_rgb && !(alpha_reg > 0.f)) o0 = o1 = o2 = 0.f;
            const int64_t dst = pg.dst_ids ? (int64_t)pg.dst_ids[pg.first + pt] : pg.first + pt;
            if (pg.raw) reinterpret_cast<float4*>(pg.raw)[dst] = make_float4(o0, o1, o2, alpha_reg);
          }
          if (pg.cmp.rgb_map) {
            // Fused compositing (raw2outputs, nerf_net_utils.py:14-59): the rows of this tile are whole rays
            // (S | 128, the chunk starts on a ray boundary).  Every thread forms the terms of its own sample --
            // alpha from sigma and the interval, the sigmoids -- and the thread that owns a ray's first sample walks
            // the ray in sample order with the transmittance in a register: the same two functions, in the same
            // order, as k_integrate.
            const int S = pg.cmp.S;
            const int s = et % S;
            const int64_t ray = (pg.first + pt - s) / S;
            float* sc = reinterpret_cast<float*>(smem_raw + COMP_OFF) + et * 5;
            if (live) {
              const float nrm = norm3(__ldg(pg.cmp.ray_d + ray * 3), __ldg(pg.cmp.ray_d + ray * 3 + 1),
                                      __ldg(pg.cmp.ray_d + ray * 3 + 2));
              const float near_ = __ldg(pg.cmp.near_ + ray), far_ = __ldg(pg.cmp.far_ + ray);
              const float z = sample_z(near_, far_, __ldg(pg.cmp.t_vals + s));
              float dist = s + 1 < S ? __fsub_rn(sample_z(near_, far_, __ldg(pg.cmp.t_vals + s + 1)), z) : 1e10f;
              dist = __fmul_rn(dist, nrm);
              const SampleTerm t = composite_sample(make_float4(o0, o1, o2, alpha_reg), dist);
              sc[0] = t.alpha;
              sc[1] = t.r;
              sc[2] = t.g;
              sc[3] = t.b;
              sc[4] = z;
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");  // the four group-0 warps
            if (live && s == 0) {
              RayAcc a = ray_acc_init();
              for (int i = 0; i < S; ++i) {
                const float* q5 = sc + i * 5;
                composite_step(a, SampleTerm{q5[0], q5[1], q5[2], q5[3]}, q5[4]);
              }
              const float bg = pg.cmp.white_bkgd ? 1.0f - a.acc : 0.f;
              pg.cmp.rgb_map[ray * 3] = pg.cmp.white_bkgd ? a.r + bg : a.r;
              pg.cmp.rgb_map[ray * 3 + 1] = pg.cmp.white_bkgd ? a.g + bg : a.g;
              pg.cmp.rgb_map[ray * 3 + 2] = pg.cmp.white_bkgd ? a.b + bg : a.b;
              pg.cmp.acc_map[ray] = a.acc;
              pg.cmp.depth_map[ray] = a.depth;
            }
            // (the terms are overwritten by the next unit's fc_4' epilogue, ~20 named barriers of all epilogue
            // warps later: the walking threads' warps take part in every one of them)
          }
        }
        // this warp is done with the accumulator of job G
        if (stats) tw[3 + ((epi == EPI_IMG || epi == EPI_MIX) ? 0 : epi == EPI_SCORES ? 1 : 2)] += clock64() - t_work;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0)
            add_release_local(cnt_epi);
          else if (pg.dbg & 1024)  // A/B knob: the former release form
            add_release_remote(cnt_epi_peer, 0);
          else
            add_relaxed_remote(cnt_epi_peer, 0);
        }
        ++G;
      }
    }
  }

  if (stats && lane == 0 && (warp == LOADER_WARP || warp == MMA_WARP || warp == EPI_WARP0 || warp == 0)) {
    // slots: loader 0-5, MMA/relay 8-13, epilogue warp 10: 16-21, mix warp 0: 24-29; slot +6 = total cycles
    const int o = warp == LOADER_WARP ? 0 : warp == MMA_WARP ? 8 : warp == EPI_WARP0 ? 16 : 24;
    for (int i = 0; i < 6; ++i) stats[o + i] = (unsigned long long)tw[i];
    stats[o + 6] = (unsigned long long)(clock64() - t_begin);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------
// host side: program construction
// ---------------------------------------------------------------------------
struct Builder {
  Program pg{};
  int V;
  int64_t tiles_per_view;  // Pp / 128
  Builder(int V_, int64_t Pp) : V(V_), tiles_per_view(Pp / 128) {
    pg.V = V_;
    pg.kp_col = 128 * V_;
    for (int v = 0; v < V_; ++v) pg.ks_col[v] = 128 * v;
  }
  uint32_t slotA(int v) const { return (uint32_t)v * SCR_ACT; }
  uint32_t slotB(int v) const { return (uint32_t)(V + v) * SCR_ACT; }
  Seg in_view(const float* buf, int C, int v) const {  // chunk-level image of a (V*Pp, C) activation
    Seg s{};
    s.img = reinterpret_cast<const unsigned char*>(buf);
    s.tile_off = v * tiles_per_view;
    s.kbs = C / 64;
    s.dep = -1;
    return s;
  }
  Seg in_point(const float* buf, int C) const { return in_view(buf, C, 0); }
  static Seg scr(uint32_t off, int C, int dep, int dep_mix = 0) {
    Seg s{};
    s.scratch_off = off;
    s.kbs = C / 64;
    s.dep = dep;
    s.dep_mix = dep_mix;
    return s;
  }
  // returns the job index
  int add(std::initializer_list<Seg> segs, const unsigned char* wimg, const float* bias, int N, int relu, int epi,
          uint32_t out_off, int view = 0, int col = -1, int wait_back = 2) {
    const int j = pg.njobs++;
    Job& jb = pg.job[j];
    jb = Job{};
    for (const Seg& s : segs) {
      jb.seg[jb.nseg++] = s;
      jb.nkb += s.kbs;
    }
    jb.wimg = wimg;
    jb.bias = bias;
    jb.N = N;
    jb.relu = relu;
    jb.epi = epi;
    jb.out_off = out_off;
    jb.view = view;
    jb.tmem_col = col >= 0 ? col : (j & 1) * 256;
    jb.wait_back = wait_back;
    jb.reader = -1;
    return j;
  }
};

}  // namespace chain

// Bytes of scratch one launch over P points touches (2 CTAs per 256-point unit, at most one CTA per SM).
size_t chain_scratch_bytes(int64_t P, int V, int num_sms) {
  const int64_t units = pad_points(P) / 256;
  const int64_t ctas = 2 * units < num_sms ? 2 * units : num_sms;
  return (size_t)ctas * chain::scratch_stride(V);
}
bool chain_supported(int V) { return V >= 1 && V <= chain::CHAIN_MAX_V; }

namespace chain {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// rows of 64 fp16 over [base, base + bytes), box = 64 x box_rows, no swizzle (the images are pre-swizzled)
static int make_rows_map(CUtensorMap* m, const void* base, size_t bytes, int box_rows) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
      set_error("mlp_forward_chain: cuTensorMapEncodeTiled is not available");
      return TH_ECUDA;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[2] = {64, (cuuint64_t)(bytes / 128)};
  const cuuint64_t gstride[1] = {128};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mlp_forward_chain: cuTensorMapEncodeTiled failed (%d) base %p bytes %zu box %d", (int)r, base, bytes,
              box_rows);
    return TH_ECUDA;
  }
  return TH_OK;
}
}  // namespace chain

// The whole per-point network for one chunk in ONE launch.  Inputs (rep, pix, pix_mean, vd)
// are the tile images the feature kernel wrote; `scratch` holds chain_scratch_bytes().
int mlp_forward_chain(const MlpRun& run, const MlpBuffers& b, const PackedHeader& h, unsigned char* scratch,
                      float* alpha, cudaStream_t st) {
  using namespace chain;
  ProfScope prof_(PROF_GEMM, st);
  const int64_t P = run.P, Pp = pad_points(P);
  const int V = run.V;
  if (!chain_supported(V)) {
    set_error("mlp_forward_chain: V=%d needs more than 512 TMEM columns for the key embeds", V);
    return TH_EUNSUPPORTED;
  }
  auto wimg = [&](uint64_t off) { return run.weights + off; };
  auto wf = [&](uint64_t off) { return reinterpret_cast<const float*>(run.weights + off); };
  Builder B(V, Pp);
  int j_s[TH_MAX_VIEWS], j_x[TH_MAX_VIEWS], j_n1[TH_MAX_VIEWS], j_int[TH_MAX_VIEWS], j_g[TH_MAX_VIEWS];
  // Pre-mapped inputs (experimental): the feature kernel has already blended alpha_res_0 F + b, so X_v
  // arrives as a 256-wide image and is only copied into the scratch (identity weights, exact: hi + lo
  // of the operand times 1.0); P2_v (128 wide, behind X in the pix block) and R (128 wide, in the
  // pix_mean block) enter view_fc' / fc_4' through identity blocks of W_gvfp / W_tp.
  const bool pre = run.premapped != 0;
  const float* p2_img = b.pix + (size_t)V * Pp * 256;
  // front: X_v = relu(alpha_res_0 pix_v) -> slot B ; S_v = relu(fc_0 rep_v) -> slot A
  // TH_PREMAP_COPY=1 keeps the first form of the pre-mapped program (X copied into slot B by identity jobs)
  static const bool pre_copy = getenv("TH_PREMAP_COPY") && atoi(getenv("TH_PREMAP_COPY"));
  const bool x_in_chunk = pre && !pre_copy;
  for (int v = 0; v < V; ++v)
    j_x[v] = x_in_chunk ? -1
             : pre      ? B.add({B.in_view(b.pix, 256, v)}, wimg(h.h_xid), nullptr, 256, 1, EPI_IMG, B.slotB(v))
                        : B.add({B.in_view(b.pix, PIX_LD, v)}, wimg(h.h_ar0), wf(h.ar0_b), 256, 1, EPI_IMG, B.slotB(v));
  for (int v = 0; v < V; ++v)
    j_s[v] = B.add({B.in_view(b.rep, REP_LD, v)}, wimg(h.h_fc0), wf(h.fc0_b), 256, 1, EPI_IMG, B.slotA(v));
  // key embeds: KS_v = key_embed_1 S_v stays in TMEM; KP_v = key_embed_0 X_v is consumed by the score epilogue
  if (x_in_chunk && V == 3) {
    // KS_0 is issued right after S_2 (accumulator columns 0..255, still being drained by its epilogue): give the
    // first two kept key embeds the columns of S_1's accumulator, which was drained a job earlier
    // (TH_CHAIN_STATS: KS_0 waited 10 kcycles per unit for TMEM at columns 0..127)
    B.pg.ks_col[0] = 256;
    B.pg.ks_col[1] = 384;
    B.pg.ks_col[2] = 0;
    B.pg.kp_col = 128;
  }
  // Two forms of the attention mix in the pre-mapped program (both parity-green, equally fast within noise,
  // profiles/README.md r2g-r2j): in place on the fp16 operands by the six mix warps (default: least DRAM traffic), or
  // on the accumulator side in TMEM (TH_CHAIN_MIX=tmem): see the job list below
  // run.premapped: 1 = default (env TH_CHAIN_MIX = "tmem" | "inplace" decides, read per call so that tests can
  // toggle it), 2 = TMEM-side mix, 3 = in-place mix
  const char* mix_env = getenv("TH_CHAIN_MIX");
  const bool tmix = x_in_chunk && (run.premapped == 2 || (run.premapped == 1 && mix_env && !strcmp(mix_env, "tmem")));
  uint64_t mix_y_img[MAX_JOBS] = {0};  // EPI_MIX jobs: weight image of the Y jobs they combine (for acc_scale2)
  int j_ks[TH_MAX_VIEWS];
  for (int v = 0; v < V; ++v)
    j_ks[v] = B.add({Builder::scr(B.slotA(v), 256, j_s[v])}, wimg(h.h_k1), nullptr, 128, 0, EPI_KEEP, 0, v,
                    B.pg.ks_col[v], 2);
  for (int v = 0; v < V; ++v) {
    Seg x = x_in_chunk ? B.in_view(b.pix, 256, v) : Builder::scr(B.slotB(v), 256, j_x[v]);
    // X is read again: by the Y jobs (TMEM-side mix), or -- in-place mix -- by the mix warps ~50 kcycles later, which
    // otherwise find it evicted (the loader's default for chunk inputs is evict-first).  TH_CHAIN_XKEEP=0: A/B knob.
    static const bool xkeep = !(getenv("TH_CHAIN_XKEEP") && atoi(getenv("TH_CHAIN_XKEEP")) == 0);
    x.keep = (tmix || (x_in_chunk && xkeep)) ? 1 : 0;
    const int j = B.add({x}, wimg(h.h_k0), wf(h.k0_b), 128, 0, EPI_SCORES, 0, v, B.pg.kp_col, v == 0 ? 2 : 1);
    B.pg.job[j].bias2 = wf(h.k1_b);
    if (v == V - 1)
      for (int u = 0; u < V; ++u) B.pg.job[j_ks[u]].reader = j;  // the kept key embeds are read by every score job
  }
  if (tmix) {
    // N1_j = relu(W_s S_j + sum_i A[i][j] (W_x X_i) + b): the X part of fc_1' applied BEFORE the mix, so the mix
    // becomes an fp32 combination of accumulators in the epilogue instead of a rewrite of fp16 operands that the
    // tensor pipe has to wait for (TH_CHAIN_STATS, in-place mix: fc_1' of view 0 waited 51 of the unit's 274
    // kcycles; without any mix the launch is 18 % shorter).  512 TMEM columns hold three kept Y_i = W_x X_i and one
    // Z_j = W_s S_j of 128 output channels each, so fc_1' runs as two halves of 128 rows: per half
    //   Y_0h Y_1h Y_2h (N = 128, K = 256, kept at the key-embed columns), then Z_0h Z_1h Z_2h (N = 128, K = 256, at the
    //   score column) whose epilogue (EPI_MIX) writes channels [128h, 128h + 128) of N1_j into slot B(j).
    // Same MACs as fc_1' with K = 512; X stays pristine in the chunk image (no in-place writes, no mix fences).
    B.pg.has_mix = 0;
    const uint64_t hs[2] = {h.h_fc1s0, h.h_fc1s1}, hx[2] = {h.h_fc1x0, h.h_fc1x1}, bs[2] = {h.fc1s0_b, h.fc1s1_b};
    int j_z[TH_MAX_VIEWS][2];
    for (int hh = 0; hh < 2; ++hh) {
      int j_y[TH_MAX_VIEWS];
      for (int i = 0; i < V; ++i) {
        Seg x = B.in_view(b.pix, 256, i);
        x.keep = hh == 0 ? 1 : 0;
        j_y[i] = B.add({x}, wimg(hx[hh]), nullptr, 128, 0, EPI_KEEP, 0, i, B.pg.ks_col[i], 2);
      }
      for (int jv = 0; jv < V; ++jv) {
        j_z[jv][hh] = B.add({Builder::scr(B.slotA(jv), 256, j_s[jv])}, wimg(hs[hh]), wf(bs[hh]), 128, 1, EPI_MIX,
                            B.slotB(jv) + 2u * (uint32_t)hh * TILE_IMG, jv, B.pg.kp_col, 1);
        mix_y_img[j_z[jv][hh]] = hx[hh];
      }
      for (int i = 0; i < V; ++i) B.pg.job[j_y[i]].reader = j_z[V - 1][hh];
    }
    // INTER_v = relu(fc_2 N1_v): N1_v from slot B(v) (two halves, two producers), INTER_v over the dead S_v
    for (int v = 0; v < V; ++v)
      j_int[v] = B.add({Builder::scr(B.slotB(v), 128, j_z[v][0]), Builder::scr(B.slotB(v) + 2u * TILE_IMG, 128, j_z[v][1])},
                       wimg(h.h_fc2), wf(h.fc2_b), 256, 1, EPI_IMG, B.slotA(v));
  }
  // Deferred tail (pre-mapped program with the in-place mix; OPT-IN with TH_CHAIN_DEFER=1): the jobs after fc_2 --
  // fc_3 + alpha head, view_fc' per view, fc_4' + rgb head, ~40 kcycles of tensor work per unit -- are issued one
  // iteration LATE, between the score jobs and fc_1' of the NEXT unit, i.e. into the ~45 kcycles in which the tensor
  // pipe waits for the mix warps (TH_CHAIN_STATS).  INTER of the previous unit must then outlive S / N1 of the
  // current one: slot set A exists twice and alternates with the unit's parity.  Parity-green (also under role
  // jitter) and it does close the window (fc_1' of view 0 waits 8 instead of 46 kcycles), but the launch is NOT
  // faster: the kernel moves ~68 KB per point through L2 at ~85 % of the chip's L2 throughput cap, so the tail's
  // operand stream and the mix now contend (mix 50 -> 68 kcycles per unit, fc_3 waits 20 kcycles for operands), and the
  // larger scratch costs L2 hits (71 % -> 59 %, DRAM 6.8 -> 9.3 GB per launch).  Measured in profiles/README.md (r2h-r2j).
  // TH_CHAIN_DEFER=2, "deep" deferral: EVERYTHING behind the scores -- fc_1', fc_2 and the tail -- runs one iteration
  // late, i.e. iteration `it` issues the front of unit `it` (fc_0, key embeds, scores) and then fc_1' ... fc_4' of unit
  // `it - 1`, whose in-place mix has had the whole previous back half to finish: the tensor pipe never waits for the
  // mix warps.  Same job order as the plain program, only the unit the back half works on changes; the attention
  // table alternates with the unit's parity (the scores of unit `it` are written while the mix of unit `it - 1` may
  // still read its own).
  const char* defer_env = getenv("TH_CHAIN_DEFER");
  const int defer_mode = (x_in_chunk && !tmix && defer_env) ? atoi(defer_env) : 0;
  const bool defer = defer_mode == 1, deep = defer_mode == 2;
  auto add_fc3 = [&]() {
    const int j = B.pg.njobs++;
    Job& jb = B.pg.job[j];
    jb = Job{};
    for (int v = 0; v < V; ++v) {
      jb.seg[jb.nseg++] = Builder::scr(B.slotA(v), 256, j_int[v]);
      jb.nkb += 4;
    }
    jb.wimg = wimg(h.h_fc3m);
    jb.bias = wf(h.fc3m_b);
    jb.N = 256;
    jb.relu = 1;
    jb.epi = EPI_ALPHA;
    jb.tmem_col = (j & 1) * 256;
    jb.wait_back = 2;
    jb.reader = -1;
  };
  auto add_tail = [&]() {
    if (run.alpha_only) {
      add_fc3();
      return;
    }
    auto add_gvf = [&](int v) {
      j_g[v] = pre ? B.add({Builder::scr(B.slotA(v), 256, j_int[v]), B.in_view(p2_img, 128, v), B.in_point(b.vd, 64)},
                           wimg(h.h_gvfp), wf(h.gvfp_b), 128, 1, EPI_IMG, B.slotB(v))
                   : B.add({Builder::scr(B.slotA(v), 256, j_int[v]), B.in_view(b.pix, PIX_LD, v),
                            B.in_point(b.vd, 64)},
                           wimg(h.h_gvf), wf(h.gvf_b), 128, 1, EPI_IMG, B.slotB(v));
    };
    for (int v = 0; v + 1 < V; ++v) add_gvf(v);
    add_fc3();
    add_gvf(V - 1);
    {  // T = relu([G_0 | ... | mean pix or R] W_t^T)
      const int j = B.pg.njobs++;
      Job& jb = B.pg.job[j];
      jb = Job{};
      for (int v = 0; v < V; ++v) {
        jb.seg[jb.nseg++] = Builder::scr(B.slotB(v), 128, j_g[v]);
        jb.nkb += 2;
      }
      const int c_tail = pre ? 128 : PIX_LD;  // R (pre-mapped) or the view mean of pix
      jb.seg[jb.nseg++] = B.in_point(b.pix_mean, c_tail);
      jb.nkb += c_tail / 64;
      jb.wimg = wimg(pre ? h.h_tp : h.h_t);
      jb.bias = wf(pre ? h.tp_b : h.t_b);
      jb.N = 128;
      jb.relu = 1;
      jb.epi = EPI_RGB;
      jb.tmem_col = (j & 1) * 256;
      jb.wait_back = 2;
      jb.reader = -1;
    }
  };
  const int first_tail = B.pg.njobs;
  int n_tail = 0;
  if (defer) {
    n_tail = run.alpha_only ? 1 : V + 2;
    for (int v = 0; v < V; ++v) j_int[v] = first_tail + n_tail + V + v;  // the indices fc_2's jobs WILL get
    add_tail();
    if (B.pg.njobs != first_tail + n_tail) {
      set_error("mlp_forward_chain: internal error (tail job count)");
      return TH_EINVAL;
    }
  }
  if (!tmix) {
    B.pg.has_mix = 1;
    // N1_v = relu([S_v | XT_v] W_fc1f^T) in place of S_v; INTER_v = relu(fc_2 N1_v) in place again
    for (int v = 0; v < V; ++v) {
      Seg xt = x_in_chunk ? B.in_view(b.pix, 256, v) : Builder::scr(B.slotB(v), 256, -1, 1);
      xt.dep_mix = 1;  // mixed in place (scratch slot B, or the chunk image), released per k-block
      j_n1[v] = B.add({Builder::scr(B.slotA(v), 256, -1), xt}, wimg(h.h_fc1f), wf(h.fc1f_b), 256, 1, EPI_IMG,
                      B.slotA(v), v, -1, v == 0 ? 1 : 2);
    }
    if (x_in_chunk) {
      B.pg.x_img = reinterpret_cast<unsigned char*>(const_cast<float*>(b.pix));
      B.pg.x_view_stride = (int64_t)(Pp / 128) * 4 * TILE_IMG;
    }
    for (int v = 0; v < V; ++v) {
      const int j = B.add({Builder::scr(B.slotA(v), 256, j_n1[v])}, wimg(h.h_fc2), wf(h.fc2_b), 256, 1, EPI_IMG,
                          B.slotA(v));
      if (defer && j != j_int[v]) {
        set_error("mlp_forward_chain: internal error (fc_2 job index)");
        return TH_EINVAL;
      }
      j_int[v] = j;
    }
  }
  if (!defer) add_tail();
  if (defer || deep) {
    if (deep) n_tail = B.pg.njobs - first_tail;  // fc_1' ... fc_4': every job behind the front
    B.pg.deferred_tail = 1;
    const uint32_t a_end = (uint32_t)V * SCR_ACT, b_end = 2u * a_end;
    auto remap = [&](uint32_t off, int32_t* aset) {   // slot A -> parity-alternating set; slot B -> behind both A sets
      if (off < a_end) {
        *aset = 1;
        return off;
      }
      const uint32_t v = (off - a_end) / SCR_ACT, in = (off - a_end) % SCR_ACT;
      *aset = 0;
      return b_end + v * (SCR_ACT / 2) + in;
    };
    for (int j = 0; j < B.pg.njobs; ++j) {
      Job& jb = B.pg.job[j];
      jb.shift = (j >= first_tail && j < first_tail + n_tail) ? -1 : 0;
      for (int sgi = 0; sgi < jb.nseg; ++sgi)
        if (!jb.seg[sgi].img) jb.seg[sgi].scratch_off = remap(jb.seg[sgi].scratch_off, &jb.seg[sgi].aset);
      if (jb.epi == EPI_IMG) jb.out_off = remap(jb.out_off, &jb.out_aset);
      if (jb.tmem_col == (((j - (jb.shift ? 0 : 0)) & 1) * 256) && jb.epi != EPI_KEEP && jb.epi != EPI_SCORES) {
        // auto columns stay (j & 1) * 256 of the FINAL position (Builder::add used the final index already)
      }
    }
  }
  Program& pg = B.pg;
  for (int j = 0; j < pg.njobs; ++j) {  // accumulator scales of the weight images (PackedHeader::img_inv_scale)
    pg.job[j].acc_scale = img_inv_scale_ptr(run.weights, h, (uint64_t)(pg.job[j].wimg - run.weights));
    pg.job[j].acc_scale2 = img_inv_scale_ptr(run.weights, h, mix_y_img[j] ? mix_y_img[j] : h.h_k1);
    if (!pg.job[j].acc_scale || !pg.job[j].acc_scale2) {
      set_error("mlp_forward_chain: job %d reads a weight image the blob header does not list", j);
      return TH_EINVAL;
    }
  }
  if (x_in_chunk) {
    // TMEM waits, derived: the MMA issuer may start executed job number G once the epilogue of executed job number
    // G - wait_back is done.  Walk the sequence the kernel executes (four units of one cluster, with the skipped
    // jobs of the first iteration and the draining iteration of a deferred tail), and for every job take the
    // smallest distance to the last reader of any accumulator whose columns it overwrites: its own epilogue, or
    // for a kept accumulator the last score / mix epilogue of its unit.
    const int n = pg.njobs, n_units = 4, n_iter = n_units + (pg.deferred_tail ? 1 : 0);
    const uint32_t flip_on_h = (n & 1) ? 256u : 0u;
    struct Exec { int j, it, ui; };
    std::vector<Exec> seq;
    for (int it = 0; it < n_iter; ++it)
      for (int j = 0; j < n; ++j) {
        const int ui = it + pg.job[j].shift;
        if (ui >= 0 && ui < n_units) seq.push_back({j, it, ui});
      }
    auto exec_index = [&](int j, int ui) {
      for (size_t e = 0; e < seq.size(); ++e)
        if (seq[e].j == j && seq[e].ui == ui) return (int)e;
      return -1;
    };
    std::vector<int> wb(n, 1 << 20);
    for (size_t G = 0; G < seq.size(); ++G) {
      const Job& jg = pg.job[seq[G].j];
      const uint32_t c0 = ((uint32_t)jg.tmem_col + ((seq[G].it & 1) ? flip_on_h : 0u)) & 511u, c1 = c0 + jg.N;
      for (size_t A = 0; A < G; ++A) {
        const Job& ja = pg.job[seq[A].j];
        const uint32_t a0 = ((uint32_t)ja.tmem_col + ((seq[A].it & 1) ? flip_on_h : 0u)) & 511u, a1 = a0 + ja.N;
        if (!(a0 < c1 && c0 < a1)) continue;
        const int R = ja.reader >= 0 ? exec_index(ja.reader, seq[A].ui) : (int)A;
        if (R < 0 || R >= (int)G) {
          set_error("mlp_forward_chain: job %d overwrites TMEM columns of job %d before its reader has run", seq[G].j,
                    seq[A].j);
          return TH_EINVAL;
        }
        if ((int)G - R < wb[seq[G].j]) wb[seq[G].j] = (int)G - R;
      }
    }
    for (int j = 0; j < n; ++j) pg.job[j].wait_back = wb[j] < 1 ? 1 : wb[j];
  }
  if (run.program_dump) {  // host-side test hook: the program as built, nothing launched
    pg.afc_w = wf(h.afc_w);
    pg.afc_b = wf(h.afc_b);
    pg.rgb_w = wf(h.rgb_w);
    pg.rgb_b = wf(h.rgb_b);
    pg.alpha_only = run.alpha_only;
    pg.num_units = (int)(Pp / 256);
    *static_cast<Program*>(run.program_dump) = pg;
    return TH_OK;
  }
  {
    // A operands: everything from the first chunk input (rep) to the end of the view-direction image
    // lies in one workspace block (mlp_carve), the scratch included
    const unsigned char* a0 = reinterpret_cast<const unsigned char*>(b.rep);
    const unsigned char* a1 = reinterpret_cast<const unsigned char*>(b.vd) + (size_t)Pp * 2 * VD_LD * 4;
    int rc;
    if ((rc = make_rows_map(&pg.tm_a, a0, (size_t)(a1 - a0), 256))) return rc;
    if ((rc = make_rows_map(&pg.tm_b256, run.weights, (size_t)h.total_bytes, 256))) return rc;
    if ((rc = make_rows_map(&pg.tm_b128, run.weights, (size_t)h.total_bytes, 128))) return rc;
    pg.a_base = a0;
    pg.w_base = run.weights;
  }
  pg.scratch = scratch;
  pg.afc_w = wf(h.afc_w);
  pg.afc_b = wf(h.afc_b);
  pg.rgb_w = wf(h.rgb_w);
  pg.rgb_b = wf(h.rgb_b);
  pg.alpha = alpha;
  pg.raw = run.raw;
  pg.alpha_out = run.alpha_out;
  pg.dst_ids = run.dst_ids;
  pg.first = run.first;
  pg.P = P;
  pg.zero_rgb = run.zero_rgb_if_transparent;
  pg.alpha_only = run.alpha_only;
  pg.num_units = (int)(Pp / 256);
  pg.cmp = run.cmp;
  if (pg.cmp.rgb_map && (run.alpha_only || run.dst_ids || pg.cmp.S < 1 || 128 % pg.cmp.S || run.first % pg.cmp.S ||
                         P % pg.cmp.S)) {
    set_error("mlp_forward_chain: fused compositing needs dense rays, S | 128 and chunks of whole rays (S %d, first %lld, P %lld)",
              pg.cmp.S, (long long)run.first, (long long)P);
    return TH_EINVAL;
  }

  int num_sms = 0;
  if (device_sm_count(&num_sms)) return TH_ECUDA;
  // per-device attribute: set on every launch (cheap) rather than caching a process-global flag
  TH_CUDA(cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  const int nclusters = pg.num_units < num_sms / 2 ? pg.num_units : num_sms / 2;
  static const bool want_stats = getenv("TH_CHAIN_STATS") != nullptr;
  const char* dbg_env = getenv("TH_CHAIN_DBG");  // read per launch so that a test can toggle it
  pg.dbg = dbg_env ? atoi(dbg_env) : 0;
  static unsigned long long* d_stats_dev[64] = {nullptr};  // debug only (TH_CHAIN_STATS), one buffer per device
  unsigned long long* d_stats = nullptr;
  if (want_stats) {
    int dev = 0;
    TH_CUDA(cudaGetDevice(&dev));
    dev &= 63;
    if (!d_stats_dev[dev]) TH_CUDA(cudaMalloc(&d_stats_dev[dev], (size_t)num_sms * STATS_PER_CTA * 8));
    d_stats = d_stats_dev[dev];
    TH_CUDA(cudaMemsetAsync(d_stats, 0, (size_t)num_sms * STATS_PER_CTA * 8, st));
    pg.stats = d_stats;
  }
  // TH_CHAIN_L2WIN=1 (experiment, profiles/README.md): an access-policy window over the per-CTA scratch asks L2 to
  // keep those lines (persisting) while the chunk inputs stream past them.
  static const int l2win = getenv("TH_CHAIN_L2WIN") ? atoi(getenv("TH_CHAIN_L2WIN")) : 0;
  if (l2win) {
    int dev = 0, max_persist = 0, max_win = 0;
    TH_CUDA(cudaGetDevice(&dev));
    TH_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
    TH_CUDA(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev));
    const size_t scr = (size_t)(2 * nclusters) * scratch_stride(V);
    const size_t persist = scr < (size_t)max_persist ? scr : (size_t)max_persist;
    static bool limit_set[64] = {false};
    if (!limit_set[dev & 63]) {
      TH_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist));
      limit_set[dev & 63] = true;
      fprintf(stderr, "[chain] L2 window: scratch %zu MB, persisting limit %zu MB (max %d MB), window max %d MB\n",
              scr >> 20, persist >> 20, max_persist >> 20, max_win >> 20);
    }
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeAccessPolicyWindow;
    attr.val.accessPolicyWindow.base_ptr = scratch;
    attr.val.accessPolicyWindow.num_bytes = scr < (size_t)max_win ? scr : (size_t)max_win;
    attr.val.accessPolicyWindow.hitRatio = (float)((double)persist / (double)attr.val.accessPolicyWindow.num_bytes);
    if (attr.val.accessPolicyWindow.hitRatio > 1.f) attr.val.accessPolicyWindow.hitRatio = 1.f;
    attr.val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    TH_CUDA(cudaLaunchKernelEx(&cfg, k_chain, pg));
  } else {
    k_chain<<<2 * nclusters, NUM_THREADS, SMEM_BYTES, st>>>(pg);
  }
  TH_LAUNCHED();
  if (want_stats) {
    static int printed = 0;
    std::vector<unsigned long long> hs((size_t)num_sms * STATS_PER_CTA);
    TH_CUDA(cudaMemcpyAsync(hs.data(), d_stats, hs.size() * 8, cudaMemcpyDeviceToHost, st));
    TH_CUDA(cudaStreamSynchronize(st));
    if (printed++ < 4 || (printed % 64) == 0) {
      const char* names[4][7] = {
          {"dep", "mix", "empty", "issue", "fence", "-", "total"},
          {"epi(tmem)", "full", "pfull", "-", "-", "-", "total"},
          {"bar", "tfull", "store", "w_img", "w_scores", "w_heads", "total"},
          {"scores", "work", "-", "-", "-", "-", "total"}};
      const char* roles[4] = {"loader", "mma/relay", "epilogue w10", "mix w0"};
      for (int r = 0; r < 4; ++r) {
        fprintf(stderr, "[chain stats] %-12s", roles[r]);
        for (int i = 0; i < 7; ++i) {
          if (names[r][i][0] == '-') continue;
          double lead = 0, peer = 0;
          for (int c = 0; c < 2 * nclusters; ++c) (c & 1 ? peer : lead) += (double)hs[(size_t)c * STATS_PER_CTA + r * 8 + i];
          fprintf(stderr, " %s=%.0f/%.0f", names[r][i], lead / nclusters / 1e3, peer / nclusters / 1e3);
        }
        fprintf(stderr, "  (kcycles per CTA, leader/peer; units=%d clusters=%d)\n", pg.num_units, nclusters);
      }
      fprintf(stderr, "[chain stats] MMA waits per job (kcycles per unit: tmem/full/pfull):");
      for (int j = 0; j < pg.njobs; ++j) {
        double w[3] = {0, 0, 0};
        for (int c = 0; c < 2 * nclusters; c += 2)
          for (int k = 0; k < 3; ++k) w[k] += (double)hs[(size_t)c * STATS_PER_CTA + 32 + 3 * j + k];
        const double units = (double)pg.num_units;
        fprintf(stderr, " j%d[N%d K%d]=%.1f/%.1f/%.1f", j, pg.job[j].N, pg.job[j].nkb * 64, w[0] / units / 1e3,
                w[1] / units / 1e3, w[2] / units / 1e3);
      }
      fprintf(stderr, "\n");
    }
  }
  return TH_OK;
}

}  // namespace th

// Host-only test hook (no device needed): the job program mlp_forward_chain builds for a chunk of
// n_points, as a table of int64 so that a CPU test can interpret it against the oracle
// (tests/test_chain_program.py).  Addresses are reported relative to a fictitious workspace base
// (the chunk block of mlp_carve) and to the start of the packed weight blob.
//   header: [njobs, V, has_mix, alpha_only, Pp, scr_act_bytes, tile_img_bytes, deferred_tail]
//   job (16 + 6 * MAX_SEG words): N, relu, epi, out_off, tmem_col, wait_back, view, nseg, nkb, wimg_off,
//        bias_off (-1 = none), bias2_off (-1 = none), reader (-1 = own epilogue), shift, out_aset, 0, then per segment:
//        kind (0 scratch / 1 chunk image), offset (scratch bytes / image byte offset from the chunk base),
//        tile_off, kbs, dep, dep_mix | aset << 1
extern "C" int64_t th_debug_chain_program(const void* packed_host, int32_t n_views, int64_t n_points,
                                          int32_t alpha_only, int32_t premapped, int64_t* table, int64_t capacity) {
  using namespace th;
  using namespace th::chain;
  if (!packed_host || !table || n_points < 1 || !chain_supported(n_views)) return TH_EINVAL;
  PackedHeader h;
  memcpy(&h, packed_host, sizeof(h));
  if (h.magic != PACK_MAGIC || h.n_views != n_views) return TH_EINVAL;
  float* base = reinterpret_cast<float*>(uintptr_t(1) << 32);  // never dereferenced
  MlpBuffers b;
  mlp_carve(base, n_points, n_views, &b);
  MlpRun run{};
  run.weights = static_cast<const unsigned char*>(packed_host);
  run.P = n_points;
  run.V = n_views;
  run.alpha_only = alpha_only;
  run.premapped = premapped;
  run.use_tensor_cores = 1;
  run.inputs_are_images = 1;
  Program pg{};
  run.program_dump = &pg;
  const int rc = mlp_forward_chain(run, b, h, reinterpret_cast<unsigned char*>(b.s), nullptr, nullptr);
  if (rc) return rc;
  const int64_t words = 8 + (int64_t)pg.njobs * (16 + 6 * MAX_SEG);
  if (capacity < words) return TH_EWORKSPACE;
  const unsigned char* wbase = run.weights;
  const unsigned char* cbase = reinterpret_cast<const unsigned char*>(b.rep);
  auto woff = [&](const void* p) { return p ? (int64_t)(static_cast<const unsigned char*>(p) - wbase) : (int64_t)-1; };
  int64_t* t = table;
  t[0] = pg.njobs; t[1] = pg.V; t[2] = pg.has_mix; t[3] = pg.alpha_only; t[4] = pad_points(n_points);
  t[5] = SCR_ACT; t[6] = TILE_IMG; t[7] = pg.deferred_tail;
  t += 8;
  for (int j = 0; j < pg.njobs; ++j) {
    const Job& jb = pg.job[j];
    t[0] = jb.N; t[1] = jb.relu; t[2] = jb.epi; t[3] = jb.out_off; t[4] = jb.tmem_col; t[5] = jb.wait_back;
    t[6] = jb.view; t[7] = jb.nseg; t[8] = jb.nkb; t[9] = woff(jb.wimg); t[10] = woff(jb.bias); t[11] = woff(jb.bias2);
    t[12] = jb.reader; t[13] = jb.shift; t[14] = jb.out_aset; t[15] = 0;
    for (int sgi = 0; sgi < MAX_SEG; ++sgi) {
      const Seg& sg = jb.seg[sgi];
      int64_t* q = t + 16 + 6 * sgi;
      q[0] = sg.img ? 1 : 0;
      q[1] = sg.img ? (int64_t)(sg.img - cbase) : (int64_t)sg.scratch_off;
      q[2] = sg.tile_off; q[3] = sg.kbs; q[4] = sg.dep; q[5] = sg.dep_mix | (sg.aset << 1);
    }
    t += 16 + 6 * MAX_SEG;
  }
  return words;
}
