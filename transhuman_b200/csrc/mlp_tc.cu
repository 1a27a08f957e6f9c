// tcgen05 GEMM for the per-point network's layers (rows a9/a10), sm_100a.
//
//   C[M,N] = act( sum_seg A_seg[M,K_seg] . W[:, seg]^T + bias ),  fp32 in / fp32 out
//
// Precision: a single-pass TF32/BF16 product misses the <= 1e-4 RGB bar
// (SURVEY finding 2), so every operand is split into two fp16 terms,
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 significant bits), and
// three tensor-core products are accumulated in fp32 in TMEM:
//     D += A_hi B_hi ;  D += A_lo B_hi ;  D += A_hi B_lo        (lo*lo ~ 2^-22 dropped)
// Weights are split once on the host (th_pack_weights) and stored as ready-made
// shared-memory tile images (128B-swizzled, K-major), so a k-block of B is one
// cp.async.bulk.  An A segment is either fp32 rows (split on the fly by the
// producer warps) or an activation that a previous GEMM's epilogue already wrote
// in the same tile-image format (one 32 KB cp.async.bulk per k-block, no LSU work).
//
// Persistent, warp-specialised 2-CTA cluster (cta_group::2) per TPC; each CTA owns
// 128 rows of a 256-row super-tile and half of every weight tile; BN = N (128 or
// 256: A is read once), BK = 64:
//   warps 0-7   A producers: coalesced fp32 loads -> hi/lo fp16 -> swizzled smem
//                            (two groups alternating k-blocks, register prefetched)
//   warp  8     loader:      cp.async.bulk of weight tile images and image-format A
//   warp  9     MMA issuer (leader CTA) / stage-full relay (peer CTA)
//   warps 10-13 epilogue:    tcgen05.ld -> bias/ReLU -> fp32 rows (smem-staged
//                            full-line stores) or hi/lo tile image (smem-staged 4 KB
//                            bulk stores); overlaps the next tile's mainloop through
//                            a double-buffered TMEM accumulator
// Pipelines: smem full/empty mbarriers per stage, TMEM full/empty per buffer.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace th {
namespace tc {

constexpr int NUM_THREADS = 448;                // 8 producer + 1 loader + 1 MMA + 4 epilogue warps
constexpr int EPI_LD = 36;                     // staging row stride in floats (16 B aligned, conflict-free float4)

struct TcArgs {
  GemmArgs g;
  const unsigned char* wimg;  // per k-block: [hi tile image (N x 128 B) | lo tile image]
  int nkb;                    // k-blocks over all segments (each segment rounded up to 64)
  int num_tiles;
};

// ===========================================================================
// The kernel (cta_group::2): a cluster of two CTAs (one TPC) works on a
// 256-row super-tile.  Each CTA stages its own 128 rows of A and HALF of the
// weight tile (N/2 rows); the leader CTA issues tcgen05.mma.cta_group::2 (M=256),
// whose operand fetch reads both halves, so every SM reads and loads only half of
// B: shared-memory operand traffic per SM drops from 96 to 64 B/clk, weight L2
// traffic halves, and a stage shrinks to 64 KB (3 stages).  D stays per CTA
// (128 rows x N columns of its own TMEM), so the epilogue is unchanged.
// Cross-CTA synchronisation:
//   full   : per-CTA barrier (producers + B bytes); the peer's MMA-warp lane 0
//            relays each completed phase to the leader's `pfull` barrier
//   empty  : tcgen05.commit multicast (mask 0b11) arrives in both CTAs
//   tfull  : same multicast commit
//   tempty : leader's epilogue threads arrive locally, the peer's remotely
// ===========================================================================
template <int N>
struct Cfg2 {
  static constexpr int HALF_BYTES = (N / 2) * BK * 2;           // this CTA's half of one weight plane
  static constexpr int PLANE_BYTES = N * BK * 2;                // a full hi (or lo) plane in the image
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * HALF_BYTES;
  static constexpr int NSTAGE = N == 256 ? 3 : 4;
  static constexpr int CTRL_BYTES = 256;
  static constexpr int STAGING_BYTES = 4 * 8192;                // per epilogue warp: 32x36 fp32 or [hi 4 KB | lo 4 KB]
  static constexpr size_t SMEM = 1024 + (size_t)NSTAGE * STAGE_BYTES + CTRL_BYTES + N * 4 + STAGING_BYTES + 64;
};

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) k_gemm_tc2(const TcArgs a) {
  using C = Cfg2<N>;
  constexpr int NSTAGE = C::NSTAGE;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  constexpr int TMEM_COLS = 2 * N;

  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t ctrl = base + NSTAGE * STAGE_BYTES;
  const uint32_t bar_full = ctrl, bar_empty = ctrl + 32, bar_pfull = ctrl + 64, bar_tfull = ctrl + 96,
                 bar_tempty = ctrl + 112, bar_ptempty = ctrl + 128;
  unsigned char* ctrl_ptr = base_ptr + NSTAGE * STAGE_BYTES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl_ptr + 160);
  float* s_bias = reinterpret_cast<float*>(ctrl_ptr + C::CTRL_BYTES);
  unsigned char* s_stage = reinterpret_cast<unsigned char*>(s_bias + N);  // 4 x 8 KB, 1 KB-aligned offsets not needed

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int num_super = (a.num_tiles + 1) >> 1;  // 256-row super-tiles
  // fp32-row A segments need the producer warps; with every segment in tile-image format
  // (the fused path) the loader alone fills a stage and the producers retire at once
  bool any_f32 = false;
  for (int i = 0; i < a.g.nseg; ++i) any_f32 |= (a.g.seg[i].img == nullptr);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, (any_f32 ? 128 : 0) + 1);  // producer threads + the loader's expect_tx arrive
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_pfull + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 128);
      mbar_init(bar_ptempty + 8 * b, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < N; i += NUM_THREADS) s_bias[i] = a.g.bias ? a.g.bias[i] : 0.f;
  const float acc_scale = a.g.acc_scale ? __ldg(a.g.acc_scale) : 1.0f;  // 2^-e of the weight image (read from the blob)
  const int n_store = a.g.n_store > 0 ? a.g.n_store : N;
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t M = a.g.M;

  if (warp < 8) {
    // ===================== A producers (fp32-row segments only) =====================
    // Two groups of 4 warps alternate k-blocks, each software pipelined (the global loads of a
    // group's next k-block are in flight while it waits for a free stage).  Lane mapping per
    // pass (16 rows x 64 columns per warp): row = lane>>1, 8 float4 loads at columns
    // 8*i + 4*(lane&1); the 8-byte swizzled stores of a half-warp hit 8 distinct 16-byte chunks.
    const int grp = warp >> 2, wq = warp & 3;
    float4 v[2][8];
    // channel-major fp32 operand (GemmSeg::col_stride != 0, the pre-map GEMM over an NCHW map): a warp owns 32
    // consecutive rows (pixels, lane = row) and all 64 columns of the k-block, so that every load instruction
    // is one coalesced 128-byte line and a thread holds whole 16-byte chunks of its row
    auto load_block_cm = [&](int st, int sgi, int kin) {
      const GemmSeg sg = a.g.seg[sgi];
      const int64_t m = (int64_t)st * (2 * BM) + rank * BM + wq * 32 + lane;
      const bool row_ok = m < M;
      const float* src = sg.ptr + (int64_t)kin * sg.col_stride + (row_ok ? m : 0);
      float* vf = reinterpret_cast<float*>(&v[0][0]);
#pragma unroll
      for (int c = 0; c < 64; ++c)
        vf[c] = (row_ok && kin + c < sg.K) ? __ldg(src + (int64_t)c * sg.col_stride) : 0.f;
    };
    auto load_block = [&](int st, int sgi, int kin) {
      const GemmSeg sg = a.g.seg[sgi];
      if (sg.col_stride) {
        load_block_cm(st, sgi, kin);
        return;
      }
      const int64_t m0 = (int64_t)st * (2 * BM) + rank * BM;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int r = pass * 64 + wq * 16 + (lane >> 1);
        const int64_t m = m0 + r;
        const bool row_ok = m < M;
        const int64_t row = sg.row_mod ? m % sg.row_mod : m;
        const float* src = sg.ptr + row * sg.ld + kin + 4 * (lane & 1);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[pass][i] = (row_ok && kin + 8 * i + 4 * (lane & 1) < sg.K)
                           ? __ldg(reinterpret_cast<const float4*>(src + 8 * i))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    int st = cluster_id, sgi = 0, kin = 0;
    auto advance = [&]() {
      kin += BK;
      if (kin >= a.g.seg[sgi].K) {
        kin = 0;
        if (++sgi == a.g.nseg) {
          sgi = 0;
          st += nclusters;
        }
      }
    };
    int kcount = 0;
    if (!any_f32) st = num_super;  // nothing to produce
    if (grp == 1 && st < num_super) {
      advance();
      kcount = 1;
    }
    if (st < num_super && !a.g.seg[sgi].img) load_block(st, sgi, kin);
    while (st < num_super) {
      const int s = kcount % NSTAGE;
      const uint32_t ph = (kcount / NSTAGE) & 1;
      const bool is_img = a.g.seg[sgi].img != nullptr;  // the loader warp brings image-format k-blocks
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (!is_img) {
        unsigned char* a_hi = base_ptr + s * STAGE_BYTES;
        unsigned char* a_lo = a_hi + A_TILE_BYTES;
        if (a.g.seg[sgi].col_stride) {
          const float* vf = reinterpret_cast<const float*>(&v[0][0]);
          const int r = wq * 32 + lane;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 hi, lo;
            split2(vf[8 * i + 0], vf[8 * i + 1], hi.x, lo.x);
            split2(vf[8 * i + 2], vf[8 * i + 3], hi.y, lo.y);
            split2(vf[8 * i + 4], vf[8 * i + 5], hi.z, lo.z);
            split2(vf[8 * i + 6], vf[8 * i + 7], hi.w, lo.w);
            const int off = r * 128 + ((i ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(a_hi + off) = hi;
            *reinterpret_cast<uint4*>(a_lo + off) = lo;
          }
        } else {
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const int r = pass * 64 + wq * 16 + (lane >> 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint2 hi, lo;
              split2(v[pass][i].x, v[pass][i].y, hi.x, lo.x);
              split2(v[pass][i].z, v[pass][i].w, hi.y, lo.y);
              const int off = r * 128 + ((i ^ (r & 7)) << 4) + ((lane & 1) << 3);
              *reinterpret_cast<uint2*>(a_hi + off) = hi;
              *reinterpret_cast<uint2*>(a_lo + off) = lo;
            }
          }
        }
        fence_proxy_async();
      }
      mbar_arrive(bar_full + 8 * s);
      advance();
      if (st < num_super) advance();
      kcount += 2;
      if (st < num_super && !a.g.seg[sgi].img) load_block(st, sgi, kin);
    }
  } else if (warp == 8) {
    // ===================== loader: this CTA's half of every weight plane, and A k-blocks
    // that are already in tile-image format (one 32 KB bulk copy each) =====================
    if (lane == 0) {
      int kcount = 0;
      for (int st = cluster_id; st < num_super; st += nclusters) {
        const int64_t tile = 2 * (int64_t)st + rank;  // this CTA's 128-row tile
        int kb = 0;
        for (int sgi = 0; sgi < a.g.nseg; ++sgi) {
          const GemmSeg sg = a.g.seg[sgi];
          const int seg_kbs = (sg.K + BK - 1) / BK;
          for (int kk = 0; kk < seg_kbs; ++kk, ++kb, ++kcount) {
            const int s = kcount % NSTAGE;
            const uint32_t ph = (kcount / NSTAGE) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const uint32_t sa = base + s * STAGE_BYTES;
            mbar_arrive_expect_tx(bar_full + 8 * s, 2 * C::HALF_BYTES + (sg.img ? 2 * A_TILE_BYTES : 0));
            if (sg.img) {
              const int64_t tl = sg.img_tile_mod ? tile % sg.img_tile_mod : tile;
              bulk_g2s(sa, sg.img + ((size_t)tl * seg_kbs + kk) * (2 * A_TILE_BYTES), 2 * A_TILE_BYTES,
                       bar_full + 8 * s);
            }
            // this CTA's half of the k-block: [hi | lo], contiguous in the packed image
            const unsigned char* src = a.wimg + (size_t)kb * (2 * C::PLANE_BYTES) + (size_t)rank * (2 * C::HALF_BYTES);
            bulk_g2s(sa + 2 * A_TILE_BYTES, src, 2 * C::HALF_BYTES, bar_full + 8 * s);
          }
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      if (rank == 0) {
        // ===================== MMA issuer (leader CTA) =====================
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
        int kcount = 0, it = 0;
        for (int st = cluster_id; st < num_super; st += nclusters, ++it) {
          const int ab = it & 1;
          const uint32_t aph = (it >> 1) & 1;
          mbar_wait(bar_tempty + 8 * ab, aph ^ 1);   // both epilogues have drained this accumulator
          mbar_wait(bar_ptempty + 8 * ab, aph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + ab * N;
          for (int kb = 0; kb < a.nkb; ++kb, ++kcount) {
            const int s = kcount % NSTAGE;
            const uint32_t ph = (kcount / NSTAGE) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            mbar_wait(bar_pfull + 8 * s, ph);
            tc_fence_after();
            const uint32_t sa = base + s * STAGE_BYTES;
            const uint64_t d_ahi = umma_desc(sa), d_alo = umma_desc(sa + A_TILE_BYTES);
            const uint64_t d_bhi = umma_desc(sa + 2 * A_TILE_BYTES),
                           d_blo = umma_desc(sa + 2 * A_TILE_BYTES + C::HALF_BYTES);
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              const uint64_t adv = (uint64_t)((ks * 32) >> 4);
              umma_f16_2cta(d_tmem, d_ahi + adv, d_bhi + adv, idesc, (kb | ks) ? 1u : 0u);
              umma_f16_2cta(d_tmem, d_alo + adv, d_bhi + adv, idesc, 1u);
              umma_f16_2cta(d_tmem, d_ahi + adv, d_blo + adv, idesc, 1u);
            }
            umma_commit_2cta(bar_empty + 8 * s);
          }
          umma_commit_2cta(bar_tfull + 8 * ab);
        }
      } else {
        // ===================== relay (peer CTA): forward "stage full" to the leader =====================
        int kcount = 0;
        for (int st = cluster_id; st < num_super; st += nclusters) {
          for (int kb = 0; kb < a.nkb; ++kb, ++kcount) {
            const int s = kcount % NSTAGE;
            const uint32_t ph = (kcount / NSTAGE) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            mbar_arrive_remote(bar_pfull + 8 * s, 0);
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    unsigned char* stage_b = s_stage + q * 8192;
    float* stage = reinterpret_cast<float*>(stage_b);
    int it = 0;
    for (int st = cluster_id; st < num_super; st += nclusters, ++it) {
      const int ab = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * ab, aph);
      tc_fence_after();
      const int64_t tile = 2 * (int64_t)st + rank;
      const int64_t mbase = tile * BM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * N;
      if (a.g.C_img) {
        // hi/lo fp16 tile image: per 64-wide k-block this warp owns rows [32q, 32q+32) of both
        // planes = two contiguous 4 KB slabs; stage them swizzled, then two bulk stores.
        const int r = q * 32 + lane;  // row inside the 128-row tile
#pragma unroll 1
        for (int kb = 0; kb < N / BK; ++kb) {
          if (lane == 0) bulk_wait_read0();  // previous slabs have been read out of the staging buffer
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld32(taddr + kb * BK + h * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                x[e] = fmaf(__uint_as_float(v[j + e]), acc_scale, s_bias[kb * BK + h * 32 + j + e]);
                if (a.g.relu) x[e] = fmaxf(x[e], 0.f);
              }
              uint4 hi, lo;
              split2(x[0], x[1], hi.x, lo.x);
              split2(x[2], x[3], hi.y, lo.y);
              split2(x[4], x[5], hi.z, lo.z);
              split2(x[6], x[7], hi.w, lo.w);
              const int chunk = h * 4 + (j >> 3);  // 16-byte chunk of the 128-byte row
              const int off = lane * 128 + ((chunk ^ (r & 7)) << 4);
              *reinterpret_cast<uint4*>(stage_b + off) = hi;
              *reinterpret_cast<uint4*>(stage_b + 4096 + off) = lo;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            unsigned char* dst = a.g.C_img + ((size_t)tile * (N / BK) + kb) * (2 * A_TILE_BYTES) + (size_t)q * 4096;
            bulk_s2g(dst, smem_u32(stage_b), 4096);
            bulk_s2g(dst + A_TILE_BYTES, smem_u32(stage_b + 4096), 4096);
            bulk_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = fmaf(__uint_as_float(v[j + 0]), acc_scale, s_bias[c0 + j + 0]);
            o.y = fmaf(__uint_as_float(v[j + 1]), acc_scale, s_bias[c0 + j + 1]);
            o.z = fmaf(__uint_as_float(v[j + 2]), acc_scale, s_bias[c0 + j + 2]);
            o.w = fmaf(__uint_as_float(v[j + 3]), acc_scale, s_bias[c0 + j + 3]);
            if (a.g.relu) {
              o.x = fmaxf(o.x, 0.f);
              o.y = fmaxf(o.y, 0.f);
              o.z = fmaxf(o.z, 0.f);
              o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(stage + lane * EPI_LD + j) = o;
          }
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) {
            const int row = rr + (lane >> 3);
            const int64_t m = mbase + row;
            const float4 o = *reinterpret_cast<const float4*>(stage + row * EPI_LD + (lane & 7) * 4);
            if (m < M && c0 + (lane & 7) * 4 < n_store)
              *reinterpret_cast<float4*>(a.g.C + m * a.g.ldc + c0 + (lane & 7) * 4) = o;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      if (rank == 0)
        mbar_arrive(bar_tempty + 8 * ab);
      else
        mbar_arrive_remote(bar_ptempty + 8 * ab, 0);
    }
    if (a.g.C_img && lane == 0) bulk_wait_all0();  // image stores complete before the kernel ends
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

}  // namespace tc

int tc_image_kblocks(const GemmArgs& a) {
  int n = 0;
  for (int s = 0; s < a.nseg; ++s) n += (a.seg[s].K + tc::BK - 1) / tc::BK;
  return n;
}

int launch_gemm_tc(const GemmArgs& a, const void* w_image, cudaStream_t st, int prof_cat) {
  ProfScope prof_(prof_cat, st);
  if (a.M <= 0) return TH_OK;
  if (a.N != 128 && a.N != 256) {
    set_error("gemm_tc: N=%d unsupported (128 or 256)", a.N);
    return TH_EINVAL;
  }
  for (int s = 0; s < a.nseg; ++s) {
    const GemmSeg& g = a.seg[s];
    const bool ok = g.img ? (g.K % tc::BK == 0 && (reinterpret_cast<uintptr_t>(g.img) & 1023) == 0)
                          : (g.K % 16 == 0 && g.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g.ptr) & 15) == 0);
    if (!ok) {
      set_error("gemm_tc: segment %d K=%d ld=%d unsupported", s, g.K, g.ld);
      return TH_EINVAL;
    }
  }
  if ((a.C_img != nullptr) == (a.C != nullptr) || (a.C_img && (a.M % (2 * tc::BM) != 0 ||
                                                                (reinterpret_cast<uintptr_t>(a.C_img) & 1023)))) {
    set_error("gemm_tc: exactly one of C / C_img, image output needs M %% 256 == 0");
    return TH_EINVAL;
  }
  int num_sms = 0;
  if (device_sm_count(&num_sms)) return TH_ECUDA;
  tc::TcArgs t;
  t.g = a;
  t.wimg = static_cast<const unsigned char*>(w_image);
  t.nkb = tc_image_kblocks(a);
  t.num_tiles = (int)cdiv(a.M, tc::BM);
  const int num_super = (t.num_tiles + 1) / 2;
  const int nclusters = num_super < num_sms / 2 ? num_super : num_sms / 2;
  // function attributes are per device: set them on every launch (cheap) instead of caching a flag
  if (a.N == 256) {
    TH_CUDA(cudaFuncSetAttribute(tc::k_gemm_tc2<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)tc::Cfg2<256>::SMEM));
    tc::k_gemm_tc2<256><<<2 * nclusters, tc::NUM_THREADS, tc::Cfg2<256>::SMEM, st>>>(t);
  } else {
    TH_CUDA(cudaFuncSetAttribute(tc::k_gemm_tc2<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)tc::Cfg2<128>::SMEM));
    tc::k_gemm_tc2<128><<<2 * nclusters, tc::NUM_THREADS, tc::Cfg2<128>::SMEM, st>>>(t);
  }
  TH_LAUNCHED();
  return TH_OK;
}

// Pre-mapped feature maps (kernels.cuh): per view two N = 256 GEMMs over the (H*W, 384) channel-major map,
// [alpha_res_0 + b] -> channels 0..255 and [V1 rgb_res_0 ; fc_4 rgb_res_1 / V] -> channels 256..511 of the
// channel-last output.  0.31 TFLOP per 512 x 512 x 3 frame; bound by reading the map twice and writing 1.6 GB.
int launch_premap(const float* src_nchw, const unsigned char* weights, const PackedHeader& hdr, float* dst, int n_views,
                  int h, int w, cudaStream_t st) {
  const int64_t HW = (int64_t)h * w;
  for (int v = 0; v < n_views; ++v)
    for (int half = 0; half < 2; ++half) {
      GemmArgs g{};
      g.nseg = 1;
      g.seg[0].ptr = src_nchw + (int64_t)v * TH_C_PIX * HW;
      g.seg[0].K = TH_C_PIX;
      g.seg[0].ld = 4;  // unused in channel-major mode (must pass the alignment check)
      g.seg[0].col_stride = HW;
      g.bias = reinterpret_cast<const float*>(weights + (half ? hdr.preb_b : hdr.ar0_b));
      g.C = dst + (int64_t)v * HW * 512 + half * 256;
      g.ldc = 512;
      g.M = HW;
      g.N = 256;
      g.relu = 0;
      g.acc_scale = img_inv_scale_ptr(weights, hdr, half ? hdr.h_preb : hdr.h_ar0);
      int rc = launch_gemm_tc(g, weights + (half ? hdr.h_preb : hdr.h_ar0), st, PROF_PREMAP);
      if (rc) return rc;
    }
  return TH_OK;
}

}  // namespace th
