// Mesh extraction from the density cube (second half): if_mesh_renderer.py:98-104 calls the CPU package
// `mcubes.marching_cubes(cube, cfg.mesh_th)`; th_marching_cubes does the same step on the device, on the cube
// th_query_density left there: one vertex per cut lattice edge at a + (iso - v_a) / (v_b - v_a) (index coordinates,
// fp32), triangles from the 256-case table of mc_table.h (derived by tools/gen_mc_table.py; PyMCubes' own table is
// not available: see oracle/marching_cubes.py for what that means for parity).  Indexed mesh, no duplicate vertices:
//   k_mc_edge_counts / k_mc_vertices : slot = 3 * lattice point + axis; flags recomputed from the volume, per-block
//                                      counts -> single-block scan -> vertex id of every cut edge + its position
//   k_mc_tri_counts  / k_mc_triangles: cube case -> triangle count, scan, triangles as vertex ids through the slots
// Orders are fixed (vertices by slot, triangles by cube then table), so the result is deterministic.
#include "kernels.cuh"
#include "mc_table.h"

namespace th {

__constant__ int8_t MC_NTRI[256];
__constant__ int8_t MC_TRI[256][16];
// edge e of a cube = (offset of its lower lattice point, axis)
__constant__ int8_t MC_EDGE[12][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {0, 1, 0, 0}, {0, 0, 0, 1}, {0, 0, 1, 0}, {1, 0, 1, 1},
                                      {0, 1, 1, 0}, {0, 0, 1, 1}, {0, 0, 0, 2}, {1, 0, 0, 2}, {1, 1, 0, 2}, {0, 1, 0, 2}};

struct McVol {
  const float* vol;
  int nx, ny, nz;
  float iso;
};
constexpr int MC_PER_THREAD = 16, MC_PER_BLOCK = 256 * MC_PER_THREAD;

// is the edge of `slot` cut?  (i, j, k, axis) and the two values on request
__device__ __forceinline__ bool mc_edge(const McVol& v, int64_t slot, int64_t n_slots, int* ijk, int* axis, float* va,
                                        float* vb) {
  if (slot >= n_slots) return false;
  const int64_t p = slot / 3;
  const int a = (int)(slot - 3 * p);
  const int k = (int)(p % v.nz), j = (int)((p / v.nz) % v.ny), i = (int)(p / ((int64_t)v.nz * v.ny));
  const int lim = a == 0 ? v.nx : (a == 1 ? v.ny : v.nz), c = a == 0 ? i : (a == 1 ? j : k);
  if (c >= lim - 1) return false;
  const int64_t stride = a == 0 ? (int64_t)v.ny * v.nz : (a == 1 ? v.nz : 1);
  const float x = v.vol[p], y = v.vol[p + stride];
  if ((x > v.iso) == (y > v.iso)) return false;
  ijk[0] = i, ijk[1] = j, ijk[2] = k;
  *axis = a, *va = x, *vb = y;
  return true;
}

__device__ __forceinline__ int block_sum_256(int c, int* ws) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  return ws[0] + ws[1] + ws[2] + ws[3] + ws[4] + ws[5] + ws[6] + ws[7];
}
// exclusive prefix of c over the block's 256 threads
__device__ __forceinline__ int block_prefix_256(int c, int* ws) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  int before = x - c;
  for (int w = 0; w < warp; ++w) before += ws[w];
  return before;
}

__global__ void __launch_bounds__(256) k_mc_edge_counts(McVol v, int64_t n_slots, int32_t* __restrict__ counts) {
  __shared__ int ws[8];
  const int64_t s0 = blockIdx.x * (int64_t)MC_PER_BLOCK + threadIdx.x * MC_PER_THREAD;
  int c = 0, ijk[3], axis;
  float va, vb;
  for (int i = 0; i < MC_PER_THREAD; ++i) c += mc_edge(v, s0 + i, n_slots, ijk, &axis, &va, &vb) ? 1 : 0;
  const int total = block_sum_256(c, ws);
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_mc_vertices(McVol v, int64_t n_slots, const int32_t* __restrict__ offsets,
                                                     int32_t* __restrict__ vid, float* __restrict__ verts,
                                                     int64_t max_verts) {
  __shared__ int ws[8];
  const int64_t s0 = blockIdx.x * (int64_t)MC_PER_BLOCK + threadIdx.x * MC_PER_THREAD;
  int c = 0, ijk[3], axis;
  float va, vb;
  for (int i = 0; i < MC_PER_THREAD; ++i) c += mc_edge(v, s0 + i, n_slots, ijk, &axis, &va, &vb) ? 1 : 0;
  int id = offsets[blockIdx.x] + block_prefix_256(c, ws);
  for (int i = 0; i < MC_PER_THREAD; ++i) {
    const int64_t slot = s0 + i;
    if (slot >= n_slots) break;
    if (mc_edge(v, slot, n_slots, ijk, &axis, &va, &vb)) {
      vid[slot] = id;
      if (verts && id < max_verts) {
        const float t = __fdiv_rn(__fsub_rn(v.iso, va), __fsub_rn(vb, va));
        float p[3] = {(float)ijk[0], (float)ijk[1], (float)ijk[2]};
        p[axis] = __fadd_rn(p[axis], t);
        verts[3 * (int64_t)id] = p[0], verts[3 * (int64_t)id + 1] = p[1], verts[3 * (int64_t)id + 2] = p[2];
      }
      ++id;
    } else {
      vid[slot] = -1;
    }
  }
}

__device__ __forceinline__ int mc_case(const McVol& v, int64_t cube, int64_t n_cubes, int* ijk) {
  if (cube >= n_cubes) return 0;
  const int cz = v.nz - 1, cy = v.ny - 1;
  const int k = (int)(cube % cz), j = (int)((cube / cz) % cy), i = (int)(cube / ((int64_t)cz * cy));
  ijk[0] = i, ijk[1] = j, ijk[2] = k;
  const int64_t sy = v.nz, sx = (int64_t)v.ny * v.nz, p = i * sx + j * sy + k;
  const float iso = v.iso;
  int m = 0;
  m |= (v.vol[p] > iso) ? 1 : 0;
  m |= (v.vol[p + sx] > iso) ? 2 : 0;
  m |= (v.vol[p + sx + sy] > iso) ? 4 : 0;
  m |= (v.vol[p + sy] > iso) ? 8 : 0;
  m |= (v.vol[p + 1] > iso) ? 16 : 0;
  m |= (v.vol[p + sx + 1] > iso) ? 32 : 0;
  m |= (v.vol[p + sx + sy + 1] > iso) ? 64 : 0;
  m |= (v.vol[p + sy + 1] > iso) ? 128 : 0;
  return m;
}

__global__ void __launch_bounds__(256) k_mc_tri_counts(McVol v, int64_t n_cubes, int32_t* __restrict__ counts) {
  __shared__ int ws[8];
  const int64_t c0 = blockIdx.x * (int64_t)MC_PER_BLOCK + threadIdx.x * MC_PER_THREAD;
  int c = 0, ijk[3];
  for (int i = 0; i < MC_PER_THREAD; ++i) c += MC_NTRI[mc_case(v, c0 + i, n_cubes, ijk)];
  const int total = block_sum_256(c, ws);
  if (threadIdx.x == 0) counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_mc_triangles(McVol v, int64_t n_cubes, const int32_t* __restrict__ offsets,
                                                      const int32_t* __restrict__ vid, int32_t* __restrict__ tris,
                                                      int64_t max_tris) {
  __shared__ int ws[8];
  const int64_t c0 = blockIdx.x * (int64_t)MC_PER_BLOCK + threadIdx.x * MC_PER_THREAD;
  int c = 0, ijk[3];
  for (int i = 0; i < MC_PER_THREAD; ++i) c += MC_NTRI[mc_case(v, c0 + i, n_cubes, ijk)];
  int64_t t = offsets[blockIdx.x] + block_prefix_256(c, ws);
  for (int i = 0; i < MC_PER_THREAD; ++i) {
    const int m = mc_case(v, c0 + i, n_cubes, ijk);
    const int n = MC_NTRI[m];
    for (int q = 0; q < n; ++q, ++t) {
      if (t >= max_tris) continue;
#pragma unroll
      for (int e3 = 0; e3 < 3; ++e3) {
        const int e = MC_TRI[m][3 * q + e3];
        const int64_t p = ((int64_t)(ijk[0] + MC_EDGE[e][0]) * v.ny + (ijk[1] + MC_EDGE[e][1])) * v.nz + (ijk[2] + MC_EDGE[e][2]);
        tris[3 * t + e3] = vid[3 * p + MC_EDGE[e][3]];
      }
    }
  }
}

size_t marching_cubes_workspace_bytes(int nx, int ny, int nz) {
  const int64_t n_slots = 3LL * nx * ny * nz;
  return align_up((size_t)cdiv(n_slots, MC_PER_BLOCK) * 4, 256) + align_up((size_t)n_slots * 4, 256) + 256;
}

// counts_dev[0] = vertices, [1] = triangles (device, unsigned long long); buffers may be null (count only)
int launch_marching_cubes(const float* vol, int nx, int ny, int nz, float iso, float* verts, int64_t max_verts,
                          int32_t* tris, int64_t max_tris, unsigned long long* counts_dev, void* workspace,
                          cudaStream_t st) {
  ProfScope prof_(PROF_PROLOGUE, st);
  static bool tables[64] = {false};
  int dev = 0;
  TH_CUDA(cudaGetDevice(&dev));
  if (!tables[dev & 63]) {  // per device, once
    TH_CUDA(cudaMemcpyToSymbol(MC_NTRI, MC_NTRI_HOST, sizeof(MC_NTRI_HOST)));
    TH_CUDA(cudaMemcpyToSymbol(MC_TRI, MC_TRI_HOST, sizeof(MC_TRI_HOST)));
    tables[dev & 63] = true;
  }
  const McVol v{vol, nx, ny, nz, iso};
  const int64_t n_slots = 3LL * nx * ny * nz, n_cubes = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
  int32_t* counts = static_cast<int32_t*>(workspace);
  int32_t* vid = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(workspace) +
                                            align_up((size_t)cdiv(n_slots, MC_PER_BLOCK) * 4, 256));
  const int nb_e = (int)cdiv(n_slots, MC_PER_BLOCK), nb_c = (int)cdiv(n_cubes, MC_PER_BLOCK);
  k_mc_edge_counts<<<nb_e, 256, 0, st>>>(v, n_slots, counts);
  TH_LAUNCHED();
  int rc = launch_scan_counts(counts, nb_e, &counts_dev[0], st);
  if (rc) return rc;
  k_mc_vertices<<<nb_e, 256, 0, st>>>(v, n_slots, counts, vid, verts, max_verts);
  TH_LAUNCHED();
  if (n_cubes <= 0) {
    TH_CUDA(cudaMemsetAsync(&counts_dev[1], 0, 8, st));
    return TH_OK;
  }
  k_mc_tri_counts<<<nb_c, 256, 0, st>>>(v, n_cubes, counts);
  TH_LAUNCHED();
  if ((rc = launch_scan_counts(counts, nb_c, &counts_dev[1], st))) return rc;
  if (tris) {
    k_mc_triangles<<<nb_c, 256, 0, st>>>(v, n_cubes, counts, vid, tris, max_tris);
    TH_LAUNCHED();
  }
  return TH_OK;
}

}  // namespace th
