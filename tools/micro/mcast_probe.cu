// Probe: does a cluster multicast of a bulk copy reduce L2 -> SM sector traffic compared with every
// CTA loading the same bytes itself?  Each CTA streams the same 8 MB blob (like the chain kernel's
// weights) through a 2 x 32 KB shared-memory ring.  MODE 0: every CTA loads all of it (unicast).
// MODE 1: clusters of 4, CTA r loads the chunks with (chunk & 3) == r and multicasts them to all 4.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mcast_probe mcast_probe.cu
// Run under: ncu --metrics lts__t_sectors.sum,lts__t_sectors_srcunit_tex.sum,gpu__time_duration.sum
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int CHUNK = 32768;

template <int MODE>
__global__ void __launch_bounds__(128, 1) probe(const unsigned char* blob, int nchunks, int reps, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bar = base + 2 * CHUNK;  // two full barriers
  const uint32_t rank = MODE ? ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE) cluster_sync();
  float acc = 0.f;
  int n = 0;
  for (int rep = 0; rep < reps; ++rep)
    for (int c = 0; c < nchunks; ++c, ++n) {
      const int s = n & 1;
      const uint32_t ph = (n >> 1) & 1;
      // everyone must be done reading slot s (from iteration n - 2) before it is refilled
      if (MODE) cluster_sync(); else __syncthreads();
      if (threadIdx.x == 0) {
        mbar_expect(bar + 8 * s, CHUNK);
        if (MODE == 0) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                           "r"(base + s * CHUNK), "l"(blob + (size_t)c * CHUNK), "r"(CHUNK), "r"(bar + 8 * s) : "memory");
        } else if ((uint32_t)(c & 3) == rank) {
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
                  "r"(base + s * CHUNK), "l"(blob + (size_t)c * CHUNK), "r"(CHUNK), "r"(bar + 8 * s), "h"((uint16_t)0xF) : "memory");
        }
      }
      mbar_wait(bar + 8 * s, ph);
      acc += reinterpret_cast<const float*>(smem + s * CHUNK)[threadIdx.x];
    }
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  const int nchunks = 256, reps = 4;  // 8 MB x 4
  unsigned char* blob;
  float* sink;
  cudaMalloc(&blob, (size_t)nchunks * CHUNK);
  cudaMemset(blob, 1, (size_t)nchunks * CHUNK);
  cudaMalloc(&sink, 4);
  const int smem = 2 * CHUNK + 64;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) {
    probe<0><<<148, 128, smem>>>(blob, nchunks, reps, sink);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe<1>, (const unsigned char*)blob, nchunks, reps, sink);
    if (e != cudaSuccess) printf("cluster launch: %s\n", cudaGetErrorString(e));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return 0;
}
