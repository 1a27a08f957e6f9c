// Issue-rate probe for the instructions the chain kernel's epilogue and mix are made of
// (fp16 <-> fp32 conversions, packed fp32x2 arithmetic, FMNMX) on sm_100a.
// Every warp runs ILP independent dependency chains of ONE instruction kind; the probe prints
// cycles per warp-instruction per SM sub-partition for 1, 2 and 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

enum Kind { K_H2F, K_F2FP, K_FADD2, K_FFMA2, K_FMNMX, K_FFMA, K_SPLIT, K_JOIN, K_SPLIT_TRUNC, NKIND };
static const char* kind_name[NKIND] = {"cvt.f32.f16 (HADD2.F32)", "cvt.rn.f16x2.f32 (F2FP)", "add.f32x2 (FADD2)",
                                       "fma.f32x2 (FFMA2)",      "max.f32 (FMNMX)",         "fma.f32 (FFMA)",
                                       "split2 (5 instr / pair)", "join2 (5 instr / pair)",
                                       "split2 via LOP3 truncation (5 instr / pair)"};
constexpr int ILP = 8, ITERS = 2048;

template <int KIND>
__global__ void probe(uint32_t* out, long long* cyc) {
  uint32_t r[ILP];
  float f[ILP], g[ILP];
  unsigned long long d[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    r[i] = 0x3c003c00u + threadIdx.x + i;
    f[i] = 1.0f + 0.001f * (threadIdx.x + i);
    g[i] = 0.5f + 0.002f * i;
    d[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(g[i]);
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == K_H2F) {
        asm volatile("{.reg .f16 h, l; mov.b32 {h, l}, %1; cvt.f32.f16 %0, h;}" : "=f"(f[i]) : "r"(r[i]));
        r[i] ^= __float_as_uint(f[i]);
      } else if (KIND == K_F2FP) {
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(f[i]), "f"(g[i]));
        f[i] = __uint_as_float(r[i] | 0x3c000000u);
      } else if (KIND == K_FADD2) {
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(d[(i + 1) % ILP]));
      } else if (KIND == K_FFMA2) {
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(d[(i + 1) % ILP]));
      } else if (KIND == K_FMNMX) {
        asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
        g[i] = f[i] - 1.0f;
      } else if (KIND == K_FFMA) {
        asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[i]));
      } else if (KIND == K_SPLIT) {
        // the epilogue's split: hi = fp16x2(x), lo = fp16x2(x - hi)
        __half2 h = __floats2half2_rn(f[i], g[i]);
        const float2 hf = __half22float2(h);
        unsigned long long a = ((unsigned long long)__float_as_uint(g[i]) << 32) | __float_as_uint(f[i]);
        unsigned long long b = ((unsigned long long)__float_as_uint(hf.y) << 32) | __float_as_uint(hf.x), c;
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
        __half2 l = __floats2half2_rn(__uint_as_float((uint32_t)c), __uint_as_float((uint32_t)(c >> 32)));
        r[i] ^= *reinterpret_cast<uint32_t*>(&h) + *reinterpret_cast<uint32_t*>(&l);
        f[i] += 1.0f;
      } else if (KIND == K_JOIN) {
        const float2 a = __half22float2(*reinterpret_cast<__half2*>(&r[i]));
        const float2 b = __half22float2(*reinterpret_cast<__half2*>(&r[(i + 1) % ILP]));
        f[i] += a.x + b.x;
        g[i] += a.y + b.y;
        r[i] += 0x00010001u;
      } else if (KIND == K_SPLIT_TRUNC) {
        // hi = x with the mantissa cut to 10 bits (exact in fp16), lo = fp16(x - hi)
        const float hx = __uint_as_float(__float_as_uint(f[i]) & 0xffffe000u);
        const float hy = __uint_as_float(__float_as_uint(g[i]) & 0xffffe000u);
        __half2 h = __floats2half2_rn(hx, hy);
        unsigned long long a = ((unsigned long long)__float_as_uint(g[i]) << 32) | __float_as_uint(f[i]);
        unsigned long long b = ((unsigned long long)__float_as_uint(hy) << 32) | __float_as_uint(hx), c;
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
        __half2 l = __floats2half2_rn(__uint_as_float((uint32_t)c), __uint_as_float((uint32_t)(c >> 32)));
        r[i] ^= *reinterpret_cast<uint32_t*>(&h) + *reinterpret_cast<uint32_t*>(&l);
        f[i] += 1.0f;
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc ^= r[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]) ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(uint32_t* out, long long* cyc, int sms) {
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 128 * wps;
    probe<KIND><<<sms, threads>>>(out, cyc);
    probe<KIND><<<sms, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    // warp-level "instructions" (chain steps) per sub-partition = ITERS * ILP * wps
    printf("%-46s warps/SMSP %d: %7.2f cycles per chain step per SMSP (%.2f per warp)\n", kind_name[KIND], wps,
           avg / ((double)ITERS * ILP * wps), avg / ((double)ITERS * ILP));
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)sms * 512 * 4);
  cudaMalloc(&cyc, 256 * sizeof(long long));
  run<K_H2F>(out, cyc, sms);
  run<K_F2FP>(out, cyc, sms);
  run<K_FADD2>(out, cyc, sms);
  run<K_FFMA2>(out, cyc, sms);
  run<K_FMNMX>(out, cyc, sms);
  run<K_FFMA>(out, cyc, sms);
  run<K_SPLIT>(out, cyc, sms);
  run<K_JOIN>(out, cyc, sms);
  run<K_SPLIT_TRUNC>(out, cyc, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
