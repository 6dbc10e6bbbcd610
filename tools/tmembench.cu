// Microbenchmark: TMEM -> register read bandwidth (tcgen05.ld.32x32b) on sm_100a, with and without
// the FMNMX3 reduction of the Chamfer filter.  It decides whether a tcgen05/TMEM version of the
// filter scan could beat the legacy-HMMA one (64 filter values per clock per SM, tensor bound):
// every filter value has to leave TMEM through tcgen05.ld once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmembench.bin tmembench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr));
}

// MODE 0: loads only (values xor-folded every 8th load to keep them alive cheaply)
// MODE 1: loads + 16 FMNMX3 per 32 values (the filter reduction)
template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, long long* cyc) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // lane quadrant of this warp
  float rm = 1e30f, acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c += 2) {
      float v[32], w[32];
      tmem_ld32(base + ((it & 3) * 128 + c * 32), v);
      tmem_ld32(base + ((it & 3) * 128 + c * 32 + 32), w);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (MODE == 1) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) rm = fmin3(rm, v[e], v[e + 1]);
#pragma unroll
        for (int e = 0; e < 32; e += 2) acc = fmin3(acc, w[e], w[e + 1]);
      } else {
        acc += v[0] + v[31] + w[0] + w[31];
      }
    }
  }
  const long long t1 = clock64();
  if (rm + acc == 123.456f) out[0] = rm;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

template <int MODE>
static void run(const char* name, int threads, int sms) {
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, 4));
  CK(cudaMalloc(&cyc, 8));
  const int iters = 1 << 13;
  for (int rep = 0; rep < 2; rep++) {
    bench<MODE><<<sms, threads>>>(out, iters, cyc);
    CK(cudaDeviceSynchronize());
  }
  long long h = 0;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const double clk_per_ld = (double)h / (iters * 4.0);  // one warp: 32 lanes x 32 columns x 4 B = 4096 B
  const int warps = threads / 32;
  printf("%-26s warps/SM=%d  clk per x32 load per warp = %6.2f  ->  %.1f B/clk/SM = %.1f values/clk/SM\n", name, warps,
         clk_per_ld, 4096.0 * warps / clk_per_ld, 1024.0 * warps / clk_per_ld);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  run<0>("tcgen05.ld only", 128, sms);
  run<0>("tcgen05.ld only", 256, sms);
  run<1>("tcgen05.ld + 16 FMNMX3", 128, sms);
  run<1>("tcgen05.ld + 16 FMNMX3", 256, sms);
  return 0;
}
