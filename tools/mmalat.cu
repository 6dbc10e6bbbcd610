// Microbenchmark: latency of ONE tcgen05.mma (M=128, N=256 or 128, K=16, bf16, operands in shared memory in the
// canonical no-swizzle K-major layout of nn_distance_fwd_umma.cu) from issue to mbarrier completion, alone and while
// four other warps keep draining another accumulator with tcgen05.ld.  Decides how much of a scan step is MMA time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mmalat.bin mmalat.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr) : "memory");
}
template <int N, int LOADERS, int BURST = 1>
__global__ void __launch_bounds__(256) bench(float* out, int iters, long long* res) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ unsigned long long bar;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * 32 + 2048 * 32) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;  // any finite bf16s
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  if (warp < 4) {
    if (tid == 0) {
      const uint32_t desc_hi = (uint32_t)(256 >> 4) | (1u << 14);
      const uint32_t a_lo = ((smem_u32(smem) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t b_lo = ((smem_u32(smem + 4096) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      long long tot = 0, mx = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; it++) {
        const long long t0 = clock64();
#pragma unroll 1
        for (int q = 0; q < BURST; q++)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (q & 1) * 256),
                     "l"(((uint64_t)desc_hi << 32) | a_lo), "l"(((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)((it + q) & 3) * (N * 32 / 16))), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        }
        const long long t1 = clock64();
        phase ^= 1;
        tot += t1 - t0;
        if (t1 - t0 > mx) mx = t1 - t0;
        for (int w = 0; w < 40; w++) asm volatile("nanosleep.u32 20;");  // gap between MMAs
      }
      if (blockIdx.x == 0) { res[0] = tot; res[1] = mx; }
      stop = 1;
    }
  } else if (LOADERS == 2) {  // warps 4-7 poll an mbarrier that never completes (what waiting warps of a pipeline do)
    __shared__ unsigned long long idle_bar;
    if (tid == 128) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&idle_bar)));
    asm volatile("bar.sync 1, 128;");
    uint32_t done = 0;
    while (!stop) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&idle_bar)), "r"(0u) : "memory");
    }
    if (done == 123) out[0] = 1.f;
  } else if (LOADERS == 3) {  // same with test_wait (pure spin)
    __shared__ unsigned long long idle_bar2;
    if (tid == 128) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&idle_bar2)));
    asm volatile("bar.sync 1, 128;");
    uint32_t done = 0;
    while (!stop) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&idle_bar2)), "r"(0u) : "memory");
    }
    if (done == 123) out[0] = 1.f;
  } else if (LOADERS) {  // warps 4-7 drain the other accumulator all the time
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    float acc = 0.f;
    while (!stop) {
      float v[32];
#pragma unroll
      for (int e = 0; e < 8; e++) { tmem_ld32(base + 32 * e, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += v[0] + v[31]; }
    }
    if (acc == 123.456f) out[0] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int N, int LOADERS, int BURST = 1>
static void run(const char* name, int sms) {
  float* out; long long* res;
  CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&res, 16));
  const int iters = 2000;
  const size_t sm = 128 * 32 + 2048 * 32;
  CK(cudaFuncSetAttribute(bench<N, LOADERS, BURST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  for (int rep = 0; rep < 2; rep++) { bench<N, LOADERS, BURST><<<sms, 256, sm>>>(out, iters, res); CK(cudaDeviceSynchronize()); }
  long long h[2]; CK(cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost));
  printf("%-44s issue -> barrier complete: avg %6.1f clk, max %lld\n", name, (double)h[0] / iters, h[1]);
  cudaFree(out); cudaFree(res);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  run<256, 0>("M128 N256 K16, alone", p.multiProcessorCount);
  run<256, 1>("M128 N256 K16, 4 warps draining TMEM", p.multiProcessorCount);
  run<128, 0>("M128 N128 K16, alone", p.multiProcessorCount);
  run<128, 1>("M128 N128 K16, 4 warps draining TMEM", p.multiProcessorCount);
  run<64, 0>("M128 N64 K16, alone", p.multiProcessorCount);
  run<256, 2>("M128 N256 K16, 4 warps polling try_wait", p.multiProcessorCount);
  run<256, 3>("M128 N256 K16, 4 warps polling test_wait", p.multiProcessorCount);
  run<256, 2, 16>("M128 N256 K16 x16, 4 warps polling try_wait", p.multiProcessorCount);
  run<256, 0, 2>("M128 N256 K16 x2 per commit", p.multiProcessorCount);
  run<256, 0, 4>("M128 N256 K16 x4 per commit", p.multiProcessorCount);
  run<256, 0, 16>("M128 N256 K16 x16 per commit", p.multiProcessorCount);
  run<256, 1, 16>("M128 N256 K16 x16 per commit, 4 warps draining", p.multiProcessorCount);
  run<128, 0, 16>("M128 N128 K16 x16 per commit", p.multiProcessorCount);
  return 0;
}
