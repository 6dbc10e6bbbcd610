// Microbenchmark of the TMEM drain of nn_fwd_umma_kernel: tcgen05.ld.32x32b.x32 + the 4-chain FMNMX3 fold + key
// tracking, per 32-column tile, for 4 / 8 / 16 warps per SM.  Answers: what does one tile cost a scheduler when
// (a) only the loads run, (b) only the fold runs (values already in registers), (c) both, double-buffered.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o drainbench.bin drainbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
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
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fold(const float (&v)[32], int tile, float& c1, float& c2, float& c3) {
  float m0 = fmin3(v[0], v[1], v[2]), m1 = fmin3(v[8], v[9], v[10]);
  float m2 = fmin3(v[16], v[17], v[18]), m3 = fmin3(v[24], v[25], v[26]);
  m0 = fmin3(m0, v[3], v[4]); m1 = fmin3(m1, v[11], v[12]); m2 = fmin3(m2, v[19], v[20]); m3 = fmin3(m3, v[27], v[28]);
  m0 = fmin3(m0, v[5], v[6]); m1 = fmin3(m1, v[13], v[14]); m2 = fmin3(m2, v[21], v[22]); m3 = fmin3(m3, v[29], v[30]);
  m0 = fmin3(m0, v[7], m1); m2 = fmin3(m2, v[23], m3);
  const float m = fmin3(fmin3(m0, v[15], v[31]), m2, m2);
  const float key = __int_as_float((__float_as_int(m) & ~63) | tile);
  c3 = fminf(c3, fmaxf(c2, key)); c2 = fminf(c2, fmaxf(c1, key)); c1 = fminf(c1, key);
}
// MODE 0: loads only; 1: fold only (registers perturbed cheaply so nothing is hoisted); 2: double-buffered ld + fold
// (the kernel's tc_drain); 3: single buffer ld -> wait -> fold
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// MMA > 0: an extra warp (the last one) keeps issuing M128 N256 K16 bf16 MMAs into the accumulators the other warps
// read (values do not matter), one per MMA clocks (1 = back to back), until the drain warps are done.
template <int MODE, int MMA>
__global__ void __launch_bounds__(544) bench(float* out, int iters, long long* cyc) {
  extern __shared__ __align__(128) unsigned char opsm[];
  __shared__ unsigned long long mbar;
  __shared__ volatile int stop;
  const int nscan = (blockDim.x >> 5) - (MMA ? 1 : 0);
  if (MMA) {
    for (int i = threadIdx.x; i < (128 * 32 + 2048 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(opsm)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); stop = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (MMA && warp == nscan) {
    if ((threadIdx.x & 31) == 0) {
      const uint32_t desc_hi = (uint32_t)(256 >> 4) | (1u << 14);
      const uint32_t a_lo = ((smem_u32(opsm) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t b_lo = ((smem_u32(opsm + 4096) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t phase = 0; unsigned n = 0;
      while (!stop) {
        const long long t0 = clock64();
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_base + (n & 1) * 256),
                     "l"(((uint64_t)desc_hi << 32) | a_lo), "l"(((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)(n & 7) * (256 * 32 / 16))), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
        phase ^= 1; n++;
        while (MMA > 1 && clock64() - t0 < MMA && !stop) {}
      }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    return;
  }
  const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 256;
  float c1 = 3e38f, c2 = 3e38f, c3 = 3e38f, acc = 0.f;
  float v[32], w[32];
#pragma unroll
  for (int e = 0; e < 32; e++) { v[e] = (float)(threadIdx.x * 37 + e); w[e] = (float)(threadIdx.x * 11 + e * 3); }
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {   // one iteration = one 256-column step = 8 tiles
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        tmem_ld32(base + 32 * e, v); tmem_ld32(base + 32 * e + 32, w); ld_wait();
        acc += v[0] + v[31] + w[0] + w[31];
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        fold(v, e, c1, c2, c3); fold(w, e + 1, c1, c2, c3);
        v[(e * 5) & 31] += c1; w[(e * 7 + 3) & 31] += c2;  // keep the inputs changing
      }
    } else if (MODE == 2) {
      tmem_ld32(base, v); ld_wait();
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        tmem_ld32(base + 32 * (e + 1), w); fold(v, e, c1, c2, c3); ld_wait();
        if (e + 2 < 8) { tmem_ld32(base + 32 * (e + 2), v); fold(w, e + 1, c1, c2, c3); ld_wait(); }
        else fold(w, e + 1, c1, c2, c3);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; e++) { tmem_ld32(base + 32 * e, v); ld_wait(); fold(v, e, c1, c2, c3); }
    }
  }
  const long long t1 = clock64();
  if (MMA) { asm volatile("bar.sync 1, %0;" ::"r"(nscan * 32)); if (threadIdx.x == 0) stop = 1; }
  if (c1 + c2 + c3 + acc + v[5] + w[7] == 123.456f) out[0] = c1;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
template <int MODE, int MMA = 0>
static void run(const char* name, int threads, int sms) {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8));
  const int iters = 1 << 12;
  const size_t sm = MMA ? 128 * 32 + 2048 * 32 : 0;
  if (MMA) CK(cudaFuncSetAttribute(bench<MODE, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  for (int rep = 0; rep < 2; rep++) { bench<MODE, MMA><<<sms, threads + (MMA ? 32 : 0), sm>>>(out, iters, cyc); CK(cudaDeviceSynchronize()); }
  long long h = 0; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const int warps = threads / 32;
  const double per_tile_warp = (double)h / (iters * 8.0);
  printf("%-34s warps/SM=%2d  clk per tile per warp = %6.1f  per scheduler-tile = %6.1f  -> %.1f values/clk/SM\n", name, warps,
         per_tile_warp, per_tile_warp / (warps / 4.0), 1024.0 * warps / per_tile_warp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  for (int th : {128, 256, 512}) {
    run<0>("tcgen05.ld x32 only", th, sms);
    run<1>("fold only (16 FMNMX3 + key + 5 FMNMX)", th, sms);
    run<2>("ld + fold, double-buffered", th, sms);
    run<3>("ld -> wait -> fold", th, sms);
  }
  printf("-- with an MMA stream (M128 N256 K16) into the accumulators being read --\n");
  run<0, 1>("ld only, MMAs back to back", 512, sms);
  run<1, 1>("fold only, MMAs back to back", 512, sms);
  run<3, 1>("ld -> wait -> fold, MMAs back to back", 512, sms);
  run<0, 400>("ld only, one MMA per 400 clk", 512, sms);
  run<3, 400>("ld -> wait -> fold, one MMA per 400 clk", 512, sms);
  run<3, 800>("ld -> wait -> fold, one MMA per 800 clk", 512, sms);
  return 0;
}
