// Microbenchmark: the scan / issuer pair of nn_fwd_umma_kernel alone (16 drain warps, one issuing thread, two
// 256-column accumulators, acc_full / acc_empty mbarriers), to separate the cost of the handshake from the cost of
// the drain.  Reports clocks per 256-column step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scanbench.bin scanbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
#define LD32(base, off, v, o) asm volatile( \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
      "%29,%30,%31}, [%32];" \
      : "=f"(v[o+0]), "=f"(v[o+1]), "=f"(v[o+2]), "=f"(v[o+3]), "=f"(v[o+4]), "=f"(v[o+5]), "=f"(v[o+6]), "=f"(v[o+7]), "=f"(v[o+8]), \
        "=f"(v[o+9]), "=f"(v[o+10]), "=f"(v[o+11]), "=f"(v[o+12]), "=f"(v[o+13]), "=f"(v[o+14]), "=f"(v[o+15]), "=f"(v[o+16]), \
        "=f"(v[o+17]), "=f"(v[o+18]), "=f"(v[o+19]), "=f"(v[o+20]), "=f"(v[o+21]), "=f"(v[o+22]), "=f"(v[o+23]), "=f"(v[o+24]), \
        "=f"(v[o+25]), "=f"(v[o+26]), "=f"(v[o+27]), "=f"(v[o+28]), "=f"(v[o+29]), "=f"(v[o+30]), "=f"(v[o+31]) \
      : "r"((base) + (off)) : "memory")
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tile_min(const float (&v)[64], int o) {
  float m0 = fmin3(v[o + 0], v[o + 1], v[o + 2]), m1 = fmin3(v[o + 8], v[o + 9], v[o + 10]);
  float m2 = fmin3(v[o + 16], v[o + 17], v[o + 18]), m3 = fmin3(v[o + 24], v[o + 25], v[o + 26]);
  m0 = fmin3(m0, v[o + 3], v[o + 4]); m1 = fmin3(m1, v[o + 11], v[o + 12]); m2 = fmin3(m2, v[o + 19], v[o + 20]); m3 = fmin3(m3, v[o + 27], v[o + 28]);
  m0 = fmin3(m0, v[o + 5], v[o + 6]); m1 = fmin3(m1, v[o + 13], v[o + 14]); m2 = fmin3(m2, v[o + 21], v[o + 22]); m3 = fmin3(m3, v[o + 29], v[o + 30]);
  m0 = fmin3(m0, v[o + 7], m1); m2 = fmin3(m2, v[o + 23], m3);
  return fmin3(fmin3(m0, v[o + 15], v[o + 31]), m2, m2);
}
__device__ __forceinline__ void fold2(const float (&v)[64], int tile, float& c1, float& c2, float& c3) {
  const float ma = tile_min(v, 0), mb = tile_min(v, 32);
  const float ka = __int_as_float((__float_as_int(ma) & ~63) | tile), kb = __int_as_float((__float_as_int(mb) & ~63) | (tile + 1));
  const float lo = fminf(ka, kb), hi = fmaxf(ka, kb);
  const float x = fmaxf(c1, lo), y = fminf(c2, hi);
  c3 = fminf(fmaxf(x, y), c3); c2 = fminf(x, y); c1 = fminf(c1, lo);
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// V bit 0: scan warps skip the loads and the fold (handshake only)      bit 1: no handshake at all (free-running)
//   bit 2: scan warps wait with a test_wait spin instead of try_wait    bit 3: hand back AFTER the fold
//   bit 4: the issuer waits with try_wait instead of spinning           bit 6: one chain at a time, the issuer times
//   issue -> acc_full visible to it -> acc_empty complete              bit 7: barriers 128 B apart
//   bit 9: the scan warps poll a counter in shared memory that the issuer bumps when IT sees acc_full (chain mode)
//   bit 8: four independent pipelines, one per column quarter: N=64 MMAs, acc_empty counts that quarter's four warps
template <int V, int NSCAN>
__global__ void __launch_bounds__(768) bench(float* out, int iters, long long* res) {
  extern __shared__ __align__(128) unsigned char opsm[];
  __shared__ __align__(128) unsigned long long bars[256];  // full[2], empty[2]; 16 words apart with V & 128
  __shared__ uint32_t tmem_base;
  __shared__ volatile unsigned full_count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int BS = (V & 128) ? 16 : 1;   // barrier stride in 8-byte words
  constexpr uint32_t BB = BS * 8;
  for (int i = tid; i < (128 * 32 + 2048 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(opsm)[i] = 0x3c003c00u + i;
  constexpr bool PG = (V & 256) != 0;
  if (PG && tid == 0) {
    for (int i = 0; i < 8; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i * BS])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[(8 + i) * BS])), "r"(NSCAN / 4));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 0) full_count = 0;
  if (!PG && tid == 0) {
    for (int i = 0; i < 2; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i * BS])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[(2 + i) * BS])), "r"(NSCAN));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
  uint32_t bfull = smem_u32(&bars[0]), bempty = smem_u32(&bars[(PG ? 8 : 2) * BS]);
  if (warp < NSCAN) {
    constexpr int G = NSCAN / 4;            // column groups
    constexpr int COLS = 256 / G;           // columns per warp and step (64 for 16 warps, 128 for 8)
    const int g = warp >> 2;
    uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * COLS);
    if (PG) { bfull += g * 2 * BB; bempty += g * 2 * BB; }
    asm volatile("" : "+r"(trow), "+r"(bfull), "+r"(bempty));
    float c1 = 3e38f, c2 = 3e38f, c3 = 3e38f;
    bool ready = false;
    const long long t0 = clock64();
#pragma unroll 1
    for (unsigned ns = 0; ns < (unsigned)iters; ns++) {
      const uint32_t bo = (ns & 1) * BB;
      if (!(V & 2)) {
        if (V & 512) {
          while (full_count <= ns) {}
          __threadfence_block();
        } else if (!ready) {
          if (V & 4) { while (!mbar_test(bfull + bo, (ns >> 1) & 1)) {} }
          else { while (!mbar_try(bfull + bo, (ns >> 1) & 1)) {} }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (V & 1) {
        if (!(V & 2)) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bempty + bo);
          ready = mbar_test(bfull + (bo ^ BB), ((ns + 1) >> 1) & 1);
        }
      } else {
#pragma unroll
        for (int h = 0; h < COLS / 64; h++) {
          float v[64];
          LD32(trow + (ns & 1) * 256, h * 64, v, 0);
          LD32(trow + (ns & 1) * 256, h * 64 + 32, v, 32);
          ld_wait();
          if (!(V & 2) && !(V & 8) && h == COLS / 64 - 1) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bempty + bo);
            ready = mbar_test(bfull + (bo ^ BB), ((ns + 1) >> 1) & 1);
          }
          fold2(v, (int)(ns & 7) * 8 + g * 2 + h * 2, c1, c2, c3);
        }
        if (!(V & 2) && (V & 8)) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bempty + bo);
          ready = mbar_test(bfull + (bo ^ BB), ((ns + 1) >> 1) & 1);
        }
      }
    }
    const long long t1 = clock64();
    if (c1 + c2 + c3 == 123.456f) out[0] = c1;
    if (tid == 0 && blockIdx.x == 0) res[0] = t1 - t0;
  } else if (warp == NSCAN) {
    if (lane == 0) {
      const uint32_t desc_hi = (uint32_t)(256 >> 4) | (1u << 14);
      const uint32_t a_lo = ((smem_u32(opsm) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t b_lo = ((smem_u32(opsm + 4096) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      long long lat_full = 0, lat_empty = 0, tissue = 0;
      if (PG) {
        const uint32_t idesc64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        unsigned sg[4] = {0, 0, 0, 0};
        unsigned left = 4 * (unsigned)iters;
        while (left) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const unsigned s = sg[g];
            if (s >= (unsigned)iters) continue;
            if (s >= 2 && !mbar_test(bempty + (g * 2 + (s & 1)) * BB, ((s >> 1) - 1) & 1)) continue;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (s & 1) * 256 + g * 64),
                         "l"(((uint64_t)desc_hi << 32) | a_lo), "l"(((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)((s & 7) * 256 + g * 64) * (32 / 16))), "r"(idesc64), "r"(0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bfull + (g * 2 + (s & 1)) * BB) : "memory");
            sg[g] = s + 1;
            left--;
          }
        }
      } else
      for (unsigned s = 0; s < (unsigned)iters; s++) {
        if (V & 64) {
          if (s >= 1) {  // previous step completely handed back before the next is issued
            const long long t0 = clock64();
            while (!mbar_test(bfull + ((s - 1) & 1) * BB, ((s - 1) >> 1) & 1)) {}
            const long long t1 = clock64();
            if (V & 512) { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __threadfence_block(); full_count = s; }
            while (!mbar_test(bempty + ((s - 1) & 1) * BB, ((s - 1) >> 1) & 1)) {}
            const long long t2 = clock64();
            lat_full += t1 - tissue; lat_empty += t2 - t1; (void)t0;
          }
        } else
        if (!(V & 2) && s >= 2) {
          if (V & 16) { while (!mbar_try(bempty + (s & 1) * BB, ((s >> 1) - 1) & 1)) {} }
          else { while (!mbar_test(bempty + (s & 1) * BB, ((s >> 1) - 1) & 1)) {} }
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tissue = clock64();
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (s & 1) * 256),
                     "l"(((uint64_t)desc_hi << 32) | a_lo), "l"(((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)(s & 7) * (256 * 32 / 16))), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bfull + (s & 1) * BB) : "memory");
        if (V & 2) { const long long t0 = clock64(); while (clock64() - t0 < 300) {} }
      }
      if ((V & 64) && (V & 512)) {  // the last step still has to be announced
        const unsigned s = (unsigned)iters;
        while (!mbar_test(bfull + ((s - 1) & 1) * BB, ((s - 1) >> 1) & 1)) {}
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __threadfence_block(); full_count = s;
      }
      if ((V & 64) && blockIdx.x == 0) { res[1] = lat_full; res[2] = lat_empty; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int V, int NSCAN>
static void run(const char* name, int sms, int extra_warps = 0) {
  float* out; long long* res;
  CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&res, 24)); CK(cudaMemset(res, 0, 24));
  const int iters = 4000;
  const size_t sm = 128 * 32 + 2048 * 32;
  CK(cudaFuncSetAttribute(bench<V, NSCAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  for (int rep = 0; rep < 2; rep++) { bench<V, NSCAN><<<sms, (NSCAN + 1 + extra_warps) * 32, sm>>>(out, iters, res); CK(cudaDeviceSynchronize()); }
  long long h[3] = {0, 0, 0}; CK(cudaMemcpy(h, res, 24, cudaMemcpyDeviceToHost));
  printf("%-72s %7.1f clk per 256-column step", name, (double)h[0] / iters);
  if (V & 64) printf("   issue -> full %6.1f, full -> all handed back %6.1f", (double)h[1] / (iters - 1), (double)h[2] / (iters - 1));
  printf("\n"); fflush(stdout);
  cudaFree(out); cudaFree(res);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  run<0, 16>("16 scan warps, handshake, ld x2 -> hand back -> fold2 (the kernel)", sms);
  run<1, 16>("16 scan warps, handshake only (no loads, no fold)", sms);
  run<2, 16>("16 scan warps, no handshake (free-running loads + fold, MMA per 300 clk)", sms);
  run<4, 16>("16 scan warps, handshake, scan waits by spinning", sms);
  run<5, 16>("16 scan warps, handshake only, scan waits by spinning", sms);
  run<8, 16>("16 scan warps, handshake, hand back after the fold", sms);
  run<16, 16>("16 scan warps, handshake, issuer waits with try_wait", sms);
  run<17, 16>("16 scan warps, handshake only, issuer waits with try_wait", sms);
  run<0, 16>("16 scan warps, handshake + 7 idle warps parked at the closing barrier", sms, 7);
  run<65, 16>("one chain, handshake only", sms);
  run<64, 16>("one chain, loads + fold", sms);
  run<69, 16>("one chain, handshake only, scan spins", sms);
  run<128, 16>("16 scan warps, handshake, barriers 128 B apart", sms);
  run<129, 16>("16 scan warps, handshake only, barriers 128 B apart", sms);
  run<193, 16>("one chain, handshake only, barriers 128 B apart", sms);
  run<65, 4>("one chain, handshake only, 4 scan warps", sms);
  run<65, 8>("one chain, handshake only, 8 scan warps", sms);
  run<577, 16>("one chain, handshake only, 16 scan warps poll a shared-memory counter", sms);
  run<576, 16>("one chain, loads + fold, 16 scan warps poll a shared-memory counter", sms);
  run<256, 16>("four pipelines (N=64 MMAs, one per column quarter), loads + fold", sms);
  run<257, 16>("four pipelines, handshake only", sms);
  run<384, 16>("four pipelines, loads + fold, barriers 128 B apart", sms);
  run<0, 8>("8 scan warps (128 columns each), handshake", sms);
  run<1, 8>("8 scan warps, handshake only", sms);
  run<2, 8>("8 scan warps, no handshake", sms);
  return 0;
}
