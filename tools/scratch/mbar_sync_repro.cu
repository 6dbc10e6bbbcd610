// Minimal reproduction of the barrier pattern of nn_fwd_umma_kernel for compute-sanitizer --tool synccheck:
// twelve mbarriers initialised by thread 0 (fence.mbarrier_init + __syncthreads), four "helper" warps arrive on
// barrier 4 (count 4) once per step, sixteen "scan" warps test_wait on it two steps later; the barriers sit at the
// kernel's own shared-memory offset, next to a TMEM allocation slot, a shared-memory counter that is bumped with
// atomicAdd and cleared by a polling thread, and the helpers meet at a named barrier -- every ingredient of the kernel
// around its mbarriers except the tensor-core work itself.  Clean under synccheck in every one of these forms
// (profiles/r02_sanitizer.txt).  Development tool: nvcc -arch=sm_100a -o mbar_sync_repro.bin mbar_sync_repro.cu
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__global__ void __launch_bounds__(768) repro(int steps, int* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + 217472);  // the offset the barriers have in nn_fwd_umma_kernel
  const uint32_t bar0 = s32(bars);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 2; i++) {
      mbar_init(bar0 + 8u * i, 4);
      mbar_init(bar0 + 8u * (2 + i), 16);
      mbar_init(bar0 + 8u * (4 + i), 4);
      mbar_init(bar0 + 8u * (6 + i), 3);
      mbar_init(bar0 + 8u * (8 + i), 1);
      mbar_init(bar0 + 8u * (10 + i), 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 217472 + 12 * 8 + 8);
  volatile int* done_cnt = reinterpret_cast<volatile int*>(smem + 217472 + 12 * 8);
  if (tid == 0) { done_cnt[0] = 0; done_cnt[1] = 0; }
  if (warp == 0) {  // TMEM allocation writes its base address next to the barriers, as in the kernel
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  int spins = 0;
  if (warp < 16) {  // scan: step k writes "keys", arrives on keys_full(k&1), waits keys_empty(k&1) from step 2 on
    for (int k = 0; k < steps; k++) {
      if (k >= 2) while (!mbar_test(bar0 + 8u * (4 + (k & 1)), ((k >> 1) - 1) & 1)) spins++;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8u * (2 + (k & 1)));
    }
  } else if (warp < 20) {  // helper: waits keys_full(k&1), arrives keys_empty(k&1)
    for (int k = 0; k < steps; k++) {
      while (!mbar_test(bar0 + 8u * (2 + (k & 1)), (k >> 1) & 1)) spins++;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar0 + 8u * (4 + (k & 1)));
      asm volatile("bar.sync 2, 128;" ::: "memory");                                   // helper group barrier
      if ((tid & 127) == 0) atomicAdd(const_cast<int*>(done_cnt) + (k & 1), 1);      // refined-jobs counter
    }
  } else if (warp < 23) {  // stager: polls and clears the counters, as in the kernel
    if ((tid - 640) == 0)
      for (int k = 0; k < steps; k++) {
        while (done_cnt[k & 1] == 0) __nanosleep(100);
        done_cnt[k & 1] = 0;
      }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  if (tid == 0) out[blockIdx.x] = spins;
}
int main() {
  int* out;
  cudaMalloc(&out, 64 * sizeof(int));
  cudaFuncSetAttribute(repro, cudaFuncAttributeMaxDynamicSharedMemorySize, 217584);
  repro<<<8, 768, 217584>>>(8, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("repro: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
