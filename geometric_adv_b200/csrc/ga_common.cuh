// Shared helpers for libga_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ga_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libga_b200 is written for sm_100a (B200) only"
#endif

namespace ga {

// ---- host side -------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);
void count_launch(int n = 1);
void note_kernel(const char* name);

// Completion tickets (core.cu): a small per-device array through which the forward search tells the
// gradient kernel, per batch element, that its dist/idx rows are final -- see nn_distance_bwd.cu.
constexpr int kTicketSlots = 4096;  // + 8 words of debug counters behind the slots (ga_debug_ticket_stats)
unsigned long long* ticket_buffer(cudaStream_t st);  // nullptr if it cannot be provided right now
unsigned long long next_call_id();
int* frame_hint(cudaStream_t st, volatile int** host_view);  // device view of the per-device hint word, or nullptr
struct LastForward {  // the calling thread's most recent ticketed forward launch
  cudaStream_t stream;
  const int* idx1;
  const int* idx2;
  int b, n, m, expected;
  unsigned long long call_id;
  unsigned long long* ticket;
};
LastForward& last_forward();
// Streamed ingest (FwdArgs::ready, nn_search.cuh): the host entry arms the calling thread's next forward launch.
struct ReadyArm {
  const int* flags;
  int per;
  int* abort_word;
  int pdl;  // launch the forward as a programmatic dependent of the kernel before it (the SM-driven ingest)
};
extern thread_local ReadyArm t_ready_arm;                // nn_distance_fwd_mma.cu
bool fwd_ready_supported(int b, int n, int m);          // nn_distance_fwd.cu: would that launch honour the flags?
int sm_count();

#define GA_CUDA_TRY(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return ga::cuda_fail(_e, #expr); \
  } while (0)

#define GA_LAUNCH_CHECK(what)                                 \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return ga::cuda_fail(_e, what);    \
    ga::count_launch();                                       \
    ga::note_kernel(what);                                    \
  } while (0)

static inline cudaStream_t as_stream(ga_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device side -----------------------------------------------------------
// Packed fp32x2 FMA (sm_100 FFMA2): one issue slot for two FMAs.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// Three-input minimum (sm_100 FMNMX3); NaN operands are ignored like fminf.
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// Squared distance in the two pinned arithmetics (see ga_b200.h).  x = target - query.
template <int MODE>
__device__ __forceinline__ float sqdist(float tx, float ty, float tz, float qx, float qy, float qz) {
  float x = __fsub_rn(tx, qx), y = __fsub_rn(ty, qy), z = __fsub_rn(tz, qz);
  if (MODE == GA_MODE_CPU_EXACT) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  } else {
    return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
  }
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ int ticket_slot(unsigned long long call_id, int batch) {
  return (int)((call_id * 2654435761ull + (unsigned long long)batch) & (unsigned long long)(kTicketSlots - 1));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ga
