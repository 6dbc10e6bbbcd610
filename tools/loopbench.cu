// Development tool: the filter inner loop in isolation, several formulations, vs warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../geometric_adv_b200/csrc/nn_tiles.cuh"
namespace ga { void set_error(const char*, ...) {} int cuda_fail(cudaError_t e, const char*) { return (int)e; } void count_launch(int) {} int sm_count() { return 148; } }
using namespace ga;

constexpr int CH = 2048;

// V1: library filter_scan (FFMA2 broadcast form)
template <int Q, int T>
__global__ void k_v1(float* out, int reps, float seed) {
  extern __shared__ float4 tgt[];
  for (int i = threadIdx.x; i < CH + 2 * kPipeU; i += blockDim.x) tgt[i] = make_float4(i * 1e-3f, seed, i * 2e-3f, 0.5f);
  __syncthreads();
  float ax[Q], ay[Q], az[Q], acc[Q];
  for (int j = 0; j < Q; j++) { ax[j] = seed + threadIdx.x + j; ay[j] = ax[j] * 2; az[j] = ax[j] * 3; acc[j] = 1e30f; }
  for (int r = 0; r < reps; r++)
    filter_scan<Q, T>(tgt, CH / T, ax, ay, az, [&](int, const float(&tm)[Q]) {
#pragma unroll
      for (int j = 0; j < Q; j++) acc[j] = fminf(acc[j], tm[j]);
    });
  float t = 0;
  for (int j = 0; j < Q; j++) t += acc[j];
  if (t == 1.2345f) out[0] = t;
}

// V2: scalar FFMA, targets as float4 {x,y,z,n}, 2 targets per step
template <int Q, int T>
__global__ void k_v2(float* out, int reps, float seed) {
  extern __shared__ float4 tgt[];
  for (int i = threadIdx.x; i < CH + 2 * kPipeU; i += blockDim.x) tgt[i] = make_float4(i * 1e-3f, seed, i * 2e-3f, 0.5f);
  __syncthreads();
  float ax[Q], ay[Q], az[Q], acc[Q];
  for (int j = 0; j < Q; j++) { ax[j] = seed + threadIdx.x + j; ay[j] = ax[j] * 2; az[j] = ax[j] * 3; acc[j] = 1e30f; }
  constexpr int U = 4;
  for (int r = 0; r < reps; r++) {
    float4 buf[2][2 * U];
#pragma unroll
    for (int e = 0; e < 2 * U; e++) buf[0][e] = tgt[e];
#pragma unroll 1
    for (int tile = 0; tile < CH / T; tile++) {
      const float4* tp = tgt + tile * T;
#pragma unroll
      for (int blk = 0; blk < (T / 2) / U; blk++) {
#pragma unroll
        for (int e = 0; e < 2 * U; e++) buf[(blk + 1) & 1][e] = tp[(blk + 1) * 2 * U + e];
#pragma unroll
        for (int pp = 0; pp < U; pp++) {
          const float4 u = buf[blk & 1][2 * pp], v = buf[blk & 1][2 * pp + 1];
#pragma unroll
          for (int j = 0; j < Q; j++) {
            float f0 = fmaf(az[j], u.z, u.w); f0 = fmaf(ay[j], u.y, f0); f0 = fmaf(ax[j], u.x, f0);
            float f1 = fmaf(az[j], v.z, v.w); f1 = fmaf(ay[j], v.y, f1); f1 = fmaf(ax[j], v.x, f1);
            acc[j] = fmin3(acc[j], f0, f1);
          }
        }
      }
    }
  }
  float t = 0;
  for (int j = 0; j < Q; j++) t += acc[j];
  if (t == 1.2345f) out[0] = t;
}

// V3: registers only (no LDS): FFMA2 x3 + FMNMX3, operands loop-invariant, Q chains
template <int Q>
__global__ void k_v3(float* out, int reps, float seed) {
  float ax[Q], ay[Q], az[Q], acc[Q];
  for (int j = 0; j < Q; j++) { ax[j] = seed + threadIdx.x + j; ay[j] = ax[j] * 2; az[j] = ax[j] * 3; acc[j] = 1e30f; }
  float4 u = make_float4(seed, seed * 2, seed * 3, seed * 4), v = make_float4(seed * 5, seed * 6, seed * 7, seed * 8);
  for (int r = 0; r < reps * 1024; r++) {
#pragma unroll
    for (int j = 0; j < Q; j++) {
      float2 f = ffma2(make_float2(az[j], az[j]), make_float2(v.x, v.y), make_float2(v.z, v.w));
      f = ffma2(make_float2(ay[j], ay[j]), make_float2(u.z, u.w), f);
      f = ffma2(make_float2(ax[j], ax[j]), make_float2(u.x, u.y), f);
      acc[j] = fmin3(acc[j], f.x, f.y);
    }
    u.x += 1e-7f; v.z -= 1e-7f;  // keep it loop-variant
  }
  float t = 0;
  for (int j = 0; j < Q; j++) t += acc[j];
  if (t == 1.2345f) out[0] = t;
}

template <class F>
void run(const char* name, F launch, int blocks, int threads, int q, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double warps_per_smsp = (double)blocks * threads / 32.0 / (148 * 4);
  double steps = (double)reps * 1024;           // pair-steps per warp
  double cyc = best * 1e-3 * 1.965e9;          // assumes max clock
  double cyc_per_step_smsp = cyc / (steps * warps_per_smsp);
  double ideal = 6.0 * q;                       // FMA-pipe cycles per pair-step (3 FFMA2 x 2 cyc per query)
  printf("%-14s Q=%d warps/SMSP=%.2f  %7.3f ms  %6.2f cyc/pair-step/SMSP  FMA-util %5.1f%%  (%s)\n", name, q,
         warps_per_smsp, best, cyc_per_step_smsp, 100.0 * ideal / cyc_per_step_smsp, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out; cudaMalloc(&out, 4);
  const int reps = 64;
  const size_t smem = (CH + 2 * kPipeU) * 16;
#define SETS(K) cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
  for (int thr : {64, 128}) for (int occ : {1, 2, 3, 4, 6}) {
    int blocks = 148 * occ;
    if (thr * occ > 1024 || occ * smem > 220 * 1024) continue;
    printf("--- %d threads x %d CTA/SM\n", thr, occ);
    SETS((k_v1<2, 32>)); run("v1 ffma2 T32", [&] { k_v1<2, 32><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 2, reps);
    SETS((k_v1<4, 32>)); run("v1 ffma2 T32", [&] { k_v1<4, 32><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 4, reps);
    SETS((k_v1<4, 64>)); run("v1 ffma2 T64", [&] { k_v1<4, 64><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 4, reps);
    SETS((k_v1<6, 32>)); run("v1 ffma2 T32", [&] { k_v1<6, 32><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 6, reps);
    SETS((k_v1<8, 32>)); run("v1 ffma2 T32", [&] { k_v1<8, 32><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 8, reps);
    SETS((k_v2<4, 32>)); run("v2 ffma  T32", [&] { k_v2<4, 32><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 4, reps);
    SETS((k_v2<8, 32>)); run("v2 ffma  T32", [&] { k_v2<8, 32><<<blocks, thr, smem>>>(out, reps, 1.f); }, blocks, thr, 8, reps);
    run("v3 regs only", [&] { k_v3<4><<<blocks, thr>>>(out, reps, 1.f); }, blocks, thr, 4, reps);
    run("v3 regs only", [&] { k_v3<8><<<blocks, thr>>>(out, reps, 1.f); }, blocks, thr, 8, reps);
  }
  return 0;
}
