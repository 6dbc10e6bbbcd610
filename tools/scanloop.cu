// Microbenchmark of the REAL tensor-core scan loop (mma_scan_range of nn_mma.cuh and candidate rewrites of it):
// W warps per SM each scan a staged 2048-target cloud `reps` times; reports clocks per HMMA per warp and the
// HMMA rate per SM.  Development tool.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../include -I../geometric_adv_b200/csrc -o scanloop.bin scanloop.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nn_mma.cuh"

using namespace ga;

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e = (x);                                                        \
    if (e != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

// V0: the plain loop (round 1): ptxas' own schedule
__device__ __forceinline__ void scan_v0(const MmaRows& R, const uint2* __restrict__ bfrag, int blk0, int blk1, int lane,
                                        MmaTrack& tr) {
#pragma unroll
  for (int r = 0; r < 8; r++) tr.c1[r] = tr.c2[r] = tr.c3[r] = kMmaBig;
#pragma unroll 1
  for (int blk = blk0; blk < blk1; blk++) {
    float rm[8];
#pragma unroll
    for (int r = 0; r < 8; r++) rm[r] = kMmaBig;
    const uint2* bp = bfrag + (size_t)blk * 16 * 32 + lane;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint2 bf = bp[j * 32];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float c[4];
        mma16816(c, R.a[i], bf.x, bf.y);
        rm[2 * i] = fmin3(rm[2 * i], c[0], c[1]);
        rm[2 * i + 1] = fmin3(rm[2 * i + 1], c[2], c[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const float key = __int_as_float((__float_as_int(rm[r]) & ~15) | blk);
      tr.c3[r] = fminf(tr.c3[r], fmaxf(tr.c2[r], key));
      tr.c2[r] = fminf(tr.c2[r], fmaxf(tr.c1[r], key));
      tr.c1[r] = fminf(tr.c1[r], key);
    }
  }
}

// V2: NSTAGE accumulator sets, one step per set, rolled loop over groups of NSTAGE steps; the HMMAs of a set are
// issued right after that set has been folded, so a result is consumed NSTAGE - 1 steps after its issue.
template <int NSTAGE>
__device__ __forceinline__ void scan_vs(const MmaRows& R, const uint2* __restrict__ bfrag, int blk0, int blk1, int lane,
                                        MmaTrack& tr) {
  static_assert(16 % NSTAGE == 0, "a block is a whole number of groups");
#pragma unroll
  for (int r = 0; r < 8; r++) tr.c1[r] = tr.c2[r] = tr.c3[r] = kMmaBig;
  if (blk0 >= blk1) return;
  const uint2* bp = bfrag + (size_t)blk0 * 16 * 32 + lane;
  const int steps = (blk1 - blk0) * 16;
  float acc[NSTAGE][4][4];
  uint2 bf[NSTAGE];
#pragma unroll
  for (int g = 0; g < NSTAGE; g++) {
    bf[g] = bp[(size_t)g * 32];
#pragma unroll
    for (int i = 0; i < 4; i++) mma16816(acc[g][i], R.a[i], bf[g].x, bf[g].y);
  }
#pragma unroll
  for (int g = 0; g < NSTAGE; g++) bf[g] = bp[(size_t)(NSTAGE + g) * 32];  // a range has >= 16 steps
  int s = 0;
#pragma unroll 1
  for (int blk = blk0; blk < blk1; blk++) {
    float rm[8];
#pragma unroll
    for (int r = 0; r < 8; r++) rm[r] = kMmaBig;
#pragma unroll 1
    for (int it = 0; it < 16 / NSTAGE; it++, s += NSTAGE) {
      const bool more = s + NSTAGE < steps;        // the next group exists
      const bool more2 = s + 2 * NSTAGE < steps;   // and the one after it
#pragma unroll
      for (int g = 0; g < NSTAGE; g++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          rm[2 * i] = fmin3(rm[2 * i], acc[g][i][0], acc[g][i][1]);
          rm[2 * i + 1] = fmin3(rm[2 * i + 1], acc[g][i][2], acc[g][i][3]);
        }
        if (more) {
#pragma unroll
          for (int i = 0; i < 4; i++) mma16816(acc[g][i], R.a[i], bf[g].x, bf[g].y);
          if (more2) bf[g] = bp[(size_t)(s + 2 * NSTAGE + g) * 32];
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const float key = __int_as_float((__float_as_int(rm[r]) & ~15) | blk);
      tr.c3[r] = fminf(tr.c3[r], fmaxf(tr.c2[r], key));
      tr.c2[r] = fminf(tr.c2[r], fmaxf(tr.c1[r], key));
      tr.c1[r] = fminf(tr.c1[r], key);
    }
  }
}

template <int VAR>
__global__ void __launch_bounds__(512, 1) bench(const float* __restrict__ q, const uint2* __restrict__ bsrc, int reps,
                                                long long* cyc, float* out) {
  extern __shared__ uint2 bfrag[];
  for (int i = threadIdx.x; i < 2048 * 4 + 64; i += blockDim.x) bfrag[i] = bsrc[i % (2048 * 4)];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  MmaRows R;
  mma_load_rows(R, q, 2048, 64 * warp, lane);
  float sink = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    MmaTrack tr;
    if (VAR == 0) scan_v0(R, bfrag, 0, 16, lane, tr);
    if (VAR == 1) mma_scan_range<4>(R, bfrag, 0, 16, lane, tr);
    if (VAR == 2) scan_vs<2>(R, bfrag, 0, 16, lane, tr);
    if (VAR == 4) scan_vs<4>(R, bfrag, 0, 16, lane, tr);
#pragma unroll
    for (int k = 0; k < 8; k++) sink += tr.c1[k] + tr.c2[k] + tr.c3[k];
    R.a[0][0] ^= (r & 1);  // keep the repetitions distinct
  }
  const long long t1 = clock64();
  if (sink == 123.456f) out[0] = sink;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int VAR>
static void run(const char* name, int warps, int sms, const float* q, const uint2* bsrc) {
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, 4));
  CK(cudaMalloc(&cyc, 8));
  const int reps = 8;
  const size_t smem = (2048 * 4 + 64) * sizeof(uint2);
  CK(cudaFuncSetAttribute(bench<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; rep++) bench<VAR><<<sms, warps * 32, smem>>>(q, bsrc, reps, cyc, out);
  CK(cudaDeviceSynchronize());
  long long h = 0;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const double per_hmma = (double)h / (reps * 1024.0);
  printf("%-44s warps/SM=%2d  clk per HMMA per warp = %6.2f  per scheduler = %6.2f  HMMA/clk/SM = %.3f\n", name, warps,
         per_hmma, per_hmma / ((warps + 3) / 4), warps / per_hmma);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  std::vector<float> hq(2048 * 3);
  std::vector<uint2> hb(2048 * 4);
  unsigned st = 12345;
  auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (st >> 8) * (1.0f / 16777216.0f) - 0.5f; };
  for (auto& v : hq) v = rnd();
  for (auto& v : hb) {  // bf16 pairs of plausible magnitude
    auto bf = [&](float f) { unsigned u; memcpy(&u, &f, 4); return u >> 16; };
    v = make_uint2(bf(rnd()) | (bf(rnd() * 0.004f) << 16), bf(rnd()) | (bf(rnd()) << 16));
  }
  float* q;
  uint2* b;
  CK(cudaMalloc(&q, hq.size() * 4));
  CK(cudaMalloc(&b, hb.size() * 8));
  CK(cudaMemcpy(q, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
  for (int w : {1, 4, 8, 12, 16}) run<0>("v0 plain loop (ptxas schedule)", w, sms, q, b);
  for (int w : {1, 4, 8, 12, 16}) run<1>("v1 mma_scan_range of nn_mma.cuh", w, sms, q, b);
  for (int w : {1, 4, 8, 12, 16}) run<2>("v2 two accumulator sets, rolled", w, sms, q, b);
  for (int w : {1, 4, 8, 12, 16}) run<4>("v4 four accumulator sets, rolled", w, sms, q, b);
  return 0;
}
