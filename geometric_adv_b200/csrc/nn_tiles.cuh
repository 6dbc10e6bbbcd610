// Shared building blocks of the Chamfer forward and kNN kernels: staging of a
// target cloud chunk in shared memory, and the FFMA2/FMNMX3 filter scan.
//
// Shared-memory layout ("pair-SoA"): two consecutive targets t0,t1 occupy two
// float4: {x0,x1,y0,y1} {z0,z1,n0,n1} with n = |t|^2, so that one LDS.128 pair
// feeds three packed FMAs that evaluate the filter
//     f(q,t) = |t|^2 - 2 q.t
// for two targets at once.  Padding targets are (0,0,0,+inf): their filter value
// is +inf (or NaN) and never wins a minimum.
#pragma once
#include "ga_common.cuh"

namespace ga {

constexpr int kPipeU = 4;  // pair-steps per software-pipeline block

// Stage targets [c0, c0 + ntile*T) of `tpts` (nt points, AoS xyz) into `tgt`.
// Returns the CTA-wide max |coordinate| of the staged chunk (NaNs dropped).
// Contains two __syncthreads(): one before overwriting `tgt`, one after.
template <int THREADS, int T>
__device__ __forceinline__ float stage_targets(float4* __restrict__ tgt, float* __restrict__ red,
                                               const float* __restrict__ tpts, int c0, int nt, int ntile,
                                               int tid) {
  const float kInf = __int_as_float(0x7f800000);
  __syncthreads();  // previous contents fully consumed
  float lmax = 0.0f;
  // SB pairs per thread per round: all 6*SB loads are issued before any is used
  constexpr int SB = 4;
  const int npair = ntile * (T / 2);
  for (int p0 = tid; p0 < npair; p0 += THREADS * SB) {
    float c[SB][6];
#pragma unroll
    for (int s = 0; s < SB; s++) {
      const int p = p0 + s * THREADS;
      const long long g = (long long)c0 + 2 * p;
#pragma unroll
      for (int e = 0; e < 6; e++) {
        const bool ok = p < npair && g + (e >= 3) < nt;
        c[s][e] = ok ? __ldg(tpts + g * 3 + e) : 0.0f;
      }
    }
#pragma unroll
    for (int s = 0; s < SB; s++) {
      const int p = p0 + s * THREADS;
      if (p >= npair) break;
      const int g = c0 + 2 * p;
      float n0 = kInf, n1 = kInf;
      if (g < nt) {
        n0 = fmaf(c[s][2], c[s][2], fmaf(c[s][1], c[s][1], c[s][0] * c[s][0]));
        lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(c[s][0]), fabsf(c[s][1])), fabsf(c[s][2])));
      }
      if (g + 1 < nt) {
        n1 = fmaf(c[s][5], c[s][5], fmaf(c[s][4], c[s][4], c[s][3] * c[s][3]));
        lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(c[s][3]), fabsf(c[s][4])), fabsf(c[s][5])));
      }
      tgt[2 * p] = make_float4(c[s][0], c[s][3], c[s][1], c[s][4]);
      tgt[2 * p + 1] = make_float4(c[s][2], c[s][5], n0, n1);
    }
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  float bm = 0.0f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; w++) bm = fmaxf(bm, red[w]);
  return bm;
}

// Filter values of one staged pair for one query.
__device__ __forceinline__ float2 filter_pair(const float4 u, const float4 v, float ax2, float ay2, float az2) {
  float2 f = ffma2(make_float2(az2, az2), make_float2(v.x, v.y), make_float2(v.z, v.w));
  f = ffma2(make_float2(ay2, ay2), make_float2(u.z, u.w), f);
  f = ffma2(make_float2(ax2, ax2), make_float2(u.x, u.y), f);
  return f;
}

// Filter scan over `ntile` tiles of T targets for Q queries per thread.  After
// each tile, op(tile, tm) receives the per-query minima of the filter over that
// tile.  Loads run one pipeline block ahead of the FMAs (the block after the
// last tile reads the kPipeU-pair pad behind the staged chunk).
template <int Q, int T, class TileOp>
__device__ __forceinline__ void filter_scan(const float4* __restrict__ tgt, int ntile, const float (&ax2)[Q],
                                            const float (&ay2)[Q], const float (&az2)[Q], TileOp&& op) {
  constexpr int U = kPipeU;
  constexpr int NB = (T / 2) / U;
  static_assert((T / 2) % (2 * U) == 0, "tile must hold an even number of pipeline blocks");
  const float kInf = __int_as_float(0x7f800000);
  float4 buf[2][2 * U];
#pragma unroll
  for (int e = 0; e < 2 * U; e++) buf[0][e] = tgt[e];
#pragma unroll 1
  for (int tile = 0; tile < ntile; tile++) {
    float tm[Q];
#pragma unroll
    for (int j = 0; j < Q; j++) tm[j] = kInf;
    const float4* tp = tgt + (size_t)tile * T;
#pragma unroll
    for (int blk = 0; blk < NB; blk++) {
#pragma unroll
      for (int e = 0; e < 2 * U; e++) buf[(blk + 1) & 1][e] = tp[(blk + 1) * 2 * U + e];
#pragma unroll
      for (int pp = 0; pp < U; pp++) {
        const float4 u = buf[blk & 1][2 * pp];
        const float4 v = buf[blk & 1][2 * pp + 1];
#pragma unroll
        for (int j = 0; j < Q; j++) {
          float2 f = ffma2(make_float2(az2[j], az2[j]), make_float2(v.x, v.y), make_float2(v.z, v.w));
          f = ffma2(make_float2(ay2[j], ay2[j]), make_float2(u.z, u.w), f);
          f = ffma2(make_float2(ax2[j], ax2[j]), make_float2(u.x, u.y), f);
          tm[j] = fmin3(tm[j], f.x, f.y);
        }
      }
    }
    op(tile, tm);
  }
}

// Second scan for kNN: same pipelined walk as filter_scan, but every target whose filter value
// is <= thr[j] is appended (chunk-local index, 16 bit) to query j's queue.  Appends are
// predicated stores through a running pointer (3-4 instructions per target, no branch, the warp
// stays in step).  The queue has cap + T rows: a tile can add at most T entries, and the fill
// level is checked once per tile; a query that passes `cap` stops collecting and reports
// cnt = cap + 1 (overflow: the caller falls back to the exact warp path).
template <int Q, int T>
__device__ __forceinline__ void filter_collect(const float4* __restrict__ tgt, int ntile, const float (&ax2)[Q],
                                               const float (&ay2)[Q], const float (&az2)[Q], float (&thr)[Q],
                                               int (&cnt)[Q], unsigned short* __restrict__ queue, int qstride,
                                               int jstride, int cap) {
  constexpr int U = kPipeU;
  constexpr int NB = (T / 2) / U;
  unsigned short* qp[Q];
  bool over[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    qp[j] = queue + j * jstride;
    over[j] = false;
  }
  float4 buf[2][2 * U];
#pragma unroll
  for (int e = 0; e < 2 * U; e++) buf[0][e] = tgt[e];
#pragma unroll 1
  for (int tile = 0; tile < ntile; tile++) {
    const float4* tp = tgt + (size_t)tile * T;
    const int gb = tile * T;
#pragma unroll
    for (int blk = 0; blk < NB; blk++) {
#pragma unroll
      for (int e = 0; e < 2 * U; e++) buf[(blk + 1) & 1][e] = tp[(blk + 1) * 2 * U + e];
#pragma unroll
      for (int pp = 0; pp < U; pp++) {
        const float4 u = buf[blk & 1][2 * pp];
        const float4 v = buf[blk & 1][2 * pp + 1];
#pragma unroll
        for (int j = 0; j < Q; j++) {
          const float2 f = filter_pair(u, v, ax2[j], ay2[j], az2[j]);
          if (!(f.x > thr[j])) {  // NaN filter value / threshold counts as a candidate
            *qp[j] = (unsigned short)(gb + (blk * U + pp) * 2);
            qp[j] += qstride;
          }
          if (!(f.y > thr[j])) {
            *qp[j] = (unsigned short)(gb + (blk * U + pp) * 2 + 1);
            qp[j] += qstride;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < Q; j++) {
      if ((int)(qp[j] - (queue + j * jstride)) > cap * qstride) {  // too many: stop collecting
        over[j] = true;
        qp[j] = queue + j * jstride;
        thr[j] = -__int_as_float(0x7f800000);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < Q; j++) cnt[j] = over[j] ? cap + 1 : (int)(qp[j] - (queue + j * jstride)) / qstride;
}

// Window of the filter (see nn_distance_fwd.cu): W = 128u (A+Bm)^2 + denormal slack.
__device__ __forceinline__ float filter_window(float qabs, float bm) {
  const float s = (qabs + bm) * 1.0001f;
  return fmaf(4.0f * s * s, 1.9073486328125e-06f /* 2^-19 */, 1e-41f);
}

// max |coordinate| of a query; +inf when a coordinate is NaN (fmaxf drops NaNs).
__device__ __forceinline__ float query_abs(float x, float y, float z) {
  float a = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
  if (x != x || y != y || z != z) a = __int_as_float(0x7f800000);
  return a;
}

}  // namespace ga
