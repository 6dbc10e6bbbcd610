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

// Frame of the filter.  The window of both filters scales with s^2, s = max|q_c| + max|t_c| measured from the ORIGIN
// of the frame the filter is evaluated in; the reference's own rounding error does not (it subtracts first:
// |d - D| <= 5u D).  A cloud that sits far from the origin relative to its size therefore loses the filter's
// selectivity (unit cubes at offset 2: 3.7x slower, at offset 10: 10x, profiles/r02_tune_offset.txt).  The kernels that
// take a Frame evaluate the FILTER on q' = fl(q - c), t' = fl(t - c) for a centre c chosen per CTA from the bounding box
// of the first staged chunk; everything exact (reference arithmetic, d(0) seed, outputs) keeps the original
// coordinates, read from global memory.  Any c is valid: with D' = |q' - t'|^2, |D' - D| <= 6u s'^2 (one rounding of
// relative size u per shifted coordinate), which adds 12u s'^2 to the chains of nn_distance_fwd.cu (66.2u -> 78.2u
// against W = 128u) and nn_mma.cuh (882u -> 894u against 1024u; 1458u -> 1470u against 2048u), s' measured in the
// shifted frame.  c = 0 reproduces the unshifted kernel bit for bit (x - 0 = x) and is kept unless no point of the
// chunk comes near the origin (min |t|_inf > max |t|_inf / 4: a shell around the origin passes this, a cloud that
// contains it does not) AND the box centre is more than an eighth of the box half-extent away from the origin (and
// everything is finite).
struct Frame {
  float cx, cy, cz;
  bool shifted;  // c != 0: exact evaluations must fetch the original target from global memory
};

// float <-> int whose signed order is the float order (no NaNs): for redux.sync min / max
__device__ __forceinline__ int float_order(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// The frame is chosen in three steps, the first of which is all that centred data pays for:
//  1. stage_targets_min: stage_targets with ONE more value riding along in the reduction, the smallest |t|_inf of
//     the chunk (one FMNMX per point, one warp reduction).  A cloud whose points all keep away from the origin
//     (frame_candidate: min |t|_inf > max |t|_inf / 4) is a candidate;
//  2. frame_from_box (candidates only; uniform branch): bounding box of the staged chunk, centre c = box centre if it
//     really is off the origin;
//  3. shift_staged / stage_targets_shifted (shifted frames only): the staged points become fl(t - c).
// nn_fwd_mma_kernel sits at its 128-register cap: any trace of steps 2-3 in it costs the common path 2-3 us of 56 at
// B=50, whether inlined, behind a branch or in a function of their own (profiles/r02_ab_fwd.txt).  It therefore
// exists twice: a plain kernel that only DETECTS such a cloud (steps 1-2, then a word in mapped host memory), and
// a frame kernel with steps 1-3, which its launcher uses once the plain one has reported in.

// stage_targets + the chunk's smallest |t|_inf (NaN points dropped; +inf for an empty chunk) in mn when want_min
// (else mn = 0).  red: [0, 8) max per warp, [8, 16) min per warp.
template <int THREADS, int T>
__device__ __forceinline__ float stage_targets_min(float4* __restrict__ tgt, float* __restrict__ red,
                                                   const float* __restrict__ tpts, int c0, int nt, int ntile, int tid,
                                                   bool want_min, float& mn) {
  static_assert(THREADS / 32 <= 8, "red holds 8 warp values per kind");
  const float kInf = __int_as_float(0x7f800000);
  __syncthreads();  // previous contents fully consumed
  float lmax = 0.0f, lmin = kInf;
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
        const float a = fmaxf(fmaxf(fabsf(c[s][0]), fabsf(c[s][1])), fabsf(c[s][2]));
        lmax = fmaxf(lmax, a);
        lmin = fminf(lmin, a);
      }
      if (g + 1 < nt) {
        n1 = fmaf(c[s][5], c[s][5], fmaf(c[s][4], c[s][4], c[s][3] * c[s][3]));
        const float a = fmaxf(fmaxf(fabsf(c[s][3]), fabsf(c[s][4])), fabsf(c[s][5]));
        lmax = fmaxf(lmax, a);
        lmin = fminf(lmin, a);
      }
      tgt[2 * p] = make_float4(c[s][0], c[s][3], c[s][1], c[s][4]);
      tgt[2 * p + 1] = make_float4(c[s][2], c[s][5], n0, n1);
    }
  }
  lmax = warp_max(lmax);
  // non-negative floats (and +inf) order as their bit patterns
  if (want_min) lmin = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(lmin)));
  if ((tid & 31) == 0) {
    red[tid >> 5] = lmax;
    if (want_min) red[8 + (tid >> 5)] = lmin;
  }
  __syncthreads();
  float bm = 0.0f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; w++) bm = fmaxf(bm, red[w]);
  mn = 0.0f;
  if (want_min) {
    mn = kInf;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) mn = fminf(mn, red[8 + w]);
  }
  return bm;
}

// Everything finite and no point near the origin; an empty chunk (mn = +inf, bm = 0) fails bm > 0.
__device__ __forceinline__ bool frame_candidate(float mn, float bm) {
  return mn > 0.25f * bm && bm < 1e30f && bm > 0.0f;
}

// Bounding box of the chunk staged by stage_targets(_min) (behind its closing barrier), and the frame chosen from
// it; false (and c = 0) if the box centre is within an eighth of the box half-extent of the origin.  Every thread
// of the CTA must call it.  red: [16, 64) boxes per warp.
template <int THREADS, int T>
__device__ __forceinline__ bool frame_from_box(const float4* __restrict__ tgt, float* __restrict__ red, int c0, int nt,
                                               int ntile, int tid, Frame& fr) {
  const float kInf = __int_as_float(0x7f800000);
  const int npair = ntile * (T / 2);
  float lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf};
  for (int p = tid; p < npair; p += THREADS) {
    const int g = c0 + 2 * p;
    const float4 u = tgt[2 * p], v = tgt[2 * p + 1];
    if (g < nt) {
      lo[0] = fminf(lo[0], u.x), hi[0] = fmaxf(hi[0], u.x);
      lo[1] = fminf(lo[1], u.z), hi[1] = fmaxf(hi[1], u.z);
      lo[2] = fminf(lo[2], v.x), hi[2] = fmaxf(hi[2], v.x);
    }
    if (g + 1 < nt) {
      lo[0] = fminf(lo[0], u.y), hi[0] = fmaxf(hi[0], u.y);
      lo[1] = fminf(lo[1], u.w), hi[1] = fmaxf(hi[1], u.w);
      lo[2] = fminf(lo[2], v.y), hi[2] = fmaxf(hi[2], v.y);
    }
  }
#pragma unroll
  for (int e = 0; e < 3; e++) {
    lo[e] = order_float(__reduce_min_sync(0xffffffffu, float_order(lo[e])));
    hi[e] = order_float(__reduce_max_sync(0xffffffffu, float_order(hi[e])));
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int e = 0; e < 3; e++) {
      red[16 + (tid >> 5) * 6 + e] = lo[e];
      red[16 + (tid >> 5) * 6 + 3 + e] = hi[e];
    }
  }
  __syncthreads();
  float mid[3], ext = 0.0f, off = 0.0f;
#pragma unroll
  for (int e = 0; e < 3; e++) {
    float l = kInf, h = -kInf;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
      l = fminf(l, red[16 + w * 6 + e]);
      h = fmaxf(h, red[16 + w * 6 + 3 + e]);
    }
    mid[e] = 0.5f * l + 0.5f * h;
    ext = fmaxf(ext, 0.5f * h - 0.5f * l);
    off = fmaxf(off, fabsf(mid[e]));
  }
  // finite box, centre well away from the origin; an empty or infinite box fails the comparisons: c = 0
  const bool shift = off > 0.125f * ext && off < 1e30f && ext < 1e30f;
  fr.cx = shift ? mid[0] : 0.0f;
  fr.cy = shift ? mid[1] : 0.0f;
  fr.cz = shift ? mid[2] : 0.0f;
  fr.shifted = shift;
  return shift;
}

// Move the chunk staged by stage_targets(_min) into the frame fr, in place.  Every thread rewrites the pairs it staged
// itself (no barrier before this pass); returns the CTA-wide max |t'_c|.  red: [64, 72) max per warp.
template <int THREADS, int T>
__device__ __forceinline__ float shift_staged(float4* __restrict__ tgt, float* __restrict__ red, int c0, int nt,
                                              int ntile, int tid, const Frame& fr) {
  const int npair = ntile * (T / 2);
  float lmax = 0.0f;
  for (int p = tid; p < npair; p += THREADS) {
    const int g = c0 + 2 * p;
    float4 u = tgt[2 * p], v = tgt[2 * p + 1];
    if (g < nt) {
      u.x -= fr.cx, u.z -= fr.cy, v.x -= fr.cz;
      v.z = fmaf(v.x, v.x, fmaf(u.z, u.z, u.x * u.x));
      lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(u.x), fabsf(u.z)), fabsf(v.x)));
    }
    if (g + 1 < nt) {
      u.y -= fr.cx, u.w -= fr.cy, v.y -= fr.cz;
      v.w = fmaf(v.y, v.y, fmaf(u.w, u.w, u.y * u.y));
      lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(u.y), fabsf(u.w)), fabsf(v.y)));
    }
    tgt[2 * p] = u;
    tgt[2 * p + 1] = v;
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[64 + (tid >> 5)] = lmax;
  __syncthreads();
  float bm = 0.0f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; w++) bm = fmaxf(bm, red[64 + w]);
  return bm;
}

// stage_targets in the frame fr: stages fl(t - c) and |fl(t - c)|^2; returns the CTA-wide max |t'_c| of the chunk.
template <int THREADS, int T>
__device__ __forceinline__ float stage_targets_shifted(float4* __restrict__ tgt, float* __restrict__ red,
                                                       const float* __restrict__ tpts, int c0, int nt, int ntile,
                                                       int tid, const Frame& fr) {
  const float kInf = __int_as_float(0x7f800000);
  __syncthreads();  // previous contents fully consumed
  float lmax = 0.0f;
  const int npair = ntile * (T / 2);
  for (int p = tid; p < npair; p += THREADS) {
    const long long g = (long long)c0 + 2 * p;
    float c[6];
#pragma unroll
    for (int e = 0; e < 6; e++) c[e] = g + (e >= 3) < nt ? __ldg(tpts + g * 3 + e) : 0.0f;
    float n0 = kInf, n1 = kInf;
    if (g < nt) {
      c[0] -= fr.cx, c[1] -= fr.cy, c[2] -= fr.cz;
      n0 = fmaf(c[2], c[2], fmaf(c[1], c[1], c[0] * c[0]));
      lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(c[0]), fabsf(c[1])), fabsf(c[2])));
    }
    if (g + 1 < nt) {
      c[3] -= fr.cx, c[4] -= fr.cy, c[5] -= fr.cz;
      n1 = fmaf(c[5], c[5], fmaf(c[4], c[4], c[3] * c[3]));
      lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(c[3]), fabsf(c[4])), fabsf(c[5])));
    }
    tgt[2 * p] = make_float4(c[0], c[3], c[1], c[4]);
    tgt[2 * p + 1] = make_float4(c[2], c[5], n0, n1);
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
  // running append positions as 32-bit shared-memory addresses: a generic pointer costs a 64-bit add per append
  uint32_t qp[Q], qb[Q];
  bool over[Q];
  const uint32_t rowb = 2u * (uint32_t)qstride;  // bytes from one queue row to the next
#pragma unroll
  for (int j = 0; j < Q; j++) {
    qb[j] = (uint32_t)__cvta_generic_to_shared(queue + j * jstride);
    qp[j] = qb[j];
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
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(qp[j]), "h"((unsigned short)(gb + (blk * U + pp) * 2)) : "memory");
            qp[j] += rowb;
          }
          if (!(f.y > thr[j])) {
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(qp[j]), "h"((unsigned short)(gb + (blk * U + pp) * 2 + 1)) : "memory");
            qp[j] += rowb;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < Q; j++) {
      if (qp[j] - qb[j] > (uint32_t)cap * rowb) {  // too many: stop collecting
        over[j] = true;
        qp[j] = qb[j];
        thr[j] = -__int_as_float(0x7f800000);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < Q; j++) cnt[j] = over[j] ? cap + 1 : (int)((qp[j] - qb[j]) / rowb);
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
