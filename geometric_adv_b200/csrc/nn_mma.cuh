// Tensor-core filter for the Chamfer forward search (nn_distance_fwd_mma.cu).
//
// The filter f(q,t) = |t|^2 - 2 q.t of nn_tiles.cuh is a K=4 contraction.  Here it is evaluated
// by the warp-level MMA m16n8k16 (bf16 inputs, fp32 accumulate) on split operands:
//   v = v1 + v2 + v3 exactly, v1 = bf16(v), v2 = bf16(v - v1), v3 = bf16(v - v1 - v2)
//   h(q,t) = sum_c (Q1 t1 + Q1 t2 + Q2 t1 + Q2 t2)_c + (n1 + n2 + n3),   Q = -2q, n = fl(|t|^2)
// 15 of the 16 K slots are used; products of two bf16 are exact in fp32.  What is dropped is
// Q1 t3 + Q3 t1 (+ terms below 2^-26): per coordinate <= 2 * 2^-18 |Q_c t_c|.
//
// Error bound (u = 2^-24, A = max|q_c|, Bm = max|t_c|, s = A + Bm, g = D - |q|^2 the real value):
//   dropped terms   <= 4 * 2^-18 * sum_c |q_c t_c| <= 4 * 2^-18 * 3 A Bm <= 3 * 2^-18 s^2 = 192u s^2
//   accumulation    16 terms of magnitude <= 3 s^2, hardware adder of >= 24 bits: budget 128u s^2
//   norm rounding   <= 9u Bm^2
//   |h - g| <= e2 = 330u s^2; the scan overwrites 4 low bits of a tile minimum with the tile's
//   block number (MmaTrack): + 96u s^2.
// With |d - D| <= e0 = 15u s^2 (reference arithmetic) the reference argmin k* satisfies
//   h(k*) <= min_j h(j) + 2 e0 + 2 (e2 + 96u s^2) = min h + 882u s^2,
// and with |f - g| <= e1 = 18.1u s^2 (the fp32 filter of nn_tiles.cuh)
//   f(k*) <= min h + 2 e0 + e1 + e2 + 96u s^2 = min h + 474u s^2.
// One window W2 = 1024u s^2 = 2^-14 s^2 (s inflated by 1.0001, plus an absolute term for
// flushed denormals) is used for both tests.  tests/test_mma_filter_gpu.py measures |h - g| on
// the device (ga_debug_mma_filter) and checks it against e2.
// Non-finite or overflowing data give W2 = inf/NaN: every comparison "!(x > thr)" passes, the
// query is served by the exact warp scan, and the result is still the reference's.
#pragma once
#include "nn_search.cuh"

namespace ga {

constexpr int kMmaQW = 64;    // queries per warp: 4 m-tiles of 16 rows
constexpr int kMmaBlk = 128;  // targets per MMA block = 16 n-tiles; quad lane t owns targets [32t, 32t+32) of it
constexpr int kMmaT = 32;     // refine tile (the 32 contiguous targets a quad lane owns in a block)

// v rounded to bf16 (nearest even), as a float
__device__ __forceinline__ float bf16r(float v) {
  unsigned short h;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __uint_as_float((uint32_t)h << 16);
}
// {bf16(lo), bf16(hi)}: lo in bits 0-15 (the lower K index of an MMA operand register)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}

// Window of the tensor-core filter: 2^-14 s^2 (see the bound above).
__device__ __forceinline__ float mma_window(float qabs, float bm) {
  const float s = (qabs + bm) * 1.0001f;
  return fmaf(s * s, 6.103515625e-05f /* 2^-14 */, 1e-35f);
}

// K-slot assignment.  Quad lane t (= lane & 3) holds K slots {2t, 2t+1, 2t+8, 2t+9} of both operands:
//   t = 0,1,2 (coordinate c = x,y,z):  A = (C1, C1, C2, 1)   B = (c1, c2, c1, n_{t+1})
//   t = 3:                             A = (X2, Y2, Z2, 0)   B = (x2, y2, z2, 0)
// where C = -2 q_c.  One staged target is 32 bytes: the 8 bytes of lanes t = 0..3 in order.

// Target chunk-local index of column `c` (0..7) of n-tile `j` (0..15) of block `blk`.
__device__ __forceinline__ int mma_col_target(int blk, int j, int c) {
  return blk * kMmaBlk + 32 * (c >> 1) + 2 * j + (c & 1);
}

// B fragment (32 bytes) of staged target `tau` of a chunk with `cn` valid targets, from the
// pair-SoA staging of nn_tiles.cuh; targets >= cn are padding (h = 1e38: far away but finite, so
// that a key never becomes a NaN bit pattern).
__device__ __forceinline__ void make_bfrag(uint4* __restrict__ dst, const float4* __restrict__ tgt, int tau, int cn) {
  const bool ok = tau < cn;
  const float* pu = reinterpret_cast<const float*>(tgt + 2 * (tau >> 1)) + (tau & 1);
  const float x = ok ? pu[0] : 0.0f, y = ok ? pu[2] : 0.0f, z = ok ? pu[4] : 0.0f;
  const float n = ok ? pu[6] : 1.0e38f;
  const float xr = x - bf16r(x), yr = y - bf16r(y), zr = z - bf16r(z);
  const float nr = n - bf16r(n);
  const float nr2 = nr - bf16r(nr);
  dst[0] = make_uint4(pack_bf16(x, xr), pack_bf16(x, n), pack_bf16(y, yr), pack_bf16(y, nr));
  dst[1] = make_uint4(pack_bf16(z, zr), pack_bf16(z, nr2), pack_bf16(xr, yr), pack_bf16(zr, 0.0f));
}

// Build the B fragments of `nblk` blocks (whole CTA).  Call between two __syncthreads().
template <int THREADS>
__device__ __forceinline__ void stage_bfrag(uint4* __restrict__ bfrag, const float4* __restrict__ tgt, int nblk,
                                            int cn, int tid) {
  const int total = nblk * kMmaBlk;
  for (int p = tid; p < total; p += THREADS) {
    const int c = p & 7, j = (p >> 3) & 15, blk = p >> 7;
    make_bfrag(bfrag + 2 * p, tgt, mma_col_target(blk, j, c), cn);
  }
}

// One WARP stages MMA block `blk` of a cloud (targets [128 blk, +128) of `tpts`, nt points):
// pair-SoA first, then the B fragments built from it.  No CTA barrier: the block is self-contained.
// Returns this lane's partial max |coordinate| (NaNs dropped).
__device__ __forceinline__ float stage_block_warp(float4* __restrict__ tgt, uint4* __restrict__ bfrag,
                                                  const float* __restrict__ tpts, int nt, int blk, int lane) {
  const float kInf = __int_as_float(0x7f800000);
  float lmax = 0.0f;
  float c[2][6];
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int g = 2 * (64 * blk + 32 * s + lane);
#pragma unroll
    for (int e = 0; e < 6; e++) c[s][e] = g + (e >= 3) < nt ? __ldg(tpts + (size_t)g * 3 + e) : 0.0f;
  }
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int p = 64 * blk + 32 * s + lane;
    const int g = 2 * p;
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
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int q = lane + 32 * i;  // item of this block: column c = q & 7 of n-tile j = q >> 3
    make_bfrag(bfrag + 2 * (blk * kMmaBlk + q), tgt, mma_col_target(blk, q >> 3, q & 7), nt);
  }
  return lmax;
}

// Per-lane view of the warp's 64 queries: 8 rows (m-tile i = r>>1, half h = r&1: local query
// 16 i + (lane>>2) + 8 h), their A fragments and max |coordinate|.
struct MmaRows {
  uint32_t a[4][4];
  float qabs[8];
};

// Quad lane t < 3 loads and splits coordinate t of each of its rows (A slots C1, C1, C2, 1); lane 3
// needs the three residuals (X2, Y2, Z2, 0), which it gets from lanes 0..2 by shuffle.  The max
// |coordinate| of a row is a quad reduction (+inf when a coordinate is NaN, as query_abs).
// The raw coordinate (t < 3 ? t : 2) of each of the lane's 8 rows: issued early, used by mma_build_rows.
__device__ __forceinline__ void mma_fetch_rows(float (&qraw)[8], const float* __restrict__ qpts, int nq, int qbase,
                                               int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int cc = t < 3 ? t : 2;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int qi = qbase + 16 * (r >> 1) + g + 8 * (r & 1);
    const int qs = qi < nq ? qi : 0;
    qraw[r] = __ldg(qpts + (size_t)qs * 3 + cc);
  }
}
// shift: this lane's component (coordinate t < 3 ? t : 2) of the frame centre (Frame, nn_tiles.cuh); 0 = no shift.
__device__ __forceinline__ void mma_build_rows(MmaRows& R, const float (&qraw)[8], int lane, float shift) {
  const int t = lane & 3;
  const unsigned qbaseLane = lane & ~3;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i = r >> 1, h = r & 1;
    const float q = qraw[r] - shift;
    const float C = -2.0f * q;
    const float Cr = C - bf16r(C);
    const float Xr = __shfl_sync(0xffffffffu, Cr, qbaseLane), Yr = __shfl_sync(0xffffffffu, Cr, qbaseLane + 1),
                Zr = __shfl_sync(0xffffffffu, Cr, qbaseLane + 2);
    const uint32_t lo_c = pack_bf16(C, C), hi_c = pack_bf16(Cr, 1.0f);
    const uint32_t lo_3 = pack_bf16(Xr, Yr), hi_3 = pack_bf16(Zr, 0.0f);
    R.a[i][h] = t < 3 ? lo_c : lo_3;
    R.a[i][2 + h] = t < 3 ? hi_c : hi_3;
    float a = q != q ? __int_as_float(0x7f800000) : fabsf(q);
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 2));
    R.qabs[r] = a;
  }
}
__device__ __forceinline__ void mma_load_rows(MmaRows& R, const float* __restrict__ qpts, int nq, int qbase,
                                              int lane, float shift = 0.0f) {
  float qraw[8];
  mma_fetch_rows(qraw, qpts, nq, qbase, lane);
  mma_build_rows(R, qraw, lane, shift);
}

// Sentinel for "no value": finite, so that a key never turns into a NaN bit pattern.
constexpr float kMmaBig = 3.0e38f;

// The three smallest tile keys per row (this lane's tiles only).  A key is the tile minimum of h
// with its low 4 mantissa bits replaced by the block number (the tile is 4 blk + t): the three
// smallest tiles AND their identities come out of five FMNMX per tile and row.  Overwriting 4
// bits moves a value by < 16 ulp <= 96u s^2, which the window W2 absorbs (nn_mma.cuh header:
// 2 e0 + 2 (e2 + 96u) = 882u < 1024u).
struct MmaTrack {
  float c1[8], c2[8], c3[8];
};
__device__ __forceinline__ int mma_key_tile(float key, int t) { return ((__float_as_int(key) & 15) << 2) + t; }

// Tensor-core filter scan over the staged blocks [blk0, blk1) (blk1 <= 16).  Lane (g,t) sees, for each
// of its 8 rows, the minimum of h over the 32 contiguous targets [128 blk + 32 t, +32) of every block,
// i.e. over refine tile 4 blk + t, without any cross-lane traffic.  KEYBITS = 4: the key carries the
// block number (the tile is 4 blk + t, t known from the lane); KEYBITS = 6: it carries the whole tile
// number 4 blk + t, so that keys of different lanes can be merged (quad_merge; 64 ulp <= 384u s^2 of
// perturbation: callers use mma_window_wide).
// BATCH: the four MMAs of a B fragment are issued into four accumulator quads of their own before any of them is
// folded.  In a kernel whose job loop keeps many values live (all_pairs.cu) ptxas otherwise puts every HMMA of the
// scan on ONE quad with a NOP behind it (1.04 G NOPs for 1.07 G HMMAs, 32 % of the samples: ncu of
// all_pairs_directed_mma_kernel, round 2) -- whether the scan is inlined or sits behind a call.
template <int KEYBITS, bool BATCH = false>
__device__ __forceinline__ void mma_scan_range(const MmaRows& R, const uint2* __restrict__ bfrag, int blk0, int blk1,
                                               int lane, MmaTrack& tr) {
  static_assert(KEYBITS == 4 || KEYBITS == 6, "key layouts");
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
      if constexpr (BATCH) {
        float c[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++) mma16816(c[i], R.a[i], bf.x, bf.y);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          rm[2 * i] = fmin3(rm[2 * i], c[i][0], c[i][1]);
          rm[2 * i + 1] = fmin3(rm[2 * i + 1], c[i][2], c[i][3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          float c[4];
          mma16816(c, R.a[i], bf.x, bf.y);
          rm[2 * i] = fmin3(rm[2 * i], c[0], c[1]);
          rm[2 * i + 1] = fmin3(rm[2 * i + 1], c[2], c[3]);
        }
      }
    }
    const int id = KEYBITS == 4 ? blk : 4 * blk + (lane & 3);
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const float key = __int_as_float((__float_as_int(rm[r]) & ~((1 << KEYBITS) - 1)) | id);
      tr.c3[r] = fminf(tr.c3[r], fmaxf(tr.c2[r], key));
      tr.c2[r] = fminf(tr.c2[r], fmaxf(tr.c1[r], key));
      tr.c1[r] = fminf(tr.c1[r], key);
    }
  }
}
__device__ __forceinline__ void mma_scan(const MmaRows& R, const uint2* __restrict__ bfrag, int nblk, int lane,
                                         MmaTrack& tr) {
  mma_scan_range<4>(R, bfrag, 0, nblk, lane, tr);
}

// Window for 6-bit keys: 2 e0 + 2 (e2 + 384u) = 1458u < 2048u = 2^-13 s^2 (see the header).
__device__ __forceinline__ float mma_window_wide(float qabs, float bm) {
  const float s = (qabs + bm) * 1.0001f;
  return fmaf(s * s, 1.220703125e-04f /* 2^-13 */, 1e-35f);
}

// (a1 <= a2 <= a3) := the three smallest of two sorted triples.
__device__ __forceinline__ void merge3(float& a1, float& a2, float& a3, float b1, float b2, float b3) {
  const float lo = fmaxf(a1, b1), hi = fminf(a2, b2);
  a3 = fminf(fmaxf(lo, hi), fminf(a3, b3));
  a2 = fminf(lo, hi);
  a1 = fminf(a1, b1);
}

// After this every lane of a quad holds, per row, the three smallest 6-bit keys over the quad's four
// tile columns.
__device__ __forceinline__ void quad_merge(MmaTrack& tr) {
#pragma unroll
  for (int m = 1; m <= 2; m <<= 1) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const float b1 = __shfl_xor_sync(0xffffffffu, tr.c1[r], m), b2 = __shfl_xor_sync(0xffffffffu, tr.c2[r], m),
                  b3 = __shfl_xor_sync(0xffffffffu, tr.c3[r], m);
      merge3(tr.c1[r], tr.c2[r], tr.c3[r], b1, b2, b3);
    }
  }
}

// Candidate bit mask of one 32-target tile for one query: bit k set <=> the fp32 filter of target
// g0 + k is <= thr (NaN counts as a candidate).  Lanes of a warp usually walk DIFFERENT tiles,
// which are a multiple of 512 B apart: each lane starts at its own pair (conflict-free), collects
// the bits in walk order with immediate masks, and rotates the word back once at the end.
__device__ __forceinline__ unsigned tile_candidate_mask(const float4* __restrict__ tp, float ax2, float ay2,
                                                        float az2, float thr) {
  constexpr int T = kMmaT;
  const int rot = threadIdx.x & (T / 2 - 1);
  unsigned m = 0;
#pragma unroll
  for (int i = 0; i < T / 2; i++) {
    const int pp = (i + rot) & (T / 2 - 1);
    const float2 f = filter_pair(tp[2 * pp], tp[2 * pp + 1], ax2, ay2, az2);
    m |= (!(f.x > thr) ? (1u << (2 * i)) : 0u) | (!(f.y > thr) ? (2u << (2 * i)) : 0u);
  }
  return __funnelshift_l(m, m, 2 * rot);  // walk position i holds pair (i + rot) mod 16
}

// Two tiles for two queries in one walk: twice the independent loads and FMAs in flight (a lane's two
// queries are otherwise served one after the other, each a chain of LDS -> FFMA2 -> compare).
__device__ __forceinline__ void tile_candidate_mask2(const float4* __restrict__ tpa, const float4* __restrict__ tpb,
                                                     const float (&ax2)[2], const float (&ay2)[2],
                                                     const float (&az2)[2], const float (&thr)[2], unsigned& ma,
                                                     unsigned& mb) {
  constexpr int T = kMmaT;
  const int rot = threadIdx.x & (T / 2 - 1);
  ma = 0;
  mb = 0;
#pragma unroll
  for (int i = 0; i < T / 2; i++) {
    const int pp = (i + rot) & (T / 2 - 1);
    const float2 fa = filter_pair(tpa[2 * pp], tpa[2 * pp + 1], ax2[0], ay2[0], az2[0]);
    const float2 fb = filter_pair(tpb[2 * pp], tpb[2 * pp + 1], ax2[1], ay2[1], az2[1]);
    ma |= (!(fa.x > thr[0]) ? (1u << (2 * i)) : 0u) | (!(fa.y > thr[0]) ? (2u << (2 * i)) : 0u);
    mb |= (!(fb.x > thr[1]) ? (1u << (2 * i)) : 0u) | (!(fb.y > thr[1]) ? (2u << (2 * i)) : 0u);
  }
  ma = __funnelshift_l(ma, ma, 2 * rot);
  mb = __funnelshift_l(mb, mb, 2 * rot);
}

// Refine for the tensor-core scan.  Per query: cnt qualifying tiles (ta, tb valid for cnt <= 2).
//  1. first tile, one lane per query: fp32 filter -> candidate mask -> one or two exact evaluations
//     (all lanes in step); more than two candidates -> exact warp scan of the chunk;
//  2. second tiles are rare (a few per warp): the whole warp evaluates such a tile exactly, one
//     target per lane, and merges by (value, index);
//  3. cnt > 2 / many candidates / non-finite window: warp_exact_scan (nn_search.cuh).
// torig != nullptr: staged copy and s.ax2.. are in a shifted frame; exact evaluations read the original target.
template <int MODE, int Q = 2>
__device__ __forceinline__ void refine_tiles(QueryState<Q>& s, const float4* __restrict__ tgt, int c0, int nt,
                                             int ntile, const int (&cnt)[Q], const int (&ta)[Q], const int (&tb)[Q],
                                             const float (&thr)[Q], const float* __restrict__ torig = nullptr) {
  constexpr int T = kMmaT;
  const int lane = threadIdx.x & 31;
  bool hard[Q], second[Q];
  unsigned masks[Q];
  int g0s[Q];
  bool use[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    hard[j] = s.valid[j] && cnt[j] > 2;
    second[j] = s.valid[j] && cnt[j] == 2;
    use[j] = s.valid[j] && cnt[j] >= 1 && cnt[j] <= 2;
    g0s[j] = c0 + (use[j] ? ta[j] : 0) * T;  // an unused slot walks tile 0 and drops the mask
    masks[j] = 0;
  }
  if constexpr (Q == 2) {
    tile_candidate_mask2(tgt + (size_t)(g0s[0] - c0), tgt + (size_t)(g0s[1] - c0), s.ax2, s.ay2, s.az2, thr, masks[0],
                         masks[1]);
  } else {
#pragma unroll
    for (int j = 0; j < Q; j++)
      masks[j] = tile_candidate_mask(tgt + (size_t)(g0s[j] - c0), s.ax2[j], s.ay2[j], s.az2[j], thr[j]);
  }
#pragma unroll
  for (int j = 0; j < Q; j++) {
    const int g0 = g0s[j];
    const int rem = nt - g0;  // >= 1: only staged tiles are listed (and tile 0 exists)
    unsigned mask = use[j] ? masks[j] : 0u;
    mask &= rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    const int n = __popc(mask);
    if (n >= 1 && n <= 2)
      eval_candidate<MODE>(tgt, c0, g0 + __ffs(mask) - 1, s.qx[j], s.qy[j], s.qz[j], s.best[j], s.besti[j], torig);
    if (n == 2)
      eval_candidate<MODE>(tgt, c0, g0 + 31 - __clz(mask), s.qx[j], s.qy[j], s.qz[j], s.best[j], s.besti[j], torig);
    if (n > 2) {
      hard[j] = true;
      second[j] = false;
    }
  }
#pragma unroll
  for (int j = 0; j < Q; j++) {
    unsigned pending = __ballot_sync(0xffffffffu, second[j]);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const float bqx = __shfl_sync(0xffffffffu, s.qx[j], src), bqy = __shfl_sync(0xffffffffu, s.qy[j], src),
                  bqz = __shfl_sync(0xffffffffu, s.qz[j], src);
      const int tile = __shfl_sync(0xffffffffu, tb[j], src);
      const int gl = tile * T + lane;  // chunk-local target of this lane
      const float* pu = reinterpret_cast<const float*>(tgt + 2 * (gl >> 1)) + (gl & 1);
      float b = __int_as_float(0x7f800000);
      int bi = 0x7fffffff;
      if (c0 + gl < nt) {
        float tx = pu[0], ty = pu[2], tz = pu[4];
        if (torig != nullptr) {
          tx = __ldg(torig + (size_t)(c0 + gl) * 3);
          ty = __ldg(torig + (size_t)(c0 + gl) * 3 + 1);
          tz = __ldg(torig + (size_t)(c0 + gl) * 3 + 2);
        }
        const float d = sqdist<MODE>(tx, ty, tz, bqx, bqy, bqz);
        if (d < b) {  // NaN is never selected
          b = d;
          bi = c0 + gl;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, b, o);
        const int obi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < b || (ob == b && obi < bi)) {
          b = ob;
          bi = obi;
        }
      }
      if (lane == src && (b < s.best[j] || (b == s.best[j] && bi < s.besti[j]))) {
        s.best[j] = b;
        s.besti[j] = bi;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < Q; j++) {
    unsigned pending = __ballot_sync(0xffffffffu, hard[j]);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const float bqx = __shfl_sync(0xffffffffu, s.qx[j], src), bqy = __shfl_sync(0xffffffffu, s.qy[j], src),
                  bqz = __shfl_sync(0xffffffffu, s.qz[j], src);
      const float bax = __shfl_sync(0xffffffffu, s.ax2[j], src), bay = __shfl_sync(0xffffffffu, s.ay2[j], src),
                  baz = __shfl_sync(0xffffffffu, s.az2[j], src);
      const float bthr = __shfl_sync(0xffffffffu, thr[j], src);
      float b;
      int bi;
      warp_exact_scan<MODE>(tgt, c0, nt, ntile * (T / 2), bqx, bqy, bqz, bax, bay, baz, bthr, b, bi, lane, torig);
      if (lane == src && (b < s.best[j] || (b == s.best[j] && bi < s.besti[j]))) {
        s.best[j] = b;
        s.besti[j] = bi;
      }
    }
  }
}

// ---- per-warp job pieces shared by nn_distance_fwd_mma.cu and all_pairs.cu ------------------------

// The two queries a lane refines and writes: m-tile t = lane&3, rows g = lane>>2 and g+8 of the
// warp's 64 queries, i.e. local queries 16 t + g + 8 j.
// (cx, cy, cz): centre of the filter's frame; qx.. and d0 stay in original coordinates, the filter side is shifted.
template <int MODE>
__device__ __forceinline__ void mma_init_queries(QueryState<2>& s, const float* __restrict__ qpts, int nq, int qbase,
                                                 const float* __restrict__ tpts, int lane, float cx = 0.0f,
                                                 float cy = 0.0f, float cz = 0.0f) {
  const float kInf = __int_as_float(0x7f800000);
  const int g = lane >> 2, t = lane & 3;
  const float t0x = __ldg(tpts), t0y = __ldg(tpts + 1), t0z = __ldg(tpts + 2);
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int qi = qbase + 16 * t + g + 8 * j;
    s.valid[j] = qi < nq;
    const int qs = s.valid[j] ? qi : 0;
    s.qx[j] = __ldg(qpts + (size_t)qs * 3);
    s.qy[j] = __ldg(qpts + (size_t)qs * 3 + 1);
    s.qz[j] = __ldg(qpts + (size_t)qs * 3 + 2);
    s.qabs[j] = query_abs(s.qx[j] - cx, s.qy[j] - cy, s.qz[j] - cz);
    s.ax2[j] = -2.0f * (s.qx[j] - cx);
    s.ay2[j] = -2.0f * (s.qy[j] - cy);
    s.az2[j] = -2.0f * (s.qz[j] - cz);
    s.d0[j] = sqdist<MODE>(t0x, t0y, t0z, s.qx[j], s.qy[j], s.qz[j]);
    s.best[j] = kInf;
    s.besti[j] = 0;
    s.m1g[j] = kInf;
  }
}
// One staged chunk (targets [c0, c0 + cn) of a cloud with nt points) for one warp: tensor-core
// scan, per-query lists of the qualifying tiles, exact refine.  mrun: running row minima of h over
// the chunks seen so far (same value in the 4 lanes of a quad); wcnt / wtile: this warp's lists.
template <int MODE, bool SCAN_BATCH = false>
__device__ __forceinline__ void mma_chunk(const MmaRows& R, QueryState<2>& s, float (&mrun)[8],
                                          const float4* __restrict__ tgt, const uint4* __restrict__ bfrag, int c0, int nt,
                                          int cn, float bm_run, int* __restrict__ wcnt,
                                          unsigned short* __restrict__ wtile, int lane,
                                          const float* __restrict__ torig = nullptr) {
  constexpr int T = kMmaT;
  const int g = lane >> 2, t = lane & 3;
  const int ntile = (cn + T - 1) / T;
  const int nblk = (cn + kMmaBlk - 1) / kMmaBlk;
  wcnt[16 * t + g] = 0;
  wcnt[16 * t + g + 8] = 0;
  __syncwarp();

  MmaTrack tr;
  if constexpr (SCAN_BATCH)
    mma_scan_range<4, true>(R, reinterpret_cast<const uint2*>(bfrag), 0, nblk, lane, tr);
  else
    mma_scan(R, reinterpret_cast<const uint2*>(bfrag), nblk, lane, tr);

  // Row minimum over the quad, window, and the qualifying tiles of this lane -> per-query lists.
  float mythr[2] = {0.0f, 0.0f};
#pragma unroll
  for (int r = 0; r < 8; r++) {
    float m = tr.c1[r];
    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    mrun[r] = fminf(mrun[r], m);
    const float thr = mrun[r] + mma_window(R.qabs[r], bm_run);
    if (t == (r >> 1)) mythr[r & 1] = thr;
    const int ql = 16 * (r >> 1) + g + 8 * (r & 1);
    if (!(tr.c1[r] > thr)) {
      const int slot = atomicAdd(&wcnt[ql], 1);
      if (slot < 2) wtile[2 * ql + slot] = (unsigned short)mma_key_tile(tr.c1[r], t);
    }
    if (!(tr.c2[r] > thr)) {
      const int slot = atomicAdd(&wcnt[ql], 1);
      if (slot < 2) wtile[2 * ql + slot] = (unsigned short)mma_key_tile(tr.c2[r], t);
    }
    if (!(tr.c3[r] > thr)) atomicAdd(&wcnt[ql], 3);  // a third tile of this lane: exact scan
  }
  __syncwarp();

  int cnt[2], ta[2], tb[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int ql = 16 * t + g + 8 * j;
    cnt[j] = wcnt[ql];
    ta[j] = wtile[2 * ql];
    tb[j] = wtile[2 * ql + 1];
    // a tile id beyond the staged tiles can only come from padding under a non-finite window
    if ((cnt[j] >= 1 && ta[j] >= ntile) || (cnt[j] >= 2 && tb[j] >= ntile)) cnt[j] = 3;
  }
  refine_tiles<MODE>(s, tgt, c0, nt, ntile, cnt, ta, tb, mythr, torig);
  __syncwarp();  // lists are reused by the next chunk / job
}


// Results of a warp's 64 queries.  A lane holds local queries 16 t + g + 8 j; they are transposed by shuffle so that
// lane L writes local queries L and 32 + L: four full 128-byte lines per warp and array instead of sixteen 32-byte
// pieces (what matters when mdist / midx are pinned host buffers: the stores are posted writes over PCIe).
// Every lane of the warp must call it.  nq: queries of the cloud (rows beyond it are not written).
__device__ __forceinline__ void mma_write(const QueryState<2>& s, int qbase, int nq, int lane, float* __restrict__ odist,
                                          int* __restrict__ oidx, float* __restrict__ mdist, int* __restrict__ midx,
                                          size_t moff) {
  float d[2];
  int i[2];
#pragma unroll
  for (int j = 0; j < 2; j++) finish_query<2>(s, j, d[j], i[j]);
  const int jj = (lane >> 3) & 1;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int src = (lane & 7) * 4 + (lane >> 4) + 2 * h;  // holder of local query 32 h + lane, in its slot jj
    const float d0 = __shfl_sync(0xffffffffu, d[0], src), d1 = __shfl_sync(0xffffffffu, d[1], src);
    const int i0 = __shfl_sync(0xffffffffu, i[0], src), i1 = __shfl_sync(0xffffffffu, i[1], src);
    const int qi = qbase + 32 * h + lane;
    if (qi < nq) {
      const float dv = jj ? d1 : d0;
      const int iv = jj ? i1 : i0;
      odist[qi] = dv;
      oidx[qi] = iv;
      if (mdist != nullptr) {
        mdist[moff + qi] = dv;
        midx[moff + qi] = iv;
      }
    }
  }
}

}  // namespace ga
