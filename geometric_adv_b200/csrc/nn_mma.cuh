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
//   |h - g| <= e2 = 330u s^2.
// With |d - D| <= e0 = 15u s^2 (reference arithmetic) the reference argmin k* satisfies
//   h(k*) <= min_j h(j) + 2 e0 + 2 e2 = min h + 690u s^2,
// and with |f - g| <= e1 = 18.1u s^2 (the fp32 filter of nn_tiles.cuh)
//   f(k*) <= min h + 2 e0 + e1 + e2 = min h + 378u s^2.
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

// Build the B fragments of `nblk` blocks from the pair-SoA staging of nn_tiles.cuh
// (targets >= cn are padding: h = +inf).  Call between two __syncthreads().
template <int THREADS>
__device__ __forceinline__ void stage_bfrag(uint4* __restrict__ bfrag, const float4* __restrict__ tgt, int nblk,
                                            int cn, int tid) {
  const float kInf = __int_as_float(0x7f800000);
  const int total = nblk * kMmaBlk;
  for (int p = tid; p < total; p += THREADS) {
    const int c = p & 7, j = (p >> 3) & 15, blk = p >> 7;
    const int tau = mma_col_target(blk, j, c);
    const bool ok = tau < cn;
    const float* pu = reinterpret_cast<const float*>(tgt + 2 * (tau >> 1)) + (tau & 1);
    const float x = ok ? pu[0] : 0.0f, y = ok ? pu[2] : 0.0f, z = ok ? pu[4] : 0.0f;
    const float n = ok ? pu[6] : kInf;
    const float xr = x - bf16r(x), yr = y - bf16r(y), zr = z - bf16r(z);
    float nr = n - bf16r(n);
    float nr2 = nr - bf16r(nr);
    if (!ok) nr = nr2 = 0.0f;  // keep the padding a clean +inf
    bfrag[2 * p] = make_uint4(pack_bf16(x, xr), pack_bf16(x, n), pack_bf16(y, yr), pack_bf16(y, nr));
    bfrag[2 * p + 1] = make_uint4(pack_bf16(z, zr), pack_bf16(z, nr2), pack_bf16(xr, yr), pack_bf16(zr, 0.0f));
  }
}

// A-fragment registers of one query row for quad lane t.
__device__ __forceinline__ void mma_afrag_row(float qx, float qy, float qz, int t, uint32_t& lo, uint32_t& hi) {
  const float X = -2.0f * qx, Y = -2.0f * qy, Z = -2.0f * qz;
  const float Xr = X - bf16r(X), Yr = Y - bf16r(Y), Zr = Z - bf16r(Z);
  const float C = t == 0 ? X : (t == 1 ? Y : Z);
  const float Cr = t == 0 ? Xr : (t == 1 ? Yr : Zr);
  const uint32_t lo_c = pack_bf16(C, C), hi_c = pack_bf16(Cr, 1.0f);
  const uint32_t lo_3 = pack_bf16(Xr, Yr), hi_3 = pack_bf16(Zr, 0.0f);
  lo = t < 3 ? lo_c : lo_3;
  hi = t < 3 ? hi_c : hi_3;
}

// Per-lane view of the warp's 64 queries: 8 rows (m-tile i = r>>1, half h = r&1: local query
// 16 i + (lane>>2) + 8 h), their A fragments and max |coordinate|.
struct MmaRows {
  uint32_t a[4][4];
  float qabs[8];
};

__device__ __forceinline__ void mma_load_rows(MmaRows& R, const float* __restrict__ qpts, int nq, int qbase,
                                              int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i = r >> 1, h = r & 1;
    const int qi = qbase + 16 * i + g + 8 * h;
    const int qs = qi < nq ? qi : 0;
    const float x = __ldg(qpts + (size_t)qs * 3), y = __ldg(qpts + (size_t)qs * 3 + 1),
                z = __ldg(qpts + (size_t)qs * 3 + 2);
    R.qabs[r] = query_abs(x, y, z);
    mma_afrag_row(x, y, z, t, R.a[i][h], R.a[i][2 + h]);
  }
}

// The three smallest tile minima per row and the tiles of the two smallest (this lane's tiles only).
struct MmaTrack {
  float c1[8], c2[8], c3[8];
  int i1[8], i2[8];
};

// Tensor-core filter scan over `nblk` staged blocks.  Lane (g,t) sees, for each of its 8 rows,
// the minimum of h over the 32 contiguous targets [128 blk + 32 t, +32) of every block, i.e.
// over refine tile 4 blk + t, without any cross-lane traffic.
__device__ __forceinline__ void mma_scan(const MmaRows& R, const uint2* __restrict__ bfrag, int nblk, int lane,
                                         MmaTrack& tr) {
  const float kInf = __int_as_float(0x7f800000);
  const int t = lane & 3;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    tr.c1[r] = tr.c2[r] = tr.c3[r] = kInf;
    tr.i1[r] = tr.i2[r] = 0;
  }
#pragma unroll 1
  for (int blk = 0; blk < nblk; blk++) {
    float rm[8];
#pragma unroll
    for (int r = 0; r < 8; r++) rm[r] = kInf;
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
    const int tile = blk * 4 + t;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const float tm = rm[r];
      const bool lt1 = tm < tr.c1[r], lt2 = tm < tr.c2[r];
      tr.c3[r] = fminf(tr.c3[r], fmaxf(tr.c2[r], tm));
      tr.i2[r] = lt1 ? tr.i1[r] : (lt2 ? tile : tr.i2[r]);
      tr.c2[r] = fminf(tr.c2[r], fmaxf(tr.c1[r], tm));
      tr.i1[r] = lt1 ? tile : tr.i1[r];
      tr.c1[r] = fminf(tr.c1[r], tm);
    }
  }
}

// Refine: exact evaluation of up to two tiles per query (tile ids ta, tb; cnt of them valid),
// candidates selected by the fp32 filter against thr; cnt > 2 or more than two candidates send the
// query to the cooperative exact scan.  Same structure as search_phase2.
template <int T, int MODE, int Q>
__device__ __forceinline__ void refine_tiles(QueryState<Q>& s, const float4* __restrict__ tgt, int c0, int nt,
                                             int ntile, const int (&cnt)[Q], const int (&ta)[Q], const int (&tb)[Q],
                                             const float (&thr)[Q]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < Q; j++) {
    bool hard = false;
    if (s.valid[j]) {
      hard = cnt[j] > 2;
      if (!hard) {
        int n = 0, ca = 0, cb = 0;
        if (cnt[j] >= 1)
          scan_tile_candidates<T>(tgt + (size_t)ta[j] * T, c0 + ta[j] * T, nt, s.ax2[j], s.ay2[j], s.az2[j], thr[j], n,
                                  ca, cb);
        if (cnt[j] >= 2)
          scan_tile_candidates<T>(tgt + (size_t)tb[j] * T, c0 + tb[j] * T, nt, s.ax2[j], s.ay2[j], s.az2[j], thr[j], n,
                                  ca, cb);
        if (n >= 1 && n <= 2) eval_candidate<MODE>(tgt, c0, ca, s.qx[j], s.qy[j], s.qz[j], s.best[j], s.besti[j]);
        if (n == 2) eval_candidate<MODE>(tgt, c0, cb, s.qx[j], s.qy[j], s.qz[j], s.best[j], s.besti[j]);
        hard = n > 2;
      }
    }
    unsigned pending = __ballot_sync(0xffffffffu, hard);
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
      warp_exact_scan<MODE>(tgt, c0, nt, ntile * (T / 2), bqx, bqy, bqz, bax, bay, baz, bthr, b, bi, lane);
      if (lane == src && (b < s.best[j] || (b == s.best[j] && bi < s.besti[j]))) {
        s.best[j] = b;
        s.besti[j] = bi;
      }
    }
  }
}

}  // namespace ga
