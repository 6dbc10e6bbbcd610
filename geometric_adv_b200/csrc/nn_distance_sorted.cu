// Chamfer forward with exact spatial pruning ("sorted" path, used when the caller provides a
// workspace; see ga_nn_distance_fwd_ws in ga_b200.h).
//
// The plain kernel (nn_distance_fwd.cu) evaluates the filter for every (query, target) pair:
// ~4 issue cycles per evaluation is the floor of that formulation on the FP32 pipe.  This
// path keeps the same filter / window / refine logic (same bit-exact results) but skips whole
// tiles of targets that provably cannot hold a candidate:
//
//  1. cloud_sort_kernel (one CTA per cloud): order the points along a Morton curve, write them
//     as staged pair-SoA tiles (ready to copy into shared memory), the permutation back to
//     the caller's indices, and an axis-aligned box per tile of 32 points.
//  2. nn_fwd_sorted_kernel: queries are taken in Morton order too, so the 128 queries of a
//     warp occupy a small region.  Per warp: box-to-box lower bounds LB to all target tiles,
//     tiles sorted by LB, scanned in that order; the walk stops at the first tile with
//         LB_safe > max over the warp's queries of (filter minimum so far + |q|^2) + 2W,
//     because no target in that tile (or any later one) can come within the filter window of
//     any of the warp's queries.  Everything is warp-uniform: no divergence is introduced.
//
// Soundness of the skip.  For a target t in the tile and a query q of the warp the real squared
// distance D >= LB.  The filter obeys f(q,t) >= D - |q|^2 - 18.1u s^2 (nn_distance_fwd.cu), and
// nq = fl(|q|^2) is within 9u s^2 of |q|^2, so f(q,t) >= LB - nq - 27.1u s^2.  A target can only
// matter if f(q,t) <= m_q + W (m_q = filter minimum of q, which only decreases).  With
// LB_safe = LB (1 - 2^-18) and the threshold max_q(m_q + nq) + 2W, W = 128u s^2, the skip
// condition implies f(q,t) > m_q + W for every q of the warp.  Non-finite clouds get infinite
// boxes (LB = 0: nothing is skipped).  Ties are still broken by the caller's ORIGINAL index.
#include <atomic>

#include "nn_search.cuh"

namespace ga {

constexpr int kSortMaxPts = 2048;  // clouds up to this size take the sorted path
constexpr int kSortT = 32;         // targets per tile
constexpr int kSortThreads = 256;

// Workspace of one cloud (npad = n rounded up to a multiple of 64, so tiles pair up):
//   header   16 floats: [0] max|coord|, [1..3] original point 0, [4] finite flag
//   boxes    (npad/32) x 8 floats: min x,y,z,_, max x,y,z,_
//   perm     npad x u16 (sorted position -> original index)
//   pts      (npad/2 + kPipeU) pairs x 32 B, pair-SoA with norms
struct CloudWs {
  float* header;
  float* boxes;
  unsigned short* perm;
  float4* pts;
};

__host__ __device__ inline int sort_npad(int n) { return (n + 63) & ~63; }
__host__ __device__ inline size_t cloud_ws_bytes(int n) {
  const size_t npad = (size_t)sort_npad(n);
  size_t b = 64 + (npad / kSortT) * 32 + npad * 2 + (npad / 2 + kPipeU) * 32;
  return (b + 255) & ~(size_t)255;
}
__host__ __device__ inline CloudWs cloud_ws(void* base, int n) {
  const size_t npad = (size_t)sort_npad(n);
  char* p = reinterpret_cast<char*>(base);
  CloudWs w;
  w.header = reinterpret_cast<float*>(p);
  w.boxes = reinterpret_cast<float*>(p + 64);
  w.perm = reinterpret_cast<unsigned short*>(p + 64 + (npad / kSortT) * 32);
  w.pts = reinterpret_cast<float4*>(p + 64 + (npad / kSortT) * 32 + npad * 2);
  return w;
}

__device__ __forceinline__ unsigned spread7(unsigned v) {  // 7 bits -> every third bit
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) r |= ((v >> i) & 1u) << (3 * i);
  return r;
}

struct SortArgs {
  int b, n, m;
  const float* xyz1;
  const float* xyz2;
  char* ws1;  // b clouds of n points
  char* ws2;  // b clouds of m points
  size_t stride1, stride2;
};

// ---- 1. Morton ordering of one cloud per CTA -------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) cloud_sort_kernel(const SortArgs a) {
  __shared__ float sx[kSortMaxPts], sy[kSortMaxPts], sz[kSortMaxPts];
  __shared__ unsigned keys[kSortMaxPts];
  __shared__ float red[6][kSortThreads / 32];
  __shared__ int bad_any;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int which = blockIdx.x & 1, batch = blockIdx.x >> 1;
  const int n = which ? a.m : a.n;
  if (n == 0) return;
  const float* pts = (which ? a.xyz2 : a.xyz1) + (size_t)batch * n * 3;
  CloudWs w = cloud_ws((which ? a.ws2 + (size_t)batch * a.stride2 : a.ws1 + (size_t)batch * a.stride1), n);
  const float kInf = __int_as_float(0x7f800000);
  if (tid == 0) bad_any = 0;
  float mn[3] = {kInf, kInf, kInf}, mx[3] = {-kInf, -kInf, -kInf};
  bool bad = false;
  for (int i = tid; i < n; i += kSortThreads) {
    const float x = __ldg(pts + (size_t)i * 3), y = __ldg(pts + (size_t)i * 3 + 1), z = __ldg(pts + (size_t)i * 3 + 2);
    sx[i] = x; sy[i] = y; sz[i] = z;
    bad |= !(fabsf(x) < kInf) || !(fabsf(y) < kInf) || !(fabsf(z) < kInf);
    mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
    mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
  }
  __syncthreads();
  if (bad) bad_any = 1;
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) {
      red[c][warp] = mn[c];
      red[3 + c][warp] = mx[c];
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < kSortThreads / 32; k++) {
      mn[c] = fminf(mn[c], red[c][k]);
      mx[c] = fmaxf(mx[c], red[3 + c][k]);
    }
  const bool finite = bad_any == 0;
  // sort size: next power of two >= n
  int S = 64;
  while (S < n) S <<= 1;
  float sc[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float ext = mx[c] - mn[c];
    sc[c] = (finite && ext > 0.f) ? 127.99f / ext : 0.f;
  }
  for (int i = tid; i < S; i += kSortThreads) {
    unsigned key = 0xffffffffu;
    if (i < n) {
      unsigned qx = 0, qy = 0, qz = 0;
      if (finite) {
        qx = min(127u, (unsigned)((sx[i] - mn[0]) * sc[0]));
        qy = min(127u, (unsigned)((sy[i] - mn[1]) * sc[1]));
        qz = min(127u, (unsigned)((sz[i] - mn[2]) * sc[2]));
      }
      key = ((spread7(qx) | (spread7(qy) << 1) | (spread7(qz) << 2)) << 11) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  // bitonic sort of S composite keys (unique, so the order is total and deterministic)
  for (int k = 2; k <= S; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < S / 2; t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // index with bit j cleared
        const int l = i | j;
        const unsigned ka = keys[i], kb = keys[l];
        const bool up = (i & k) == 0;
        if ((ka > kb) == up) {
          keys[i] = kb;
          keys[l] = ka;
        }
      }
      __syncthreads();
    }
  }
  // ---- write the workspace ---------------------------------------------------------------
  const int npad = sort_npad(n);
  if (tid == 0) {
    float bm = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) bm = fmaxf(bm, fmaxf(fabsf(mn[c]), fabsf(mx[c])));
    if (!finite) bm = kInf;
    w.header[0] = bm;
    w.header[1] = sx[0];
    w.header[2] = sy[0];
    w.header[3] = sz[0];
    w.header[4] = finite ? 1.f : 0.f;
  }
  for (int p = tid; p < npad / 2 + kPipeU; p += kSortThreads) {
    float c[2][3], nn[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int g = 2 * p + h;
      c[h][0] = c[h][1] = c[h][2] = 0.f;
      nn[h] = kInf;
      if (g < n) {
        const int o = (int)(keys[g] & 2047u);
        c[h][0] = sx[o]; c[h][1] = sy[o]; c[h][2] = sz[o];
        nn[h] = fmaf(c[h][2], c[h][2], fmaf(c[h][1], c[h][1], c[h][0] * c[h][0]));
      }
    }
    w.pts[2 * p] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
    w.pts[2 * p + 1] = make_float4(c[0][2], c[1][2], nn[0], nn[1]);
  }
  for (int g = tid; g < npad; g += kSortThreads) w.perm[g] = g < n ? (unsigned short)(keys[g] & 2047u) : 0;
  for (int t = tid; t < npad / kSortT; t += kSortThreads) {
    float bn[3] = {kInf, kInf, kInf}, bx[3] = {-kInf, -kInf, -kInf};
    for (int g = t * kSortT; g < min(n, (t + 1) * kSortT); g++) {
      const int o = (int)(keys[g] & 2047u);
      bn[0] = fminf(bn[0], sx[o]); bn[1] = fminf(bn[1], sy[o]); bn[2] = fminf(bn[2], sz[o]);
      bx[0] = fmaxf(bx[0], sx[o]); bx[1] = fmaxf(bx[1], sy[o]); bx[2] = fmaxf(bx[2], sz[o]);
    }
    if (!finite) {  // unbounded boxes: lower bound 0, the tile is never skipped
#pragma unroll
      for (int c = 0; c < 3; c++) { bn[c] = -kInf; bx[c] = kInf; }
    }
    // an empty (all-padding) tile keeps min=+inf, max=-inf: its lower bound is +inf
    float* bo = w.boxes + (size_t)t * 8;
    bo[0] = bn[0]; bo[1] = bn[1]; bo[2] = bn[2]; bo[3] = 0.f;
    bo[4] = bx[0]; bo[5] = bx[1]; bo[6] = bx[2]; bo[7] = 0.f;
  }
}

// ---- 2. pruned search ----------------------------------------------------------------------
struct SortedFwdArgs {
  int b, n, m;
  const char* ws1;
  const char* ws2;
  size_t stride1, stride2;
  float* dist1;
  int* idx1;
  float* dist2;
  int* idx2;
  int tiles1, tiles2;
};

template <int THREADS, int Q>
struct SortedCfg {
  static constexpr int kThreads = THREADS, kQ = Q, kQT = THREADS * Q, kT = kSortT;
  // tgt pairs (+pad) | boxes[64][8] | perm[2048] u16 | order scratch
  static constexpr size_t kOffBoxes = ((size_t)kSortMaxPts / 2 + kPipeU) * 32;
  static constexpr size_t kOffPerm = kOffBoxes + (kSortMaxPts / kSortT) * 32;
  static constexpr size_t kSmem = kOffPerm + (size_t)kSortMaxPts * 2;
};

// filter minima of ONE tile for Q queries (loads double-buffered inside the tile)
template <int Q, int T>
__device__ __forceinline__ void scan_one_tile(const float4* __restrict__ tp, const float (&ax2)[Q],
                                              const float (&ay2)[Q], const float (&az2)[Q], float (&tm)[Q]) {
  constexpr int U = kPipeU;
  constexpr int NB = (T / 2) / U;
  const float kInf = __int_as_float(0x7f800000);
#pragma unroll
  for (int j = 0; j < Q; j++) tm[j] = kInf;
  float4 buf[2][2 * U];
#pragma unroll
  for (int e = 0; e < 2 * U; e++) buf[0][e] = tp[e];
#pragma unroll
  for (int blk = 0; blk < NB; blk++) {
    if (blk + 1 < NB) {
#pragma unroll
      for (int e = 0; e < 2 * U; e++) buf[(blk + 1) & 1][e] = tp[(blk + 1) * 2 * U + e];
    }
#pragma unroll
    for (int pp = 0; pp < U; pp++) {
      const float4 u = buf[blk & 1][2 * pp];
      const float4 v = buf[blk & 1][2 * pp + 1];
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const float2 f = filter_pair(u, v, ax2[j], ay2[j], az2[j]);
        tm[j] = fmin3(tm[j], f.x, f.y);
      }
    }
  }
}

// reference arithmetic for one staged target at SORTED position g; ties by ORIGINAL index
template <int MODE>
__device__ __forceinline__ void eval_sorted(const float4* __restrict__ tgt, const unsigned short* __restrict__ perm,
                                            int g, float qx, float qy, float qz, float& best, int& besti) {
  const float* pu = reinterpret_cast<const float*>(tgt + 2 * (g >> 1));
  const int h = g & 1;
  const float d = sqdist<MODE>(pu[h], pu[2 + h], pu[4 + h], qx, qy, qz);
  const int o = perm[g];
  if (d < best || (d == best && o < besti)) {
    best = d;
    besti = o;
  }
}

template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::kThreads) nn_fwd_sorted_kernel(const SortedFwdArgs a) {
  constexpr int THREADS = Cfg::kThreads, Q = Cfg::kQ, QT = Cfg::kQT, T = Cfg::kT;
  const float kInf = __int_as_float(0x7f800000);
  extern __shared__ float4 smem_f4[];
  unsigned char* smem_raw = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* boxes = reinterpret_cast<float*>(smem_raw + Cfg::kOffBoxes);
  unsigned short* perm = reinterpret_cast<unsigned short*>(smem_raw + Cfg::kOffPerm);

  const int tid = threadIdx.x, lane = tid & 31;
  const int jpb = a.tiles1 + a.tiles2;
  const int batch = blockIdx.x / jpb;
  const int r = blockIdx.x - batch * jpb;
  const bool rev = r >= a.tiles1;
  const int qtile = rev ? r - a.tiles1 : r;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const CloudWs wq = cloud_ws(const_cast<char*>((rev ? a.ws2 + (size_t)batch * a.stride2 : a.ws1 + (size_t)batch * a.stride1)), nq);
  const CloudWs wt = cloud_ws(const_cast<char*>((rev ? a.ws1 + (size_t)batch * a.stride1 : a.ws2 + (size_t)batch * a.stride2)), nt);
  float* odist = (rev ? a.dist2 : a.dist1) + (size_t)batch * nq;
  int* oidx = (rev ? a.idx2 : a.idx1) + (size_t)batch * nq;
  const int ntile = sort_npad(nt) / T;  // even

  // ---- stage the pre-formatted target cloud: straight 16-byte copies -------------------------
  {
    const int nvec = sort_npad(nt) + 2 * kPipeU;  // float4 count: 2 per pair
    for (int i = tid; i < nvec; i += THREADS) tgt[i] = __ldg(wt.pts + i);
    const float4* bsrc = reinterpret_cast<const float4*>(wt.boxes);
    float4* bdst = reinterpret_cast<float4*>(boxes);
    for (int i = tid; i < ntile * 2; i += THREADS) bdst[i] = __ldg(bsrc + i);
    const uint32_t* psrc = reinterpret_cast<const uint32_t*>(wt.perm);
    uint32_t* pdst = reinterpret_cast<uint32_t*>(perm);
    for (int i = tid; i < sort_npad(nt) / 2; i += THREADS) pdst[i] = __ldg(psrc + i);
  }
  const float bm = __ldg(wt.header);
  const float t0x = __ldg(wt.header + 1), t0y = __ldg(wt.header + 2), t0z = __ldg(wt.header + 3);

  // ---- queries in Morton order -------------------------------------------------------------
  float qx[Q], qy[Q], qz[Q], qabs[Q], ax2[Q], ay2[Q], az2[Q], nqv[Q], d0[Q];
  bool valid[Q];
  float rmin[3] = {kInf, kInf, kInf}, rmax[3] = {-kInf, -kInf, -kInf};
#pragma unroll
  for (int j = 0; j < Q; j++) {
    // a warp owns Q*32 CONSECUTIVE sorted queries, so that its region is compact
    const int qi = qtile * QT + (tid >> 5) * (32 * Q) + j * 32 + lane;
    valid[j] = qi < nq;
    const int qs = valid[j] ? qi : 0;
    const float* pu = reinterpret_cast<const float*>(wq.pts + 2 * (qs >> 1));
    const int h = qs & 1;
    qx[j] = __ldg(pu + h);
    qy[j] = __ldg(pu + 2 + h);
    qz[j] = __ldg(pu + 4 + h);
    qabs[j] = query_abs(qx[j], qy[j], qz[j]);
    ax2[j] = -2.0f * qx[j];
    ay2[j] = -2.0f * qy[j];
    az2[j] = -2.0f * qz[j];
    nqv[j] = fmaf(qz[j], qz[j], fmaf(qy[j], qy[j], qx[j] * qx[j]));
    d0[j] = sqdist<MODE>(t0x, t0y, t0z, qx[j], qy[j], qz[j]);
    if (valid[j]) {
      rmin[0] = fminf(rmin[0], qx[j]); rmin[1] = fminf(rmin[1], qy[j]); rmin[2] = fminf(rmin[2], qz[j]);
      rmax[0] = fmaxf(rmax[0], qx[j]); rmax[1] = fmaxf(rmax[1], qy[j]); rmax[2] = fmaxf(rmax[2], qz[j]);
    }
  }
  bool qbad = false;
#pragma unroll
  for (int j = 0; j < Q; j++) qbad |= valid[j] && !(qabs[j] < kInf);
  qbad = __any_sync(0xffffffffu, qbad);
#pragma unroll
  for (int c = 0; c < 3; c++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rmin[c] = fminf(rmin[c], __shfl_xor_sync(0xffffffffu, rmin[c], o));
      rmax[c] = fmaxf(rmax[c], __shfl_xor_sync(0xffffffffu, rmax[c], o));
    }
  }
  __syncthreads();  // staged data visible

  // ---- lower bounds to every tile, sorted ascending (two tiles per lane, 64-wide bitonic) ------
  float lb[2];
  int tl[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int t = lane + 32 * e;
    float v = kInf;
    if (t < ntile) {
      const float* bo = boxes + t * 8;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float g = fmaxf(0.f, fmaxf(bo[c] - rmax[c], rmin[c] - bo[4 + c]));  // NaN gaps drop to 0
        s = fmaf(g, g, s);
      }
      v = qbad ? 0.f : s * 0.99999619f;  // (1 - 2^-18): stays a lower bound after rounding
      if (!(v == v)) v = 0.f;
    }
    lb[e] = v;
    tl[e] = t;
  }
  // element index of (lane, e) is lane + 32 e; compare-exchange network over 64 elements
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      if (jj == 32) {  // partner is the other element of the same lane
        const bool up = true;  // k == 64: whole sequence ascending
        if ((lb[0] > lb[1] || (lb[0] == lb[1] && tl[0] > tl[1])) == up) {
          const float tv = lb[0]; lb[0] = lb[1]; lb[1] = tv;
          const int ti = tl[0]; tl[0] = tl[1]; tl[1] = ti;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = lane + 32 * e;
          const float ov = __shfl_xor_sync(0xffffffffu, lb[e], jj);
          const int ot = __shfl_xor_sync(0xffffffffu, tl[e], jj);
          const bool up = (i & k) == 0;
          const bool lower = (i & jj) == 0;  // this element keeps the smaller one when ascending
          const bool mine_gt = lb[e] > ov || (lb[e] == ov && tl[e] > ot);
          const bool take = (lower == up) ? mine_gt : !mine_gt;
          if (take) {
            lb[e] = ov;
            tl[e] = ot;
          }
        }
      }
    }
  }

  // ---- walk the tiles in ascending lower-bound order ------------------------------------------
  float c1[Q], c2[Q], c3[Q];
  int i1[Q], i2[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    c1[j] = c2[j] = c3[j] = kInf;
    i1[j] = i2[j] = 0;
  }
  float wq2[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) wq2[j] = 2.0f * filter_window(qabs[j], bm);
  for (int it = 0; it < ntile; it++) {
    const float lbt = __shfl_sync(0xffffffffu, it < 32 ? lb[0] : lb[1], it & 31);
    const int tile = __shfl_sync(0xffffffffu, it < 32 ? tl[0] : tl[1], it & 31);
    float ub = -kInf;
#pragma unroll
    for (int j = 0; j < Q; j++)
      if (valid[j]) ub = fmaxf(ub, (c1[j] + nqv[j]) + wq2[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ub = fmaxf(ub, __shfl_xor_sync(0xffffffffu, ub, o));
    if (lbt > ub) break;  // warp-uniform; NaN bounds never break
    float tm[Q];
    scan_one_tile<Q, T>(tgt + (size_t)tile * T, ax2, ay2, az2, tm);
#pragma unroll
    for (int j = 0; j < Q; j++) {
      const bool lt1 = tm[j] < c1[j], lt2 = tm[j] < c2[j];
      c3[j] = fminf(c3[j], fmaxf(c2[j], tm[j]));
      i2[j] = lt1 ? i1[j] : (lt2 ? tile : i2[j]);
      c2[j] = fminf(c2[j], fmaxf(c1[j], tm[j]));
      i1[j] = lt1 ? tile : i1[j];
      c1[j] = fminf(c1[j], tm[j]);
    }
  }

  // ---- refine (as in nn_search.cuh, with original-index tie breaking) --------------------------
  float best[Q];
  int besti[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    best[j] = kInf;
    besti[j] = 0;
    float thr = 0.f;
    bool hard = false;
    if (valid[j]) {
      thr = c1[j] + filter_window(qabs[j], bm);
      hard = !(c3[j] > thr);
      if (!hard) {
        int cnt = 0, ca = 0, cb = 0;
        if (!(c1[j] > thr))
          scan_tile_candidates<T>(tgt + (size_t)i1[j] * T, i1[j] * T, nt, ax2[j], ay2[j], az2[j], thr, cnt, ca, cb);
        if (!(c2[j] > thr))
          scan_tile_candidates<T>(tgt + (size_t)i2[j] * T, i2[j] * T, nt, ax2[j], ay2[j], az2[j], thr, cnt, ca, cb);
        if (cnt >= 1 && cnt <= 2) eval_sorted<MODE>(tgt, perm, ca, qx[j], qy[j], qz[j], best[j], besti[j]);
        if (cnt == 2) eval_sorted<MODE>(tgt, perm, cb, qx[j], qy[j], qz[j], best[j], besti[j]);
        hard = cnt > 2;
      }
    }
    unsigned pending = __ballot_sync(0xffffffffu, hard);
    while (pending) {  // rare: the whole warp scans the staged cloud for one query
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const float bqx = __shfl_sync(0xffffffffu, qx[j], src), bqy = __shfl_sync(0xffffffffu, qy[j], src),
                  bqz = __shfl_sync(0xffffffffu, qz[j], src);
      const float bax = __shfl_sync(0xffffffffu, ax2[j], src), bay = __shfl_sync(0xffffffffu, ay2[j], src),
                  baz = __shfl_sync(0xffffffffu, az2[j], src);
      const float bthr = __shfl_sync(0xffffffffu, thr, src);
      float b = kInf;
      int bi = 0x7fffffff;
      for (int p = lane; p < ntile * (T / 2); p += 32) {
        const float2 f = filter_pair(tgt[2 * p], tgt[2 * p + 1], bax, bay, baz);
        if (!(f.x > bthr) && 2 * p < nt) eval_sorted<MODE>(tgt, perm, 2 * p, bqx, bqy, bqz, b, bi);
        if (!(f.y > bthr) && 2 * p + 1 < nt) eval_sorted<MODE>(tgt, perm, 2 * p + 1, bqx, bqy, bqz, b, bi);
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
      if (lane == src && bi != 0x7fffffff) {
        best[j] = b;
        besti[j] = bi;
      }
    }
  }

#pragma unroll
  for (int j = 0; j < Q; j++) {
    if (!valid[j]) continue;
    const int qi = qtile * QT + (tid >> 5) * (32 * Q) + j * 32 + lane;
    const int qo = __ldg(wq.perm + qi);
    const bool seed_nan = d0[j] != d0[j];
    odist[qo] = seed_nan ? d0[j] : best[j];
    oidx[qo] = seed_nan ? 0 : besti[j];
  }
}

int g_sorted_variant = 0;

template <class Cfg>
static int launch_sorted(const SortedFwdArgs& a0, int mode, cudaStream_t st) {
  SortedFwdArgs a = a0;
  a.tiles1 = (a.n + Cfg::kQT - 1) / Cfg::kQT;
  a.tiles2 = (a.m + Cfg::kQT - 1) / Cfg::kQT;
  const long long jobs = (long long)a.b * (a.tiles1 + a.tiles2);
  if (jobs > 0x7fffffffLL) {
    set_error("ga_nn_distance_fwd_ws: problem too large for one launch");
    return GA_ERR_UNSUPPORTED;
  }
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_sorted_kernel<Cfg, GA_MODE_CPU_EXACT>
                                     : nn_fwd_sorted_kernel<Cfg, GA_MODE_GPU_REF>;
  GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
  k<<<(unsigned)jobs, Cfg::kThreads, Cfg::kSmem, st>>>(a);
  GA_LAUNCH_CHECK("nn_fwd_sorted_kernel");
  return GA_OK;
}

}  // namespace ga

using namespace ga;

extern "C" {

size_t ga_nn_distance_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0 || n > kSortMaxPts || m > kSortMaxPts) return 0;
  return (size_t)b * (cloud_ws_bytes(n) + cloud_ws_bytes(m));
}

int ga_nn_distance_fwd_ws(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                          float* dist2, int* idx2, int mode, void* workspace, size_t workspace_bytes,
                          ga_stream_t stream) {
  const size_t need = ga_nn_distance_workspace_bytes(b, n, m);
  // small clouds gain nothing from pruning; big ones do not fit the single-chunk kernel
  if (workspace == nullptr || need == 0 || workspace_bytes < need || n < 256 || m < 256 ||
      (mode != GA_MODE_CPU_EXACT && mode != GA_MODE_GPU_REF))
    return ga_nn_distance_fwd(b, n, m, xyz1, xyz2, dist1, idx1, dist2, idx2, mode, stream);
  cudaStream_t st = as_stream(stream);
  SortArgs sa;
  sa.b = b; sa.n = n; sa.m = m; sa.xyz1 = xyz1; sa.xyz2 = xyz2;
  sa.ws1 = reinterpret_cast<char*>(workspace);
  sa.stride1 = cloud_ws_bytes(n);
  sa.ws2 = sa.ws1 + (size_t)b * sa.stride1;
  sa.stride2 = cloud_ws_bytes(m);
  cloud_sort_kernel<<<2 * b, kSortThreads, 0, st>>>(sa);
  GA_LAUNCH_CHECK("cloud_sort_kernel");
  SortedFwdArgs fa;
  fa.b = b; fa.n = n; fa.m = m;
  fa.ws1 = sa.ws1; fa.ws2 = sa.ws2; fa.stride1 = sa.stride1; fa.stride2 = sa.stride2;
  fa.dist1 = dist1; fa.idx1 = idx1; fa.dist2 = dist2; fa.idx2 = idx2;
  fa.tiles1 = fa.tiles2 = 0;
  switch (g_sorted_variant) {
    case 1: return launch_sorted<SortedCfg<64, 4>>(fa, mode, st);
    case 2: return launch_sorted<SortedCfg<128, 2>>(fa, mode, st);
    case 3: return launch_sorted<SortedCfg<64, 2>>(fa, mode, st);
    case 4: return launch_sorted<SortedCfg<256, 2>>(fa, mode, st);
    default: return launch_sorted<SortedCfg<128, 4>>(fa, mode, st);
  }
}

}  // extern "C"
