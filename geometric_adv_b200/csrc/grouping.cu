// grouping ops: knn_point (fused, no dense matrix), the legacy dense selection
// sort, group_point, and the defense's per-point kNN distances.
//
// Reference: knn_point = TF-side dense (b,m,n) squared-distance tensor
// (tf_grouping.py:66-68) + selection_sort_gpu (tf_grouping_g.cu:83-123, one thread
// per row, k passes over the row in global memory) + tf.slice.  Here: one kernel,
// O(n + k) memory per query.
//
// Algorithm (same filter as nn_distance_fwd.cu, f(q,t) = |t|^2 - 2 q.t):
//   pass A  FFMA2/FMNMX3 filter scan; minima per group of tiles feed a (k+1)-entry
//           sorted list whose last entry tau bounds the (k+1)-th smallest filter
//           value from above (k+1 different groups each hold a point with f <= tau).
//   pass B1 second filter scan: points with f <= tau + W (about k+4 per query) are
//           appended to a per-query queue in shared memory, branch-free.
//   pass B2 the queued points are evaluated in the reference arithmetic
//           ((dx*dx+dy*dy)+dz*dz, dx = data - query) and inserted into an exact,
//           sorted (k+1)-entry list, all lanes in step.  W is the rounding window of
//           the filter, so the exact k+1 nearest are always among the candidates.
//   ties    the reference's selection sort is unstable on exact ties.  If the exact
//           list shows a tie (or NaN / missing entries), the query is replayed by a
//           warp that simulates the selection sort on the only elements that can
//           take part in it (positions < k, values < v_k, first k values == v_k).
#include "nn_mma.cuh"

namespace ga {

// ---------------------------------------------------------------------------
// Exact replay of the reference selection sort for one query, by one warp.
// pts: data set (n points), q: query, vk: k-th smallest exact squared distance.
// cv/ci: shared scratch, >= 3k entries; on return cv[0..k) / ci[0..k) hold the
// reference's first k columns.
//
// Only these elements can take part in the first k rounds of the selection sort:
// positions < k (they get displaced), values < v_k (they get selected), and the
// first k elements at positions >= k with value == v_k (selected in position
// order).  Everything else is never selected and never moved, so the literal sort
// on the compacted row gives the reference's result.
// ---------------------------------------------------------------------------
__device__ void selection_replay(const float* __restrict__ pts, int n, float qx, float qy, float qz, int k,
                                 float vk, float* cv, int* ci, int lane) {
  const unsigned lt_mask = (1u << lane) - 1u;
  int cnt = 0, eqcnt = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool in = i < n;
    float d = 0.f;
    if (in)
      d = sqdist<GA_MODE_CPU_EXACT>(__ldg(pts + (size_t)i * 3), __ldg(pts + (size_t)i * 3 + 1),
                                    __ldg(pts + (size_t)i * 3 + 2), qx, qy, qz);
    const bool low = in && i < k;
    const bool lt = in && d < vk;
    const bool eq = in && !low && !lt && d == vk;
    const unsigned meq = __ballot_sync(0xffffffffu, eq);
    const bool take = low || lt || (eq && eqcnt + __popc(meq & lt_mask) < k);
    const unsigned mt = __ballot_sync(0xffffffffu, take);
    if (take) {
      const int pos = cnt + __popc(mt & lt_mask);
      cv[pos] = d;
      ci[pos] = i;
    }
    cnt += __popc(mt);
    eqcnt = min(k, eqcnt + __popc(meq));
  }
  __syncwarp();
  if (lane == 0) {
    for (int s = 0; s < k; s++) {
      int mn = s;
      for (int t = s + 1; t < cnt; t++)
        if (cv[t] < cv[mn]) mn = t;
      const float tv = cv[mn];
      const int ti = ci[mn];
      cv[mn] = cv[s];
      ci[mn] = ci[s];
      cv[s] = tv;
      ci[s] = ti;
    }
  }
  __syncwarp();
}

// k-th smallest exact squared distance of one query by k rounds of "next smallest
// (value, position)", one warp.  O(k n); used when the fast path cannot decide.
__device__ float kth_smallest_warp(const float* __restrict__ pts, int n, float qx, float qy, float qz, int k,
                                   int lane) {
  const float kInf = __int_as_float(0x7f800000);
  float pv = -kInf;
  int pp = -1;
  for (int s = 0; s < k; s++) {
    float bv = kInf;
    int bp = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      const float d = sqdist<GA_MODE_CPU_EXACT>(__ldg(pts + (size_t)i * 3), __ldg(pts + (size_t)i * 3 + 1),
                                                __ldg(pts + (size_t)i * 3 + 2), qx, qy, qz);
      const bool after = d > pv || (d == pv && i > pp);
      if (after && (d < bv || (d == bv && i < bp))) {
        bv = d;
        bp = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov < bv || (ov == bv && op < bp)) {
        bv = ov;
        bp = op;
      }
    }
    if (bp == 0x7fffffff) return kInf;  // fewer than k comparable distances
    pv = bv;
    pp = bp;
  }
  return pv;
}

// Write one query's result from the replay scratch, applying the defense epilogue.
__device__ __forceinline__ void write_replayed(const float* cv, const int* ci, int k, int skip, int do_sqrt,
                                               float* vo, int* io, int lane) {
  for (int s = skip + lane; s < k; s += 32) {
    vo[s - skip] = do_sqrt ? __fsqrt_rn(cv[s]) : cv[s];
    if (io) io[s - skip] = ci[s];
  }
  __syncwarp();
}

struct KnnArgs {
  int b, n, m, k;      // k = neighbours searched (already k+1 for the defense epilogue)
  const float* xyz1;   // data set (b,n,3)
  const float* xyz2;   // queries  (b,m,3)
  float* val;          // (b,m,k-skip)
  int* idx;            // (b,m,k-skip) or nullptr
  int skip;            // leading neighbours dropped (defense: 1)
  int do_sqrt;         // defense: sqrt of the squared distance
  int qtiles;          // query tiles per cloud
};

constexpr int kKnnQueue = 32;  // candidate slots per query

template <int THREADS, int Q, int T, int CH, int KL, bool MULTI>
struct KnnCfg {
  static constexpr int kThreads = THREADS, kQ = Q, kQT = THREADS * Q, kT = T, kCH = CH, kKL = KL;
  static constexpr bool kMulti = MULTI;  // data sets larger than CH: lists persist across chunks
  static constexpr int kTiles = CH / T;
  static constexpr int kWarps = THREADS / 32;
  static_assert((T & (T - 1)) == 0, "tile size must be a power of two");
  // tgt | pad | red[32] | queue[C+T][QT] u16 | flag_q[QT] | flag_vk[QT] | flag_cnt | replay scratch
  static constexpr size_t kOffRed = (size_t)CH * 16 + (size_t)kPipeU * 32;
  static constexpr size_t kOffQueue = kOffRed + 32 * 4;
  static constexpr size_t kOffFlagQ = kOffQueue + (((size_t)(kKnnQueue + T) * kQT * 2 + 15) & ~(size_t)15);
  static constexpr size_t kOffFlagV = kOffFlagQ + (size_t)kQT * 4;
  static constexpr size_t kOffCnt = kOffFlagV + (size_t)kQT * 4;
  static constexpr size_t kOffScratch = kOffCnt + 16;
  static constexpr size_t kSmem = kOffScratch + (size_t)kWarps * 3 * KL * 8;
};

// sorted insert of (d, g) into an ascending (value, index) list held in registers
template <int KL>
__device__ __forceinline__ void list_insert(float (&Lv)[KL], int (&Li)[KL], float d, int g) {
  if (d < Lv[KL - 1] || (d == Lv[KL - 1] && g < Li[KL - 1])) {
#pragma unroll
    for (int s = KL - 1; s >= 1; s--) {
      const bool up = d < Lv[s - 1] || (d == Lv[s - 1] && g < Li[s - 1]);
      const bool here = d < Lv[s] || (d == Lv[s] && g < Li[s]);
      Li[s] = up ? Li[s - 1] : (here ? g : Li[s]);
      Lv[s] = up ? Lv[s - 1] : (here ? d : Lv[s]);
    }
    if (d < Lv[0] || (d == Lv[0] && g < Li[0])) {
      Lv[0] = d;
      Li[0] = g;
    }
  }
}

// Body of knn_kernel as a device function: knn_slab_kernel (below) runs it for clouds it does not take itself.
template <class Cfg>
__device__ __forceinline__ void knn_body(const KnnArgs& a) {
  constexpr int THREADS = Cfg::kThreads, Q = Cfg::kQ, QT = Cfg::kQT, T = Cfg::kT, CH = Cfg::kCH, KL = Cfg::kKL;
  constexpr bool MULTI = Cfg::kMulti;
  constexpr int LQ = MULTI ? Q : 1;  // persistent lists only when chunks have to be merged
  const float kInf = __int_as_float(0x7f800000);
  const float kNaN = __int_as_float(0x7fc00000);
  extern __shared__ float4 smem_f4[];
  unsigned char* smem_raw = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem_raw + Cfg::kOffRed);
  unsigned short* queue = reinterpret_cast<unsigned short*>(smem_raw + Cfg::kOffQueue);
  int* flag_q = reinterpret_cast<int*>(smem_raw + Cfg::kOffFlagQ);
  float* flag_vk = reinterpret_cast<float*>(smem_raw + Cfg::kOffFlagV);
  int* flag_cnt = reinterpret_cast<int*>(smem_raw + Cfg::kOffCnt);
  unsigned char* scratch = smem_raw + Cfg::kOffScratch;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int batch = blockIdx.x / a.qtiles;
  const int qtile = blockIdx.x - batch * a.qtiles;
  const int n = a.n, k = a.k;
  const float* pts = a.xyz1 + (size_t)batch * n * 3;
  const float* qpts = a.xyz2 + (size_t)batch * a.m * 3;
  const int kout = k - a.skip;
  const bool need_idx = a.idx != nullptr;
  if (tid == 0) *flag_cnt = 0;

  float qx[Q], qy[Q], qz[Q], qabs[Q], ax2[Q], ay2[Q], az2[Q];
  bool valid[Q], replay[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    const int qi = qtile * QT + j * THREADS + tid;
    valid[j] = qi < a.m;
    const int qs = valid[j] ? qi : 0;
    qx[j] = __ldg(qpts + (size_t)qs * 3);
    qy[j] = __ldg(qpts + (size_t)qs * 3 + 1);
    qz[j] = __ldg(qpts + (size_t)qs * 3 + 2);
    qabs[j] = query_abs(qx[j], qy[j], qz[j]);
    ax2[j] = -2.0f * qx[j];
    ay2[j] = -2.0f * qy[j];
    az2[j] = -2.0f * qz[j];
    // A NaN distance at a position < k is "selected" by the reference (nothing compares
    // below NaN): only the replay reproduces that.
    bool nan_low = false;
    if (need_idx)
      for (int i = 0; i < k; i++) {
        const float d = sqdist<GA_MODE_CPU_EXACT>(__ldg(pts + (size_t)i * 3), __ldg(pts + (size_t)i * 3 + 1),
                                                  __ldg(pts + (size_t)i * 3 + 2), qx[j], qy[j], qz[j]);
        nan_low |= d != d;
      }
    replay[j] = nan_low;
  }

  // exact sorted lists (value, index); persistent across chunks only in the MULTI build
  float PLv[LQ][KL];
  int PLi[LQ][KL];
#pragma unroll
  for (int j = 0; j < LQ; j++)
#pragma unroll
    for (int s = 0; s < KL; s++) {
      PLv[j][s] = kInf;
      PLi[j][s] = -1;
    }
  float tau[Q];
  bool overflow[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    tau[j] = kInf;
    overflow[j] = false;
  }
  float bm_run = 0.0f;

  for (int c0 = 0; c0 < (MULTI ? n : 1); c0 += CH) {
    const int cn = min(CH, n - c0);
    const int ntile = (cn + T - 1) / T;
    bm_run = fmaxf(bm_run, stage_targets<THREADS, T>(tgt, red, pts, c0, n, ntile, tid));

    // ---- pass A: filter scan, tile minima to smem, k+1 smallest group minima -----
    {
      int tpg = ntile / (2 * (k + 1));  // tiles per group; >= k+1 groups give a finite tau
      if (tpg < 1) tpg = 1;
      float S[Q][KL];
      float gmin[Q];
#pragma unroll
      for (int j = 0; j < Q; j++) {
        gmin[j] = kInf;
#pragma unroll
        for (int s = 0; s < KL; s++) S[j][s] = kInf;
      }
      int left = tpg;
      filter_scan<Q, T>(tgt, ntile, ax2, ay2, az2, [&](int tile, const float(&tm)[Q]) {
#pragma unroll
        for (int j = 0; j < Q; j++) gmin[j] = fminf(gmin[j], tm[j]);
        if (--left == 0 || tile == ntile - 1) {
          left = tpg;
#pragma unroll
          for (int j = 0; j < Q; j++) {
            float v = gmin[j];
            gmin[j] = kInf;
#pragma unroll
            for (int s = 0; s < KL; s++) {  // sorted insert by compare-exchange chain
              const float lo = fminf(S[j][s], v);
              v = fmaxf(S[j][s], v);
              S[j][s] = lo;
            }
          }
        }
      });
#pragma unroll
      for (int j = 0; j < Q; j++) {
        float tk = kInf;
#pragma unroll
        for (int s = 0; s < KL; s++)
          if (s == k) tk = S[j][s];  // (k+1)-th smallest group minimum
        tau[j] = fminf(tau[j], tk);
      }
    }

    // ---- pass B1: second scan, candidates (f <= tau + W) appended to per-query queues ----
    // (the union of the tiles the 32 lanes of a warp need is essentially every tile, so the
    // scan is a plain broadcast walk like pass A, not a per-lane walk over qualifying tiles)
    float thr[Q];
    int cnts[Q];
#pragma unroll
    for (int j = 0; j < Q; j++) {
      // invalid / already overflowed slots collect nothing
      thr[j] = (valid[j] && !overflow[j]) ? tau[j] + filter_window(qabs[j], bm_run) : -kInf;
      cnts[j] = 0;
    }
    filter_collect<Q, T>(tgt, ntile, ax2, ay2, az2, thr, cnts, queue + tid, QT, THREADS, kKnnQueue);

    // ---- pass B2, per query slot: drain the queue into the exact sorted list ------------
#pragma unroll
    for (int j = 0; j < Q; j++) {
      if (!valid[j] || overflow[j]) continue;
      const unsigned short* myq = queue + j * THREADS + tid;
      const int cnt = cnts[j];
      if (cnt > kKnnQueue) {  // dense neighbourhood / degenerate data: exact warp path below
        overflow[j] = true;
        continue;
      }
      if (!need_idx && !MULTI) {
        // values only (ga_knn_dists): equal distances need no order, so a branch-free compare-exchange chain
        // replaces the index-ordered insert (a third of knn_slab_kernel's instructions before it did the same)
        float V[KL];
#pragma unroll
        for (int s = 0; s < KL; s++) V[s] = kInf;
        for (int c = 0; c < cnt; c++) {
          const int gl = myq[c * QT];
          const float* pu = reinterpret_cast<const float*>(tgt + 2 * (gl >> 1));
          const int h = gl & 1;
          float v = sqdist<GA_MODE_CPU_EXACT>(pu[h], pu[2 + h], pu[4 + h], qx[j], qy[j], qz[j]);
          v = (v == v && c0 + gl < n) ? v : kInf;  // NaN is never selected beyond position k; padding
#pragma unroll
          for (int s = 0; s < KL; s++) {
            const float lo = fminf(V[s], v);
            v = fmaxf(V[s], v);
            V[s] = lo;
          }
        }
        const int qi = qtile * QT + j * THREADS + tid;
        float* vo = a.val + ((size_t)batch * a.m + qi) * kout;
#pragma unroll
        for (int s = 0; s < KL; s++)
          if (s >= a.skip && s < k) vo[s - a.skip] = a.do_sqrt ? __fsqrt_rn(V[s]) : V[s];
        continue;
      }
      // B2: drain the queue into the exact sorted list (all lanes step through their own
      // queue together; the insert is order-independent)
      float TLv[KL];
      int TLi[KL];
#pragma unroll
      for (int s = 0; s < KL; s++) {
        TLv[s] = MULTI ? PLv[MULTI ? j : 0][s] : kInf;
        TLi[s] = MULTI ? PLi[MULTI ? j : 0][s] : -1;
      }
      for (int c = 0; c < cnt; c++) {
        const int gl = myq[c * QT];  // chunk-local index
        if (c0 + gl >= n) continue;  // padding can only get here when the threshold is not finite
        const float* pu = reinterpret_cast<const float*>(tgt + 2 * (gl >> 1));
        const int h = gl & 1;
        const float d = sqdist<GA_MODE_CPU_EXACT>(pu[h], pu[2 + h], pu[4 + h], qx[j], qy[j], qz[j]);
        if (d == d) list_insert<KL>(TLv, TLi, d, c0 + gl);  // NaN is never selected beyond position k
      }
      if (MULTI) {
#pragma unroll
        for (int s = 0; s < KL; s++) {
          PLv[MULTI ? j : 0][s] = TLv[s];
          PLi[MULTI ? j : 0][s] = TLi[s];
        }
        if (c0 + CH < n) continue;  // more chunks to merge
      }
      // ---- output / tie detection ----
      const int qi = qtile * QT + j * THREADS + tid;
      bool rp = replay[j];
      float vk = kInf;
#pragma unroll
      for (int s = 0; s < KL; s++) {
        if (s == k - 1) {
          vk = TLv[s];
          rp |= TLi[s] < 0;  // fewer than k finite distances
        }
        if (s + 1 < KL && s < k) rp |= (TLv[s] == TLv[s + 1]) && TLi[s + 1] >= 0;
      }
      if (need_idx && rp) {
        const int slot = atomicAdd(flag_cnt, 1);
        flag_q[slot] = j * THREADS + tid;
        flag_vk[slot] = TLi[0] < 0 ? kInf : vk;
        continue;
      }
      float* vo = a.val + ((size_t)batch * a.m + qi) * kout;
      int* io = need_idx ? a.idx + ((size_t)batch * a.m + qi) * kout : nullptr;
#pragma unroll
      for (int s = 0; s < KL; s++) {
        if (s >= a.skip && s < k) {
          vo[s - a.skip] = a.do_sqrt ? __fsqrt_rn(TLv[s]) : TLv[s];
          if (need_idx) io[s - a.skip] = TLi[s];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < Q; j++) {
    if (valid[j] && overflow[j]) {
      const int slot = atomicAdd(flag_cnt, 1);
      flag_q[slot] = j * THREADS + tid;
      flag_vk[slot] = kNaN;  // v_k unknown: the warp path computes it
    }
  }
  __syncthreads();
  // ---- flagged queries, one warp each: exact v_k if needed, then the replay ---------
  const int nflag = *flag_cnt;
  float* cv = reinterpret_cast<float*>(scratch + (size_t)warp * 3 * KL * 8);
  int* ci = reinterpret_cast<int*>(cv + 3 * KL);
  for (int f = warp; f < nflag; f += THREADS / 32) {
    const int qi = qtile * QT + flag_q[f];
    const float x = __ldg(qpts + (size_t)qi * 3), y = __ldg(qpts + (size_t)qi * 3 + 1),
                z = __ldg(qpts + (size_t)qi * 3 + 2);
    float vk = flag_vk[f];
    if (vk != vk) vk = kth_smallest_warp(pts, n, x, y, z, k, lane);
    selection_replay(pts, n, x, y, z, k, vk, cv, ci, lane);
    write_replayed(cv, ci, k, a.skip, a.do_sqrt, a.val + ((size_t)batch * a.m + qi) * kout,
                   need_idx ? a.idx + ((size_t)batch * a.m + qi) * kout : nullptr, lane);
  }
}

// ---------------------------------------------------------------------------
// knn_mma_kernel: the same search with both filter scans on the tensor cores (k + 1 <= 12, data sets of
// 512..2048 points, large batches: BASELINE config 5).
//
// A CTA of 16 warps stages one data set (pair-SoA + the B fragments of nn_mma.cuh, 96 KB) and serves 1024 queries,
// one 64-query job per warp:
//   scan 1  mma_scan_range<6>: every lane keeps the three smallest tile keys of the 16 tiles it owns per row; the
//           quad's 4 x 3 = 12 keys belong to 12 different tiles, so their maximum tau bounds the 12-th smallest
//           filter value of the row from above (12 >= k + 1 certified points).
//   scan 2  the same MMA loop against thr = tau + W: a 64-bit mask per row of the tiles whose minimum passes
//           (one compare per tile and row instead of the key tracking); 16 shuffles hand every lane the masks of the
//           two queries it owns (rows 2t, 2t+1 of its quad: local queries 16 t + g + 8 j, as in nn_mma.cuh).
//   walk    one query per lane at a time: the ~13-20 qualifying tiles are walked with the fp32 filter, candidates
//           (f <= thr) go to the query's queue in shared memory (predicated stores, the warp stays in step).
//           Lanes walk DIFFERENT tiles (512 B apart: same banks), so each lane starts at its own pair (lane & 15)
//           and the upper half-warp loads the {z,n} half of a pair first: one LDS.128 of the warp then touches
//           every bank once.  The two halves need different operands for the same FMA chain, hence the per-lane
//           coefficient quadruple (cA1, cA2, cB1, cB2) = (ax, ay, az, 1) or (az, 1, ax, ay).
//   drain   the queue is evaluated in the reference arithmetic and inserted into the exact sorted (k+1)-list;
//           ties, NaN and overflow go to the replay exactly as in knn_kernel.
// Bound: a point p of the reference's first k + 1 has d(p) <= max_c d(c) over the 12 certified points c, hence
// (nn_mma.cuh: e0 reference rounding, e1 fp32 filter, e2 MMA filter, 384u key perturbation)
//   h(p)   <= tau + 384u + 2 e2 + 2 e0 = tau + 1074u s^2   (tile mask)
//   f32(p) <= tau + 384u + e2 + 2 e0 + e1 = tau + 763u s^2  (walk)
// both below W = mma_window_wide = 2048u s^2.
// ---------------------------------------------------------------------------
constexpr int kKnnMmaWarps = 16;
constexpr int kKnnMmaThreads = kKnnMmaWarps * 32;
constexpr int kKnnMmaQT = kKnnMmaWarps * kMmaQW;  // 1024 queries per CTA
constexpr int kKnnMmaCH = 2048;
constexpr int kKnnMmaKL = 12;
constexpr int kKnnMmaEnt = 20;  // queue entries (tile, 32-bit candidate mask) per query: ~13 tiles are walked (max ~20), most have a candidate
constexpr size_t kKnnMmaOffRed = (size_t)kKnnMmaCH * 16 + (size_t)kPipeU * 32;
constexpr size_t kKnnMmaOffB = kKnnMmaOffRed + 32 * 4;
constexpr size_t kKnnMmaOffQueue = kKnnMmaOffB + (size_t)kKnnMmaCH * 32;
constexpr size_t kKnnMmaOffQTile = kKnnMmaOffQueue + (size_t)kKnnMmaWarps * 2 * kKnnMmaEnt * 32 * 4;  // masks: [warp][2][ent][32] u32
constexpr size_t kKnnMmaOffFlagQ = kKnnMmaOffQTile + (size_t)kKnnMmaWarps * 2 * kKnnMmaEnt * 32;       // tiles: [warp][2][ent][32] u8
constexpr size_t kKnnMmaOffFlagV = kKnnMmaOffFlagQ + (size_t)kKnnMmaQT * 4;
constexpr size_t kKnnMmaOffCnt = kKnnMmaOffFlagV + (size_t)kKnnMmaQT * 4;
constexpr size_t kKnnMmaOffScratch = kKnnMmaOffCnt + 16;
constexpr size_t kKnnMmaSmem = kKnnMmaOffScratch + (size_t)kKnnMmaWarps * 3 * kKnnMmaKL * 8;
static_assert(kKnnMmaOffB % 16 == 0 && kKnnMmaOffQueue % 16 == 0 && kKnnMmaSmem <= 232448, "shared memory layout");

__global__ void __launch_bounds__(kKnnMmaThreads, 1) knn_mma_kernel(const KnnArgs a) {
  constexpr int KL = kKnnMmaKL, T = kMmaT;
  const float kInf = __int_as_float(0x7f800000);
  const float kNaN = __int_as_float(0x7fc00000);
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem + kKnnMmaOffRed);
  uint4* bfrag = reinterpret_cast<uint4*>(smem + kKnnMmaOffB);
  int* flag_q = reinterpret_cast<int*>(smem + kKnnMmaOffFlagQ);
  float* flag_vk = reinterpret_cast<float*>(smem + kKnnMmaOffFlagV);
  int* flag_cnt = reinterpret_cast<int*>(smem + kKnnMmaOffCnt);
  unsigned char* scratch = smem + kKnnMmaOffScratch;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int batch = blockIdx.x / a.qtiles;
  const int qtile = blockIdx.x - batch * a.qtiles;
  const int n = a.n, k = a.k;
  const float* pts = a.xyz1 + (size_t)batch * n * 3;
  const float* qpts = a.xyz2 + (size_t)batch * a.m * 3;
  const int kout = k - a.skip;
  const bool need_idx = a.idx != nullptr;
  if (tid == 0) *flag_cnt = 0;

  // ---- stage the data set: warp w builds MMA block w (128 points) ----
  const int nblk = (n + kMmaBlk - 1) / kMmaBlk, ntile = (n + T - 1) / T;
  {
    float lmax = 0.0f;
    if (warp < nblk) lmax = stage_block_warp(tgt, bfrag, pts, n, warp, lane);
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
  }
  __syncthreads();
  float bm = 0.0f;
#pragma unroll
  for (int w = 0; w < kKnnMmaWarps; w++) bm = fmaxf(bm, red[w]);

  const int qbase = qtile * kKnnMmaQT + warp * kMmaQW;
  if (qbase < a.m) {
    const int g = lane >> 2, t = lane & 3;
    const unsigned qb = lane & ~3;
    MmaRows R;
    mma_load_rows(R, qpts, a.m, qbase, lane);
    // ---- scan 1: the four smallest tile minima per row among this lane's tiles -> tau per row ----
    float thr[8];
    {
      float c1[8], c2[8], c3[8], c4[8];
#pragma unroll
      for (int r = 0; r < 8; r++) c1[r] = c2[r] = c3[r] = c4[r] = kMmaBig;
      const uint2* bfr = reinterpret_cast<const uint2*>(bfrag);
#pragma unroll 1
      for (int blk = 0; blk < nblk; blk++) {
        float rm[8];
#pragma unroll
        for (int r = 0; r < 8; r++) rm[r] = kMmaBig;
        const uint2* bp = bfr + (size_t)blk * 16 * 32 + lane;
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
          const float v = rm[r];
          c4[r] = fminf(c4[r], fmaxf(c3[r], v));
          c3[r] = fminf(c3[r], fmaxf(c2[r], v));
          c2[r] = fminf(c2[r], fmaxf(c1[r], v));
          c1[r] = fminf(c1[r], v);
        }
      }
      // tau = the 12-th smallest of the quad's 4 x 4 values (12 different tiles hold a point with h <= tau, which
      // covers every k + 1 <= 12): merge with the neighbour lane (bitonic, 4 + 4), then the 12-th of two sorted
      // eights = min over i of max(A[i - 1], B[11 - i]), i = 4..8
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const float b0 = __shfl_xor_sync(0xffffffffu, c1[r], 1), b1 = __shfl_xor_sync(0xffffffffu, c2[r], 1),
                    b2 = __shfl_xor_sync(0xffffffffu, c3[r], 1), b3 = __shfl_xor_sync(0xffffffffu, c4[r], 1);
        const float l0 = fminf(c1[r], b3), l1 = fminf(c2[r], b2), l2 = fminf(c3[r], b1), l3 = fminf(c4[r], b0);
        float h0 = fmaxf(c1[r], b3), h1 = fmaxf(c2[r], b2), h2 = fmaxf(c3[r], b1), h3 = fmaxf(c4[r], b0);
        const float A3 = fmaxf(fmaxf(l0, l1), fmaxf(l2, l3));  // the 4-th smallest of the eight
        // h0..h3 is bitonic: two compare-exchange stages sort it
        float t0 = fminf(h0, h2), t2 = fmaxf(h0, h2), t1 = fminf(h1, h3), t3 = fmaxf(h1, h3);
        const float A4 = fminf(t0, t1), A5 = fmaxf(t0, t1), A6 = fminf(t2, t3), A7 = fmaxf(t2, t3);
        const float B3 = __shfl_xor_sync(0xffffffffu, A3, 2), B4 = __shfl_xor_sync(0xffffffffu, A4, 2),
                    B5 = __shfl_xor_sync(0xffffffffu, A5, 2), B6 = __shfl_xor_sync(0xffffffffu, A6, 2),
                    B7 = __shfl_xor_sync(0xffffffffu, A7, 2);
        float tau = fminf(fmaxf(A3, B7), fmaxf(A4, B6));
        tau = fminf(tau, fmaxf(A5, B5));
        tau = fminf(tau, fminf(fmaxf(A6, B4), fmaxf(A7, B3)));
        thr[r] = tau + mma_window_wide(R.qabs[r], bm);
      }
    }
    // ---- scan 2: tile masks (bit blk of mk[r]: tile 4 blk + t of row r passes) ----
    uint32_t pk[4];
    {
      uint32_t mk[8];
#pragma unroll
      for (int r = 0; r < 8; r++) mk[r] = 0;
      const uint2* bfr = reinterpret_cast<const uint2*>(bfrag);
#pragma unroll 1
      for (int blk = 0; blk < nblk; blk++) {
        float rm[8];
#pragma unroll
        for (int r = 0; r < 8; r++) rm[r] = kMmaBig;
        const uint2* bp = bfr + (size_t)blk * 16 * 32 + lane;
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
        for (int r = 0; r < 8; r++) mk[r] |= !(rm[r] > thr[r]) ? (1u << blk) : 0u;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) pk[i] = mk[2 * i] | (mk[2 * i + 1] << 16);
    }
    // masks of this lane's two queries: w0 = tile columns 0 (low half) and 1 (high half), w1 = columns 2 and 3
    uint32_t w0a = 0, w1a = 0, w0b = 0, w1b = 0;  // a: query j = 0, b: j = 1
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t p0 = __shfl_sync(0xffffffffu, pk[i], qb), p1 = __shfl_sync(0xffffffffu, pk[i], qb + 1),
                     p2 = __shfl_sync(0xffffffffu, pk[i], qb + 2), p3 = __shfl_sync(0xffffffffu, pk[i], qb + 3);
      if (i == t) {
        w0a = (p0 & 0xffffu) | (p1 << 16);
        w1a = (p2 & 0xffffu) | (p3 << 16);
        w0b = (p0 >> 16) | (p1 & 0xffff0000u);
        w1b = (p2 >> 16) | (p3 & 0xffff0000u);
      }
    }
    const float thqa = t == 0 ? thr[0] : (t == 1 ? thr[2] : (t == 2 ? thr[4] : thr[6]));
    const float thqb = t == 0 ? thr[1] : (t == 1 ? thr[3] : (t == 2 ? thr[5] : thr[7]));
    // ---- walk: both queries of the lane in one loop; per tile a 32-bit candidate mask, queued as (tile, mask) ----
    // queue of a query: kKnnMmaEnt entries, entry e of lane l at word e * 32 + l (masks) / byte e * 32 + l (tiles)
    uint32_t* qmask = reinterpret_cast<uint32_t*>(smem + kKnnMmaOffQueue) + (size_t)warp * 2 * kKnnMmaEnt * 32 + lane;
    unsigned char* qtile_ = smem + kKnnMmaOffQTile + (size_t)warp * 2 * kKnnMmaEnt * 32 + lane;
    const int hi = lane >> 4, rot = lane & 15;
    float qx[2], qy[2], qz[2];
    bool valid[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int qi = qbase + 16 * t + g + 8 * j;
      valid[j] = qi < a.m;
      const int qs = valid[j] ? qi : 0;
      qx[j] = __ldg(qpts + (size_t)qs * 3);
      qy[j] = __ldg(qpts + (size_t)qs * 3 + 1);
      qz[j] = __ldg(qpts + (size_t)qs * 3 + 2);
    }
    int ne[2] = {0, 0};  // queued entries per query (kKnnMmaEnt + 1 = overflow)
    {
      // the upper half-warp loads the {z,n} half of a pair first (every LDS.128 of the warp touches every bank once)
      // and runs the same FMA chain on swapped operands: coefficients (cA1, cA2, cB1, cB2) = (ax, ay, az, 1) or
      // (az, 1, ax, ay)
      float2 cA1[2], cA2[2], cB1[2], cB2[2];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const float ax = -2.0f * qx[j], ay = -2.0f * qy[j], az = -2.0f * qz[j];
        cA1[j] = hi ? make_float2(az, az) : make_float2(ax, ax);
        cA2[j] = hi ? make_float2(1.0f, 1.0f) : make_float2(ay, ay);
        cB1[j] = hi ? make_float2(ax, ax) : make_float2(az, az);
        cB2[j] = hi ? make_float2(ay, ay) : make_float2(1.0f, 1.0f);
      }
      uint32_t w0[2] = {valid[0] ? w0a : 0u, valid[1] ? w0b : 0u}, w1[2] = {valid[0] ? w1a : 0u, valid[1] ? w1b : 0u};
      const float thq[2] = {thqa, thqb};
      const float4* tbase = tgt + hi;          // first load of a pair: {x,y} half (lower lanes) or {z,n} half (upper)
      const float4* tbase2 = tgt + (hi ^ 1);
      const int ro = 2 * rot;
      while (__any_sync(0xffffffffu, (w0[0] | w1[0] | w0[1] | w1[1]) != 0u)) {
        int tile[2];
        bool act[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          act[j] = true;
          tile[j] = 0;
          if (w0[j]) {
            const int p = __ffs(w0[j]) - 1;
            w0[j] &= w0[j] - 1;
            tile[j] = 4 * (p & 15) + (p >> 4);
          } else if (w1[j]) {
            const int p = __ffs(w1[j]) - 1;
            w1[j] &= w1[j] - 1;
            tile[j] = 4 * (p & 15) + 2 + (p >> 4);
          } else {
            act[j] = false;
          }
          act[j] = act[j] && tile[j] < ntile && ne[j] <= kKnnMmaEnt;
        }
        const float4* pa0 = tbase + tile[0] * T;
        const float4* pb0 = tbase2 + tile[0] * T;
        const float4* pa1 = tbase + tile[1] * T;
        const float4* pb1 = tbase2 + tile[1] * T;
        uint32_t m0 = 0, m1 = 0;
#pragma unroll
        for (int i = 0; i < T / 2; i++) {
          const int o = (2 * i + ro) & (T - 2);  // pair (i + rot) mod 16, in float4 units
          const float4 A0 = pa0[o], B0 = pb0[o], A1 = pa1[o], B1 = pb1[o];
          float2 f0 = ffma2(cB2[0], make_float2(B0.z, B0.w), make_float2(0.0f, 0.0f));
          float2 f1 = ffma2(cB2[1], make_float2(B1.z, B1.w), make_float2(0.0f, 0.0f));
          f0 = ffma2(cB1[0], make_float2(B0.x, B0.y), f0);
          f1 = ffma2(cB1[1], make_float2(B1.x, B1.y), f1);
          f0 = ffma2(cA2[0], make_float2(A0.z, A0.w), f0);
          f1 = ffma2(cA2[1], make_float2(A1.z, A1.w), f1);
          f0 = ffma2(cA1[0], make_float2(A0.x, A0.y), f0);
          f1 = ffma2(cA1[1], make_float2(A1.x, A1.y), f1);
          // NaN filter value / threshold counts as a candidate
          m0 |= (!(f0.x > thq[0]) ? (1u << (2 * i)) : 0u) | (!(f0.y > thq[0]) ? (2u << (2 * i)) : 0u);
          m1 |= (!(f1.x > thq[1]) ? (1u << (2 * i)) : 0u) | (!(f1.y > thq[1]) ? (2u << (2 * i)) : 0u);
        }
        // walk position i holds pair (i + rot) mod 16: rotate the word back
        m0 = __funnelshift_l(m0, m0, ro);
        m1 = __funnelshift_l(m1, m1, ro);
        if (act[0] && m0) {
          if (ne[0] < kKnnMmaEnt) {
            qmask[ne[0] * 32] = m0;
            qtile_[ne[0] * 32] = (unsigned char)tile[0];
          }
          ne[0]++;
        }
        if (act[1] && m1) {
          if (ne[1] < kKnnMmaEnt) {
            qmask[(kKnnMmaEnt + ne[1]) * 32] = m1;
            qtile_[(kKnnMmaEnt + ne[1]) * 32] = (unsigned char)tile[1];
          }
          ne[1]++;
        }
      }
    }
    // ---- drain, one query at a time: exact distances of the queued candidates -> sorted (k+1)-list ----
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
      if (!valid[j]) continue;
      const int lq = 16 * t + g + 8 * j;
      const int qi = qbase + lq;
      if (ne[j] > kKnnMmaEnt) {  // more tiles with candidates than the queue holds: exact warp path below
        const int slot = atomicAdd(flag_cnt, 1);
        flag_q[slot] = warp * kMmaQW + lq;
        flag_vk[slot] = kNaN;  // v_k unknown: the warp path computes it
        continue;
      }
      // A NaN distance at a position < k is "selected" by the reference: only the replay reproduces that.
      bool rp = false;
      if (need_idx)
        for (int i = 0; i < k; i++) {
          const float d = sqdist<GA_MODE_CPU_EXACT>(__ldg(pts + (size_t)i * 3), __ldg(pts + (size_t)i * 3 + 1),
                                                    __ldg(pts + (size_t)i * 3 + 2), qx[j], qy[j], qz[j]);
          rp |= d != d;
        }
      float TLv[KL];
      int TLi[KL];
#pragma unroll
      for (int s = 0; s < KL; s++) {
        TLv[s] = kInf;
        TLi[s] = -1;
      }
      const uint32_t* em = qmask + j * kKnnMmaEnt * 32;
      const unsigned char* et = qtile_ + j * kKnnMmaEnt * 32;
      for (int e = 0; e < ne[j]; e++) {
        uint32_t m = em[e * 32];
        const int gb = (int)et[e * 32] * T;
        while (m) {
          const int gl = gb + __ffs(m) - 1;
          m &= m - 1;
          if (gl >= n) continue;  // padding can only get here when the threshold is not finite
          const float* pu = reinterpret_cast<const float*>(tgt + 2 * (gl >> 1));
          const int h = gl & 1;
          const float d = sqdist<GA_MODE_CPU_EXACT>(pu[h], pu[2 + h], pu[4 + h], qx[j], qy[j], qz[j]);
          if (d == d) list_insert<KL>(TLv, TLi, d, gl);  // NaN is never selected beyond position k
        }
      }
      // ---- output / tie detection (as knn_kernel) ----
      float vk = kInf;
#pragma unroll
      for (int s = 0; s < KL; s++) {
        if (s == k - 1) {
          vk = TLv[s];
          rp |= TLi[s] < 0;  // fewer than k finite distances
        }
        if (s + 1 < KL && s < k) rp |= (TLv[s] == TLv[s + 1]) && TLi[s + 1] >= 0;
      }
      if (need_idx && rp) {
        const int slot = atomicAdd(flag_cnt, 1);
        flag_q[slot] = warp * kMmaQW + lq;
        flag_vk[slot] = TLi[0] < 0 ? kInf : vk;
        continue;
      }
      float* vo = a.val + ((size_t)batch * a.m + qi) * kout;
      int* io = need_idx ? a.idx + ((size_t)batch * a.m + qi) * kout : nullptr;
#pragma unroll
      for (int s = 0; s < KL; s++) {
        if (s >= a.skip && s < k) {
          vo[s - a.skip] = a.do_sqrt ? __fsqrt_rn(TLv[s]) : TLv[s];
          if (need_idx) io[s - a.skip] = TLi[s];
        }
      }
    }
  }
  __syncthreads();
  // ---- flagged queries, one warp each: exact v_k if needed, then the replay ----
  const int nflag = *flag_cnt;
  float* cv = reinterpret_cast<float*>(scratch + (size_t)warp * 3 * KL * 8);
  int* ci = reinterpret_cast<int*>(cv + 3 * KL);
  for (int f = warp; f < nflag; f += kKnnMmaWarps) {
    const int qi = qtile * kKnnMmaQT + flag_q[f];
    const float x = __ldg(qpts + (size_t)qi * 3), y = __ldg(qpts + (size_t)qi * 3 + 1),
                z = __ldg(qpts + (size_t)qi * 3 + 2);
    float vk = flag_vk[f];
    if (vk != vk) vk = kth_smallest_warp(pts, n, x, y, z, k, lane);
    selection_replay(pts, n, x, y, z, k, vk, cv, ci, lane);
    write_replayed(cv, ci, k, a.skip, a.do_sqrt, a.val + ((size_t)batch * a.m + qi) * kout,
                   need_idx ? a.idx + ((size_t)batch * a.m + qi) * kout : nullptr, lane);
  }
}

static int launch_knn_mma(const KnnArgs& a0, cudaStream_t st) {
  KnnArgs a = a0;
  a.qtiles = (a.m + kKnnMmaQT - 1) / kKnnMmaQT;
  const long long ctas = (long long)a.b * a.qtiles;
  if (ctas > 0x7fffffffLL) {
    set_error("ga_knn: problem too large for one launch");
    return GA_ERR_UNSUPPORTED;
  }
  GA_CUDA_TRY(cudaFuncSetAttribute(knn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kKnnMmaSmem));
  knn_mma_kernel<<<(unsigned)ctas, kKnnMmaThreads, kKnnMmaSmem, st>>>(a);
  GA_LAUNCH_CHECK("knn_mma_kernel");
  return GA_OK;
}

// ---------------------------------------------------------------------------
// Generic k (k > 32): one warp per query, exact v_k then the replay.  O(k n) per
// query; rare path (tf_grouping.py's own self-test uses k = 64).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) knn_generic_kernel(const KnnArgs a) {
  extern __shared__ float4 smem_f4[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long q = (long long)blockIdx.x * 4 + warp;
  if (q >= (long long)a.b * a.m) return;
  const int batch = (int)(q / a.m), qi = (int)(q - (long long)batch * a.m);
  const int n = a.n, k = a.k;
  const float* pts = a.xyz1 + (size_t)batch * n * 3;
  const float* qp = a.xyz2 + ((size_t)batch * a.m + qi) * 3;
  const float qx = __ldg(qp), qy = __ldg(qp + 1), qz = __ldg(qp + 2);
  const float vk = kth_smallest_warp(pts, n, qx, qy, qz, k, lane);
  float* cv = reinterpret_cast<float*>(smem_f4) + (size_t)warp * 6 * k;
  int* ci = reinterpret_cast<int*>(cv + 3 * k);
  const int kout = k - a.skip;
  selection_replay(pts, n, qx, qy, qz, k, vk, cv, ci, lane);
  write_replayed(cv, ci, k, a.skip, a.do_sqrt, a.val + ((size_t)batch * a.m + qi) * kout,
                 a.idx ? a.idx + ((size_t)batch * a.m + qi) * kout : nullptr, lane);
}

// ---------------------------------------------------------------------------
// Legacy dense selection sort, one warp per row (tf_grouping_g.cu:83-123).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) selection_sort_kernel(long long rows, int n, int k,
                                                            const float* __restrict__ dist,
                                                            int* __restrict__ outi, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* src = dist + (size_t)row * n;
  float* v = out + (size_t)row * n;
  int* ix = outi + (size_t)row * n;
  for (int s = lane; s < n; s += 32) {
    v[s] = src[s];
    ix[s] = s;
  }
  __syncwarp();
  const float kInf = __int_as_float(0x7f800000);
  for (int s = 0; s < k && s < n; s++) {
    const float head = v[s];
    float bv = kInf;
    int bp = 0x7fffffff;
    for (int t = s + lane; t < n; t += 32) {
      const float x = v[t];
      if (x < bv) {
        bv = x;
        bp = t;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov < bv || (ov == bv && op < bp)) {
        bv = ov;
        bp = op;
      }
    }
    // "min = s; if (p[t] < p[min]) min = t": a NaN head is never displaced, an
    // all-inf/NaN tail leaves min = s.
    int mn = (head != head || bp == 0x7fffffff) ? s : bp;
    if (mn != s && lane == 0) {
      const float tv = v[mn];
      v[mn] = v[s];
      v[s] = tv;
      const int ti = ix[mn];
      ix[mn] = ix[s];
      ix[s] = ti;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) group_point_kernel(long long total, int n, int c, int m, int nsample,
                                                         const float* __restrict__ points,
                                                         const int* __restrict__ idx, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= total) return;
  const int l = (int)(e % c);
  const long long r = e / c;  // (batch, j, s) flattened
  const long long batch = r / ((long long)m * nsample);
  const int ii = __ldg(idx + r);
  out[e] = __ldg(points + ((size_t)batch * n + ii) * c + l);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads) knn_kernel(const KnnArgs a) {
  knn_body<Cfg>(a);
}

// ---- knn_slab_kernel: the defense epilogue's kNN distances (ga_knn_dists) with the scans pruned to an x-slab -------
//
// ga_knn_dists returns VALUES only (sqrt of the k+1 smallest squared distances of every point to its own cloud, the
// first dropped: get_knn_dists_per_point.py:78-81), so neither the order of equal distances nor the identity of the
// neighbours matters, and the cloud may be reordered.  Per CTA (= cloud, 512 queries) the cloud is counting-sorted
// by x into 1024 bins (every CTA of a cloud sorts it again: ~7 % of a CTA), the points of a bin ordered by index so
// that the CTAs of a cloud agree on the sorted positions they split the queries by, and staged in that order (the raw
// cloud arrives by one TMA bulk copy);
// a warp owns 64 queries that are neighbours in the sorted order, i.e. a thin slab [xa, xb] of the cloud.
//   A0  filter scan (nn_tiles.cuh) over the warp's own positions +- kSlabSeed targets (the same count at the ends of
//       the cloud, where one side is cut off and a half ball reaches further): the K-th smallest minimum
//       over tiles of 16 targets is tau, K distinct targets with filter value <= tau.  Hence the K-th smallest
//       exact distance of query j is at most U_j = tau_j + |q_j|^2 + error terms.
//   r   r^2 >= max_j U_j over the warp.  A target whose x lies outside [xa - r, xb + r] (directed rounding) has
//       |dx| > r to every query of the warp, so its reference distance fl(dx^2 + ..) >= fl(r^2) > U_j: it is not among
//       the K smallest of any of them.  The bin of a coordinate is a monotonic function of it, so the targets inside
//       the window are a contiguous range of the sorted order, found from the bin starts.
//   B1  second scan over that range only: targets with f <= tau + W (the window of knn_kernel) go to the query's queue;
//   B2  queued targets in the reference arithmetic into an exact sorted list of K values, written as sqrt.
// On uniform clouds of 2048 points A0 sees 25 % and B1 ~34 % of the cloud where knn_kernel scans all of it twice
// (config 5: 0.48 ms against 0.97; DESIGN.md 4.3 has the measurements and the two profiler findings on the way).
// A query whose queue overflows (dense clusters) is served by the whole warp with K selection passes over the
// cloud.  A cloud with a non-finite or huge coordinate, or fewer than 512 points, is handed to knn_body unchanged
// (NaN distances take part in the reference's selection sort in a way only its replay reproduces).
constexpr int kSlabBins = 1024;
constexpr int kSlabSeed = 224;   // targets on either side of a warp's own positions in the first scan
constexpr int kSlabQueue = 40;   // candidate slots per query
constexpr int kSlabT = 16;
constexpr int kSlabMaxBin = 32;  // fullest bin the slab path takes (its points get an insertion sort by index)

using SlabBase = KnnCfg<256, 2, 32, 2048, 12, false>;
struct SlabCfg {
  static constexpr int kThreads = 256, kQ = 2, kQT = 512, kCH = 2048, kKL = 12;
  // tgt (+ pipeline pad) | hist/starts int[1026] | qidx u16[2048] | red[64] | queue u16[(40 + 16)][512]
  static constexpr size_t kOffHist = (size_t)kCH * 16 + (size_t)kPipeU * 32;
  static constexpr size_t kOffQidx = kOffHist + 1032 * 4;
  static constexpr size_t kOffRed = kOffQidx + (size_t)kCH * 2;
  static constexpr size_t kOffQueue = kOffRed + 64 * 4;
  static constexpr size_t kSmemSlab = kOffQueue + (size_t)(kSlabQueue + kSlabT) * kQT * 2;
  static constexpr size_t kSmem = kSmemSlab > SlabBase::kSmem ? kSmemSlab : SlabBase::kSmem;
};

// A query whose queue overflowed, served by the whole warp from the staged slab [w0, w1) (every one of its K nearest
// lies there): each lane filters a strided share of the slab with the query's threshold and keeps the exact distances
// of the survivors (a handful per warp) in registers; K rounds of a warp minimum then peel off the K smallest values
// (duplicates one at a time).  More than kSlowSlots survivors in one lane (clouds of coincident points): K selection
// passes over the slab in (value, position) order instead.  sqrt of ranks skip..K-1 to `row`.
constexpr int kSlowSlots = 8;
__device__ void slab_slow_query(const float4* __restrict__ tgt, int w0, int w1, int n, float qx, float qy, float qz,
                                float ax2, float ay2, float az2, float thr, int K, int skip, float* __restrict__ row,
                                int lane) {
  const float kInf = __int_as_float(0x7f800000);
  float c[kSlowSlots];
#pragma unroll
  for (int s = 0; s < kSlowSlots; s++) c[s] = kInf;
  int nc = 0;
  for (int p = (w0 >> 1) + lane; p < (w1 >> 1); p += 32) {
    const float4 u = tgt[2 * p], v = tgt[2 * p + 1];
    const float2 f = filter_pair(u, v, ax2, ay2, az2);
    if (!(f.x > thr) && 2 * p < n) {
      const float d = sqdist<GA_MODE_CPU_EXACT>(u.x, u.z, v.x, qx, qy, qz);
#pragma unroll
      for (int s = 0; s < kSlowSlots; s++) c[s] = s == nc ? d : c[s];
      nc++;
    }
    if (!(f.y > thr) && 2 * p + 1 < n) {
      const float d = sqdist<GA_MODE_CPU_EXACT>(u.y, u.w, v.y, qx, qy, qz);
#pragma unroll
      for (int s = 0; s < kSlowSlots; s++) c[s] = s == nc ? d : c[s];
      nc++;
    }
  }
  if (__any_sync(0xffffffffu, nc > kSlowSlots)) {
    const float* tf = reinterpret_cast<const float*>(tgt);
    const int hi = min(w1, n);
    float pv = -kInf;
    int pp = -1;
    for (int s = 0; s < K; s++) {
      float bv = kInf;
      int bp = 0x7fffffff;
      for (int i = w0 + lane; i < hi; i += 32) {
        const float* pu = tf + 8 * (i >> 1) + (i & 1);
        const float d = sqdist<GA_MODE_CPU_EXACT>(pu[0], pu[2], pu[4], qx, qy, qz);
        const bool after = d > pv || (d == pv && i > pp);
        if (after && (d < bv || (d == bv && i < bp))) {
          bv = d;
          bp = i;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (ov < bv || (ov == bv && op < bp)) {
          bv = ov;
          bp = op;
        }
      }
      pv = bv;
      pp = bp;
      if (lane == 0 && s >= skip) row[s - skip] = __fsqrt_rn(pv);
    }
    return;
  }
  float mine = kInf;  // lane s keeps the s-th smallest (K <= 32)
  for (int s = 0; s < K; s++) {
    float m = c[0];
#pragma unroll
    for (int e = 1; e < kSlowSlots; e++) m = fminf(m, c[e]);
    float wm = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wm = fminf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
    const unsigned holders = __ballot_sync(0xffffffffu, m == wm);
    if (lane == __ffs(holders) - 1) {  // one copy leaves
      bool done = false;
#pragma unroll
      for (int e = 0; e < kSlowSlots; e++) {
        const bool hit = !done && c[e] == wm;
        c[e] = hit ? kInf : c[e];
        done |= hit;
      }
    }
    if (lane == s) mine = wm;
  }
  if (lane >= skip && lane < K) row[lane - skip] = __fsqrt_rn(mine);  // one store for the row
}

#ifdef GA_SLAB_DEBUG
__device__ unsigned long long g_slab_dbg[16];
#define SLAB_DBG(i, v) atomicAdd(&g_slab_dbg[i], (unsigned long long)(v))
#else
#define SLAB_DBG(i, v)
#endif
__global__ void __launch_bounds__(256, 2) knn_slab_kernel(const KnnArgs a) {
  constexpr int THREADS = SlabCfg::kThreads, QT = SlabCfg::kQT, KL = SlabCfg::kKL, T = kSlabT, CH = SlabCfg::kCH;
  constexpr int PPT = CH / THREADS;  // points per thread in the sort
  const float kInf = __int_as_float(0x7f800000);
  extern __shared__ float4 smem_f4[];
  unsigned char* smem_raw = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  int* hist = reinterpret_cast<int*>(smem_raw + SlabCfg::kOffHist);
  unsigned short* qidx = reinterpret_cast<unsigned short*>(smem_raw + SlabCfg::kOffQidx);
  float* red = reinterpret_cast<float*>(smem_raw + SlabCfg::kOffRed);
  unsigned short* queue = reinterpret_cast<unsigned short*>(smem_raw + SlabCfg::kOffQueue);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int batch = blockIdx.x / a.qtiles;
  const int qtile = blockIdx.x - batch * a.qtiles;
  const int n = a.n, K = a.k;
  const float* pts = a.xyz1 + (size_t)batch * n * 3;
#ifdef GA_SLAB_DEBUG
  long long tk0 = clock64();
#define SLAB_CLK(i) do { if (tid == 0) { const long long t_ = clock64(); SLAB_DBG(i, t_ - tk0); tk0 = t_; } } while (0)
#else
#define SLAB_CLK(i)
#endif

  // ---- the cloud: one bulk copy (TMA, cp.async.bulk global -> shared, completion on an mbarrier) into the queue
  // region, which is not needed before the second scan; clouds that are not whole 16-byte units fall back to loads
  __shared__ __align__(8) unsigned long long raw_bar;
  const float* raw = reinterpret_cast<const float*>(queue);
  const unsigned raw_bytes = (unsigned)n * 12u;
  const bool bulk = (raw_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(pts) & 15) == 0;  // uniform over the CTA
  if (bulk) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&raw_bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(raw_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(queue)),
                   "l"(pts), "r"(raw_bytes), "r"(bar)
                   : "memory");
    }
    __syncthreads();  // the barrier is initialised for everybody
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; spin++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(bar) : "memory");
    if (!done) __trap();  // a copy that never lands: fail loudly instead of hanging
  }
  // ---- coordinates to registers, x range, max |coordinate|, finiteness ----
  float px[PPT], py[PPT], pz[PPT];
  float xmin = kInf, xmax = -kInf, amax = 0.0f;
  bool bad = false;
#pragma unroll
  for (int e = 0; e < PPT; e++) {
    const int i = tid + e * THREADS;
    const bool in = i < n;
    if (bulk) {
      px[e] = in ? raw[3 * i] : 0.0f;
      py[e] = in ? raw[3 * i + 1] : 0.0f;
      pz[e] = in ? raw[3 * i + 2] : 0.0f;
    } else {
      px[e] = in ? __ldg(pts + (size_t)i * 3) : 0.0f;
      py[e] = in ? __ldg(pts + (size_t)i * 3 + 1) : 0.0f;
      pz[e] = in ? __ldg(pts + (size_t)i * 3 + 2) : 0.0f;
    }
    if (in) {
      xmin = fminf(xmin, px[e]);
      xmax = fmaxf(xmax, px[e]);
      const float m = fmaxf(fmaxf(fabsf(px[e]), fabsf(py[e])), fabsf(pz[e]));
      amax = fmaxf(amax, m);
      // NaN, inf or a distance that could overflow (fmaxf drops NaNs: test every coordinate)
      bad |= !(fabsf(px[e]) < 1.0e15f) || !(fabsf(py[e]) < 1.0e15f) || !(fabsf(pz[e]) < 1.0e15f);
    }
  }
  for (int i = tid; i < 1032; i += THREADS) hist[i] = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    red[warp] = xmin;
    red[8 + warp] = xmax;
    red[16 + warp] = amax;
    red[24 + warp] = bad ? 1.0f : 0.0f;
  }
  __syncthreads();
  float fbad = 0.0f;
#pragma unroll
  for (int w = 0; w < THREADS / 32; w++) {
    xmin = fminf(xmin, red[w]);
    xmax = fmaxf(xmax, red[8 + w]);
    amax = fmaxf(amax, red[16 + w]);
    fbad = fmaxf(fbad, red[24 + w]);
  }
  if (fbad != 0.0f || !(amax > 1.0e-15f)) {  // uniform over the CTAs of a cloud: every one of them read all of it
    __syncthreads();
    knn_body<SlabBase>(a);
    return;
  }

  // ---- counting sort by x bin (monotonic in x), staged as pair-SoA in sorted order ----
  const float ext = xmax - xmin;
  const float inv = ext > 0.0f ? (float)kSlabBins / ext : 0.0f;
  auto bin_of = [&](float x) {
    const float t = (x - xmin) * inv;
    int bi = t > 0.0f ? (t < (float)(kSlabBins - 1) ? (int)t : kSlabBins - 1) : 0;
    return bi;
  };
  int rank[PPT], bins[PPT];
#pragma unroll
  for (int e = 0; e < PPT; e++) {
    const int i = tid + e * THREADS;
    bins[e] = bin_of(px[e]);
    rank[e] = i < n ? atomicAdd(&hist[bins[e]], 1) : 0;
  }
  __syncthreads();
  {  // exclusive scan of the 1024 counts: 4 bins per thread, warp scan, warp totals
    const int c0 = hist[4 * tid], c1 = hist[4 * tid + 1], c2 = hist[4 * tid + 2], c3 = hist[4 * tid + 3];
    const int tot = c0 + c1 + c2 + c3;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    int cmax = __reduce_max_sync(0xffffffffu, max(max(c0, c1), max(c2, c3)));
    int* wsum = reinterpret_cast<int*>(red + 32);
    if (lane == 31) wsum[warp] = inc;
    if (lane == 0) wsum[8 + warp] = cmax;
    __syncthreads();
    int base = inc - tot;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
      if (w < warp) base += wsum[w];
      cmax = max(cmax, wsum[8 + w]);
    }
    if (cmax > kSlabMaxBin) {  // many points with (nearly) the same x: no slab structure to use.  Uniform per cloud.
      __syncthreads();
      knn_body<SlabBase>(a);
      return;
    }
    hist[4 * tid] = base;
    hist[4 * tid + 1] = base + c0;
    hist[4 * tid + 2] = base + c0 + c1;
    hist[4 * tid + 3] = base + c0 + c1 + c2;
    if (tid == THREADS - 1) hist[kSlabBins] = base + tot;  // = n
  }
  __syncthreads();
  float* tf = reinterpret_cast<float*>(tgt);
#pragma unroll
  for (int e = 0; e < PPT; e++) {
    const int i = tid + e * THREADS;
    if (i < n) {
      const int pos = hist[bins[e]] + rank[e];
      float* pu = tf + 8 * (pos >> 1) + (pos & 1);
      pu[0] = px[e];
      pu[2] = py[e];
      pu[4] = pz[e];
      pu[6] = fmaf(pz[e], pz[e], fmaf(py[e], py[e], px[e] * px[e]));
      qidx[pos] = (unsigned short)i;
    }
  }
  const int npad = (n + T - 1) / T * T;
  for (int pos = n + tid; pos < npad + 2 * kPipeU; pos += THREADS) {  // padding: far away, never a candidate
    float* pu = tf + 8 * (pos >> 1) + (pos & 1);
    pu[0] = 0.0f;
    pu[2] = 0.0f;
    pu[4] = 0.0f;
    pu[6] = kInf;
  }
  __syncthreads();
  // The atomics above order the points of a bin arbitrarily, differently in every CTA of the cloud; the CTAs split
  // the queries by sorted POSITION, so the order has to be the same everywhere: points of a bin by original index
  // (insertion sort of a handful of entries; a thread owns four bins, the words of an entry are its own).
  for (int bb = 0; bb < 4; bb++) {
    const int s0 = hist[4 * tid + bb], s1 = hist[4 * tid + bb + 1];
    for (int i = s0 + 1; i < s1; i++) {
      const unsigned short qi = qidx[i];
      const float* pi = tf + 8 * (i >> 1) + (i & 1);
      const float ex = pi[0], ey = pi[2], ez = pi[4], en = pi[6];
      int j = i - 1;
      while (j >= s0 && qidx[j] > qi) {
        const float* pj = tf + 8 * (j >> 1) + (j & 1);
        float* pd = tf + 8 * ((j + 1) >> 1) + ((j + 1) & 1);
        pd[0] = pj[0];
        pd[2] = pj[2];
        pd[4] = pj[4];
        pd[6] = pj[6];
        qidx[j + 1] = qidx[j];
        j--;
      }
      float* pd = tf + 8 * ((j + 1) >> 1) + ((j + 1) & 1);
      pd[0] = ex;
      pd[2] = ey;
      pd[4] = ez;
      pd[6] = en;
      qidx[j + 1] = qi;
    }
  }
  __syncthreads();

  // ---- the warp's 64 queries: sorted positions base .. base + 63 (no barrier below) ----
  const int base = qtile * QT + warp * 64;
  if (base >= n) return;
#ifdef GA_SLAB_DEBUG
  const long long tw0 = clock64();
  long long tslow0 = 0;
  int nslow = 0;
#endif
  {
  float qx[2], qy[2], qz[2], qabs[2], ax2[2], ay2[2], az2[2], qn[2];
  bool valid[2];
  float xa = kInf, xb = -kInf;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int pos = base + 32 * j + lane;
    valid[j] = pos < n;
    const int ps = valid[j] ? pos : base;
    const float* pu = tf + 8 * (ps >> 1) + (ps & 1);
    qx[j] = pu[0];
    qy[j] = pu[2];
    qz[j] = pu[4];
    qn[j] = pu[6];
    qabs[j] = fmaxf(fmaxf(fabsf(qx[j]), fabsf(qy[j])), fabsf(qz[j]));
    ax2[j] = -2.0f * qx[j];
    ay2[j] = -2.0f * qy[j];
    az2[j] = -2.0f * qz[j];
    xa = fminf(xa, qx[j]);
    xb = fmaxf(xb, qx[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xa = fminf(xa, __shfl_xor_sync(0xffffffffu, xa, o));
    xb = fmaxf(xb, __shfl_xor_sync(0xffffffffu, xb, o));
  }

  // ---- A0: K-th smallest tile minimum around the warp's own positions ----
  float tau[2];
  {
    // the same number of targets at the ends of the cloud, where one side is cut off (a half ball reaches further:
    // without this the end warps' tau is loose and every one of their queues overflows)
    int t0 = max(0, base - kSlabSeed) & ~(T - 1);
    int t1 = min(npad, (base + 64 + kSlabSeed + T - 1) & ~(T - 1));
    if (t0 == 0) t1 = min(npad, max(t1, 64 + 2 * kSlabSeed));
    if (t1 == npad) t0 = max(0, min(t0, npad - (64 + 2 * kSlabSeed)));
    float S[2][KL];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int s = 0; s < KL; s++) S[j][s] = kInf;
    filter_scan<2, T>(tgt + t0, (t1 - t0) / T, ax2, ay2, az2, [&](int, const float(&tm)[2]) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        float v = tm[j];
#pragma unroll
        for (int s = 0; s < KL; s++) {  // sorted insert by compare-exchange chain
          const float lo = fminf(S[j][s], v);
          v = fmaxf(S[j][s], v);
          S[j][s] = lo;
        }
      }
    });
#pragma unroll
    for (int j = 0; j < 2; j++) {
      float tk = kInf;
#pragma unroll
      for (int s = 0; s < KL; s++)
        if (s == K - 1) tk = S[j][s];
      tau[j] = tk;
    }
  }

  SLAB_CLK(9);
  // ---- the slab that can hold a K-th neighbour of any of the warp's queries ----
  int w0 = 0, w1 = npad;
  {
    float U = 0.0f;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const float s = (qabs[j] + amax) * 1.0001f;
      const float uj = fmaf(fmaxf(tau[j] + qn[j], 0.0f), 1.0009765625f, s * s * 7.62939453125e-06f /* 2^-17 */);
      U = fmaxf(U, valid[j] ? uj : 0.0f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) U = fmaxf(U, __shfl_xor_sync(0xffffffffu, U, o));
    const float r = __fmul_ru(__fsqrt_ru(U), 1.000001f);
    if (__fmul_rn(r, r) > U && r < kInf) {  // otherwise (tau = +inf: fewer than K tiles around) keep the whole cloud
      const float lo = __fsub_rd(xa, r), hi = __fadd_ru(xb, r);
      const int b0 = bin_of(lo), b1 = bin_of(hi);
      w0 = hist[b0] & ~(T - 1);
      w1 = min(npad, (hist[b1 + 1] + T - 1) & ~(T - 1));
    }
  }

  // ---- B1: candidates of the slab into the queues ----
  float thr[2];
  int cnts[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    thr[j] = valid[j] ? tau[j] + filter_window(qabs[j], amax) : -kInf;
    cnts[j] = 0;
  }
  filter_collect<2, T>(tgt + w0, (w1 - w0) / T, ax2, ay2, az2, thr, cnts, queue + tid, QT, THREADS, kSlabQueue);

  SLAB_CLK(10);
#ifdef GA_SLAB_DEBUG
  if (lane == 0) {
    SLAB_DBG(0, 1);
    SLAB_DBG(1, w1 - w0);
    SLAB_DBG(2, w1 - w0 == npad);
  }
#endif
  // ---- B2: exact list of the K smallest values ----
  const int kout = K - a.skip;
  bool over[2];
  {
    // values only: a branch-free compare-exchange chain keeps the KL smallest (equal values need no order); the
    // lane's two queries are drained in one loop, two independent chains in flight
    float TLv[2][KL];
    int take[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      over[j] = valid[j] && cnts[j] > kSlabQueue;
      take[j] = (valid[j] && !over[j]) ? cnts[j] : 0;
#pragma unroll
      for (int s = 0; s < KL; s++) TLv[j][s] = kInf;
    }
    const int cmaxq = max(take[0], take[1]);
    for (int c = 0; c < cmaxq; c++) {
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const bool on = c < take[j];
        const int gl = on ? w0 + queue[c * QT + j * THREADS + tid] : 0;
        const float* pu = tf + 8 * (gl >> 1) + (gl & 1);
        float v = sqdist<GA_MODE_CPU_EXACT>(pu[0], pu[2], pu[4], qx[j], qy[j], qz[j]);
        v = (on && gl < n) ? v : kInf;  // padding is only reachable with a non-finite threshold
#pragma unroll
        for (int s = 0; s < KL; s++) {
          const float lo = fminf(TLv[j][s], v);
          v = fmaxf(TLv[j][s], v);
          TLv[j][s] = lo;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (!valid[j] || over[j]) continue;
      bool shortlist = false;
#pragma unroll
      for (int s = 0; s < KL; s++)
        if (s == K - 1) shortlist = !(TLv[j][s] < kInf);
      if (shortlist) {  // cannot happen for finite data (K targets with f <= tau lie in the slab); be safe
        SLAB_DBG(7, 1);
        over[j] = true;
        continue;
      }
      float* vo = a.val + ((size_t)batch * n + qidx[base + 32 * j + lane]) * kout;
#pragma unroll
      for (int s = 0; s < KL; s++)
        if (s >= a.skip && s < K) vo[s - a.skip] = a.do_sqrt ? __fsqrt_rn(TLv[j][s]) : TLv[j][s];
    }
  }
  SLAB_CLK(11);
#ifdef GA_SLAB_DEBUG
  tslow0 = clock64();
#endif
  // ---- overflowed queues (dense clusters): the whole warp serves the query ----
#pragma unroll
  for (int j = 0; j < 2; j++) {
    unsigned pending = __ballot_sync(0xffffffffu, over[j]);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
#ifdef GA_SLAB_DEBUG
      nslow++;
#endif
      const float bx = __shfl_sync(0xffffffffu, qx[j], src), by = __shfl_sync(0xffffffffu, qy[j], src),
                  bz = __shfl_sync(0xffffffffu, qz[j], src);
      // the threshold filter_collect started with (it overwrites thr[] of an overflowed query)
      const float bthr = __shfl_sync(0xffffffffu, tau[j] + filter_window(qabs[j], amax), src);
      slab_slow_query(tgt, w0, w1, n, bx, by, bz, -2.0f * bx, -2.0f * by, -2.0f * bz, bthr, K, a.skip,
                      a.val + ((size_t)batch * n + qidx[base + 32 * j + src]) * kout, lane);
    }
  }
  SLAB_CLK(12);
  }
#ifdef GA_SLAB_DEBUG
  if (lane == 0) {
    const long long t1 = clock64();
    atomicAdd(&g_slab_dbg[13], (unsigned long long)(t1 - tw0));
    atomicMax(&g_slab_dbg[14], (unsigned long long)(t1 - tw0));
    atomicMax(&g_slab_dbg[15], (unsigned long long)(t1 - tslow0));
    atomicMax(&g_slab_dbg[6], (unsigned long long)nslow);
  }
#endif
}

#ifdef GA_SLAB_DEBUG
extern "C" int ga_debug_slab_stats(unsigned long long* out8) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_slab_dbg, 128);
  unsigned long long z[16] = {};
  cudaMemcpyToSymbol(g_slab_dbg, z, 128);
  return 0;
}
#endif
int g_knn_slab = 1;  // tuning hook (key 28): 0 = ga_knn_dists never takes knn_slab_kernel, 1 = from half a wave of CTAs, 2 = always

static int launch_knn_slab(const KnnArgs& a0, cudaStream_t st) {
  KnnArgs a = a0;
  a.qtiles = (a.m + SlabCfg::kQT - 1) / SlabCfg::kQT;
  const long long ctas = (long long)a.b * a.qtiles;
  if (ctas > 0x7fffffffLL) {
    set_error("ga_knn: problem too large for one launch");
    return GA_ERR_UNSUPPORTED;
  }
  GA_CUDA_TRY(cudaFuncSetAttribute(knn_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SlabCfg::kSmem));
  knn_slab_kernel<<<(unsigned)ctas, SlabCfg::kThreads, SlabCfg::kSmem, st>>>(a);
  GA_LAUNCH_CHECK("knn_slab_kernel");
  return GA_OK;
}

template <class Cfg>
static int launch_knn(const KnnArgs& a0, cudaStream_t st) {
  KnnArgs a = a0;
  a.qtiles = (a.m + Cfg::kQT - 1) / Cfg::kQT;
  const long long ctas = (long long)a.b * a.qtiles;
  if (ctas > 0x7fffffffLL) {
    set_error("ga_knn: problem too large for one launch");
    return GA_ERR_UNSUPPORTED;
  }
  GA_CUDA_TRY(cudaFuncSetAttribute(knn_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
  knn_kernel<Cfg><<<(unsigned)ctas, Cfg::kThreads, Cfg::kSmem, st>>>(a);
  GA_LAUNCH_CHECK("knn_kernel");
  return GA_OK;
}

int g_knn_variant = 0;  // tuning hook

static int knn_dispatch(const KnnArgs& a, cudaStream_t st) {
  const int kl = a.k + 1;
  if (kl <= 33) {
    if (a.n <= 2048) {  // single chunk: lists live only while a query slot is drained
      if (kl <= 12) {
        int variant = g_knn_variant;
        // values only, self-kNN, sqrt epilogue (ga_knn_dists): the slab-pruned kernel; knn_point proper needs the
        // reference's tie order and keeps the full scans
        // (measured, 2048-point clouds, ms slab / full scans: B=500 0.48 / 0.97, 250 0.26 / 0.50, 125 0.17 / 0.28,
        // 50 0.105 / 0.142; from half a wave of its 512-query CTAs on; key 28 = 2 takes it at any size)
        if (variant == 0 && g_knn_slab && a.idx == nullptr && a.xyz1 == a.xyz2 && a.n == a.m && a.n >= 512 &&
            a.do_sqrt && a.skip <= 1 && a.k >= 2 &&
            (g_knn_slab >= 2 || (long long)a.b * ((a.m + 511) / 512) >= sm_count() / 2))
          return launch_knn_slab(a, st);
        // tensor-core scans (opt-in, variant 6; data sets of at least 4 MMA blocks): bit-exact, but at config 5 it
        // takes 2.18 ms against 1.11 ms for the fp32-filter kernel -- 40 K warp instructions per 64 queries (two MMA
        // scans 8 K, tile walk 9 K, drain 10 K, setup) at 35 % issue utilisation with one 16-warp CTA per SM
        // (profiles/r02_knnmma_ncu_summary.txt), where the fp32 kernel needs 34 K at twice the occupancy
        if (variant == 6 && a.n >= 512) return launch_knn_mma(a, st);
        // large batches: 256-thread CTAs (fewer, fuller CTAs; 8 % faster at B=500); otherwise 128
        if (variant == 0 && (long long)a.b * a.m >= 148LL * 4 * 512) variant = 3;
        switch (variant) {
          case 1: return launch_knn<KnnCfg<64, 4, 32, 2048, 12, false>>(a, st);
          case 2: return launch_knn<KnnCfg<128, 4, 32, 2048, 12, false>>(a, st);
          case 3: return launch_knn<KnnCfg<256, 2, 32, 2048, 12, false>>(a, st);
          case 4: return launch_knn<KnnCfg<64, 2, 32, 2048, 12, false>>(a, st);
          case 5: return launch_knn<KnnCfg<128, 1, 32, 2048, 12, false>>(a, st);
          default: return launch_knn<KnnCfg<128, 2, 32, 2048, 12, false>>(a, st);
        }
      }
      if (kl <= 17) return launch_knn<KnnCfg<128, 2, 32, 2048, 17, false>>(a, st);
      return launch_knn<KnnCfg<128, 1, 32, 2048, 33, false>>(a, st);
    }
    if (kl <= 12) return launch_knn<KnnCfg<128, 1, 32, 2048, 12, true>>(a, st);
    if (kl <= 17) return launch_knn<KnnCfg<128, 1, 32, 2048, 17, true>>(a, st);
    return launch_knn<KnnCfg<128, 1, 32, 2048, 33, true>>(a, st);
  }
  const long long queries = (long long)a.b * a.m;
  const size_t smem = (size_t)4 * 6 * a.k * 4;
  if (smem > 200 * 1024) {
    set_error("ga_knn: k = %d is too large", a.k);
    return GA_ERR_UNSUPPORTED;
  }
  if ((queries + 3) / 4 > 0x7fffffffLL) {
    set_error("ga_knn: problem too large for one launch");
    return GA_ERR_UNSUPPORTED;
  }
  GA_CUDA_TRY(cudaFuncSetAttribute(knn_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  knn_generic_kernel<<<(unsigned)((queries + 3) / 4), 128, smem, st>>>(a);
  GA_LAUNCH_CHECK("knn_generic_kernel");
  return GA_OK;
}

}  // namespace ga

using namespace ga;

extern "C" {

int ga_knn(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx,
           ga_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_knn: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (!(k > 0)) {
    set_error("SelectionSort expects positive k");  // tf_grouping.cpp:113
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (k > n) {
    // the reference would sort past the end of the row (tf_grouping_g.cu:106-108) and tf.slice would fail
    set_error("ga_knn: k = %d exceeds the data set size n = %d", k, n);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (b == 0 || m == 0) return GA_OK;
  KnnArgs a;
  a.b = b; a.n = n; a.m = m; a.k = k;
  a.xyz1 = xyz1; a.xyz2 = xyz2; a.val = val; a.idx = idx;
  a.skip = 0; a.do_sqrt = 0; a.qtiles = 0;
  return knn_dispatch(a, as_stream(stream));
}

int ga_knn_dists(int b, int n, int k, const float* pc, float* out, ga_stream_t stream) {
  if (b < 0 || n < 0) {
    set_error("ga_knn_dists: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (!(k > 0)) {
    set_error("SelectionSort expects positive k");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (k + 1 > n) {
    set_error("ga_knn_dists: k + 1 = %d exceeds the cloud size n = %d", k + 1, n);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (b == 0) return GA_OK;
  KnnArgs a;
  a.b = b; a.n = n; a.m = n; a.k = k + 1;
  a.xyz1 = pc; a.xyz2 = pc; a.val = out; a.idx = nullptr;
  a.skip = 1; a.do_sqrt = 1; a.qtiles = 0;
  return knn_dispatch(a, as_stream(stream));
}

int ga_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out, ga_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_selection_sort: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (!(k > 0)) {
    set_error("SelectionSort expects positive k");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const long long rows = (long long)b * m;
  if (rows == 0 || n == 0) return GA_OK;
  if ((rows + 3) / 4 > 0x7fffffffLL) {
    set_error("ga_selection_sort: too many rows");
    return GA_ERR_UNSUPPORTED;
  }
  selection_sort_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, as_stream(stream)>>>(rows, n, k, dist, outi, out);
  GA_LAUNCH_CHECK("selection_sort_kernel");
  return GA_OK;
}

int ga_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                   ga_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || m < 0 || nsample < 0) {
    set_error("ga_group_point: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const long long total = (long long)b * m * nsample * c;
  if (total == 0) return GA_OK;
  if ((total + 255) / 256 > 0x7fffffffLL) {
    set_error("ga_group_point: output too large");
    return GA_ERR_UNSUPPORTED;
  }
  group_point_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(total, n, c, m, nsample,
                                                                                     points, idx, out);
  GA_LAUNCH_CHECK("group_point_kernel");
  return GA_OK;
}

}  // extern "C"
