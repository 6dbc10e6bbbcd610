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
#include "nn_tiles.cuh"

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

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads) knn_kernel(const KnnArgs a) {
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
