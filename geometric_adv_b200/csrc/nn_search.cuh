// Per-thread nearest-neighbour search over a target cloud staged in shared memory:
// the filter-and-refine core shared by nn_distance_fwd.cu (writes dist/idx per query)
// and all_pairs.cu (sums dist per cloud pair).  See nn_distance_fwd.cu for the method
// and the rounding bound.
#pragma once
#include "nn_tiles.cuh"

namespace ga {

// Arguments of every Chamfer forward kernel (nn_distance_fwd.cu, nn_distance_fwd_mma.cu).
struct FwdArgs {
  int b, n, m;
  const float* xyz1;
  const float* xyz2;
  float* dist1;
  int* idx1;
  float* dist2;
  int* idx2;
  int tiles1, tiles2;  // query tiles per cloud, direction 1->2 and 2->1
  // optional second destination for every output (mapped pinned HOST memory: the host entry
  // points let the kernel stream results over PCIe while it computes, instead of a D2H copy)
  float* mdist1;
  int* midx1;
  float* mdist2;
  int* midx2;
  // completion tickets (nn_fwd_mma_kernel only; nullptr / 0 = off): see ga_common.cuh
  unsigned long long* ticket;
  unsigned long long call_id;
  int ticket_debug;
  int* frame_hint;  // device view of the launcher's off-origin hint word (nn_distance_fwd_mma.cu), or nullptr
  // Streamed ingest (nn_fwd_mma_kernel only; nullptr = off, see wait_ready below): the clouds are still arriving
  // over PCIe when the kernel starts; ready[batch / ready_per] turns non-zero once the copy engine has delivered
  // that group of batch elements.
  const int* ready = nullptr;
  int ready_per = 1;
  int* ready_abort = nullptr;  // set to 1 (mapped host word) by a CTA that gave up waiting
  unsigned ready_spin_us = 4000;  // how long a CTA waits before it gives up (the launcher scales it with the transfer)
};

// Streamed ingest, device side.  The host entry (host_api.cu) launches the search on one graph branch and the
// H2D copies on another: xyz1 / xyz2 arrive in groups of batch elements, each group followed by a 4-byte copy
// that sets its flag.  A CTA waits for the flag of its batch element before it touches the clouds (one thread
// polls with a system-scope acquire load, the others sit at the barrier), so the search of group g runs under
// the copy of group g+1 without splitting the launch.  Copy engines need no SM, hence nothing a waiting CTA
// holds can delay the data it waits for; should the flag still not come (a driver that serialises the two
// branches), the CTA gives up after ready_spin_us, raises ready_abort and returns WITHOUT results: the host
// entry sees the word after the step and redoes the step on the plain path.  Clouds must be whole 128-byte
// lines (n, m multiples of 32): a line shared by two batch elements could be cached before its second half
// has arrived.  Returns false if the CTA must leave.  Every thread of the CTA calls it.
__device__ __forceinline__ bool wait_ready(const FwdArgs& a, int batch) {
  if (a.ready == nullptr) return true;
  __shared__ int ready_ok;
  if (threadIdx.x == 0) {
    const int* f = a.ready + batch / a.ready_per;
    const unsigned long long t0 = global_ns();
    int v;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v != 0 || global_ns() - t0 > 1000ull * a.ready_spin_us) break;
      __nanosleep(200);
    }
    if (v == 0) *reinterpret_cast<volatile int*>(a.ready_abort) = 1;
    ready_ok = v != 0;
  }
  __syncthreads();
  return ready_ok != 0;
}

template <int THREADS, int Q, int T, int CH>
struct FwdCfg {
  static constexpr int kThreads = THREADS;
  static constexpr int kQ = Q;               // queries per thread
  static constexpr int kQT = THREADS * Q;    // queries per CTA
  static constexpr int kT = T;               // targets per filter tile
  static constexpr int kCH = CH;             // targets staged per chunk
  static constexpr size_t kSmem = (size_t)CH * 16 + (size_t)kPipeU * 32 + 32 * 4;
  static_assert(CH % T == 0 && (T & (T - 1)) == 0, "tile shapes");
};

// Walk one tile for one query and RECORD which targets pass the filter (first two indices +
// count, branch-free), so that all lanes stay in step; the exact evaluation happens afterwards,
// once, for every lane together.  Lanes of a warp usually refine DIFFERENT tiles; tiles are a
// multiple of 512 B apart, so walking them in step would put all 32 lanes on the same banks:
// each lane starts at its own pair instead.
template <int T>
__device__ __forceinline__ void scan_tile_candidates(const float4* __restrict__ tp, int g0, int nt, float ax2,
                                                     float ay2, float az2, float thr, int& cnt, int& ca,
                                                     int& cb) {
  const int rot = threadIdx.x & 31;
#pragma unroll 4
  for (int i = 0; i < T / 2; i++) {
    const int pp = (i + rot) & (T / 2 - 1);
    const float4 u = tp[2 * pp];
    const float4 v = tp[2 * pp + 1];
    const float2 f = filter_pair(u, v, ax2, ay2, az2);
    const int g = g0 + 2 * pp;
    const bool p0 = !(f.x > thr) && g < nt;  // NaN filter value / threshold counts as a candidate
    const bool p1 = !(f.y > thr) && g + 1 < nt;
    cb = (p0 && cnt == 1) ? g : cb;
    ca = (p0 && cnt == 0) ? g : ca;
    cnt += p0 ? 1 : 0;
    cb = (p1 && cnt == 1) ? g + 1 : cb;
    ca = (p1 && cnt == 0) ? g + 1 : ca;
    cnt += p1 ? 1 : 0;
  }
}

// Reference arithmetic for one staged target (index g in the chunk starting at c0).
// torig != nullptr: the staged copy is in a shifted frame (Frame, nn_tiles.cuh); the target's original coordinates
// are fetched from its cloud in global memory (torig = first point of the cloud).
template <int MODE>
__device__ __forceinline__ void eval_candidate(const float4* __restrict__ tgt, int c0, int g, float qx, float qy,
                                               float qz, float& best, int& besti,
                                               const float* __restrict__ torig = nullptr) {
  const int p = (g - c0) >> 1, h = (g - c0) & 1;
  const float* pu = reinterpret_cast<const float*>(tgt + 2 * p);
  float tx = pu[h], ty = pu[2 + h], tz = pu[4 + h];
  if (torig != nullptr) {
    tx = __ldg(torig + (size_t)g * 3);
    ty = __ldg(torig + (size_t)g * 3 + 1);
    tz = __ldg(torig + (size_t)g * 3 + 2);
  }
  const float d = sqdist<MODE>(tx, ty, tz, qx, qy, qz);
  if (d < best || (d == best && g < besti)) {
    best = d;
    besti = g;
  }
}

// Rare path, whole warp for ONE query: every lane filters a strided share of the staged
// chunk, survivors are evaluated in the reference arithmetic, and the 32 partial results are
// merged by (value, index).  Used when three or more tiles fall inside the window or more than
// two targets survive (exact ties, degenerate or non-finite data).  A single lane walking the
// chunk alone would take as long as the whole normal search of its warp and stall the CTA.
template <int MODE>
__device__ __forceinline__ void warp_exact_scan(const float4* __restrict__ tgt, int c0, int nt, int npair,
                                                float qx, float qy, float qz, float ax2, float ay2, float az2,
                                                float thr, float& b, int& bi, int lane,
                                                const float* __restrict__ torig = nullptr) {
  b = __int_as_float(0x7f800000);
  bi = 0x7fffffff;
  for (int p = lane; p < npair; p += 32) {
    const float4 u = tgt[2 * p];
    const float4 v = tgt[2 * p + 1];
    const float2 f = filter_pair(u, v, ax2, ay2, az2);
    const int g = c0 + 2 * p;
    if (!(f.x > thr) && g < nt) {
      float tx = u.x, ty = u.z, tz = v.x;
      if (torig != nullptr) {
        tx = __ldg(torig + (size_t)g * 3);
        ty = __ldg(torig + (size_t)g * 3 + 1);
        tz = __ldg(torig + (size_t)g * 3 + 2);
      }
      const float d = sqdist<MODE>(tx, ty, tz, qx, qy, qz);
      if (d < b || (d == b && g < bi)) {
        b = d;
        bi = g;
      }
    }
    if (!(f.y > thr) && g + 1 < nt) {
      float tx = u.y, ty = u.w, tz = v.y;
      if (torig != nullptr) {
        tx = __ldg(torig + (size_t)(g + 1) * 3);
        ty = __ldg(torig + (size_t)(g + 1) * 3 + 1);
        tz = __ldg(torig + (size_t)(g + 1) * 3 + 2);
      }
      const float d = sqdist<MODE>(tx, ty, tz, qx, qy, qz);
      if (d < b || (d == b && g + 1 < bi)) {
        b = d;
        bi = g + 1;
      }
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
}

// Q queries per thread: coordinates, filter coefficients, running results.
template <int Q>
struct QueryState {
  float qx[Q], qy[Q], qz[Q], qabs[Q], ax2[Q], ay2[Q], az2[Q];
  float d0[Q];     // reference arithmetic to target 0 ("k==0 ||" seeds best with d(0))
  float best[Q];   // exact running minimum
  int besti[Q];
  float m1g[Q];    // running minimum of the filter over all chunks seen
  bool valid[Q];
};

// Load the Q queries of this thread (query tile `qtile` of a cloud with nq points).
template <class Cfg, int MODE>
__device__ __forceinline__ void load_queries(QueryState<Cfg::kQ>& s, const float* __restrict__ qpts, int nq,
                                             int qtile, const float* __restrict__ tpts, int tid) {
  constexpr int Q = Cfg::kQ, QT = Cfg::kQT, THREADS = Cfg::kThreads;
  const float kInf = __int_as_float(0x7f800000);
  const float t0x = __ldg(tpts), t0y = __ldg(tpts + 1), t0z = __ldg(tpts + 2);
#pragma unroll
  for (int j = 0; j < Q; j++) {
    const int qi = qtile * QT + j * THREADS + tid;
    s.valid[j] = qi < nq;
    const int qs = s.valid[j] ? qi : 0;
    s.qx[j] = __ldg(qpts + (size_t)qs * 3);
    s.qy[j] = __ldg(qpts + (size_t)qs * 3 + 1);
    s.qz[j] = __ldg(qpts + (size_t)qs * 3 + 2);
    s.qabs[j] = query_abs(s.qx[j], s.qy[j], s.qz[j]);
    s.ax2[j] = -2.0f * s.qx[j];
    s.ay2[j] = -2.0f * s.qy[j];
    s.az2[j] = -2.0f * s.qz[j];
    s.d0[j] = sqdist<MODE>(t0x, t0y, t0z, s.qx[j], s.qy[j], s.qz[j]);
    s.best[j] = kInf;
    s.besti[j] = 0;
    s.m1g[j] = kInf;
  }
}

// The three smallest tile minima of the filter per query, and the tiles of the two smallest.
template <int Q>
struct TileTrack {
  float c1[Q], c2[Q], c3[Q];
  int i1[Q], i2[Q];
};

// Phase 1: filter scan over the `ntile` staged tiles; per query the three smallest tile minima.
template <class Cfg>
__device__ __forceinline__ void search_phase1(const QueryState<Cfg::kQ>& s, const float4* __restrict__ tgt,
                                              int ntile, TileTrack<Cfg::kQ>& tr) {
  constexpr int Q = Cfg::kQ, T = Cfg::kT;
  const float kInf = __int_as_float(0x7f800000);
#pragma unroll
  for (int j = 0; j < Q; j++) {
    tr.c1[j] = tr.c2[j] = tr.c3[j] = kInf;
    tr.i1[j] = tr.i2[j] = 0;
  }
  filter_scan<Q, T>(tgt, ntile, s.ax2, s.ay2, s.az2, [&](int tile, const float(&tm)[Q]) {
#pragma unroll
    for (int j = 0; j < Q; j++) {
      const bool lt1 = tm[j] < tr.c1[j], lt2 = tm[j] < tr.c2[j];
      tr.c3[j] = fminf(tr.c3[j], fmaxf(tr.c2[j], tm[j]));
      tr.i2[j] = lt1 ? tr.i1[j] : (lt2 ? tile : tr.i2[j]);
      tr.c2[j] = fminf(tr.c2[j], fmaxf(tr.c1[j], tm[j]));
      tr.i1[j] = lt1 ? tile : tr.i1[j];
      tr.c1[j] = fminf(tr.c1[j], tm[j]);
    }
  });
}

// Phase 2: refine the staged tiles whose minimum is within the window of fmin[j] (the smallest
// filter value of query j over ALL targets examined so far, possibly by other CTAs) in the
// reference arithmetic.  c0 = index of the first staged target in its cloud (nt points).
// Must be called by whole, converged warps (it uses warp collectives).
template <class Cfg, int MODE>
__device__ __forceinline__ void search_phase2(QueryState<Cfg::kQ>& s, const float4* __restrict__ tgt, int c0, int nt,
                                              int ntile, const TileTrack<Cfg::kQ>& tr, const float (&fmin)[Cfg::kQ],
                                              float bm_run) {
  constexpr int Q = Cfg::kQ, T = Cfg::kT;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < Q; j++) {
    float thr = 0.0f;
    bool hard = false;  // needs the cooperative scan
    if (s.valid[j]) {
      thr = fmin[j] + filter_window(s.qabs[j], bm_run);
      hard = !(tr.c3[j] > thr);  // three or more tiles in the window, or non-finite data
      if (!hard) {
        int cnt = 0, ca = 0, cb = 0;
        if (!(tr.c1[j] > thr))
          scan_tile_candidates<T>(tgt + (size_t)tr.i1[j] * T, c0 + tr.i1[j] * T, nt, s.ax2[j], s.ay2[j], s.az2[j],
                                  thr, cnt, ca, cb);
        if (!(tr.c2[j] > thr))
          scan_tile_candidates<T>(tgt + (size_t)tr.i2[j] * T, c0 + tr.i2[j] * T, nt, s.ax2[j], s.ay2[j], s.az2[j],
                                  thr, cnt, ca, cb);
        // the usual case: one or two survivors, evaluated by all lanes in step
        if (cnt >= 1 && cnt <= 2) eval_candidate<MODE>(tgt, c0, ca, s.qx[j], s.qy[j], s.qz[j], s.best[j], s.besti[j]);
        if (cnt == 2) eval_candidate<MODE>(tgt, c0, cb, s.qx[j], s.qy[j], s.qz[j], s.best[j], s.besti[j]);
        hard = cnt > 2;  // many survivors: exact ties or degenerate data
      }
    }
    // rare: the whole warp serves the flagged queries one at a time
    unsigned pending = __ballot_sync(0xffffffffu, hard);
    while (pending) {
      const int src = __ffs(pending) - 1;
      pending &= pending - 1;
      const float bqx = __shfl_sync(0xffffffffu, s.qx[j], src), bqy = __shfl_sync(0xffffffffu, s.qy[j], src),
                  bqz = __shfl_sync(0xffffffffu, s.qz[j], src);
      const float bax = __shfl_sync(0xffffffffu, s.ax2[j], src), bay = __shfl_sync(0xffffffffu, s.ay2[j], src),
                  baz = __shfl_sync(0xffffffffu, s.az2[j], src);
      const float bthr = __shfl_sync(0xffffffffu, thr, src);
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

// Search one staged chunk (targets [c0, c0 + ntile*T) of a cloud with nt points).
// bm_run: max |coordinate| over all targets staged so far (including this chunk).
template <class Cfg, int MODE>
__device__ __forceinline__ void search_chunk(QueryState<Cfg::kQ>& s, const float4* __restrict__ tgt, int c0, int nt,
                                             int ntile, float bm_run) {
  constexpr int Q = Cfg::kQ;
  TileTrack<Q> tr;
  search_phase1<Cfg>(s, tgt, ntile, tr);
#pragma unroll
  for (int j = 0; j < Q; j++) s.m1g[j] = fminf(s.m1g[j], tr.c1[j]);
  search_phase2<Cfg, MODE>(s, tgt, c0, nt, ntile, tr, s.m1g, bm_run);
}

// Final (dist, idx) of query slot j with the reference's NaN-seed rule.
template <int Q>
__device__ __forceinline__ void finish_query(const QueryState<Q>& s, int j, float& dist, int& idx) {
  const bool seed_nan = s.d0[j] != s.d0[j];  // reference: best = d(0) = NaN is never replaced
  dist = seed_nan ? s.d0[j] : s.best[j];
  idx = seed_nan ? 0 : s.besti[j];
}

}  // namespace ga
