// Chamfer forward: bidirectional brute-force nearest neighbour, both directions in
// one launch.  Replaces NmDistanceKernel x2 (tf_nndistance_g.cu:5-131,
// chamfer3D.cu:12-154) and, bit for bit, nnsearch (tf_nndistance.cpp:21-43).
//
// Filter and refine.  The reference arithmetic costs 8 FP32 ops + compare + two
// selects per point pair.  This kernel scans with a cheaper *filter*
//     f(q,t) = |t|^2 - 2 q.t        (3 FMAs; |q|^2 is constant per query)
// evaluated two targets at a time with packed FFMA2 and reduced with FMNMX3, which
// keeps only a running minimum per 64-target tile (no index, no select) and, per
// query, the three smallest tile minima.  A rigorous rounding bound W (below)
// guarantees that the reference argmin lies in a tile whose filter minimum is
// <= global filter minimum + W; only those tiles (almost always exactly one) are
// re-evaluated in the reference arithmetic with the reference's strict-< /
// lowest-index rule (three or more qualifying tiles: the whole chunk is rescanned).  Result: dist and idx are
// bit-identical to the selected reference arithmetic for every input, at ~3
// FMA-pipe cycles per evaluated pair instead of ~9 issue slots.
//
// Bound.  u = 2^-24, A = max|q_c|, Bm = max|t_c| over the targets seen so far,
// s = A + Bm.  Reference value d vs real D: |d - D| <= 5u*3s^2.  Filter vs real
// (D - |q|^2): <= 18.1u s^2 (norm 9u, three FMA roundings of magnitude <= 3s^2).
// If d(k) <= d(j) then f(k) <= f(j) + 66.2u s^2.  W = 128u s^2 (+ a denormal
// constant) leaves room for rounding W and the threshold themselves.  Non-finite
// or overflowing inputs make W = inf/NaN, every tile qualifies, and the kernel
// degenerates to the plain reference scan (still exact).
#include "nn_tiles.cuh"

namespace ga {

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

struct FwdArgs {
  int b, n, m;
  const float* xyz1;
  const float* xyz2;
  float* dist1;
  int* idx1;
  float* dist2;
  int* idx2;
  int tiles1, tiles2;  // query tiles per cloud, direction 1->2 and 2->1
};

// Re-evaluate one tile for one query: filter again per target, and evaluate the
// survivors in the reference arithmetic.  Lexicographic (d, index) update so the
// order in which tiles are refined does not matter.
template <int T, int MODE>
__device__ __forceinline__ void refine_tile(const float4* __restrict__ tp, int g0, int nt, float ax2, float ay2,
                                            float az2, float qx, float qy, float qz, float thr, float& best,
                                            int& besti) {
  // Lanes of a warp usually refine DIFFERENT tiles; tiles are a multiple of 1 KB apart, so
  // walking them in step would put all 32 lanes on the same banks.  Each lane starts at its
  // own pair instead (the update below is order-independent).
  const int rot = threadIdx.x & 31;
#pragma unroll 4
  for (int i = 0; i < T / 2; i++) {
    const int pp = (i + rot) & (T / 2 - 1);
    const float4 u = tp[2 * pp];
    const float4 v = tp[2 * pp + 1];
    const float2 f = filter_pair(u, v, ax2, ay2, az2);
    if (fminf(f.x, f.y) > thr) continue;  // NaN threshold falls through
    const int g = g0 + 2 * pp;
    if (!(f.x > thr) && g < nt) {
      const float d = sqdist<MODE>(u.x, u.z, v.x, qx, qy, qz);
      if (d < best || (d == best && g < besti)) {
        best = d;
        besti = g;
      }
    }
    if (!(f.y > thr) && g + 1 < nt) {
      const float d = sqdist<MODE>(u.y, u.w, v.y, qx, qy, qz);
      if (d < best || (d == best && g + 1 < besti)) {
        best = d;
        besti = g + 1;
      }
    }
  }
}

// Same walk as refine_tile, but only RECORDS which targets pass the filter (first two
// indices + count, branch-free), so that all lanes stay in step; the exact evaluation
// happens afterwards, once, for every lane together.
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
template <int MODE>
__device__ __forceinline__ void eval_candidate(const float4* __restrict__ tgt, int c0, int g, float qx, float qy,
                                               float qz, float& best, int& besti) {
  const int p = (g - c0) >> 1, h = (g - c0) & 1;
  const float* pu = reinterpret_cast<const float*>(tgt + 2 * p);
  const float d = sqdist<MODE>(pu[h], pu[2 + h], pu[4 + h], qx, qy, qz);
  if (d < best || (d == best && g < besti)) {
    best = d;
    besti = g;
  }
}

template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::kThreads) nn_fwd_kernel(const FwdArgs a) {
  constexpr int THREADS = Cfg::kThreads, Q = Cfg::kQ, QT = Cfg::kQT, T = Cfg::kT, CH = Cfg::kCH;
  const float kInf = __int_as_float(0x7f800000);
  extern __shared__ float4 smem_f4[];
  float4* tgt = smem_f4;                                             // [CH/2 (+pad)][2]
  float* red = reinterpret_cast<float*>(smem_f4 + CH + 2 * kPipeU);  // [32]

  const int tid = threadIdx.x;
  const int jpb = a.tiles1 + a.tiles2;
  const int batch = blockIdx.x / jpb;
  const int r = blockIdx.x - batch * jpb;
  const bool rev = r >= a.tiles1;
  const int qtile = rev ? r - a.tiles1 : r;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  float* odist = (rev ? a.dist2 : a.dist1) + (size_t)batch * nq;
  int* oidx = (rev ? a.idx2 : a.idx1) + (size_t)batch * nq;

  // ---- queries -------------------------------------------------------------
  float qx[Q], qy[Q], qz[Q], qabs[Q], d0[Q], ax2[Q], ay2[Q], az2[Q];
  bool valid[Q];
  const float t0x = __ldg(tpts), t0y = __ldg(tpts + 1), t0z = __ldg(tpts + 2);
#pragma unroll
  for (int j = 0; j < Q; j++) {
    const int qi = qtile * QT + j * THREADS + tid;
    valid[j] = qi < nq;
    const int qs = valid[j] ? qi : 0;
    qx[j] = __ldg(qpts + (size_t)qs * 3);
    qy[j] = __ldg(qpts + (size_t)qs * 3 + 1);
    qz[j] = __ldg(qpts + (size_t)qs * 3 + 2);
    qabs[j] = query_abs(qx[j], qy[j], qz[j]);
    ax2[j] = -2.0f * qx[j];
    ay2[j] = -2.0f * qy[j];
    az2[j] = -2.0f * qz[j];
    d0[j] = sqdist<MODE>(t0x, t0y, t0z, qx[j], qy[j], qz[j]);  // reference: "k==0 ||" seeds best with d(0)
  }
  float best[Q], m1g[Q];
  int besti[Q];
#pragma unroll
  for (int j = 0; j < Q; j++) {
    best[j] = kInf;
    besti[j] = 0;
    m1g[j] = kInf;
  }
  float bm_run = 0.0f;

  for (int c0 = 0; c0 < nt; c0 += CH) {
    const int cn = min(CH, nt - c0);
    const int ntile = (cn + T - 1) / T;
    bm_run = fmaxf(bm_run, stage_targets<THREADS, T>(tgt, red, tpts, c0, nt, ntile, tid));

    // ---- phase 1: filter scan; three smallest tile minima per query ----------
    float c1[Q], c2[Q], c3[Q];
    int i1[Q], i2[Q];
#pragma unroll
    for (int j = 0; j < Q; j++) {
      c1[j] = c2[j] = c3[j] = kInf;
      i1[j] = i2[j] = 0;
    }
    filter_scan<Q, T>(tgt, ntile, ax2, ay2, az2, [&](int tile, const float(&tm)[Q]) {
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const bool lt1 = tm[j] < c1[j], lt2 = tm[j] < c2[j];
        c3[j] = fminf(c3[j], fmaxf(c2[j], tm[j]));
        i2[j] = lt1 ? i1[j] : (lt2 ? tile : i2[j]);
        c2[j] = fminf(c2[j], fmaxf(c1[j], tm[j]));
        i1[j] = lt1 ? tile : i1[j];
        c1[j] = fminf(c1[j], tm[j]);
      }
    });

    // ---- phase 2: refine the qualifying tiles in the reference arithmetic -----
#pragma unroll
    for (int j = 0; j < Q; j++) {
      if (!valid[j]) continue;
      m1g[j] = fminf(m1g[j], c1[j]);
      const float thr = m1g[j] + filter_window(qabs[j], bm_run);
      const bool all_tiles = !(c3[j] > thr);  // three or more tiles in the window, or non-finite data
      const bool t1 = !(c1[j] > thr), t2 = !(c2[j] > thr);
      int cnt = 0, ca = 0, cb = 0;
      if (all_tiles) {
        for (int tile = 0; tile < ntile; tile++)
          scan_tile_candidates<T>(tgt + (size_t)tile * T, c0 + tile * T, nt, ax2[j], ay2[j], az2[j], thr, cnt, ca,
                                  cb);
      } else {
        if (t1)
          scan_tile_candidates<T>(tgt + (size_t)i1[j] * T, c0 + i1[j] * T, nt, ax2[j], ay2[j], az2[j], thr, cnt,
                                  ca, cb);
        if (t2)
          scan_tile_candidates<T>(tgt + (size_t)i2[j] * T, c0 + i2[j] * T, nt, ax2[j], ay2[j], az2[j], thr, cnt,
                                  ca, cb);
      }
      if (cnt <= 2) {  // the usual case: one or two survivors, evaluated by all lanes in step
        if (cnt >= 1) eval_candidate<MODE>(tgt, c0, ca, qx[j], qy[j], qz[j], best[j], besti[j]);
        if (cnt >= 2) eval_candidate<MODE>(tgt, c0, cb, qx[j], qy[j], qz[j], best[j], besti[j]);
      } else {         // many survivors (ties, degenerate or non-finite data): walk again, evaluating inline
        if (all_tiles) {
          for (int tile = 0; tile < ntile; tile++)
            refine_tile<T, MODE>(tgt + (size_t)tile * T, c0 + tile * T, nt, ax2[j], ay2[j], az2[j], qx[j], qy[j],
                                 qz[j], thr, best[j], besti[j]);
        } else {
          if (t1)
            refine_tile<T, MODE>(tgt + (size_t)i1[j] * T, c0 + i1[j] * T, nt, ax2[j], ay2[j], az2[j], qx[j],
                                 qy[j], qz[j], thr, best[j], besti[j]);
          if (t2)
            refine_tile<T, MODE>(tgt + (size_t)i2[j] * T, c0 + i2[j] * T, nt, ax2[j], ay2[j], az2[j], qx[j],
                                 qy[j], qz[j], thr, best[j], besti[j]);
        }
      }
    }
  }

#pragma unroll
  for (int j = 0; j < Q; j++) {
    if (!valid[j]) continue;
    const int qi = qtile * QT + j * THREADS + tid;
    const bool seed_nan = d0[j] != d0[j];  // reference: best = d(0) = NaN is never replaced
    odist[qi] = seed_nan ? d0[j] : best[j];
    oidx[qi] = seed_nan ? 0 : besti[j];
  }
}

template <class Cfg>
static int launch_fwd(const FwdArgs& a, int mode, cudaStream_t st) {
  const long long jobs = (long long)a.b * (a.tiles1 + a.tiles2);
  if (jobs <= 0) return GA_OK;
  if (jobs > 0x7fffffffLL) {
    set_error("ga_nn_distance_fwd: problem too large for one launch (%lld CTAs)", jobs);
    return GA_ERR_UNSUPPORTED;
  }
  auto k0 = nn_fwd_kernel<Cfg, GA_MODE_CPU_EXACT>;
  auto k1 = nn_fwd_kernel<Cfg, GA_MODE_GPU_REF>;
  auto k = mode == GA_MODE_CPU_EXACT ? k0 : k1;
  static thread_local bool attr_set[2] = {false, false};
  if (!attr_set[mode]) {
    GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    attr_set[mode] = true;
  }
  k<<<(unsigned)jobs, Cfg::kThreads, Cfg::kSmem, st>>>(a);
  GA_LAUNCH_CHECK("nn_fwd_kernel");
  return GA_OK;
}

int g_fwd_variant = 0;  // tuning hook (ga_set_tuning): 0 = default

}  // namespace ga

extern "C" int ga_nn_distance_fwd(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1,
                                  int* idx1, float* dist2, int* idx2, int mode, ga_stream_t stream) {
  using namespace ga;
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_nn_distance_fwd: negative size (b=%d n=%d m=%d)", b, n, m);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (mode != GA_MODE_CPU_EXACT && mode != GA_MODE_GPU_REF) {
    set_error("ga_nn_distance_fwd: unknown mode %d", mode);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (b == 0) return GA_OK;
  cudaStream_t st = as_stream(stream);
  // Empty target cloud: the reference loops leave best = 0, besti = 0 (tf_nndistance.cpp:27-28).
  if (m == 0 && n > 0) {
    GA_CUDA_TRY(cudaMemsetAsync(dist1, 0, sizeof(float) * (size_t)b * n, st));
    GA_CUDA_TRY(cudaMemsetAsync(idx1, 0, sizeof(int) * (size_t)b * n, st));
  }
  if (n == 0 && m > 0) {
    GA_CUDA_TRY(cudaMemsetAsync(dist2, 0, sizeof(float) * (size_t)b * m, st));
    GA_CUDA_TRY(cudaMemsetAsync(idx2, 0, sizeof(int) * (size_t)b * m, st));
  }
  if (n == 0 || m == 0) return GA_OK;

  FwdArgs a;
  a.b = b; a.n = n; a.m = m;
  a.xyz1 = xyz1; a.xyz2 = xyz2;
  a.dist1 = dist1; a.idx1 = idx1; a.dist2 = dist2; a.idx2 = idx2;
  switch (g_fwd_variant) {
#define GA_FWD_CASE(ID, TH, QQ, TT, CC)                                                          \
  case ID: {                                                                                     \
    using Cfg = FwdCfg<TH, QQ, TT, CC>;                                                          \
    a.tiles1 = (n + Cfg::kQT - 1) / Cfg::kQT;                                                    \
    a.tiles2 = (m + Cfg::kQT - 1) / Cfg::kQT;                                                    \
    return launch_fwd<Cfg>(a, mode, st);                                                         \
  }
    GA_FWD_CASE(1, 128, 4, 32, 2048)
    GA_FWD_CASE(2, 64, 4, 64, 2048)
    GA_FWD_CASE(3, 64, 2, 32, 2048)
    GA_FWD_CASE(4, 128, 2, 32, 2048)
    GA_FWD_CASE(5, 64, 8, 32, 2048)
    GA_FWD_CASE(6, 32, 8, 32, 2048)
    GA_FWD_CASE(7, 64, 4, 16, 2048)
    GA_FWD_CASE(8, 32, 4, 32, 2048)
    GA_FWD_CASE(9, 64, 4, 32, 1024)
    default:
    GA_FWD_CASE(0, 64, 4, 32, 2048)
#undef GA_FWD_CASE
  }
}
