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
// lowest-index rule (three or more qualifying tiles, or more than two surviving
// targets: the whole warp rescans the chunk for that query).  Result: dist and idx are
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
#include <atomic>

#include <cooperative_groups.h>

#include "nn_search.cuh"

namespace ga {


template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::kThreads) nn_fwd_kernel(const FwdArgs a) {
  constexpr int THREADS = Cfg::kThreads, Q = Cfg::kQ, QT = Cfg::kQT, T = Cfg::kT, CH = Cfg::kCH;
  // dependents (the gradient kernel) may be scheduled once every CTA of this grid has started
  asm volatile("griddepcontrol.launch_dependents;");
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
  float* mdist = rev ? a.mdist2 : a.mdist1;
  int* midx = rev ? a.midx2 : a.midx1;

  QueryState<Q> s;
  load_queries<Cfg, MODE>(s, qpts, nq, qtile, tpts, tid);
  float bm_run = 0.0f;
  for (int c0 = 0; c0 < nt; c0 += CH) {
    const int cn = min(CH, nt - c0);
    const int ntile = (cn + T - 1) / T;
    bm_run = fmaxf(bm_run, stage_targets<THREADS, T>(tgt, red, tpts, c0, nt, ntile, tid));
    search_chunk<Cfg, MODE>(s, tgt, c0, nt, ntile, bm_run);
  }
#pragma unroll
  for (int j = 0; j < Q; j++) {
    if (!s.valid[j]) continue;
    const int qi = qtile * QT + j * THREADS + tid;
    float d;
    int i;
    finish_query<Q>(s, j, d, i);
    odist[qi] = d;
    oidx[qi] = i;
    if (mdist != nullptr) {
      mdist[(size_t)batch * nq + qi] = d;
      midx[(size_t)batch * nq + qi] = i;
    }
  }
}

// ---- small problems: split the TARGETS of each query tile over a thread-block cluster ------
// With few cloud pairs (B=1 per-cloud losses, autoencoder.py:150-168; the attack's batch of 10)
// there are not enough query tiles to occupy 148 SMs.  Here a cluster of S CTAs shares one query
// tile; CTA r stages and scans only tiles [r*tl, (r+1)*tl) of the target cloud.  Two exchanges
// through distributed shared memory keep the result exact:
//   1. every CTA publishes its filter minimum per query (and its max |coordinate|) to all others,
//      so that all of them refine against the GLOBAL filter minimum and window;
//   2. the exact partial results (value, index) go to rank 0, which merges them by
//      (value, index) -- lowest index wins ties, as in the reference -- and writes the outputs.
constexpr int kSplitMax = 8;
using SplitCfg = FwdCfg<128, 2, 32, 2048>;   // tiny problems: most CTAs
using SplitCfg4 = FwdCfg<128, 4, 32, 2048>;  // a few hundred query tiles: FMA-bound inner loop
template <class Cfg>
constexpr size_t split_smem() {
  return Cfg::kSmem + (size_t)3 * kSplitMax * Cfg::kQT * 4 + kSplitMax * 4;
}

template <class Cfg, int MODE>
__global__ void __launch_bounds__(Cfg::kThreads) nn_fwd_split_kernel(const FwdArgs a, const int S) {
  namespace cg = cooperative_groups;
  constexpr int THREADS = Cfg::kThreads, Q = Cfg::kQ, QT = Cfg::kQT, T = Cfg::kT, CH = Cfg::kCH;
  const float kInf = __int_as_float(0x7f800000);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  // A CTA may write a peer's shared memory only once that peer has started executing: arrive now,
  // wait right before the first remote store (the local scan in between hides the barrier).
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  extern __shared__ float4 smem_f4[];
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem_f4 + CH + 2 * kPipeU);  // [32]
  float* xf = red + 32;                                              // [kSplitMax][QT] filter minima
  float* xb = xf + kSplitMax * QT;                                   // [kSplitMax][QT] exact partial values
  int* xi = reinterpret_cast<int*>(xb + kSplitMax * QT);             // [kSplitMax][QT] exact partial indices
  float* xbm = reinterpret_cast<float*>(xi + kSplitMax * QT);        // [kSplitMax] max |coordinate|

  const int tid = threadIdx.x;
  const int job = blockIdx.x / S;
  const int jpb = a.tiles1 + a.tiles2;
  const int batch = job / jpb;
  const int r = job - batch * jpb;
  const bool rev = r >= a.tiles1;
  const int qtile = rev ? r - a.tiles1 : r;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  float* odist = (rev ? a.dist2 : a.dist1) + (size_t)batch * nq;
  int* oidx = (rev ? a.idx2 : a.idx1) + (size_t)batch * nq;

  const int ntile_all = (nt + T - 1) / T;
  const int tl = (ntile_all + S - 1) / S;
  const int tile0 = min(ntile_all, rank * tl);
  const int ntl = min(tl, ntile_all - tile0);
  const int c0 = tile0 * T;

  QueryState<Q> s;
  load_queries<Cfg, MODE>(s, qpts, nq, qtile, tpts, tid);
  const float bm = stage_targets<THREADS, T>(tgt, red, tpts, c0, nt, ntl, tid);
  TileTrack<Q> tr;
  search_phase1<Cfg>(s, tgt, ntl, tr);

  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");  // every CTA of the cluster is running
  for (int rr = 0; rr < S; rr++) {  // publish this CTA's filter minima to every CTA of the cluster
    float* dst = cluster.map_shared_rank(xf, rr);
#pragma unroll
    for (int j = 0; j < Q; j++) dst[rank * QT + j * THREADS + tid] = tr.c1[j];
    if (tid == 0) cluster.map_shared_rank(xbm, rr)[rank] = bm;
  }
  cluster.sync();
  float fmin[Q];
  float bmg = 0.0f;
#pragma unroll
  for (int j = 0; j < Q; j++) fmin[j] = kInf;
  for (int rr = 0; rr < S; rr++) {
#pragma unroll
    for (int j = 0; j < Q; j++) fmin[j] = fminf(fmin[j], xf[rr * QT + j * THREADS + tid]);
    bmg = fmaxf(bmg, xbm[rr]);
  }
  search_phase2<Cfg, MODE>(s, tgt, c0, nt, ntl, tr, fmin, bmg);

  {
    float* b0 = cluster.map_shared_rank(xb, 0);
    int* i0 = cluster.map_shared_rank(xi, 0);
#pragma unroll
    for (int j = 0; j < Q; j++) {
      b0[rank * QT + j * THREADS + tid] = s.best[j];
      i0[rank * QT + j * THREADS + tid] = s.besti[j];
    }
  }
  cluster.sync();
  if (rank != 0) return;
#pragma unroll
  for (int j = 0; j < Q; j++) {
    if (!s.valid[j]) continue;
    float b = kInf;
    int bi = 0;
    for (int rr = 0; rr < S; rr++) {
      const float ob = xb[rr * QT + j * THREADS + tid];
      const int obi = xi[rr * QT + j * THREADS + tid];
      if (ob < b || (ob == b && obi < bi && ob < kInf)) {
        b = ob;
        bi = obi;
      }
    }
    s.best[j] = b;
    s.besti[j] = bi;
    const int qi = qtile * QT + j * THREADS + tid;
    float d;
    int i;
    finish_query<Q>(s, j, d, i);
    odist[qi] = d;
    oidx[qi] = i;
    float* mdist = rev ? a.mdist2 : a.mdist1;
    int* midx = rev ? a.midx2 : a.midx1;
    if (mdist != nullptr) {
      mdist[(size_t)batch * nq + qi] = d;
      midx[(size_t)batch * nq + qi] = i;
    }
  }
}

template <class Cfg>
static int launch_fwd_split(FwdArgs a, int mode, int S, cudaStream_t st) {
  constexpr size_t kSplitSmem = split_smem<Cfg>();
  a.tiles1 = (a.n + Cfg::kQT - 1) / Cfg::kQT;
  a.tiles2 = (a.m + Cfg::kQT - 1) / Cfg::kQT;
  const long long jobs = (long long)a.b * (a.tiles1 + a.tiles2);
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_split_kernel<Cfg, GA_MODE_CPU_EXACT>
                                     : nn_fwd_split_kernel<Cfg, GA_MODE_GPU_REF>;
  {
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSplitSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(jobs * S));
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = kSplitSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GA_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, a, S));
  GA_LAUNCH_CHECK("nn_fwd_split_kernel");
  return GA_OK;
}

int g_fwd_split = -1;    // tuning hook (key 5): -1 auto, 0 never, 2/4/8 force
int g_fwd_split_q = 0;   // tuning hook (key 6): 0 auto, 2 or 4 queries per thread in the split kernel

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
  {
    static std::atomic<unsigned> done_mask[2];  // per mode, one bit per device
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  k<<<(unsigned)jobs, Cfg::kThreads, Cfg::kSmem, st>>>(a);
  GA_LAUNCH_CHECK("nn_fwd_kernel");
  return GA_OK;
}

int g_fwd_variant = 0;  // tuning hook (ga_set_tuning): 0 = default
int launch_fwd_mma(const FwdArgs& a, int mode, cudaStream_t st);          // nn_distance_fwd_mma.cu
int launch_fwd_mma_persist(const FwdArgs& a, int mode, cudaStream_t st);  // nn_distance_fwd_mma.cu
int launch_fwd_mma_balanced(const FwdArgs& a, int mode, cudaStream_t st); // nn_distance_fwd_mma.cu
int launch_fwd_umma(const FwdArgs& a, int mode, cudaStream_t st);         // nn_distance_fwd_umma.cu
int launch_fwd_mma_ws(const FwdArgs& a, int mode, cudaStream_t st);       // nn_distance_fwd_mma_ws.cu
bool fwd_mma_ws_supported(int b, int n, int m);                           // nn_distance_fwd_mma_ws.cu
bool fwd_umma_supported(int n, int m);                                    // nn_distance_fwd_umma.cu
extern thread_local int t_want_tickets;                                   // nn_distance_fwd_mma.cu
int g_umma_auto = 1;  // tuning hook (key 21): 0 = never pick the tcgen05 kernel automatically

// Tensor-core kernel of choice for a batch the HMMA kernel qualifies for.  The HMMA grid kernel (512-query CTAs, two
// per SM) moves in steps of whole CTA rounds per SM: 28.6 us up to 148 CTAs, ~43 us up to 296, ~56 us up to 444
// (2048-point clouds; B = 16 costs what B = 18 does and B = 19 what B = 37 does).  The persistent tcgen05 kernel
// splits the same work into 128-query jobs, equal shares per SM, ~12 us + 4.5 us per job of an SM, so it wins on the
// near side of every HMMA step and loses on the far side (profiles/r02_tune_fwd_sweep.txt: B = 4 18.4 vs 22.4 us,
// 8 22.5 vs 26.6, 10 25.6 vs 28.6, 16 30.6 vs 28.6, 20 34.8 vs 40.9, 25 38.8 vs 41.0, 32 45.0 vs 43.9, 40 53.2 vs
// 55.3, 50 61.4 vs 57.3).  Both times scale with the size of the target clouds.  The one-call entry keeps the HMMA
// kernel: only that one checks in on the completion tickets the gradient kernel starts early from.
static bool prefer_umma(int b, int n, int m) {
  if (!fwd_umma_supported(n, m) || t_want_tickets) return false;
  const long long jobs = (long long)b * ((n + 127) / 128 + (m + 127) / 128);
  const long long ctas = (long long)b * ((n + 511) / 512 + (m + 511) / 512);
  const int sms = sm_count();
  const double scale = 0.5 * ((double)n + m) / 2048.0;  // both kernels' per-unit cost follows the target cloud size
  const double t_umma = 12.0 + 4.5 * scale * (double)((jobs + sms - 1) / sms);
  const double t_hmma = 14.5 + 14.0 * scale * (double)((ctas + sms - 1) / sms);
  return t_umma + 0.5 < t_hmma;
}

// Streamed ingest is honoured by the HMMA grid kernel only: true if the default dispatch takes that kernel for
// this shape (same conditions as below) and the clouds are whole 128-byte lines.
bool fwd_ready_supported(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0 || (n & 31) || (m & 31)) return false;
  if (g_fwd_variant != 0 && g_fwd_variant != 20) return false;
  const bool mma_ok = g_fwd_split <= 0 && n >= 256 && m >= 256 &&
                      (long long)b * ((n + 511) / 512 + (m + 511) / 512) >= 74;
  if (!mma_ok) return false;
  return g_fwd_variant == 20 || !(g_umma_auto && prefer_umma(b, n, m));
}

}  // namespace ga

namespace ga {
int nn_distance_fwd_mirrored(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                             float* dist2, int* idx2, float* mdist1, int* midx1, float* mdist2, int* midx2,
                             int mode, ga_stream_t stream) {
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
  a.mdist1 = mdist1; a.midx1 = midx1; a.mdist2 = mdist2; a.midx2 = midx2;
  a.ticket = nullptr; a.call_id = 0; a.ticket_debug = 0; a.frame_hint = nullptr;
  last_forward().call_id = 0;  // only a ticketed launch below re-arms the gradient kernel's early start
  if (t_ready_arm.flags != nullptr) {  // streamed ingest (host_api.cu checked fwd_ready_supported)
    a.ready = t_ready_arm.flags;
    a.ready_per = t_ready_arm.per;
    a.ready_abort = t_ready_arm.abort_word;
    // patience of a waiting CTA: 4 ms + the whole transfer at 1 GB/s (a twentieth of what the link does)
    a.ready_spin_us = 4000u + (unsigned)(((long long)b * ((long long)n + m) * 12) / 1000);
    return launch_fwd_mma(a, mode, st);
  }
  // Default: 4 queries per thread (FMA-pipe bound scan).  Small problems (few hundred query
  // tiles, e.g. the B=1 calls of autoencoder.py:150-168) take 2 queries per thread instead:
  // twice the CTAs to spread over the 148 SMs matters more than the LDS-bound inner loop.
  // Batches that give the tensor-core kernel at least half a wave of 512-query CTAs take it
  // (measured, profiles/r01_tune_mma.json: B=50 56 vs 78 us, B=512 401 vs 612 us, B=10 27.7 vs
  // 30.7 us, 4 x 8192 points 84 vs 113 us); below that the fp32-filter kernels win.
  const bool mma_auto = g_fwd_variant == 0 && g_fwd_split <= 0 && n >= 256 && m >= 256 &&
                        (long long)b * ((n + 511) / 512 + (m + 511) / 512) >= 74;
  // Few query tiles and clouds that fit one chunk: split the targets over a cluster.
  if (!mma_auto && n <= SplitCfg::kCH && m <= SplitCfg::kCH && g_fwd_variant == 0 && g_fwd_split != 0) {
    const long long jobs2 = (long long)b * ((n + SplitCfg::kQT - 1) / SplitCfg::kQT + (m + SplitCfg::kQT - 1) / SplitCfg::kQT);
    const long long jobs4 = (long long)b * ((n + SplitCfg4::kQT - 1) / SplitCfg4::kQT + (m + SplitCfg4::kQT - 1) / SplitCfg4::kQT);
    const int mintiles = ((n < m ? n : m) + SplitCfg::kT - 1) / SplitCfg::kT;
    // measured (profiles/r01_tune.json): 2 queries per thread wins at B=1 (12.2 vs 15.2 us) and
    // B=10 (30.7 vs 32.8 us); 4 per thread is kept selectable for tuning
    const bool q4 = g_fwd_split_q == 4;
    const long long jobs = q4 ? jobs4 : jobs2;
    int S = 1;
    if (g_fwd_split > 0) {
      S = g_fwd_split;
    } else {
      while (S < kSplitMax && jobs * S * 2 <= 148 * 4 && mintiles / (S * 2) >= 4) S *= 2;
    }
    if (S > 1) return q4 ? launch_fwd_split<SplitCfg4>(a, mode, S, st) : launch_fwd_split<SplitCfg>(a, mode, S, st);
  }
  int variant = g_fwd_variant;
  if (variant == 0) {
    const long long queries = (long long)b * ((long long)n + m);
    variant = mma_auto ? (g_umma_auto && prefer_umma(b, n, m) ? 22 : 20) : (queries < 148LL * 2 * 512 ? 4 : 1);
  }
  if (variant == 20) return launch_fwd_mma(a, mode, st);
  if (variant == 21) return launch_fwd_mma_persist(a, mode, st);
  if (variant == 22) return launch_fwd_umma(a, mode, st);
  if (variant == 23) return launch_fwd_mma_balanced(a, mode, st);
  if (variant == 24) return launch_fwd_mma_ws(a, mode, st);
  switch (variant) {
#define GA_FWD_CASE(ID, TH, QQ, TT, CC)                                                          \
  case ID: {                                                                                     \
    using Cfg = FwdCfg<TH, QQ, TT, CC>;                                                          \
    a.tiles1 = (n + Cfg::kQT - 1) / Cfg::kQT;                                                    \
    a.tiles2 = (m + Cfg::kQT - 1) / Cfg::kQT;                                                    \
    return launch_fwd<Cfg>(a, mode, st);                                                         \
  }
    GA_FWD_CASE(2, 64, 4, 64, 2048)
    GA_FWD_CASE(3, 64, 2, 32, 2048)
    GA_FWD_CASE(4, 128, 2, 32, 2048)
    GA_FWD_CASE(5, 64, 4, 32, 2048)
    GA_FWD_CASE(6, 256, 4, 32, 2048)
    GA_FWD_CASE(7, 64, 4, 16, 2048)
    GA_FWD_CASE(8, 32, 4, 32, 2048)
    GA_FWD_CASE(9, 64, 2, 32, 1024)
    GA_FWD_CASE(10, 128, 3, 32, 2048)
    GA_FWD_CASE(11, 64, 4, 32, 1024)
    GA_FWD_CASE(12, 32, 4, 32, 1024)
    GA_FWD_CASE(13, 64, 3, 32, 2048)
    GA_FWD_CASE(14, 96, 4, 32, 2048)
    GA_FWD_CASE(15, 160, 4, 32, 2048)
    default:
    GA_FWD_CASE(1, 128, 4, 32, 2048)
#undef GA_FWD_CASE
  }
}
}  // namespace ga

extern "C" int ga_nn_distance_fwd(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1,
                                  int* idx1, float* dist2, int* idx2, int mode, ga_stream_t stream) {
  return ga::nn_distance_fwd_mirrored(b, n, m, xyz1, xyz2, dist1, idx1, dist2, idx2, nullptr, nullptr, nullptr,
                                      nullptr, mode, stream);
}
