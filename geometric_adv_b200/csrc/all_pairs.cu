// All-pairs Chamfer matrix: the driver loop of
// attacker/prepare_indices_for_attack.py:104-139 as one kernel.
//
// Reference: tiles two (all,100,2048,3) host arrays, feeds 100 cloud pairs per
// sess.run through nn_distance, reduces mean(d_s_t,1)+mean(d_t_s,1), and repeats
// 44 x 438 times.  Here the directed term
//     D[a -> b] = mean over points p of cloud a of min_q |p - q|^2, q in cloud b
// is computed by CTAs that stage cloud b ONCE in shared memory and stream a block
// of source clouds through it (all clouds stay L2-resident: 2,000 x 24 KB = 49 MB);
// no per-point dist/idx array is ever written.  CD[i,j] = D[j->i] + D[i->j].
// Same filter-and-refine search as nn_distance_fwd.cu, so every per-point distance
// is bit-identical to the reference arithmetic; the per-cloud mean uses ONE fixed
// summation order in every kernel of the library (cloud_mean_sum below = the order of
// chamfer_per_cloud_kernel, reduce.cu), so the matrix is bit-identical whichever kernel,
// block size or row split produced it, and bit-identical to
// ga_chamfer_per_cloud(ga_nn_distance_fwd(...)).  (The reference's own reduce_mean order
// is unpinned: tolerance 1e-6 in the tests.)
#include <atomic>

#include "nn_mma.cuh"

namespace ga {

struct PairArgs {
  int s, n;             // clouds, points per cloud
  const float* clouds;  // (s,n,3)
  int a0, na;           // source clouds [a0, a0+na)
  int b0, nb;           // target clouds [b0, b0+nb)
  float* out;           // out[(a-a0)*stride_a + (b-b0)*stride_b] = D[a -> b]
  long long stride_a, stride_b;
  int ablk;             // source clouds per CTA
  int accumulate;       // add to out instead of overwriting (second direction of a CD row block)
};

using PairCfg = FwdCfg<128, 4, 32, 2048>;

// Canonical sum of the n per-point distances d[] (shared memory) of one cloud: 256 virtual
// threads v each add the slice d[v], d[v+256], ... in ascending order, each virtual warp folds its
// 32 slices with an xor-shuffle tree, and the 8 warp sums are added in ascending order -- exactly
// chamfer_per_cloud_kernel (reduce.cu).  Call with all THREADS (128 or 256) threads of the CTA,
// after a barrier that makes d[] visible; `red` holds 8 floats.  The result is valid in thread 0.
template <int THREADS>
__device__ __forceinline__ float cloud_mean_sum(const float* __restrict__ d, int n, float* __restrict__ red, int tid) {
  static_assert(THREADS == 128 || THREADS == 256, "virtual thread mapping");
  constexpr int V = 256 / THREADS;  // virtual threads per thread: v = tid + u * THREADS (warp w + u * THREADS / 32)
  float s[V];
#pragma unroll
  for (int u = 0; u < V; u++) {
    s[u] = 0.0f;
    for (int j = tid + u * THREADS; j < n; j += 256) s[u] += d[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    if ((tid & 31) == 0) red[(tid >> 5) + u * (THREADS / 32)] = s[u];
  }
  __syncthreads();
  float t = 0.0f;
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < 8; w++) t += red[w];
  }
  return t;
}

template <int MODE>
__global__ void __launch_bounds__(PairCfg::kThreads) all_pairs_directed_kernel(const PairArgs a) {
  using Cfg = PairCfg;
  constexpr int THREADS = Cfg::kThreads, Q = Cfg::kQ, QT = Cfg::kQT, T = Cfg::kT, CH = Cfg::kCH;
  extern __shared__ float4 smem_f4[];
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem_f4 + CH + 2 * kPipeU);  // [32]
  float* dpt = red + 32;                                               // [CH] per-point distances of one source cloud
  const int tid = threadIdx.x;
  const int n = a.n;
  const int nblk = (a.na + a.ablk - 1) / a.ablk;
  const int bj = blockIdx.x / nblk;            // target cloud (slow index: neighbours share sources in L2)
  const int ab = blockIdx.x - bj * nblk;
  const int b = a.b0 + bj;
  const float* tpts = a.clouds + (size_t)b * n * 3;
  const int ntile = (n + T - 1) / T;
  const float bm = stage_targets<THREADS, T>(tgt, red, tpts, 0, n, ntile, tid);
  const int qtiles = (n + QT - 1) / QT;
  const int a_end = min(a.na, (ab + 1) * a.ablk);
  for (int ai = ab * a.ablk; ai < a_end; ai++) {
    const float* qpts = a.clouds + (size_t)(a.a0 + ai) * n * 3;
    __syncthreads();  // dpt[] / red[] free (staging / previous cloud done)
    for (int qt = 0; qt < qtiles; qt++) {
      QueryState<Q> s;
      load_queries<Cfg, MODE>(s, qpts, n, qt, tpts, tid);
      search_chunk<Cfg, MODE>(s, tgt, 0, n, ntile, bm);
#pragma unroll
      for (int j = 0; j < Q; j++) {
        if (!s.valid[j]) continue;
        float d;
        int i;
        finish_query<Q>(s, j, d, i);
        dpt[qt * QT + j * THREADS + tid] = d;
      }
    }
    __syncthreads();
    const float t = cloud_mean_sum<THREADS>(dpt, n, red, tid);
    if (tid == 0) {
      float* o = a.out + (long long)ai * a.stride_a + (long long)bj * a.stride_b;
      const float v = t / (float)n;
      *o = a.accumulate ? *o + v : v;
    }
  }
}


// ---- tensor-core variant (default for clouds of 256..2048 points) ---------------------------------
// Same decomposition, filter scan of nn_mma.cuh: 8 warps per CTA, the target cloud staged once as
// pair-SoA + B fragments (96 KB, two CTAs per SM), every warp takes the 64-query jobs
// warp, warp + 8, ... of each source cloud.  Per-point distances are the same bits as in every other
// forward kernel; the per-cloud sum runs over (job, lane slot) in a fixed order.
constexpr int kPairMmaWarps = 8;
constexpr int kPairMmaCH = 2048;
constexpr size_t kPairMmaOffRed = (size_t)kPairMmaCH * 16 + (size_t)kPipeU * 32;
constexpr size_t kPairMmaOffB = kPairMmaOffRed + 32 * 4;
constexpr size_t kPairMmaOffCnt = kPairMmaOffB + (size_t)kPairMmaCH * 32;
constexpr size_t kPairMmaOffTile = kPairMmaOffCnt + (size_t)kPairMmaWarps * kMmaQW * 4;
constexpr size_t kPairMmaOffDpt = kPairMmaOffTile + (size_t)kPairMmaWarps * kMmaQW * 4;
constexpr size_t kPairMmaSmem = kPairMmaOffDpt + (size_t)kPairMmaCH * 4;

// One 64-query job: stores this lane's (up to two) exact distances in dpt[].  Not inlined: inside
// the caller's loops ptxas serialises the HMMAs of the scan on one accumulator quad (see persist_job).
template <int MODE>
__device__ __noinline__ void pair_job(const float* __restrict__ qpts, const float* __restrict__ tpts, int n, int qbase,
                                      const float4* __restrict__ tgt, const uint4* __restrict__ bfrag, float bm,
                                      int* wcnt, unsigned short* wtile, float* __restrict__ dpt) {
  const int lane = threadIdx.x & 31;
  MmaRows R;
  mma_load_rows(R, qpts, n, qbase, lane);
  QueryState<2> s;
  mma_init_queries<MODE>(s, qpts, n, qbase, tpts, lane);
  float mrun[8];
#pragma unroll
  for (int r = 0; r < 8; r++) mrun[r] = kMmaBig;
  mma_chunk<MODE, true>(R, s, mrun, tgt, bfrag, 0, n, n, bm, wcnt, wtile, lane);  // batched accumulators: see mma_scan_range
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (!s.valid[j]) continue;
    float d;
    int i;
    finish_query<2>(s, j, d, i);
    dpt[qbase + 16 * (lane & 3) + (lane >> 2) + 8 * j] = d;  // the lane's queries: see mma_init_queries
  }
}

template <int MODE>
__global__ void __launch_bounds__(kPairMmaWarps * 32, 2) all_pairs_directed_mma_kernel(const PairArgs a) {
  constexpr int THREADS = kPairMmaWarps * 32, T = kMmaT;
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem + kPairMmaOffRed);
  uint4* bfrag = reinterpret_cast<uint4*>(smem + kPairMmaOffB);
  float* dpt = reinterpret_cast<float*>(smem + kPairMmaOffDpt);
  const int tid = threadIdx.x, warp = tid >> 5;
  int* wcnt = reinterpret_cast<int*>(smem + kPairMmaOffCnt) + warp * kMmaQW;
  unsigned short* wtile = reinterpret_cast<unsigned short*>(smem + kPairMmaOffTile) + warp * kMmaQW * 2;
  const int n = a.n;
  const int nblk = (a.na + a.ablk - 1) / a.ablk;
  const int bj = blockIdx.x / nblk;  // target cloud (slow index: neighbours share sources in L2)
  const int ab = blockIdx.x - bj * nblk;
  const float* tpts = a.clouds + (size_t)(a.b0 + bj) * n * 3;
  const float bm = stage_targets<THREADS, T>(tgt, red, tpts, 0, n, (n + T - 1) / T, tid);
  stage_bfrag<THREADS>(bfrag, tgt, (n + kMmaBlk - 1) / kMmaBlk, n, tid);
  __syncthreads();
  const int jobs = (n + kMmaQW - 1) / kMmaQW;
  const int a_end = min(a.na, (ab + 1) * a.ablk);
  for (int ai = ab * a.ablk; ai < a_end; ai++) {
    const float* qpts = a.clouds + (size_t)(a.a0 + ai) * n * 3;
    __syncthreads();  // dpt[] / red[] free (previous cloud done)
    for (int job = warp; job < jobs; job += kPairMmaWarps)
      pair_job<MODE>(qpts, tpts, n, job * kMmaQW, tgt, bfrag, bm, wcnt, wtile, dpt);
    __syncthreads();
    const float t = cloud_mean_sum<THREADS>(dpt, n, red, tid);
    if (tid == 0) {
      float* o = a.out + (long long)ai * a.stride_a + (long long)bj * a.stride_b;
      const float v = t / (float)n;
      *o = a.accumulate ? *o + v : v;
    }
  }
}

int g_pairs_kernel = 0;  // tuning hook (key 16): 0 auto (tensor-core scan from 256 points), 1 = fp32 filter scan
int g_pairs_ablk = 0;    // tuning hook (key 19): source clouds per CTA, 0 = auto (choose_ablk)

static int launch_directed(const PairArgs& a, int mode, cudaStream_t st) {
  if (a.na <= 0 || a.nb <= 0) return GA_OK;
  const long long ctas = (long long)a.nb * ((a.na + a.ablk - 1) / a.ablk);
  if (ctas > 0x7fffffffLL) {
    set_error("ga_chamfer_all_pairs: problem too large for one launch");
    return GA_ERR_UNSUPPORTED;
  }
  if (g_pairs_kernel != 1 && a.n >= 256 && a.n <= kPairMmaCH) {
    auto km = mode == GA_MODE_CPU_EXACT ? all_pairs_directed_mma_kernel<GA_MODE_CPU_EXACT>
                                        : all_pairs_directed_mma_kernel<GA_MODE_GPU_REF>;
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairMmaSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
    km<<<(unsigned)ctas, kPairMmaWarps * 32, kPairMmaSmem, st>>>(a);
    GA_LAUNCH_CHECK("all_pairs_directed_mma_kernel");
    return GA_OK;
  }
  auto k = mode == GA_MODE_CPU_EXACT ? all_pairs_directed_kernel<GA_MODE_CPU_EXACT>
                                     : all_pairs_directed_kernel<GA_MODE_GPU_REF>;
  constexpr size_t kPairSmem = PairCfg::kSmem + (size_t)PairCfg::kCH * 4;  // + dpt[]
  GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmem));
  k<<<(unsigned)ctas, PairCfg::kThreads, kPairSmem, st>>>(a);
  GA_LAUNCH_CHECK("all_pairs_directed_kernel");
  return GA_OK;
}

static int choose_ablk(int na, int nb) {
  if (g_pairs_ablk > 0) return g_pairs_ablk;
  // enough CTAs to fill the machine several times over, as much reuse of the staged cloud as that allows
  const long long want = (long long)sm_count() * 16;
  int ablk = 16;
  while (ablk > 1 && (long long)nb * ((na + ablk - 1) / ablk) < want) ablk >>= 1;
  return ablk;
}

static int check_args(const char* who, int s, int n, const void* clouds, int row0, int rows, const void* out,
                      int mode) {
  if (s < 0 || n < 0 || row0 < 0 || rows < 0 || row0 + rows > s) {
    set_error("%s: bad sizes (s=%d n=%d row0=%d rows=%d)", who, s, n, row0, rows);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (mode != GA_MODE_CPU_EXACT && mode != GA_MODE_GPU_REF) {
    set_error("%s: unknown mode %d", who, mode);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (rows > 0 && s > 0 && (clouds == nullptr || out == nullptr)) {
    set_error("%s: null pointer", who);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (n > PairCfg::kCH) {
    set_error("%s: clouds larger than %d points are not supported by the all-pairs kernel (n=%d); "
              "use ga_nn_distance_fwd on batches", who, PairCfg::kCH, n);
    return GA_ERR_UNSUPPORTED;
  }
  return GA_OK;
}

}  // namespace ga

using namespace ga;

extern "C" {

int ga_chamfer_all_pairs_directed(int s, int n, const float* clouds, int row0, int rows, float* out, int mode,
                                  ga_stream_t stream) {
  int rc = check_args("ga_chamfer_all_pairs_directed", s, n, clouds, row0, rows, out, mode);
  if (rc != GA_OK) return rc;
  if (rows == 0 || s == 0) return GA_OK;
  cudaStream_t st = as_stream(stream);
  if (n == 0) {  // mean over nothing: the reference would produce NaN (0/0)
    GA_CUDA_TRY(cudaMemsetAsync(out, 0xff, sizeof(float) * (size_t)rows * s, st));
    return GA_OK;
  }
  PairArgs a;
  a.s = s; a.n = n; a.clouds = clouds;
  a.a0 = row0; a.na = rows; a.b0 = 0; a.nb = s;
  a.out = out; a.stride_a = s; a.stride_b = 1;
  a.ablk = choose_ablk(rows, s);
  a.accumulate = 0;
  return launch_directed(a, mode, st);
}

int ga_chamfer_all_pairs(int s, int n, const float* clouds, int row0, int rows, float* out, int mode,
                         ga_stream_t stream) {
  int rc = check_args("ga_chamfer_all_pairs", s, n, clouds, row0, rows, out, mode);
  if (rc != GA_OK) return rc;
  if (rows == 0 || s == 0) return GA_OK;
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    GA_CUDA_TRY(cudaMemsetAsync(out, 0xff, sizeof(float) * (size_t)rows * s, st));
    return GA_OK;
  }
  // out[r, j] = CD(source = clouds[j], target = clouds[row0+r]) = D[j -> i] + D[i -> j], i = row0 + r
  PairArgs a;
  a.s = s; a.n = n; a.clouds = clouds; a.out = out;
  // first D[j -> i]: sources j = all clouds, targets i = the row block
  a.a0 = 0; a.na = s; a.b0 = row0; a.nb = rows;
  a.stride_a = 1; a.stride_b = s;
  a.ablk = choose_ablk(s, rows);
  a.accumulate = 0;
  rc = launch_directed(a, mode, st);
  if (rc != GA_OK) return rc;
  // then += D[i -> j]
  a.a0 = row0; a.na = rows; a.b0 = 0; a.nb = s;
  a.stride_a = s; a.stride_b = 1;
  a.ablk = choose_ablk(rows, s);
  a.accumulate = 1;
  return launch_directed(a, mode, st);
}

}  // extern "C"
