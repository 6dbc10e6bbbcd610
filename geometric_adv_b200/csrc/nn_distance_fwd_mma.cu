// Chamfer forward with the filter scan on the tensor cores (see nn_mma.cuh for the operand
// split and the error bound).  Same contract and same bits as nn_fwd_kernel
// (tf_nndistance.cpp:21-43 / tf_nndistance_g.cu:5-131): the tensor cores only decide WHICH
// 32-target tiles can hold the reference argmin; those tiles (one per query, rarely two) are
// then walked with the fp32 filter and the survivors evaluated in the reference arithmetic
// with the reference's strict-< / lowest-index rule, exactly as in nn_search.cuh.
//
// Work decomposition: CTA = (batch element, direction, 64*WARPS queries).  A warp owns 64
// queries as four 16-row A fragments that stay in registers; the staged target chunk is
// streamed as B fragments (one conflict-free LDS.64 per 8 targets) through 4 HMMA.16816 with a
// zero accumulator, and the 16 results per lane are folded into 8 running minima with FMNMX3.
// The B columns are permuted so that each lane's results over 16 consecutive n-tiles are 32
// contiguous targets: a lane holds complete tile minima and no shuffle is needed in the scan.
#include <atomic>

#include "nn_mma.cuh"

namespace ga {

template <int WARPS, int CH>
struct MmaCfg {
  static constexpr int kWarps = WARPS;
  static constexpr int kThreads = WARPS * 32;
  static constexpr int kQT = WARPS * kMmaQW;  // queries per CTA
  static constexpr int kCH = CH;              // targets staged per chunk
  static_assert(CH % kMmaBlk == 0, "chunk must hold whole MMA blocks");
  // pair-SoA (+pipeline pad) | red[32] | B fragments | per-query tile lists (count, 2 tiles)
  static constexpr size_t kOffRed = (size_t)CH * 16 + (size_t)kPipeU * 32;
  static constexpr size_t kOffB = kOffRed + 32 * 4;
  static constexpr size_t kOffCnt = kOffB + (size_t)CH * 32;
  static constexpr size_t kOffTile = kOffCnt + (size_t)kQT * 4;
  static constexpr size_t kSmem = kOffTile + (size_t)kQT * 4;
};

template <class Cfg, int MODE, int MINB>
__global__ void __launch_bounds__(Cfg::kThreads, MINB) nn_fwd_mma_kernel(const FwdArgs a) {
  constexpr int THREADS = Cfg::kThreads, QT = Cfg::kQT, CH = Cfg::kCH, T = kMmaT;
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem + Cfg::kOffRed);
  uint4* bfrag = reinterpret_cast<uint4*>(smem + Cfg::kOffB);
  int* lcnt = reinterpret_cast<int*>(smem + Cfg::kOffCnt);
  unsigned short* ltile = reinterpret_cast<unsigned short*>(smem + Cfg::kOffTile);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int jpb = a.tiles1 + a.tiles2;
  const int batch = blockIdx.x / jpb;
  const int r0 = blockIdx.x - batch * jpb;
  const bool rev = r0 >= a.tiles1;
  const int qtile = rev ? r0 - a.tiles1 : r0;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  float* odist = (rev ? a.dist2 : a.dist1) + (size_t)batch * nq;
  int* oidx = (rev ? a.idx2 : a.idx1) + (size_t)batch * nq;
  float* mdist = rev ? a.mdist2 : a.mdist1;
  int* midx = rev ? a.midx2 : a.midx1;

  const int qbase = qtile * QT + warp * kMmaQW;  // first query of this warp
  MmaRows R;
  mma_load_rows(R, qpts, nq, qbase, lane);

  // The two queries this lane refines and writes: m-tile t, rows g and g+8.
  QueryState<2> s;
  {
    const float kInf = __int_as_float(0x7f800000);
    const float t0x = __ldg(tpts), t0y = __ldg(tpts + 1), t0z = __ldg(tpts + 2);
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int qi = qbase + 16 * t + g + 8 * j;
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
  int* wcnt = lcnt + warp * kMmaQW;
  unsigned short* wtile = ltile + warp * kMmaQW * 2;

  float mrun[8];  // running minimum of h per row over all chunks (same in the 4 lanes of a quad)
#pragma unroll
  for (int r = 0; r < 8; r++) mrun[r] = __int_as_float(0x7f800000);
  float bm_run = 0.0f;

  for (int c0 = 0; c0 < nt; c0 += CH) {
    const int cn = min(CH, nt - c0);
    const int ntile = (cn + T - 1) / T;
    const int nblk = (cn + kMmaBlk - 1) / kMmaBlk;
    bm_run = fmaxf(bm_run, stage_targets<THREADS, T>(tgt, red, tpts, c0, nt, ntile, tid));
    stage_bfrag<THREADS>(bfrag, tgt, nblk, cn, tid);
    wcnt[16 * t + g] = 0;
    wcnt[16 * t + g + 8] = 0;
    __syncthreads();

    MmaTrack tr;
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
        if (slot < 2) wtile[2 * ql + slot] = (unsigned short)tr.i1[r];
      }
      if (!(tr.c2[r] > thr)) {
        const int slot = atomicAdd(&wcnt[ql], 1);
        if (slot < 2) wtile[2 * ql + slot] = (unsigned short)tr.i2[r];
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
      // a tile id beyond the staged tiles can only come from padding (h = +inf) under a
      // non-finite window: exact scan
      if ((cnt[j] >= 1 && ta[j] >= ntile) || (cnt[j] >= 2 && tb[j] >= ntile)) cnt[j] = 3;
    }
    refine_tiles<T, MODE, 2>(s, tgt, c0, nt, ntile, cnt, ta, tb, mythr);
  }

#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (!s.valid[j]) continue;
    const int qi = qbase + 16 * t + g + 8 * j;
    float d;
    int i;
    finish_query<2>(s, j, d, i);
    odist[qi] = d;
    oidx[qi] = i;
    if (mdist != nullptr) {
      mdist[(size_t)batch * nq + qi] = d;
      midx[(size_t)batch * nq + qi] = i;
    }
  }
}

// Debug / evidence: the raw tensor-core filter values h(q,t) of one cloud pair (n queries,
// m <= 2048 targets), out[q*m + t].  Uses the same staging, fragments and column map as the
// product kernel; tests compare it with the fp64 value to check the layout and the bound e2.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads) mma_filter_dump_kernel(int n, int m, const float* __restrict__ q,
                                                                        const float* __restrict__ tp,
                                                                        float* __restrict__ out) {
  constexpr int THREADS = Cfg::kThreads, QT = Cfg::kQT, T = kMmaT;
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem + Cfg::kOffRed);
  uint4* bfrag = reinterpret_cast<uint4*>(smem + Cfg::kOffB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int qbase = blockIdx.x * QT + warp * kMmaQW;
  MmaRows R;
  mma_load_rows(R, q, n, qbase, lane);
  const int ntile = (m + T - 1) / T, nblk = (m + kMmaBlk - 1) / kMmaBlk;
  stage_targets<THREADS, T>(tgt, red, tp, 0, m, ntile, tid);
  stage_bfrag<THREADS>(bfrag, tgt, nblk, m, tid);
  __syncthreads();
  const uint2* bf2 = reinterpret_cast<const uint2*>(bfrag);
  for (int blk = 0; blk < nblk; blk++) {
    for (int j = 0; j < 16; j++) {
      const uint2 bf = bf2[((size_t)blk * 16 + j) * 32 + lane];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float c[4];
        mma16816(c, R.a[i], bf.x, bf.y);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int qi = qbase + 16 * i + g + 8 * (e >> 1);
          const int tau = mma_col_target(blk, j, 2 * t + (e & 1));
          if (qi < n && tau < m) out[(size_t)qi * m + tau] = c[e];
        }
      }
    }
  }
}

int g_mma_cfg = 0;  // tuning hook (key 7): 0 auto, 1 = 4 warps x 2 CTAs/SM, 2 = 8 warps, 3 = 4 warps x 1024-target chunks

template <class Cfg, int MINB>
static int launch_fwd_mma_cfg(FwdArgs a, int mode, cudaStream_t st) {
  a.tiles1 = (a.n + Cfg::kQT - 1) / Cfg::kQT;
  a.tiles2 = (a.m + Cfg::kQT - 1) / Cfg::kQT;
  const long long jobs = (long long)a.b * (a.tiles1 + a.tiles2);
  if (jobs <= 0) return GA_OK;
  if (jobs > 0x7fffffffLL) {
    set_error("ga_nn_distance_fwd: problem too large for one launch (%lld CTAs)", jobs);
    return GA_ERR_UNSUPPORTED;
  }
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_mma_kernel<Cfg, GA_MODE_CPU_EXACT, MINB>
                                     : nn_fwd_mma_kernel<Cfg, GA_MODE_GPU_REF, MINB>;
  {
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  k<<<(unsigned)jobs, Cfg::kThreads, Cfg::kSmem, st>>>(a);
  GA_LAUNCH_CHECK("nn_fwd_mma_kernel");
  return GA_OK;
}

int launch_fwd_mma(const FwdArgs& a, int mode, cudaStream_t st) {
  int cfg = g_mma_cfg;
  if (cfg == 0) cfg = 1;
  switch (cfg) {
    case 2:
      return launch_fwd_mma_cfg<MmaCfg<8, 2048>, 1>(a, mode, st);
    case 3:
      return launch_fwd_mma_cfg<MmaCfg<4, 1024>, 2>(a, mode, st);
    case 4:
      return launch_fwd_mma_cfg<MmaCfg<4, 1024>, 4>(a, mode, st);
    case 5:
      return launch_fwd_mma_cfg<MmaCfg<8, 2048>, 2>(a, mode, st);
    default:
      return launch_fwd_mma_cfg<MmaCfg<4, 2048>, 2>(a, mode, st);
  }
}

}  // namespace ga

extern "C" int ga_debug_mma_filter(int n, int m, const float* xyz1, const float* xyz2, float* out,
                                   ga_stream_t stream) {
  using Cfg = ga::MmaCfg<4, 2048>;
  if (n <= 0 || m <= 0 || m > Cfg::kCH) {
    ga::set_error("ga_debug_mma_filter: need n > 0 and 0 < m <= %d", Cfg::kCH);
    return GA_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = ga::as_stream(stream);
  auto k = ga::mma_filter_dump_kernel<Cfg>;
  GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
  k<<<(n + Cfg::kQT - 1) / Cfg::kQT, Cfg::kThreads, Cfg::kSmem, st>>>(n, m, xyz1, xyz2, out);
  GA_LAUNCH_CHECK("mma_filter_dump_kernel");
  return GA_OK;
}
