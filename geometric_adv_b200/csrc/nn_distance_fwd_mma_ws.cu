// Chamfer forward, tensor-core filter, WARP-SPECIALISED persistent kernel (variant 24).
//
// Same contract and same bits as nn_fwd_mma_kernel (tf_nndistance.cpp:21-43 / tf_nndistance_g.cu:5-131; filter,
// window and refine of nn_mma.cuh).  What changes is who does what, and when:
//
// nn_fwd_mma_kernel runs stage -> scan -> refine per CTA.  Its scan is bound by the legacy HMMA pipe (0.5
// HMMA.16816 per clock and SM, 0.40-0.42 with the zero accumulator), and EIGHT warps already saturate that pipe
// (tools/mmabench.cu: 0.40 with 8 warps per SM, 0.42 with 16).  Two co-resident CTAs therefore scan at half speed
// each and then refine side by side while the tensor pipe idles: the pipe is busy 42 % of the kernel at B=50 and
// 70 % at B=512.  Here one CTA of 16 warps per SM splits the roles:
//   * NS scan warps do nothing but stream B fragments through the tensor cores, job after job (a job = 64 queries
//     against one staged cloud), and leave per-query candidate lists (count, two tiles, threshold) in a 2-deep ring
//     of shared-memory slots;
//   * NH helper warps refine those lists in the reference arithmetic and write dist / idx, and between refines stage
//     the NEXT target cloud (pair-SoA + B fragments) into the other of two 96 KB buffers,
// so that the tensor pipe never waits for staging or refine after the prologue.  Every SM owns an equal, contiguous
// share of all (batch, direction, 64-query) jobs; scan warp w takes jobs j0 + w, j0 + w + NS, ... of it.
//
// Synchronisation (no CTA barrier after the prologue):
//   staged[p]      counter: helper warps that finished staging a cloud of parity p  (cloud c ready: >= NH (c/2 + 1))
//   refined[p]     counter: jobs of parity-p clouds refined and written             (buffer free: all earlier jobs)
//   full[w][s]     mbarrier: scan warp w has written a job into its slot s          (the helper sleeps on it)
//   done[w][s]     mbarrier: the helper has consumed that slot                      (scan warp: slot free again)
// The counters are monotonic (a warp may skip clouds, which a phase parity cannot express); the slot barriers have
// one producer and one consumer that use every phase in order.
// Clouds of at most 2048 points (one staged chunk per cloud).
#include <atomic>

#include "nn_mma.cuh"

namespace ga {

constexpr int kWsWarps = 16;
constexpr int kWsCH = 2048;
constexpr size_t kWsBuf = (size_t)kWsCH * 16 + (size_t)kPipeU * 32 + (size_t)kWsCH * 32;  // pair-SoA | pad | B fragments
constexpr size_t kWsOffRed = 2 * kWsBuf;                   // [2][16] float: max |coordinate| per staged block
constexpr size_t kWsOffCtl = kWsOffRed + 2 * 16 * 4;       // staged[2], refined[2] (int) | full[16][2], done[16][2] (mbarrier)
constexpr size_t kWsCtlBytes = 16 + 4 * kWsWarps * 8;
constexpr size_t kWsOffSlot = kWsOffCtl + ((kWsCtlBytes + 15) / 16) * 16;
constexpr size_t kWsSlotBytes = 64 * 4 + 64 * 2 * 2 + 64 * 4;  // cnt[64] int | tile[64][2] u16 | thr[64] float
constexpr size_t kWsSmem = kWsOffSlot + (size_t)kWsWarps * 2 * kWsSlotBytes;

// The share of one CTA, in 32-bit arithmetic (the launcher refuses J >= 2^31).
struct WsShare {
  int j0, j1, jpb, wt1, wt2, g0;  // jobs [j0, j1), jobs per batch element, per direction, first cloud (2 batch + dir)
};
struct WsJob {
  int batch, rev, qbase, c;  // c = cloud sequence number within the share (buffer c & 1)
};
__device__ __forceinline__ WsJob ws_job(const WsShare& sh, int j) {
  WsJob q;
  q.batch = j / sh.jpb;
  const int r = j - q.batch * sh.jpb;
  q.rev = r >= sh.wt1;
  q.qbase = (q.rev ? r - sh.wt1 : r) * kMmaQW;
  q.c = 2 * q.batch + q.rev - sh.g0;
  return q;
}
// jobs of cloud sequence number c that lie in the share
__device__ __forceinline__ int ws_cloud_jobs(const WsShare& sh, int c) {
  const int g = sh.g0 + c, batch = g >> 1, rev = g & 1;
  const int s = batch * sh.jpb + (rev ? sh.wt1 : 0), e = s + (rev ? sh.wt2 : sh.wt1);
  return min(e, sh.j1) - max(s, sh.j0);
}

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
// lane 0 polls until *p >= v; the warp leaves converged and ordered behind the writer's fence
__device__ __forceinline__ void ws_wait_ge(const int* p, int v, int lane) {
  if (lane == 0) {
    while (ld_volatile(p) < v) __nanosleep(20);
  }
  __syncwarp();
  __threadfence_block();
}

// Scan side of one job.  NOT inlined: inside the role loop ptxas would serialise the four HMMAs of a scan step on one
// accumulator quad (see persist_job in nn_distance_fwd_mma.cu); as a function the scan keeps its rotating quads.
__device__ __noinline__ void ws_scan_job(const float* __restrict__ qpts, int nq, int qbase,
                                         const uint2* __restrict__ bfrag, int nblk, float bm, int* __restrict__ scnt,
                                         unsigned short* __restrict__ stile, float* __restrict__ sthr, int dev) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  MmaRows R;
  if (dev & 2) {  // development: no global loads (wrong results)
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int r = 0; r < 4; r++) R.a[i][r] = 0x3f803f80u + lane + i * 7 + r + qbase;
#pragma unroll
    for (int r = 0; r < 8; r++) R.qabs[r] = 0.5f;
  } else {
    mma_load_rows(R, qpts, nq, qbase, lane);
  }
  scnt[lane] = 0;
  scnt[lane + 32] = 0;
  __syncwarp();
  MmaTrack tr;
  mma_scan(R, bfrag, nblk, lane, tr);
  // row minimum over the quad, window, qualifying tiles of this lane -> per-query lists (as mma_chunk)
#pragma unroll
  for (int r = 0; r < 8; r++) {
    float m = tr.c1[r];
    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    const float thr = m + mma_window(R.qabs[r], bm);
    const int ql = 16 * (r >> 1) + g + 8 * (r & 1);
    if (t == (r >> 1)) sthr[ql] = thr;
    if (!(tr.c1[r] > thr)) {
      const int slot = atomicAdd(&scnt[ql], 1);
      if (slot < 2) stile[2 * ql + slot] = (unsigned short)mma_key_tile(tr.c1[r], t);
    }
    if (!(tr.c2[r] > thr)) {
      const int slot = atomicAdd(&scnt[ql], 1);
      if (slot < 2) stile[2 * ql + slot] = (unsigned short)mma_key_tile(tr.c2[r], t);
    }
    if (!(tr.c3[r] > thr)) atomicAdd(&scnt[ql], 3);  // a third tile of this lane: exact scan
  }
}

// Helper side of one job: exact refine of the listed tiles and the output rows.
template <int MODE>
__device__ __noinline__ void ws_refine_job(const FwdArgs& a, int batch, bool rev, int qbase, const float4* __restrict__ tgt,
                                           const int* __restrict__ scnt, const unsigned short* __restrict__ stile,
                                           const float* __restrict__ sthr) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  QueryState<2> s;
  mma_init_queries<MODE>(s, qpts, nq, qbase, tpts, lane);
  const int ntile = (nt + kMmaT - 1) / kMmaT;
  int cnt[2], ta[2], tb[2];
  float thr[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int ql = 16 * t + g + 8 * j;
    cnt[j] = scnt[ql];
    ta[j] = stile[2 * ql];
    tb[j] = stile[2 * ql + 1];
    thr[j] = sthr[ql];
    // a tile id beyond the staged tiles can only come from padding under a non-finite window
    if ((cnt[j] >= 1 && ta[j] >= ntile) || (cnt[j] >= 2 && tb[j] >= ntile)) cnt[j] = 3;
  }
  refine_tiles<MODE>(s, tgt, 0, nt, ntile, cnt, ta, tb, thr);
  mma_write(s, qbase, nq, lane, (rev ? a.dist2 : a.dist1) + (size_t)batch * nq, (rev ? a.idx2 : a.idx1) + (size_t)batch * nq,
            rev ? a.mdist2 : a.mdist1, rev ? a.midx2 : a.midx1, (size_t)batch * nq);
}

// mbarrier helpers (CTA scope).  arrive = release, try_wait = acquire (PTX defaults).
__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ws_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ws_smem_u32(bar)) : "memory");
}
// every lane waits (warp-uniform loop); the hardware suspends the warp for up to the hint between tries
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ws_smem_u32(bar)), "r"(parity), "r"(1000000u)
        : "memory");
  } while (!ok);
}

template <int MODE, int NS>
__global__ void __launch_bounds__(kWsWarps * 32, 1)
    nn_fwd_mma_ws_kernel(const FwdArgs a, const int wt1, const int wt2, const int J, const unsigned idle_ns,
                         long long* __restrict__ dbg, const int dev) {
  constexpr int NH = kWsWarps - NS;
  static_assert(NS % NH == 0, "every helper serves NS / NH scan warps");
  constexpr int SERVE = NS / NH;
  asm volatile("griddepcontrol.launch_dependents;");
  const long long t_entry = dbg ? clock64() : 0;
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float* red = reinterpret_cast<float*>(smem + kWsOffRed);
  int* staged = reinterpret_cast<int*>(smem + kWsOffCtl);
  int* refined = staged + 2;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWsOffCtl + 16);  // [16][2] mbarriers, one arrival each
  uint64_t* done = full + 2 * kWsWarps;                                 // [16][2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  WsShare sh;
  sh.j0 = (int)((long long)J * blockIdx.x / gridDim.x);
  sh.j1 = (int)((long long)J * (blockIdx.x + 1) / gridDim.x);
  sh.jpb = wt1 + wt2;
  sh.wt1 = wt1;
  sh.wt2 = wt2;
  if (sh.j0 >= sh.j1) return;
  {
    const int b0 = sh.j0 / sh.jpb;
    sh.g0 = 2 * b0 + ((sh.j0 - b0 * sh.jpb) >= wt1 ? 1 : 0);
  }
  const int C = ws_job(sh, sh.j1 - 1).c + 1;  // clouds of this share

  auto buf_tgt = [&](int buf) { return reinterpret_cast<float4*>(smem + buf * kWsBuf); };
  auto buf_bfrag = [&](int buf) {
    return reinterpret_cast<uint4*>(smem + buf * kWsBuf + (size_t)kWsCH * 16 + (size_t)kPipeU * 32);
  };
  auto stage_block = [&](int c, int blk) {  // one warp: block `blk` of cloud sequence number c
    const int g = sh.g0 + c, batch = g >> 1, rev = g & 1;
    const int nt = rev ? a.n : a.m;
    const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
    float lmax = 0.0f;
    if (blk * kMmaBlk < nt) lmax = stage_block_warp(buf_tgt(c & 1), buf_bfrag(c & 1), tpts, nt, blk, lane);
    lmax = warp_max(lmax);
    if (lane == 0) red[(c & 1) * 16 + blk] = lmax;
  };

  // prologue: all 16 warps stage cloud 0 (one 128-target block each)
  if (tid < 4) staged[tid] = tid == 0 ? NH : 0;
  if (tid < 4 * kWsWarps) ws_mbar_init(full + tid, 1);
  stage_block(0, warp);
  __syncthreads();

  if (warp < NS) {
    // ---------------------------------------------------------------- scan warps
    unsigned char* myslots = smem + kWsOffSlot + (size_t)warp * 2 * kWsSlotBytes;
    int k = 0;
    // The two scan warps of a scheduler would run in lock step (same start, equal jobs) and reach the parts of a job
    // that do not feed the tensor pipe (query rows, lists) together; a lone warp nearly saturates its scheduler's pipe
    // (11.3 vs 10.3 clk per HMMA, tools/scanloop.cu), so the second warp starts half a job late and the pipe stays fed.
    if (warp >= 4 && idle_ns > 0) __nanosleep(idle_ns);
    long long t_wait = 0, t_work = 0, t_mark = dbg ? clock64() : 0;
    const long long t_begin = t_mark;
    for (int j = sh.j0 + warp; j < sh.j1; j += NS, k++) {
      const WsJob q = ws_job(sh, j);
      const int buf = q.c & 1, sl = k & 1, ord = k >> 1;
      ws_wait_ge(&staged[buf], NH * ((q.c >> 1) + 1), lane);         // the cloud is staged
      if (ord > 0) ws_mbar_wait(done + 2 * warp + sl, (ord - 1) & 1);  // the slot's previous job is refined
      if (dbg) {
        const long long t = clock64();
        t_wait += t - t_mark;
        t_mark = t;
      }
      float bm = 0.0f;
#pragma unroll
      for (int w = 0; w < 16; w++) bm = fmaxf(bm, red[buf * 16 + w]);
      const int nq = q.rev ? a.m : a.n;
      const int nt = q.rev ? a.n : a.m;
      const float* qpts = (q.rev ? a.xyz2 : a.xyz1) + (size_t)q.batch * nq * 3;
      unsigned char* slot = myslots + (size_t)sl * kWsSlotBytes;
      if (j + NS < sh.j1 && lane < 7) {  // the next job's 64 queries (768 B: at most 7 lines) on their way to L1
        const WsJob qn = ws_job(sh, j + NS);
        const int nqn = qn.rev ? a.m : a.n;
        const char* pn = reinterpret_cast<const char*>((qn.rev ? a.xyz2 : a.xyz1) + ((size_t)qn.batch * nqn + qn.qbase) * 3);
        const char* pe = reinterpret_cast<const char*>((qn.rev ? a.xyz2 : a.xyz1) + ((size_t)qn.batch * nqn + nqn) * 3);
        const char* pl = pn + 128 * lane;
        if (pl < pn + 768 && pl < pe) asm volatile("prefetch.global.L1 [%0];" ::"l"(pl));
      }
      ws_scan_job(qpts, nq, q.qbase, reinterpret_cast<const uint2*>(buf_bfrag(buf)), (nt + kMmaBlk - 1) / kMmaBlk, bm,
                  reinterpret_cast<int*>(slot), reinterpret_cast<unsigned short*>(slot + 256),
                  reinterpret_cast<float*>(slot + 512), dev);
      __syncwarp();
      if (lane == 0) ws_mbar_arrive(full + 2 * warp + sl);
      if (dbg) {
        const long long t = clock64();
        t_work += t - t_mark;
        t_mark = t;
      }
    }
    if (dbg && lane == 0) {
      long long* d = dbg + ((size_t)blockIdx.x * kWsWarps + warp) * 4;
      d[0] = t_wait;
      d[1] = t_work;
      d[2] = t_begin - t_entry;  // prologue
      d[3] = clock64() - t_begin;
    }
  } else {
    // ---------------------------------------------------------------- helper warps
    // Program order of a helper: its jobs in ascending order; before a job of cloud c it makes sure that its blocks
    // of cloud c + 1 are staged (which waits until the buffer's previous cloud, c - 1, is refined by everybody --
    // a dependency on strictly earlier clouds only).  All idle time is spent in the mbarrier wait for a full slot.
    const int h = warp - NS;
    int sc = 1;                              // next cloud to stage
    int cum[2] = {ws_cloud_jobs(sh, 0), 0};  // jobs of the clouds already passed, per parity
    long long t_wait = 0, t_work = 0, t_stage = 0, t_mark = dbg ? clock64() : 0;
    const long long t_begin = t_mark;
    auto lap = [&](long long& acc) {
      if (dbg) {
        const long long t = clock64();
        acc += t - t_mark;
        t_mark = t;
      }
    };
    auto stage_upto = [&](int upto) {
      while (sc < C && sc <= upto) {
        if (lane == 0) {
          while (ld_volatile(&refined[sc & 1]) < cum[sc & 1]) __nanosleep(200);
        }
        __syncwarp();
        __threadfence_block();
        for (int blk = h; blk < 16; blk += NH) stage_block(sc, blk);
        __syncwarp();
        __threadfence_block();
        if (lane == 0) atomicAdd(&staged[sc & 1], 1);
        cum[sc & 1] += ws_cloud_jobs(sh, sc);
        sc++;
        lap(t_stage);
      }
    };
    stage_upto(1);
    for (int k = 0;; k++) {
      bool any = false;
#pragma unroll
      for (int i = 0; i < SERVE; i++) {
        const int sw = h + i * NH;  // scan warp served
        const int j = sh.j0 + sw + k * NS;
        if (j >= sh.j1) break;
        any = true;
        const WsJob q = ws_job(sh, j);
        stage_upto(q.c + 1);
        const int sl = k & 1, ord = k >> 1;
        ws_mbar_wait(full + 2 * sw + sl, ord & 1);
        lap(t_wait);
        const unsigned char* slot = smem + kWsOffSlot + ((size_t)sw * 2 + sl) * kWsSlotBytes;
        if (!(dev & 1))  // development: bit 0 = no refine (wrong results)
        ws_refine_job<MODE>(a, q.batch, q.rev != 0, q.qbase, buf_tgt(q.c & 1), reinterpret_cast<const int*>(slot),
                            reinterpret_cast<const unsigned short*>(slot + 256),
                            reinterpret_cast<const float*>(slot + 512));
        __syncwarp();
        __threadfence_block();
        if (lane == 0) {
          ws_mbar_arrive(done + 2 * sw + sl);
          atomicAdd(&refined[q.c & 1], 1);
        }
        lap(t_work);
      }
      if (!any) break;
    }
    stage_upto(C - 1);
    if (dbg && lane == 0) {
      long long* d = dbg + ((size_t)blockIdx.x * kWsWarps + warp) * 4;
      d[0] = t_wait;
      d[1] = t_work;
      d[2] = t_stage;
      d[3] = clock64() - t_begin;
    }
  }
}

int g_ws_scan = 0;  // tuning hook (key 22): scan warps of the warp-specialised kernel (0 = 8; 8 or 12)
long long* g_ws_dbg = nullptr;  // development: per-warp clock totals (ga_debug_ws_trace)
int g_ws_dev = 0;  // tuning hook (key 24), development: bit 0 no refine, bit 1 no query loads in the scan warps
int g_ws_idle_ns = 0;  // tuning hook (key 23): start offset of the second scan warp of every scheduler, ns

template <int NS>
static int launch_ws(const FwdArgs& a, int mode, cudaStream_t st, int wt1, int wt2, int J) {
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_mma_ws_kernel<GA_MODE_CPU_EXACT, NS>
                                     : nn_fwd_mma_ws_kernel<GA_MODE_GPU_REF, NS>;
  {
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  extern int g_mma_grid;
  long long grid = g_mma_grid > 0 ? g_mma_grid : sm_count();
  if (grid > J) grid = J;
  k<<<(unsigned)grid, kWsWarps * 32, kWsSmem, st>>>(a, wt1, wt2, J, (unsigned)(g_ws_idle_ns >= 0 ? g_ws_idle_ns : 0), g_ws_dbg, g_ws_dev);
  GA_LAUNCH_CHECK("nn_fwd_mma_ws_kernel");
  return GA_OK;
}

bool fwd_mma_ws_supported(int b, int n, int m) {
  if (n > kWsCH || m > kWsCH || n < 1 || m < 1) return false;
  const long long J = (long long)b * ((n + kMmaQW - 1) / kMmaQW + (m + kMmaQW - 1) / kMmaQW);
  return J > 0 && J < 0x3fffffffLL;
}

int launch_fwd_mma_ws(const FwdArgs& a, int mode, cudaStream_t st) {
  if (!fwd_mma_ws_supported(a.b, a.n, a.m)) {
    set_error("nn_fwd_mma_ws_kernel: clouds of 1..%d points, fewer than 2^30 jobs", kWsCH);
    return GA_ERR_UNSUPPORTED;
  }
  const int wt1 = (a.n + kMmaQW - 1) / kMmaQW, wt2 = (a.m + kMmaQW - 1) / kMmaQW;
  const int J = a.b * (wt1 + wt2);
  if (g_ws_scan == 12) return launch_ws<12>(a, mode, st, wt1, wt2, J);
  return launch_ws<8>(a, mode, st, wt1, wt2, J);
}

}  // namespace ga

// development: per-warp clock totals of the next nn_fwd_mma_ws_kernel launches are written to `buf`
// (device memory, grid x 16 warps x 4 int64: wait, work, stage, total); nullptr switches it off.
extern "C" int ga_debug_ws_trace(long long* buf) {
  ga::g_ws_dbg = buf;
  return GA_OK;
}
