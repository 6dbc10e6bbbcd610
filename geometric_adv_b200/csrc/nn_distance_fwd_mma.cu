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
  static_assert(CH % kMmaBlk == 0 && CH / kMmaBlk <= 16, "whole MMA blocks, block number must fit 4 key bits");
  // pair-SoA (+pipeline pad) | red[96] (stage_targets_min, frame_from_box) | B fragments | per-query tile lists (count, 2 tiles)
  // The B fragments must start on a 128-byte line: a warp's LDS.64 covers 256 contiguous bytes, two wavefronts when
  // aligned and three when not (measured: 56 -> 59 us at B=50 with the fragments 64 bytes off).
  static constexpr size_t kOffRed = (size_t)CH * 16 + (size_t)kPipeU * 32;
  static constexpr size_t kOffB = kOffRed + 96 * 4;
  static_assert(kOffB % 128 == 0, "B fragments on a 128-byte line");
  static constexpr size_t kOffCnt = kOffB + (size_t)CH * 32;
  static constexpr size_t kOffTile = kOffCnt + (size_t)kQT * 4;
  static constexpr size_t kSmem = kOffTile + (size_t)kQT * 4;
};

// FRAME = false: the plain kernel; it only reports clouds that keep away from the origin (a.frame_hint).
// FRAME = true: the filter of such a cloud runs in a frame centred on it (Frame, nn_tiles.cuh).  Same bits either way.
template <class Cfg, int MODE, int MINB, bool FRAME>
__global__ void __launch_bounds__(Cfg::kThreads, MINB) nn_fwd_mma_kernel(const FwdArgs a) {
  constexpr int THREADS = Cfg::kThreads, QT = Cfg::kQT, CH = Cfg::kCH, T = kMmaT;
  // Completion tickets: clear this batch element's slot BEFORE the dependents may start (a replayed CUDA
  // graph reuses its call id: the slot still holds the previous replay's "all done").  Dependents (the
  // gradient kernel) may be scheduled once every CTA of this grid has passed this point.
  if (a.ticket != nullptr) {
    // one thread clears the slot and then triggers for the CTA; the others go straight to work (a barrier
    // here would put the ~1 us of the atomic and the fence in front of every CTA)
    // Only the FIRST CTA of a batch element clears: a sibling that starts in a later wave must not erase the
    // check-ins of siblings that have already finished (the count would never reach its target and the
    // gradient kernel would fall back to the grid dependency after its spin).  Dependents cannot run before
    // every CTA has passed its trigger below, i.e. before every clear is done.
    if (threadIdx.x == 0) {
      const int jpb0 = a.tiles1 + a.tiles2;
      if (blockIdx.x % jpb0 == 0) atomicExch(a.ticket + ticket_slot(a.call_id, (int)(blockIdx.x / jpb0)), a.call_id << 16);
      if (a.ticket_debug) atomicMax(a.ticket + kTicketSlots + 4, global_ns());  // start of the last CTA
      __threadfence();
      asm volatile("griddepcontrol.launch_dependents;");
    }
  } else {
    asm volatile("griddepcontrol.launch_dependents;");
  }
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float4* tgt = smem_f4;
  float* red = reinterpret_cast<float*>(smem + Cfg::kOffRed);
  uint4* bfrag = reinterpret_cast<uint4*>(smem + Cfg::kOffB);
  int* lcnt = reinterpret_cast<int*>(smem + Cfg::kOffCnt);
  unsigned short* ltile = reinterpret_cast<unsigned short*>(smem + Cfg::kOffTile);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jpb = a.tiles1 + a.tiles2;
  const int batch = blockIdx.x / jpb;
  const int r0 = blockIdx.x - batch * jpb;
  const bool rev = r0 >= a.tiles1;
  const int qtile = rev ? r0 - a.tiles1 : r0;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  if (!wait_ready(a, batch)) return;  // streamed ingest: this batch element's clouds are on the device from here on

  const int qbase = qtile * QT + warp * kMmaQW;  // first query of this warp
  MmaRows R;
  QueryState<2> s;
  float mrun[8];
#pragma unroll
  for (int r = 0; r < 8; r++) mrun[r] = kMmaBig;
  float bm_run = 0.0f;

  if constexpr (!FRAME) {
    // the query side is built before the staging: its loads hide behind it
    mma_load_rows(R, qpts, nq, qbase, lane);
    mma_init_queries<MODE>(s, qpts, nq, qbase, tpts, lane);
    for (int c0 = 0; c0 < nt; c0 += CH) {
      const int cn = min(CH, nt - c0);
      float mn;
      bm_run = fmaxf(bm_run, stage_targets_min<THREADS, T>(tgt, red, tpts, c0, nt, (cn + T - 1) / T, tid,
                                                            c0 == 0 && a.frame_hint != nullptr, mn));
      if (frame_candidate(mn, bm_run)) {  // uniform, rare; mn = 0 behind the first chunk
        Frame fr;
        if (frame_from_box<THREADS, T>(tgt, red, c0, nt, (cn + T - 1) / T, tid, fr) && tid == 0)
          *reinterpret_cast<volatile int*>(a.frame_hint) = 1;
      }
      stage_bfrag<THREADS>(bfrag, tgt, (cn + kMmaBlk - 1) / kMmaBlk, cn, tid);
      __syncthreads();
      mma_chunk<MODE>(R, s, mrun, tgt, bfrag, c0, nt, cn, bm_run, lcnt + warp * kMmaQW, ltile + warp * kMmaQW * 2, lane);
    }
  } else {
    // The frame is chosen behind the staging of the first chunk, and the query side is built for it after that.
    // Between its uses the frame lives in shared memory: every value kept live across the scan costs the scan loop
    // an accumulator quad.
    volatile float* frs = red + 72;
    {
      const int ntile = (min(CH, nt) + T - 1) / T;
      float mn;
      bm_run = stage_targets_min<THREADS, T>(tgt, red, tpts, 0, nt, ntile, tid, true, mn);
      Frame fr;
      fr.cx = fr.cy = fr.cz = 0.0f;
      fr.shifted = false;
      if (frame_candidate(mn, bm_run) && frame_from_box<THREADS, T>(tgt, red, 0, nt, ntile, tid, fr))
        bm_run = shift_staged<THREADS, T>(tgt, red, 0, nt, ntile, tid, fr);
      const int t = lane & 3;
      mma_load_rows(R, qpts, nq, qbase, lane, t == 0 ? fr.cx : (t == 1 ? fr.cy : fr.cz));
      mma_init_queries<MODE>(s, qpts, nq, qbase, tpts, lane, fr.cx, fr.cy, fr.cz);
      if (tid == 0) {
        frs[0] = fr.cx;
        frs[1] = fr.cy;
        frs[2] = fr.cz;
        frs[3] = fr.shifted ? 1.0f : 0.0f;
      }
    }
    for (int c0 = 0; c0 < nt; c0 += CH) {
      const int cn = min(CH, nt - c0), ntile = (cn + T - 1) / T;
      if (c0 > 0) {
        if (frs[3] != 0.0f) {
          Frame fr;
          fr.cx = frs[0];
          fr.cy = frs[1];
          fr.cz = frs[2];
          fr.shifted = true;
          bm_run = fmaxf(bm_run, stage_targets_shifted<THREADS, T>(tgt, red, tpts, c0, nt, ntile, tid, fr));
        } else {
          bm_run = fmaxf(bm_run, stage_targets<THREADS, T>(tgt, red, tpts, c0, nt, ntile, tid));
        }
      }
      stage_bfrag<THREADS>(bfrag, tgt, (cn + kMmaBlk - 1) / kMmaBlk, cn, tid);
      __syncthreads();
      // two copies of the chunk body: in the unshifted one the "original coordinates" pointer is the constant
      // nullptr and the compiler drops every trace of the shifted path from scan and refine
      if (frs[3] != 0.0f)
        mma_chunk<MODE>(R, s, mrun, tgt, bfrag, c0, nt, cn, bm_run, lcnt + warp * kMmaQW, ltile + warp * kMmaQW * 2,
                        lane, tpts);
      else
        mma_chunk<MODE>(R, s, mrun, tgt, bfrag, c0, nt, cn, bm_run, lcnt + warp * kMmaQW, ltile + warp * kMmaQW * 2,
                        lane, nullptr);
    }
  }
  mma_write(s, qbase, nq, lane, (rev ? a.dist2 : a.dist1) + (size_t)batch * nq, (rev ? a.idx2 : a.idx1) + (size_t)batch * nq,
            rev ? a.mdist2 : a.mdist1, rev ? a.midx2 : a.midx1, (size_t)batch * nq);
  // Completion ticket of this batch element: (call id << 16 | CTAs done).  The gradient kernel, started
  // early as a programmatic dependent, builds its inverse index map as soon as all CTAs of an element
  // have checked in (nn_distance_bwd.cu); a slot overwritten by another call only makes it wait longer.
  if (a.ticket != nullptr) {
    __syncthreads();  // every warp's dist / idx rows are written
    if (tid == 0) {
      __threadfence();
      unsigned long long* slot = a.ticket + ticket_slot(a.call_id, batch);
      unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(slot), assumed;
      do {
        assumed = old;
        const unsigned long long next = (assumed >> 16) == a.call_id ? assumed + 1 : ((a.call_id << 16) | 1ull);
        old = atomicCAS(slot, assumed, next);
      } while (old != assumed);
      if (a.ticket_debug) atomicMax(a.ticket + kTicketSlots + 3, global_ns());  // end of the last CTA
    }
  }
}

// ---- persistent variant: one CTA of 16 warps per SM, warps pull 64-query jobs ----------------
// The scan is bound by the HMMA pipe (0.5 HMMA.16816 per clock and SM); staging, list building and
// refine are not.  The grid of nn_fwd_mma_kernel quantises the work in CTAs of 512 queries: at B=50
// 400 CTAs on 296 slots leave a 23 % tail (sm__cycles_active.avg 84 K vs elapsed 109 K).  Here every
// SM gets an equal, contiguous share of all (batch, direction, 64-query) warp jobs (B=50: 21.6 jobs
// for 16 warps).  A share is walked in super-steps of up to TWO target clouds: both are staged (one
// 128-target block per warp and cloud: pair-SoA + B fragments, 96 KB each), one barrier, then the 16
// warps pull jobs of both clouds from a shared counter and run scan / refine / write on their own,
// so their phases drift apart and nobody waits at a cloud boundary; one barrier closes the
// super-step.  At B=50 a share crosses at most one cloud boundary: two barriers per kernel.
// Clouds of at most 2048 points.
constexpr int kPersistWarps = 16;
constexpr int kPersistCH = 2048;
constexpr size_t kPersistBuf = (size_t)kPersistCH * 16 + (size_t)kPipeU * 32 + (size_t)kPersistCH * 32;
constexpr size_t kPersistOffRed = 2 * kPersistBuf;                         // [2][16] float
constexpr size_t kPersistOffCtr = kPersistOffRed + 2 * 16 * 4;             // job counter, control block, cursor
constexpr size_t kPersistOffCnt = kPersistOffCtr + 80;                     // [16][64] int
constexpr size_t kPersistOffTile = kPersistOffCnt + kPersistWarps * kMmaQW * 4;  // [16][64][2] u16
constexpr size_t kPersistSmem = kPersistOffTile + kPersistWarps * kMmaQW * 4;

// One 64-query job of the persistent kernel.  Deliberately NOT inlined: inside the kernel's nested
// loops ptxas serialises the four HMMAs of a scan step on one accumulator quad (a NOP after every
// HMMA, 25 % slower scan); as a function of its own the scan gets the same schedule as in
// nn_fwd_mma_kernel (three rotating accumulator quads).
template <int MODE>
__device__ __noinline__ void persist_job(const FwdArgs& a, int batch, bool rev, int qbase, const unsigned char* bufp,
                                         float bm, int* wcnt, unsigned short* wtile) {
  const int lane = threadIdx.x & 31;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  const float4* tgt = reinterpret_cast<const float4*>(bufp);
  const uint4* bfrag = reinterpret_cast<const uint4*>(bufp + (size_t)kPersistCH * 16 + (size_t)kPipeU * 32);
  MmaRows R;
  mma_load_rows(R, qpts, nq, qbase, lane);
  QueryState<2> s;
  mma_init_queries<MODE>(s, qpts, nq, qbase, tpts, lane);
  float mrun[8];
#pragma unroll
  for (int r = 0; r < 8; r++) mrun[r] = kMmaBig;
  mma_chunk<MODE>(R, s, mrun, tgt, bfrag, 0, nt, nt, bm, wcnt, wtile, lane);
  mma_write(s, qbase, nq, lane, (rev ? a.dist2 : a.dist1) + (size_t)batch * nq, (rev ? a.idx2 : a.idx1) + (size_t)batch * nq,
            rev ? a.mdist2 : a.mdist1, rev ? a.midx2 : a.midx1, (size_t)batch * nq);
}

template <int MODE>
__global__ void __launch_bounds__(kPersistWarps * 32, 1)
    nn_fwd_mma_persist_kernel(const FwdArgs a, const int wt1, const int wt2, const long long J) {
  asm volatile("griddepcontrol.launch_dependents;");
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float* red = reinterpret_cast<float*>(smem + kPersistOffRed);
  int* counter = reinterpret_cast<int*>(smem + kPersistOffCtr);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* wcnt = reinterpret_cast<int*>(smem + kPersistOffCnt) + warp * kMmaQW;
  unsigned short* wtile = reinterpret_cast<unsigned short*>(smem + kPersistOffTile) + warp * kMmaQW * 2;

  // Control block of a super-step, written by thread 0 and read back from shared memory inside the job
  // loop: the 64-bit job arithmetic must not stay live across the scan (with it in registers ptxas
  // serialises the four HMMAs of a step on ONE accumulator quad and the scan runs 25 % slower).
  //   ctl[4 s + 0..3] = batch, direction, first query, jobs of segment s (0 / 1); ctl[8] = done flag
  volatile int* ctl = reinterpret_cast<volatile int*>(smem + kPersistOffCtr + 16);
  long long* jcur = reinterpret_cast<long long*>(smem + kPersistOffCtr + 64);

  auto stage = [&](int buf, int batch, bool rev) {  // this warp's 128-target block of a target cloud
    const int nt = rev ? a.n : a.m;
    const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
    float4* tgt = reinterpret_cast<float4*>(smem + buf * kPersistBuf);
    uint4* bfrag = reinterpret_cast<uint4*>(smem + buf * kPersistBuf + (size_t)kPersistCH * 16 + (size_t)kPipeU * 32);
    float lmax = 0.0f;
    if (warp * kMmaBlk < nt) lmax = stage_block_warp(tgt, bfrag, tpts, nt, warp, lane);
    lmax = warp_max(lmax);
    if (lane == 0) red[buf * 16 + warp] = lmax;
  };

  if (tid == 0) *jcur = J * blockIdx.x / gridDim.x;
  for (;;) {
    if (tid == 0) {
      const long long jpb = (long long)wt1 + wt2;
      const long long j1 = J * (blockIdx.x + 1) / gridDim.x;
      long long j = *jcur;
      ctl[8] = j >= j1;
      for (int sgm = 0; sgm < 2; sgm++) {
        int batch = 0, rev = 0, firstq = 0, njobs = 0;
        if (j < j1) {
          batch = (int)(j / jpb);
          const long long r = j - (long long)batch * jpb;
          rev = r >= wt1;
          const long long sbeg = (long long)batch * jpb + (rev ? wt1 : 0);
          const long long send = sbeg + (rev ? wt2 : wt1);
          const long long pend = send < j1 ? send : j1;
          firstq = (int)(j - sbeg) * kMmaQW;
          njobs = (int)(pend - j);
          j = pend;
        }
        ctl[4 * sgm] = batch;
        ctl[4 * sgm + 1] = rev;
        ctl[4 * sgm + 2] = firstq;
        ctl[4 * sgm + 3] = njobs;
      }
      *jcur = j;
      *counter = 0;
    }
    __syncthreads();
    if (ctl[8]) break;
    stage(0, ctl[0], ctl[1] != 0);
    if (ctl[7] > 0) stage(1, ctl[4], ctl[5] != 0);
    __syncthreads();
    for (;;) {
      int jo = 0;
      if (lane == 0) jo = atomicAdd(counter, 1);
      jo = __shfl_sync(0xffffffffu, jo, 0);
      const int na = ctl[3];
      if (jo >= na + ctl[7]) break;
      const int sgm = jo >= na ? 1 : 0;
      const int batch = ctl[4 * sgm];
      const bool rev = ctl[4 * sgm + 1] != 0;
      const int qbase = ctl[4 * sgm + 2] + (jo - (sgm ? na : 0)) * kMmaQW;
      float bm = 0.0f;
#pragma unroll
      for (int w = 0; w < 16; w++) bm = fmaxf(bm, red[sgm * 16 + w]);
      persist_job<MODE>(a, batch, rev, qbase, smem + (sgm ? kPersistBuf : 0), bm, wcnt, wtile);
    }
    __syncthreads();  // both buffers, the counter and the control block are free again
  }
}

// ---- balanced persistent variant (23) ----------------------------------------------------------
// Variant 21 balances whole 64-query jobs: 21.6 jobs for 16 warps is a full round plus a round of
// lone, latency-bound warps.  Here the unit of work is (job, 128-target block): a super-step's units
// are cut into 16 equal contiguous warp ranges, so a job may be scanned by several consecutive
// warps.  Keys carry the whole tile number (6 bits, mma_window_wide), every warp reduces its tracks
// over the quad (quad_merge) and, if it does not hold the job's last block, publishes the 64 x 3 row
// keys (768 B) in shared memory; the warp that holds the last block merges the published parts into
// its own, builds the candidate lists and refines.  A warp first publishes (at most one part), then
// runs its whole jobs, then the job it has to merge: nobody waits for a part that is not yet queued
// for publication at the very start of a peer's range.
constexpr size_t kBalOffPub = kPersistOffCnt;                                   // [16][8][24] float
constexpr size_t kBalOffFlag = kBalOffPub + (size_t)kPersistWarps * 8 * 24 * 4;  // [16] int
constexpr size_t kBalOffDesc = kBalOffFlag + 64;                                 // [16][16] int: per-warp part descriptor
constexpr size_t kBalSmem = kBalOffDesc + (size_t)kPersistWarps * 16 * 4;

struct BalSeg {  // one cloud of a super-step, as read back from the control block
  int batch, rev, firstq, njobs, nblk;
};

// Scan blocks [blk0, blk1) of one job; publish, or merge the earlier parts and finish the job.
// The part is described by 12 words in shared memory (written by the caller, read back here with
// volatile loads when they are needed): with a dozen scalar arguments live across the scan ptxas
// falls back to one accumulator quad and a NOP after most HMMAs.
//   d[0] batch, [1] direction, [2] first query, [3] blk0, [4] blk1, [5] blocks of the job,
//   [6] first warp holding a part of the job, [7] units of the super-step, [8] generation,
//   [9] max |target coordinate| (float bits), [10] staged buffer
template <int MODE>
__device__ __noinline__ void balanced_part(const FwdArgs& a, const volatile int* d, unsigned char* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* pub = reinterpret_cast<float*>(smem + kBalOffPub);
  volatile int* flags = reinterpret_cast<volatile int*>(smem + kBalOffFlag);
  const bool rev = d[1] != 0;
  const int nq = rev ? a.m : a.n;
  const int nt = rev ? a.n : a.m;
  const int qbase = d[2];
  const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)d[0] * nq * 3;
  const unsigned char* bufp = smem + (d[10] ? kPersistBuf : 0);
  const float4* tgt = reinterpret_cast<const float4*>(bufp);
  const uint2* bfrag = reinterpret_cast<const uint2*>(bufp + (size_t)kPersistCH * 16 + (size_t)kPipeU * 32);
  MmaRows R;
  mma_load_rows(R, qpts, nq, qbase, lane);
  MmaTrack tr;
  mma_scan_range<6>(R, bfrag, d[3], d[4], lane, tr);
  quad_merge(tr);
  const int gen = d[8];
  if (d[4] < d[5]) {  // not the last block of the job: hand the row keys to the warp that has it
    if (t == 0) {
      float* dst = pub + (size_t)(warp * 8 + g) * 24;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        dst[3 * r] = tr.c1[r];
        dst[3 * r + 1] = tr.c2[r];
        dst[3 * r + 2] = tr.c3[r];
      }
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) flags[warp] = gen;
    return;
  }
  const int utot = d[7];
  for (int pw = d[6]; pw < warp; pw++) {  // earlier parts of this job (blk0 > 0)
    // a warp whose range is empty (fewer units than warps) holds no part and never publishes
    if ((int)((long long)utot * pw / kPersistWarps) == (int)((long long)utot * (pw + 1) / kPersistWarps)) continue;
    if (lane == 0)
      while (flags[pw] != gen) {
      }
    __syncwarp();
    __threadfence_block();
    const volatile float* src = pub + (size_t)(pw * 8 + g) * 24;
#pragma unroll
    for (int r = 0; r < 8; r++) merge3(tr.c1[r], tr.c2[r], tr.c3[r], src[3 * r], src[3 * r + 1], src[3 * r + 2]);
  }
  // this lane's two queries are rows 2t and 2t+1 of its quad
  const int batch = d[0];
  const float bm = __int_as_float(d[9]);
  const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
  QueryState<2> s;
  mma_init_queries<MODE>(s, qpts, nq, qbase, tpts, lane);
  const int ntile = (nt + kMmaT - 1) / kMmaT;
  int cnt[2], ta[2], tb[2];
  float thr[2];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const float k1 = t == 0 ? tr.c1[j] : (t == 1 ? tr.c1[2 + j] : (t == 2 ? tr.c1[4 + j] : tr.c1[6 + j]));
    const float k2 = t == 0 ? tr.c2[j] : (t == 1 ? tr.c2[2 + j] : (t == 2 ? tr.c2[4 + j] : tr.c2[6 + j]));
    const float k3 = t == 0 ? tr.c3[j] : (t == 1 ? tr.c3[2 + j] : (t == 2 ? tr.c3[4 + j] : tr.c3[6 + j]));
    const float qa = t == 0 ? R.qabs[j] : (t == 1 ? R.qabs[2 + j] : (t == 2 ? R.qabs[4 + j] : R.qabs[6 + j]));
    thr[j] = k1 + mma_window_wide(qa, bm);
    int c = !(k1 > thr[j]) ? 1 : 0;
    c += !(k2 > thr[j]) ? 1 : 0;
    if (!(k3 > thr[j])) c = 3;
    ta[j] = __float_as_int(k1) & 63;
    tb[j] = __float_as_int(k2) & 63;
    // a tile beyond the staged tiles can only come from padding / sentinels under a non-finite window
    if ((c >= 1 && ta[j] >= ntile) || (c >= 2 && tb[j] >= ntile)) c = 3;
    cnt[j] = c;
  }
  refine_tiles<MODE>(s, tgt, 0, nt, ntile, cnt, ta, tb, thr);
  mma_write(s, qbase, nq, lane, (rev ? a.dist2 : a.dist1) + (size_t)batch * nq, (rev ? a.idx2 : a.idx1) + (size_t)batch * nq,
            rev ? a.mdist2 : a.mdist1, rev ? a.midx2 : a.midx1, (size_t)batch * nq);
}

template <int MODE>
__global__ void __launch_bounds__(kPersistWarps * 32, 1)
    nn_fwd_mma_balanced_kernel(const FwdArgs a, const int wt1, const int wt2, const long long J) {
  asm volatile("griddepcontrol.launch_dependents;");
  extern __shared__ float4 smem_f4[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(smem_f4);
  float* red = reinterpret_cast<float*>(smem + kPersistOffRed);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  volatile int* ctl = reinterpret_cast<volatile int*>(smem + kPersistOffCtr + 16);
  long long* jcur = reinterpret_cast<long long*>(smem + kPersistOffCtr + 64);
  float* pub = reinterpret_cast<float*>(smem + kBalOffPub);
  volatile int* flags = reinterpret_cast<volatile int*>(smem + kBalOffFlag);
  volatile int* desc = reinterpret_cast<volatile int*>(smem + kBalOffDesc) + warp * 16;

  auto stage = [&](int buf, int batch, bool rev) {  // this warp's 128-target block of a target cloud
    const int nt = rev ? a.n : a.m;
    const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
    float4* tgt = reinterpret_cast<float4*>(smem + buf * kPersistBuf);
    uint4* bfrag = reinterpret_cast<uint4*>(smem + buf * kPersistBuf + (size_t)kPersistCH * 16 + (size_t)kPipeU * 32);
    float lmax = 0.0f;
    if (warp * kMmaBlk < nt) lmax = stage_block_warp(tgt, bfrag, tpts, nt, warp, lane);
    lmax = warp_max(lmax);
    if (lane == 0) red[buf * 16 + warp] = lmax;
  };

  if (tid < kPersistWarps) flags[tid] = 0;
  if (tid == 0) *jcur = J * blockIdx.x / gridDim.x;
  for (int gen = 1;; gen++) {
    if (tid == 0) {
      const long long jpb = (long long)wt1 + wt2;
      const long long j1 = J * (blockIdx.x + 1) / gridDim.x;
      long long j = *jcur;
      ctl[8] = j >= j1;
      for (int sgm = 0; sgm < 2; sgm++) {
        int batch = 0, rev = 0, firstq = 0, njobs = 0;
        if (j < j1) {
          batch = (int)(j / jpb);
          const long long r = j - (long long)batch * jpb;
          rev = r >= wt1;
          const long long sbeg = (long long)batch * jpb + (rev ? wt1 : 0);
          const long long send = sbeg + (rev ? wt2 : wt1);
          const long long pend = send < j1 ? send : j1;
          firstq = (int)(j - sbeg) * kMmaQW;
          njobs = (int)(pend - j);
          j = pend;
        }
        ctl[4 * sgm] = batch;
        ctl[4 * sgm + 1] = rev;
        ctl[4 * sgm + 2] = firstq;
        ctl[4 * sgm + 3] = njobs;
      }
      *jcur = j;
    }
    __syncthreads();
    if (ctl[8]) break;
    stage(0, ctl[0], ctl[1] != 0);
    if (ctl[7] > 0) stage(1, ctl[4], ctl[5] != 0);
    __syncthreads();

    BalSeg sg[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      sg[i].batch = ctl[4 * i];
      sg[i].rev = ctl[4 * i + 1];
      sg[i].firstq = ctl[4 * i + 2];
      sg[i].njobs = ctl[4 * i + 3];
      sg[i].nblk = ((sg[i].rev ? a.n : a.m) + kMmaBlk - 1) / kMmaBlk;
    }
    const int ua = sg[0].njobs * sg[0].nblk;            // units of cloud A; cloud B follows
    const int utot = ua + sg[1].njobs * sg[1].nblk;
    auto wstart = [&](int w) { return (int)((long long)utot * w / kPersistWarps); };
    const int u0 = wstart(warp), u1 = wstart(warp + 1);
    float bmv[2] = {0.0f, 0.0f};
#pragma unroll
    for (int w = 0; w < 16; w++) {
      bmv[0] = fmaxf(bmv[0], red[w]);
      bmv[1] = fmaxf(bmv[1], red[16 + w]);
    }
    // pass 0: the part to publish (range ends inside a job); pass 1: whole jobs; pass 2: the job to merge
    for (int pass = 0; pass < 3; pass++) {
      int u = u0;
      while (u < u1) {
        const int i = u >= ua ? 1 : 0;
        const int ul = u - (i ? ua : 0);
        const int jo = ul / sg[i].nblk, blk0 = ul - jo * sg[i].nblk;
        const int blk1 = min(sg[i].nblk, blk0 + (u1 - u));
        const int kind = blk1 < sg[i].nblk ? 0 : (blk0 == 0 ? 1 : 2);
        if (kind == pass) {
          int fw = warp;  // warp that holds block 0 of this job
          if (blk0 > 0) {
            const int ub = u - blk0;
            fw = (int)(((long long)ub * kPersistWarps) / utot);
            while (fw + 1 < kPersistWarps && wstart(fw + 1) <= ub) fw++;
            while (fw > 0 && wstart(fw) > ub) fw--;
          }
          if (lane == 0) {
            desc[0] = sg[i].batch;
            desc[1] = sg[i].rev;
            desc[2] = sg[i].firstq + jo * kMmaQW;
            desc[3] = blk0;
            desc[4] = blk1;
            desc[5] = sg[i].nblk;
            desc[6] = fw;
            desc[7] = utot;
            desc[8] = gen;
            desc[9] = __float_as_int(i ? bmv[1] : bmv[0]);
            desc[10] = i;
          }
          __syncwarp();
          balanced_part<MODE>(a, desc, smem);
          __syncwarp();
        }
        u += blk1 - blk0;
      }
    }
    __syncthreads();  // buffers, control block and publish slots are free again
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

// tuning hook (key 25): 0 = plain kernel, never a frame; 1 = frame kernel always; 2 (default) = plain kernel until one
// of its launches on this device has reported a cloud away from the origin (frame_hint, core.cu), frame kernel from
// then on (setting the key again clears the report).  Both kernels return the same bits: the choice is about time
// only (unit cubes at offset 10: 569 us plain, 61 us frame; centred: 56 vs 59).  A captured graph keeps the kernel
// chosen at capture time.
int g_frame = 2;
std::atomic<int> g_frame_clear{0};  // set by ga_set_tuning(25, .): the next launch clears the device's report
int g_tickets = 1;  // tuning hook (key 18): 0 = never, 1 = for ga_nn_distance_fwd_bwd only, 2 = always + debug stamps, 3 = always
thread_local int t_want_tickets = 0;  // set by ga_nn_distance_fwd_bwd around its forward launch
thread_local ReadyArm t_ready_arm = {nullptr, 1, nullptr, 0};  // set by host_api.cu around a streamed forward launch
int g_mma_cfg = 0;  // tuning hook (key 7): 0 auto (= 5); 1-5 = (warps, chunk, CTAs/SM) combinations below

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
  a.frame_hint = nullptr;
  bool frame = g_frame == 1;
  if (g_frame == 2) {
    volatile int* seen = nullptr;
    a.frame_hint = frame_hint(st, &seen);
    if (seen != nullptr) {
      if (g_frame_clear.exchange(0, std::memory_order_relaxed)) *seen = 0;
      frame = *seen != 0;
    }
  }
  auto k = mode == GA_MODE_CPU_EXACT
               ? (frame ? nn_fwd_mma_kernel<Cfg, GA_MODE_CPU_EXACT, MINB, true>
                        : nn_fwd_mma_kernel<Cfg, GA_MODE_CPU_EXACT, MINB, false>)
               : (frame ? nn_fwd_mma_kernel<Cfg, GA_MODE_GPU_REF, MINB, true>
                        : nn_fwd_mma_kernel<Cfg, GA_MODE_GPU_REF, MINB, false>);
  {
    static std::atomic<unsigned> done_mask[4];
    const int slot = mode * 2 + (frame ? 1 : 0);
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[slot].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
      done_mask[slot].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  // arm the completion tickets (see the end of the kernel); off for batches beyond the slot count
  a.ticket = nullptr;
  a.call_id = 0;
  // The check-in costs each CTA ~1 us at its end (fence + compare-and-swap before the exit: measured 58.2 ->
  // 60.3 us at B=50), which only the one-call entry point earns back: armed there (t_want_tickets), off for a
  // plain ga_nn_distance_fwd.
  if ((g_tickets >= 2 || (g_tickets == 1 && t_want_tickets)) && a.b <= kTicketSlots && a.tiles1 + a.tiles2 < 65536) {
    a.ticket = ticket_buffer(st);
    if (a.ticket != nullptr) a.call_id = next_call_id();
  }
  a.ticket_debug = g_tickets == 2;
  if (a.ready != nullptr && t_ready_arm.pdl) {
    // behind the SM-driven ingest: start as soon as its CTAs are resident (the kernel never waits for that grid;
    // its CTAs wait for the arrival flags)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)jobs);
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    GA_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, a));
  } else {
    k<<<(unsigned)jobs, Cfg::kThreads, Cfg::kSmem, st>>>(a);
  }
  GA_LAUNCH_CHECK(frame ? "nn_fwd_mma_kernel<frame>" : "nn_fwd_mma_kernel");
  if (a.call_id != 0) {
    LastForward& lf = last_forward();
    lf.stream = st;
    lf.idx1 = a.idx1;
    lf.idx2 = a.idx2;
    lf.b = a.b;
    lf.n = a.n;
    lf.m = a.m;
    lf.expected = a.tiles1 + a.tiles2;
    lf.call_id = a.call_id;
    lf.ticket = a.ticket;
  }
  return GA_OK;
}

int g_mma_grid = 0;  // tuning hook (key 8): CTAs of the persistent kernel (0 = one per SM)

int launch_fwd_mma_persist(const FwdArgs& a, int mode, cudaStream_t st) {
  if (a.n > kPersistCH || a.m > kPersistCH) {
    set_error("nn_fwd_mma_persist_kernel: clouds of at most %d points", kPersistCH);
    return GA_ERR_UNSUPPORTED;
  }
  const int wt1 = (a.n + kMmaQW - 1) / kMmaQW, wt2 = (a.m + kMmaQW - 1) / kMmaQW;
  const long long J = (long long)a.b * (wt1 + wt2);
  if (J <= 0) return GA_OK;
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_mma_persist_kernel<GA_MODE_CPU_EXACT>
                                     : nn_fwd_mma_persist_kernel<GA_MODE_GPU_REF>;
  {
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPersistSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  long long grid = g_mma_grid > 0 ? g_mma_grid : sm_count();
  if (grid > J) grid = J;
  k<<<(unsigned)grid, kPersistWarps * 32, kPersistSmem, st>>>(a, wt1, wt2, J);
  GA_LAUNCH_CHECK("nn_fwd_mma_persist_kernel");
  return GA_OK;
}

int launch_fwd_mma_balanced(const FwdArgs& a, int mode, cudaStream_t st) {
  if (a.n > kPersistCH || a.m > kPersistCH) {
    set_error("nn_fwd_mma_balanced_kernel: clouds of at most %d points", kPersistCH);
    return GA_ERR_UNSUPPORTED;
  }
  const int wt1 = (a.n + kMmaQW - 1) / kMmaQW, wt2 = (a.m + kMmaQW - 1) / kMmaQW;
  const long long J = (long long)a.b * (wt1 + wt2);
  if (J <= 0) return GA_OK;
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_mma_balanced_kernel<GA_MODE_CPU_EXACT>
                                     : nn_fwd_mma_balanced_kernel<GA_MODE_GPU_REF>;
  {
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBalSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  long long grid = g_mma_grid > 0 ? g_mma_grid : sm_count();
  if (grid > J) grid = J;
  k<<<(unsigned)grid, kPersistWarps * 32, kBalSmem, st>>>(a, wt1, wt2, J);
  GA_LAUNCH_CHECK("nn_fwd_mma_balanced_kernel");
  return GA_OK;
}

int launch_fwd_mma(const FwdArgs& a, int mode, cudaStream_t st) {
  int cfg = g_mma_cfg;
  if (cfg == 0) cfg = 5;  // 8 warps, 128 registers, 2 CTAs per SM: fastest at every measured shape
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
