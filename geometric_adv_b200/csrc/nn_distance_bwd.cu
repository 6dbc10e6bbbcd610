// Chamfer backward: gradient of nn_distance w.r.t. both clouds, atomic-free.
// Replaces NmDistanceGradKernel x2 + 2 memsets (tf_nndistance_g.cu:132-157,
// chamfer3D.cu:155-195) and reproduces NnDistanceGradOp's CPU loops
// (tf_nndistance.cpp:122-163) bit for bit.
//
// The reference scatters with atomicAdd (order-nondeterministic).  The CPU loops
// fix a summation order per output element:
//   grad_xyz1[j]: 0 + g1[j]*(x1[j]-x2[idx1[j]])            (loop 1, direct)
//                 then -= g2[j']*(x2[j']-x1[j]) for every j' with idx2[j']==j,
//                 ascending j'                                  (loop 2, scatter)
//   grad_xyz2[k]: 0, then -= g1[j]*(x1[j]-x2[k]) for every j with idx1[j]==k,
//                 ascending j (loop 1, scatter), then += g2[k]*(x2[k]-x1[idx2[k]]).
// One CTA per (batch element, output cloud) inverts the index map with a STABLE
// counting sort in shared memory (per-warp count tables + match.any ranks), then
// one thread per output point walks its contributor list in ascending order with
// unfused __fmul_rn/__fadd_rn.  Every output element is written exactly once, so
// no memset is needed and the result is run-to-run reproducible.
#include <atomic>

#include "ga_common.cuh"

namespace ga {

constexpr int kBwdThreads = 512;
constexpr int kBwdWarps = kBwdThreads / 32;
// keys (output points) handled per pass: 2048 = one CTA per (batch element, cloud); 512 = the
// output points of a cloud are split over up to 4 CTAs, each of which walks all contributors
// but counts only the keys of its own range (4x the CTAs for the 148 SMs at attack batch sizes)

struct BwdArgs {
  int b, n, m;
  int nparts;  // CTAs per (batch element, cloud)
  const float* xyz1;
  const float* xyz2;
  const float* gd1;
  const int* idx1;
  const float* gd2;
  const int* idx2;
  float* gxyz1;
  float* gxyz2;
  // completion tickets of the forward launch that wrote idx1 / idx2 (nullptr = wait for the whole grid)
  const unsigned long long* ticket;
  unsigned long long call_id;
  int expected;
  int ticket_debug;
  int gd_final;  // fused entry point: the upstream gradients were final before the forward search was launched
};

// Partner clouds up to this size are staged in shared memory ({x, y, z, grad_dist} per point): the
// contributor walk is a chain of dependent gathers, ~30 cycles per hop from shared memory instead
// of one L2 round trip (~700 cycles) per hop from global memory.
constexpr int kBwdStageMax = 4096;

// smem: cnt[W][K] u16 | total[K] i32 | start[K] i32 | wsum[W] i32 | order[L] u16 | partner[L] float4
static size_t bwd_smem_bytes(int keys, int lmax) {
  return (size_t)kBwdWarps * keys * 2 + (size_t)keys * 4 * 2 + 64 * 4 + (((size_t)lmax * 2 + 15) & ~(size_t)15) +
         (lmax <= kBwdStageMax ? (size_t)lmax * 16 : 0);
}

template <int K, bool STAGE>
__global__ void __launch_bounds__(kBwdThreads) nn_bwd_kernel(const BwdArgs a) {
  constexpr int W = kBwdWarps;
  static_assert(K % kBwdThreads == 0, "whole keys per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned short* cnt = reinterpret_cast<unsigned short*>(smem_raw);            // [W][K]
  int* total = reinterpret_cast<int*>(smem_raw + (size_t)W * K * 2);           // [K]
  int* start = total + K;                                                      // [K]
  int* wsum = start + K;                                                       // [64]
  unsigned short* order = reinterpret_cast<unsigned short*>(wsum + 64);        // [L]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % a.nparts;
  const int batch = (blockIdx.x / a.nparts) >> 1;
  const int side = (blockIdx.x / a.nparts) & 1;
  // own = the cloud whose gradient this CTA produces (P points); oth = the other (L points)
  const int P = side ? a.m : a.n;
  const int L = side ? a.n : a.m;
  const float* own = (side ? a.xyz2 : a.xyz1) + (size_t)batch * P * 3;
  const float* oth = (side ? a.xyz1 : a.xyz2) + (size_t)batch * L * 3;
  const float* own_gd = (side ? a.gd2 : a.gd1) + (size_t)batch * P;
  const int* own_idx = (side ? a.idx2 : a.idx1) + (size_t)batch * P;
  const float* oth_gd = (side ? a.gd1 : a.gd2) + (size_t)batch * L;
  const int* oth_idx = (side ? a.idx1 : a.idx2) + (size_t)batch * L;
  float* out = (side ? a.gxyz2 : a.gxyz1) + (size_t)batch * P * 3;

  // partner cloud in shared memory (behind order[], whose length depends on max(n, m))
  float4* pcloud = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(order) +
                                           ((((size_t)(a.n > a.m ? a.n : a.m)) * 2 + 15) & ~(size_t)15));
  if (STAGE) {
    // issued first: the loads fly while the count tables are zeroed; the first __syncthreads
    // below orders the stores before every use
    for (int e0 = tid; e0 < L; e0 += 4 * kBwdThreads) {
      float px[4], py[4], pz[4], pg[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * kBwdThreads;
        const int es = e < L ? e : 0;
        px[u] = __ldg(oth + (size_t)es * 3);
        py[u] = __ldg(oth + (size_t)es * 3 + 1);
        pz[u] = __ldg(oth + (size_t)es * 3 + 2);
        pg[u] = __ldg(oth_gd + es);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * kBwdThreads;
        if (e < L) pcloud[e] = make_float4(px[u], py[u], pz[u], pg[u]);
      }
    }
  }

  const int seg = ((L + W - 1) / W + 31) & ~31;  // elements per warp, multiple of 32
  const int e_begin = warp * seg;
  const int e_end = min(L, e_begin + seg);
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int k0 = part * K; k0 < P; k0 += a.nparts * K) {
    const int kn = min(K, P - k0);
    // 1. zero the per-warp count tables
    {
      uint32_t* z = reinterpret_cast<uint32_t*>(cnt);
      for (int i = tid; i < W * K / 2; i += kBwdThreads) z[i] = 0u;
    }
    // keys of this warp's first chunks, loaded together (one global latency instead of one per
    // chunk; covers the whole segment for clouds up to 2048 points)
    constexpr int KPRE = 4;
    int kpre[KPRE];
#pragma unroll
    for (int c = 0; c < KPRE; c++) {
      const int e = e_begin + c * 32 + lane;
      kpre[c] = e < e_end ? __ldg(oth_idx + e) : -1;
    }
    __syncthreads();
    // 2. per-warp histogram of its contiguous element segment
    for (int e0 = e_begin, c = 0; e0 < e_end; e0 += 32, c++) {
      const int e = e0 + lane;
      int key = -1 - lane;
      if (e < e_end) {
        int raw;
        if (c < KPRE) {
          raw = kpre[0];
#pragma unroll
          for (int q = 1; q < KPRE; q++) raw = c == q ? kpre[q] : raw;
        } else {
          raw = __ldg(oth_idx + e);
        }
        const int kk = raw - k0;
        if (kk >= 0 && kk < kn) key = kk;
      }
      const unsigned mask = __match_any_sync(0xffffffffu, key);
      if (key >= 0 && (mask & lt_mask) == 0u) cnt[warp * K + key] += (unsigned short)__popc(mask);
      __syncwarp();
    }
    __syncthreads();
    // 3. exclusive prefix over warps per key; per-key totals
    for (int k = tid; k < K; k += kBwdThreads) {
      int run = 0;
#pragma unroll
      for (int w = 0; w < W; w++) {
        const int c = cnt[w * K + k];
        cnt[w * K + k] = (unsigned short)run;
        run += c;
      }
      total[k] = run;
    }
    __syncthreads();
    // 4. exclusive scan of totals over keys (K = 4 keys per thread at 512 threads)
    {
      constexpr int PER = K / kBwdThreads;
      int loc[PER];
      int s = 0;
#pragma unroll
      for (int i = 0; i < PER; i++) {
        loc[i] = s;
        s += total[tid * PER + i];
      }
      int incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      int woff = 0;
      for (int w = 0; w < warp; w++) woff += wsum[w];
      const int base = woff + incl - s;
#pragma unroll
      for (int i = 0; i < PER; i++) start[tid * PER + i] = base + loc[i];
    }
    __syncthreads();
    // 5. stable placement: same walk as step 2, ranks within a warp chunk from match.any
    for (int e0 = e_begin, c = 0; e0 < e_end; e0 += 32, c++) {
      const int e = e0 + lane;
      int key = -1 - lane;
      if (e < e_end) {
        int raw;
        if (c < KPRE) {
          raw = kpre[0];
#pragma unroll
          for (int q = 1; q < KPRE; q++) raw = c == q ? kpre[q] : raw;
        } else {
          raw = __ldg(oth_idx + e);
        }
        const int kk = raw - k0;
        if (kk >= 0 && kk < kn) key = kk;
      }
      const unsigned mask = __match_any_sync(0xffffffffu, key);
      if (key >= 0) {
        const int r = __popc(mask & lt_mask);
        const int off = cnt[warp * K + key];
        order[start[key] + off + r] = (unsigned short)e;
      }
      __syncwarp();
      if (key >= 0 && (mask & lt_mask) == 0u) cnt[warp * K + key] += (unsigned short)__popc(mask);
      __syncwarp();
    }
    __syncthreads();
    // 6. one thread per output point (PER points per thread, handled together so that the
    //    independent global loads of all of them are in flight at once); contributors of each
    //    point are subtracted in ascending order
    {
      constexpr int PER = K / kBwdThreads;
      float ox[PER], oy[PER], oz[PER], gown[PER];
      int j2[PER], s0[PER], cn[PER];
      bool ok[PER];
#pragma unroll
      for (int i = 0; i < PER; i++) {
        const int k = tid + i * kBwdThreads;
        ok[i] = k < kn;
        const int p = k0 + (ok[i] ? k : 0);
        ox[i] = __ldg(own + (size_t)p * 3);
        oy[i] = __ldg(own + (size_t)p * 3 + 1);
        oz[i] = __ldg(own + (size_t)p * 3 + 2);
        j2[i] = __ldg(own_idx + p);
        gown[i] = __ldg(own_gd + p);
        s0[i] = ok[i] ? start[k] : 0;
        cn[i] = ok[i] ? total[k] : 0;
      }
      float dx[PER], dy[PER], dz[PER], ax[PER], ay[PER], az[PER];
      int maxc = 0;
#pragma unroll
      for (int i = 0; i < PER; i++) {  // direct term (own loop of the reference)
        dx[i] = dy[i] = dz[i] = 0.f;
        if (ok[i] && j2[i] >= 0 && j2[i] < L) {
          const float g = __fmul_rn(gown[i], 2.0f);
          float tx, ty, tz;
          if (STAGE) {
            const float4 t4 = pcloud[j2[i]];
            tx = t4.x; ty = t4.y; tz = t4.z;
          } else {
            tx = __ldg(oth + (size_t)j2[i] * 3);
            ty = __ldg(oth + (size_t)j2[i] * 3 + 1);
            tz = __ldg(oth + (size_t)j2[i] * 3 + 2);
          }
          dx[i] = __fmul_rn(g, __fsub_rn(ox[i], tx));
          dy[i] = __fmul_rn(g, __fsub_rn(oy[i], ty));
          dz[i] = __fmul_rn(g, __fsub_rn(oz[i], tz));
        }
        ax[i] = ay[i] = az[i] = 0.f;
        if (side == 0) {  // loop 1 (direct) runs before loop 2 (scatter) for cloud 1
          ax[i] = __fadd_rn(ax[i], dx[i]);
          ay[i] = __fadd_rn(ay[i], dy[i]);
          az[i] = __fadd_rn(az[i], dz[i]);
        }
        maxc = max(maxc, cn[i]);
      }
      for (int r = 0; r < maxc; r++) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
          if (r < cn[i]) {
            const int e = order[s0[i] + r];
            float ex, ey, ez, eg;
            if (STAGE) {
              const float4 t4 = pcloud[e];
              ex = t4.x; ey = t4.y; ez = t4.z; eg = t4.w;
            } else {
              ex = __ldg(oth + (size_t)e * 3);
              ey = __ldg(oth + (size_t)e * 3 + 1);
              ez = __ldg(oth + (size_t)e * 3 + 2);
              eg = __ldg(oth_gd + e);
            }
            const float g = __fmul_rn(eg, 2.0f);
            const float tx = __fmul_rn(g, __fsub_rn(ex, ox[i]));
            const float ty = __fmul_rn(g, __fsub_rn(ey, oy[i]));
            const float tz = __fmul_rn(g, __fsub_rn(ez, oz[i]));
            ax[i] = __fsub_rn(ax[i], tx);
            ay[i] = __fsub_rn(ay[i], ty);
            az[i] = __fsub_rn(az[i], tz);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < PER; i++) {
        if (!ok[i]) continue;
        if (side == 1) {  // for cloud 2 the scatter of loop 1 comes first, its own loop 2 last
          ax[i] = __fadd_rn(ax[i], dx[i]);
          ay[i] = __fadd_rn(ay[i], dy[i]);
          az[i] = __fadd_rn(az[i], dz[i]);
        }
        const int p = k0 + tid + i * kBwdThreads;
        out[(size_t)p * 3] = ax[i];
        out[(size_t)p * 3 + 1] = ay[i];
        out[(size_t)p * 3 + 2] = az[i];
      }
    }
    __syncthreads();
  }
}


// ---- second formulation (default for partner clouds up to kBwdStageMax points) ------------------
// The stable counting sort above pays ~40 % of its time in match.any and keeps 16 count tables
// (64 KB) per CTA.  Contributor lists are short (one entry on average), so here the inverse map is
// built with plain shared-memory atomics in ARBITRARY order and the owner of an output point sorts
// its own list (insertion sort, a handful of u16 in shared memory) before it walks it: the
// summation order is again ascending source index, bit-identical to NnDistanceGradOp.  A point with
// more than kBwd2Sort contributors (collapsed clouds) is served by a sequential scan of the staged
// index array, which is ascending by construction.  Everything the walk touches (partner xyz,
// grad_dist, indices) is staged once in shared memory; ~60 KB per CTA of 256 threads, so several
// CTAs share an SM and hide each other's phase latencies.
constexpr int kBwd2Threads = 256;
constexpr int kBwd2Sort = 16;

// smem: cnt[K] i32 | start[K + 32] i32 | wsum[32] i32 | keys[L] i32 | pcloud[L] float4 | order[L] u16
static size_t bwd2_smem_bytes(int keys, int l_other_max) {
  return (size_t)keys * 4 + (size_t)(keys + 32) * 4 + 32 * 4 + (size_t)l_other_max * 4 + (size_t)l_other_max * 16 +
         (((size_t)l_other_max * 2 + 15) & ~(size_t)15);
}

template <int K>
__global__ void __launch_bounds__(kBwd2Threads) nn_bwd2_kernel(const BwdArgs a) {
  constexpr int T = kBwd2Threads, PER = K / T;
  static_assert(K % T == 0, "whole keys per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lmax = ((a.n > a.m ? a.n : a.m) + 3) & ~3;         // as the launcher: keeps pcloud 16-byte aligned
  int* cnt = reinterpret_cast<int*>(smem_raw);                 // [K] counts, then fill counters
  int* start = cnt + K;                                        // [K + 1] exclusive prefix
  int* wsum = start + K + 32;                                  // [32]
  int* keys = wsum + 32;                                       // [lmax] partner indices
  float4* pcloud = reinterpret_cast<float4*>(keys + lmax);     // [lmax] {x, y, z, grad_dist}
  unsigned short* order = reinterpret_cast<unsigned short*>(pcloud + lmax);  // [lmax]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % a.nparts;
  const int batch = (blockIdx.x / a.nparts) >> 1;
  const int side = (blockIdx.x / a.nparts) & 1;
  const int P = side ? a.m : a.n;
  const int L = side ? a.n : a.m;
  const float* own = (side ? a.xyz2 : a.xyz1) + (size_t)batch * P * 3;
  const float* oth = (side ? a.xyz1 : a.xyz2) + (size_t)batch * L * 3;
  const float* own_gd = (side ? a.gd2 : a.gd1) + (size_t)batch * P;
  const int* own_idx = (side ? a.idx2 : a.idx1) + (size_t)batch * P;
  const float* oth_gd = (side ? a.gd1 : a.gd2) + (size_t)batch * L;
  const int* oth_idx = (side ? a.idx1 : a.idx2) + (size_t)batch * L;
  float* out = (side ? a.gxyz2 : a.gxyz1) + (size_t)batch * P * 3;

  // Launched as a programmatic dependent, this grid may start while the preceding kernel of the stream is
  // still running.  Only the one-call entry (ga_nn_distance_fwd_bwd: a.ticket != nullptr) knows that this
  // kernel is the library's own forward search, which merely READS the coordinates: there they are staged
  // while the search finishes.  Behind any other kernel (ga_nn_distance_bwd is a public entry and any
  // producer of xyz / grad_dist may precede it) nothing is read before the grid dependency resolves.
  if (a.ticket == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  // Stage the partner cloud once (4 points per thread in flight).
  for (int e0 = tid; e0 < L; e0 += 4 * T) {
    float px[4], py[4], pz[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * T;
      const int es = e < L ? e : 0;
      px[u] = __ldg(oth + (size_t)es * 3);
      py[u] = __ldg(oth + (size_t)es * 3 + 1);
      pz[u] = __ldg(oth + (size_t)es * 3 + 2);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * T;
      if (e < L) pcloud[e] = make_float4(px[u], py[u], pz[u], 0.0f);
    }
  }
  // The index rows of this batch element are final once all of the forward search's CTAs for it have
  // checked in (completion ticket, written with a release fence behind their stores).  Then the inverse
  // map below is built while the rest of that grid is still running; otherwise, or when the ticket
  // never shows up (slot reused by another call), wait for the whole grid as any dependent would.
  __shared__ int early_flag;
  if (tid == 0) {
    int early = 0;
    if (a.ticket != nullptr) {
      const volatile unsigned long long* slot = a.ticket + ticket_slot(a.call_id, batch);
      const unsigned long long want = (a.call_id << 16) | (unsigned long long)a.expected;
      for (int spin = 0; spin < 400 && !early; spin++) {
        if (*slot == want) early = 1;
        else __nanosleep(100);
      }
      if (early) __threadfence();  // acquire side: the stores behind the ticket are visible below
    }
    early_flag = early;
    if (a.ticket_debug) {  // development counters: early CTAs, first / last CTA start, last inverse map done
      unsigned long long* dbg = const_cast<unsigned long long*>(a.ticket) + kTicketSlots;
      if (early) atomicAdd(dbg + 0, 1ull);
      atomicAdd(dbg + 1, 1ull);
      atomicMax(dbg + 5, global_ns());
      if (blockIdx.x == 0) dbg[6] = global_ns();
    }
  }
  __syncthreads();
  const bool early = early_flag != 0;
  if (!early && a.ticket != nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int e0 = tid; e0 < L; e0 += 4 * T) {
    int pk[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * T;
      pk[u] = *reinterpret_cast<const volatile int*>(oth_idx + (e < L ? e : 0));  // written by the forward grid
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int e = e0 + u * T;
      if (e < L) keys[e] = pk[u];
    }
  }
  bool gd_staged = false;

  for (int k0 = part * K; k0 < P; k0 += a.nparts * K) {
    const int kn = min(K, P - k0);
    for (int i = tid; i < K; i += T) cnt[i] = 0;
    __syncthreads();
    // 1. contributors per output point
    for (int e = tid; e < L; e += T) {
      const int kk = keys[e] - k0;
      if (kk >= 0 && kk < kn) atomicAdd(&cnt[kk], 1);
    }
    __syncthreads();
    // 2. exclusive prefix over the keys (PER consecutive keys per thread); counts become fill counters
    {
      int loc[PER];
      int s = 0;
#pragma unroll
      for (int i = 0; i < PER; i++) {
        loc[i] = s;
        s += cnt[tid * PER + i];
      }
      int incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) wsum[warp] = incl;
      __syncthreads();
      int woff = 0;
      for (int w = 0; w < warp; w++) woff += wsum[w];
      const int base = woff + incl - s;
#pragma unroll
      for (int i = 0; i < PER; i++) {
        start[tid * PER + i] = base + loc[i];
        cnt[tid * PER + i] = 0;
      }
      if (tid == T - 1) start[K] = base + s;
    }
    __syncthreads();
    // 3. placement, arbitrary order inside a list
    for (int e = tid; e < L; e += T) {
      const int kk = keys[e] - k0;
      if (kk >= 0 && kk < kn) order[start[kk] + atomicAdd(&cnt[kk], 1)] = (unsigned short)e;
    }
    __syncthreads();
    // upstream gradients: whatever kernel precedes this one on the stream may have written them, so
    // they are read only behind the grid dependency (a no-op if the early path was not taken)
    if (!gd_staged) {
      if (early && !a.gd_final) asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int e = tid; e < L; e += T) reinterpret_cast<float*>(pcloud + e)[3] = oth_gd[e];
      gd_staged = true;
      __syncthreads();
    }
    // 4. one thread per output point: sort its list, walk it in ascending order
    {
      float ox[PER], oy[PER], oz[PER], gown[PER];
      int j2[PER];
      bool ok[PER];
#pragma unroll
      for (int i = 0; i < PER; i++) {
        const int k = tid + i * T;
        ok[i] = k < kn;
        const int p = k0 + (ok[i] ? k : 0);
        ox[i] = __ldg(own + (size_t)p * 3);
        oy[i] = __ldg(own + (size_t)p * 3 + 1);
        oz[i] = __ldg(own + (size_t)p * 3 + 2);
        j2[i] = *reinterpret_cast<const volatile int*>(own_idx + p);
        gown[i] = own_gd[p];
      }
#pragma unroll
      for (int i = 0; i < PER; i++) {
        if (!ok[i]) continue;
        const int k = tid + i * T;
        const int s0 = start[k], cn = start[k + 1] - s0;
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (j2[i] >= 0 && j2[i] < L) {  // direct term (own loop of the reference)
          const float4 t4 = pcloud[j2[i]];
          const float g = __fmul_rn(gown[i], 2.0f);
          dx = __fmul_rn(g, __fsub_rn(ox[i], t4.x));
          dy = __fmul_rn(g, __fsub_rn(oy[i], t4.y));
          dz = __fmul_rn(g, __fsub_rn(oz[i], t4.z));
        }
        float ax = 0.f, ay = 0.f, az = 0.f;
        if (side == 0) {  // loop 1 (direct) runs before loop 2 (scatter) for cloud 1
          ax = __fadd_rn(ax, dx);
          ay = __fadd_rn(ay, dy);
          az = __fadd_rn(az, dz);
        }
        auto sub = [&](int e) {
          const float4 t4 = pcloud[e];
          const float g = __fmul_rn(t4.w, 2.0f);
          ax = __fsub_rn(ax, __fmul_rn(g, __fsub_rn(t4.x, ox[i])));
          ay = __fsub_rn(ay, __fmul_rn(g, __fsub_rn(t4.y, oy[i])));
          az = __fsub_rn(az, __fmul_rn(g, __fsub_rn(t4.z, oz[i])));
        };
        if (cn <= kBwd2Sort) {
          for (int r = 1; r < cn; r++) {  // insertion sort of a short list
            const unsigned short v = order[s0 + r];
            int q = r - 1;
            while (q >= 0 && order[s0 + q] > v) {
              order[s0 + q + 1] = order[s0 + q];
              q--;
            }
            order[s0 + q + 1] = v;
          }
          for (int r = 0; r < cn; r++) sub(order[s0 + r]);
        } else {
          const int want = k0 + k;
          for (int e = 0; e < L; e++)
            if (keys[e] == want) sub(e);
        }
        if (side == 1) {  // for cloud 2 the scatter of loop 1 comes first, its own loop 2 last
          ax = __fadd_rn(ax, dx);
          ay = __fadd_rn(ay, dy);
          az = __fadd_rn(az, dz);
        }
        const int p = k0 + k;
        out[(size_t)p * 3] = ax;
        out[(size_t)p * 3 + 1] = ay;
        out[(size_t)p * 3 + 2] = az;
      }
    }
    __syncthreads();
  }
}

// ---- third formulation (default for grids of at most one CTA per SM): one global round trip, one pass over the contributors ------------
// nn_bwd2_kernel is a chain of five phases, four of which begin with a global load (partner xyz, index rows,
// upstream gradients, own points): at B = 50 the kernel is that latency chain (13.6 us against a 3.9 us launch
// floor, long-scoreboard the top stall).  Here every global load a thread needs is issued before the first
// barrier, and the inverse index map is not a compacted list (count -> scan -> place: two passes over the
// partner's index row and three barriers) but a table of fixed buckets: contributor e of output point p goes to
// bucket[p][atomicAdd(cnt[p])] (16 entries of 16 bit; lists have one entry on average and, for distinct points, at
// most the kissing number 12).  The owner of an output
// point walks its bucket in ascending source index (selection of the next larger entry: the summation order of
// NnDistanceGradOp, tf_nndistance.cpp:126-163, whatever order the atomics produced); a point with more than 16
// contributors (collapsed clouds) scans the staged copy of the partner's index row, which is ascending by
// construction.  Two barriers; 58 KB of shared memory at 2048 points and 512 keys.
constexpr int kBwd3Threads = 256;
constexpr int kBwd3Bucket = 16;

// smem: pcloud[L] float4 | cnt[K] i32 | bucket[K][16] u16 | keys[L] i32
static size_t bwd3_smem_bytes(int keys, int lpad) {
  return (size_t)lpad * 16 + (size_t)keys * 4 + (size_t)keys * kBwd3Bucket * 2 + (size_t)lpad * 4;
}

template <int K>
__global__ void __launch_bounds__(kBwd3Threads) nn_bwd3_kernel(const BwdArgs a) {
  constexpr int T = kBwd3Threads, PER = K / T, EB = 8;  // EB: partner elements per thread and batch
  static_assert(K % T == 0, "whole keys per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lpad = ((a.n > a.m ? a.n : a.m) + 3) & ~3;
  float4* pcloud = reinterpret_cast<float4*>(smem_raw);                       // [lpad] {x, y, z, grad_dist}
  int* cnt = reinterpret_cast<int*>(pcloud + lpad);                           // [K]
  unsigned short* bucket = reinterpret_cast<unsigned short*>(cnt + K);        // [K][16]
  int* keys = reinterpret_cast<int*>(bucket + K * kBwd3Bucket);               // [lpad] the partner's index row

  const int tid = threadIdx.x;
  const int part = blockIdx.x % a.nparts;
  const int batch = (blockIdx.x / a.nparts) >> 1;
  const int side = (blockIdx.x / a.nparts) & 1;
  const int P = side ? a.m : a.n;
  const int L = side ? a.n : a.m;
  const float* own = (side ? a.xyz2 : a.xyz1) + (size_t)batch * P * 3;
  const float* oth = (side ? a.xyz1 : a.xyz2) + (size_t)batch * L * 3;
  const float* own_gd = (side ? a.gd2 : a.gd1) + (size_t)batch * P;
  const int* own_idx = (side ? a.idx2 : a.idx1) + (size_t)batch * P;
  const float* oth_gd = (side ? a.gd1 : a.gd2) + (size_t)batch * L;
  const int* oth_idx = (side ? a.idx1 : a.idx2) + (size_t)batch * L;
  float* out = (side ? a.gxyz2 : a.gxyz1) + (size_t)batch * P * 3;

  // Behind an arbitrary kernel (two-call form) nothing is read before the grid dependency resolves; behind the
  // library's own forward search (one-call entry, a.ticket != nullptr) the coordinates may be read at once, the
  // index rows once the batch element's completion ticket is in, the upstream gradients when they are final.
  const bool twocall = a.ticket == nullptr;
  if (twocall) asm volatile("griddepcontrol.wait;" ::: "memory");

  for (int k0 = part * K; k0 < P; k0 += a.nparts * K) {  // one round unless the cloud has more than nparts * K points
    const int kn = min(K, P - k0);
    // partner coordinates: global -> shared memory without passing through registers
    for (int e = tid; e < L; e += T) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(pcloud + e);
      const float* srcp = oth + (size_t)e * 3;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\tcp.async.ca.shared.global [%0+4], [%1+4], 4;\n\t"
                   "cp.async.ca.shared.global [%0+8], [%1+8], 4;" ::"r"(dst), "l"(srcp) : "memory");
    }
    // ---- own points of this thread: coordinates now, index / gradient when allowed ----
    float ox[PER], oy[PER], oz[PER], gown[PER];
    int j2[PER];
    bool ok[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int k = tid + i * T;
      ok[i] = k < kn;
      const int p = k0 + (ok[i] ? k : 0);
      ox[i] = __ldg(own + (size_t)p * 3);
      oy[i] = __ldg(own + (size_t)p * 3 + 1);
      oz[i] = __ldg(own + (size_t)p * 3 + 2);
    }
    for (int i = tid; i < K; i += T) cnt[i] = 0;
    if (!twocall) {
      // one-call entry: index rows behind the completion ticket (or the grid dependency), gradients when final
      __shared__ int early_flag;
      if (tid == 0) {
        int early = 0;
        const volatile unsigned long long* slot = a.ticket + ticket_slot(a.call_id, batch);
        const unsigned long long want = (a.call_id << 16) | (unsigned long long)a.expected;
        for (int spin = 0; spin < 400 && !early; spin++) {
          if (*slot == want) early = 1;
          else __nanosleep(100);
        }
        if (early) __threadfence();  // acquire side: the stores behind the ticket are visible below
        early_flag = early;
        if (a.ticket_debug) {
          unsigned long long* dbg = const_cast<unsigned long long*>(a.ticket) + kTicketSlots;
          if (early) atomicAdd(dbg + 0, 1ull);
          atomicAdd(dbg + 1, 1ull);
          atomicMax(dbg + 5, global_ns());
          if (blockIdx.x == 0) dbg[6] = global_ns();
        }
      }
      __syncthreads();
      if (!(early_flag != 0 && a.gd_final)) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    // upstream gradients of the partner -> pcloud[].w, first batch of its index row -> registers
    for (int e = tid; e < L; e += T) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(pcloud + e) + 12;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(oth_gd + e) : "memory");
    }
    int pk[EB];
#pragma unroll
    for (int u = 0; u < EB; u++) {
      const int e = tid + u * T;
      pk[u] = *reinterpret_cast<const volatile int*>(oth_idx + (e < L ? e : 0));  // written by the forward grid
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
      const int p = k0 + (ok[i] ? tid + i * T : 0);
      j2[i] = *reinterpret_cast<const volatile int*>(own_idx + p);
      gown[i] = own_gd[p];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // cnt[] is zero, pcloud[] complete
    for (int e0 = tid; e0 < L; e0 += EB * T) {
      if (e0 != tid) {
#pragma unroll
        for (int u = 0; u < EB; u++) {
          const int e = e0 + u * T;
          pk[u] = *reinterpret_cast<const volatile int*>(oth_idx + (e < L ? e : 0));
        }
      }
#pragma unroll
      for (int u = 0; u < EB; u++) {
        const int e = e0 + u * T;
        const int kk = pk[u] - k0;
        if (e < L) keys[e] = pk[u];
        if (e < L && kk >= 0 && kk < kn) {
          const int pos = atomicAdd(&cnt[kk], 1);
          if (pos < kBwd3Bucket) bucket[kk * kBwd3Bucket + pos] = (unsigned short)e;
        }
      }
    }
    __syncthreads();
    // ---- one thread per output point ----
#pragma unroll
    for (int i = 0; i < PER; i++) {
      if (!ok[i]) continue;
      const int k = tid + i * T;
      const int cn = cnt[k];
      float dx = 0.f, dy = 0.f, dz = 0.f;
      if (j2[i] >= 0 && j2[i] < L) {  // direct term (own loop of the reference)
        const float4 t4 = pcloud[j2[i]];
        const float g = __fmul_rn(gown[i], 2.0f);
        dx = __fmul_rn(g, __fsub_rn(ox[i], t4.x));
        dy = __fmul_rn(g, __fsub_rn(oy[i], t4.y));
        dz = __fmul_rn(g, __fsub_rn(oz[i], t4.z));
      }
      float ax = 0.f, ay = 0.f, az = 0.f;
      if (side == 0) {  // loop 1 (direct) runs before loop 2 (scatter) for cloud 1
        ax = __fadd_rn(ax, dx);
        ay = __fadd_rn(ay, dy);
        az = __fadd_rn(az, dz);
      }
      auto sub = [&](int e) {
        const float4 t4 = pcloud[e];
        const float g = __fmul_rn(t4.w, 2.0f);
        ax = __fsub_rn(ax, __fmul_rn(g, __fsub_rn(t4.x, ox[i])));
        ay = __fsub_rn(ay, __fmul_rn(g, __fsub_rn(t4.y, oy[i])));
        az = __fsub_rn(az, __fmul_rn(g, __fsub_rn(t4.z, oz[i])));
      };
      if (cn <= kBwd3Bucket) {
        if (cn > 0) {
          const uint4 bw = *reinterpret_cast<const uint4*>(bucket + k * kBwd3Bucket);
          const uint4 bv = *reinterpret_cast<const uint4*>(bucket + k * kBwd3Bucket + 8);
          const uint32_t w[8] = {bw.x, bw.y, bw.z, bw.w, bv.x, bv.y, bv.z, bv.w};
          int prev = -1;
          for (int r = 0; r < cn; r++) {  // next larger source index
            int best = 0x10000;
#pragma unroll
            for (int q = 0; q < kBwd3Bucket; q++) {
              const int v = (int)((w[q >> 1] >> (16 * (q & 1))) & 0xffffu);
              if (q < cn && v > prev && v < best) best = v;
            }
            sub(best);
            prev = best;
          }
        }
      } else {
        const int want = k0 + k;
        for (int e = 0; e < L; e++)
          if (keys[e] == want) sub(e);
      }
      if (side == 1) {  // for cloud 2 the scatter of loop 1 comes first, its own loop 2 last
        ax = __fadd_rn(ax, dx);
        ay = __fadd_rn(ay, dy);
        az = __fadd_rn(az, dz);
      }
      const int p = k0 + k;
      out[(size_t)p * 3] = ax;
      out[(size_t)p * 3 + 1] = ay;
      out[(size_t)p * 3 + 2] = az;
    }
    __syncthreads();  // a further round reuses cnt / bucket / pcloud
  }
}

extern int g_tickets;                    // nn_distance_fwd_mma.cu
extern thread_local int t_want_tickets;  // nn_distance_fwd_mma.cu
int g_pdl = 1;         // tuning hook (key 15): 0 = plain launches (no programmatic dependent launch)
int g_bwd_kernel = 0;  // tuning hook (key 14): 0 auto, 1 = stable counting sort, 2 = second formulation (compacted lists), 3 = third (buckets, one round trip)
int g_bwd_stage = 1;   // tuning hook (key 13): 0 = gather the partner cloud from global memory (no staging)
int g_bwd_split = -1;  // tuning hook (key 9): -1 auto, 0 one CTA per cloud, 1 output points split over 4 CTAs

}  // namespace ga

namespace ga {
static thread_local int t_gd_final = 0;  // set by ga_nn_distance_fwd_bwd around its gradient launch
}  // namespace ga

extern "C" int ga_nn_distance_bwd(int b, int n, int m, const float* xyz1, const float* xyz2,
                                  const float* grad_dist1, const int* idx1, const float* grad_dist2,
                                  const int* idx2, float* grad_xyz1, float* grad_xyz2, ga_stream_t stream) {
  using namespace ga;
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_nn_distance_bwd: negative size (b=%d n=%d m=%d)", b, n, m);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (b == 0 || (n == 0 && m == 0)) return GA_OK;
  cudaStream_t st = as_stream(stream);
  // A cloud whose partner is empty only ever sees idx = 0 into nothing: the reference
  // would read out of bounds.  We define the gradient as zero there.
  if (n == 0 || m == 0) {
    if (n > 0) GA_CUDA_TRY(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3, st));
    if (m > 0) GA_CUDA_TRY(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3, st));
    return GA_OK;
  }
  const int lmax = n > m ? n : m;
  if (lmax > 65536) {
    set_error("ga_nn_distance_bwd: clouds larger than 65536 points are not supported (n=%d m=%d)", n, m);
    return GA_ERR_UNSUPPORTED;
  }
  if ((long long)b * 8 > 0x7fffffffLL) {
    set_error("ga_nn_distance_bwd: batch too large");
    return GA_ERR_UNSUPPORTED;
  }
  BwdArgs a;
  a.b = b; a.n = n; a.m = m;
  a.xyz1 = xyz1; a.xyz2 = xyz2;
  a.gd1 = grad_dist1; a.idx1 = idx1; a.gd2 = grad_dist2; a.idx2 = idx2;
  a.gxyz1 = grad_xyz1; a.gxyz2 = grad_xyz2;
  a.ticket = nullptr; a.call_id = 0; a.expected = 0; a.ticket_debug = 0;
  a.gd_final = t_gd_final;
  // Up to one wave of split CTAs: split the output points (measured, 2048-point clouds: B=1 17.4 ->
  // 14.3 us, B=10 18.4 -> 14.3 us); more cloud pairs: one CTA per cloud walks the contributors only
  // once (B=50 18.4 vs 27.5 us split, B=512 82 vs 184 us).
  if (g_bwd_kernel != 1 && lmax <= kBwdStageMax) {
    // CTAs per (batch element, cloud): enough to give every SM about two CTAs, each part at least 512 keys
    int parts = 1;
    if (g_bwd_split >= 0) {
      parts = g_bwd_split == 0 ? 1 : (g_bwd_split == 1 ? 4 : g_bwd_split);
    } else {
      while (parts < 4 && (long long)2 * b * parts < 2LL * sm_count() && (lmax + parts * 2 - 1) / (parts * 2) >= 512) parts *= 2;
    }
    const int per = (lmax + parts - 1) / parts;
    a.nparts = parts;
    // early start behind the forward search that wrote these index arrays on this stream (ga_common.cuh)
    a.ticket = nullptr;
    a.call_id = 0;
    a.expected = 0;
    {
      // Tickets are honoured only inside ga_nn_distance_fwd_bwd (t_gd_final), where the library itself launched
      // the search immediately before on this stream; the record is consumed by the first gradient launch
      // that looks at it, so a stale ticket can never be matched by a later, unrelated call.
      LastForward& lf = last_forward();
      if (g_pdl && t_gd_final && lf.call_id != 0 && lf.stream == st && lf.idx1 == idx1 && lf.idx2 == idx2 && lf.b == b &&
          lf.n == n && lf.m == m) {
        a.ticket = lf.ticket;
        a.call_id = lf.call_id;
        a.expected = lf.expected;
        a.ticket_debug = g_tickets == 2;
      }
      lf.call_id = 0;
    }
    static std::atomic<unsigned> done2{0};
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done2.load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd2_kernel<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bwd2_smem_bytes(2048, kBwdStageMax)));
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd2_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bwd2_smem_bytes(1024, kBwdStageMax)));
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd2_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bwd2_smem_bytes(512, kBwdStageMax)));
      done2.fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
    const unsigned grid = (unsigned)(2 * b * parts);
    const int lpad = (lmax + 3) & ~3;  // keeps the float4 array behind the int array 16-byte aligned
    // One CTA per SM or fewer: the kernel is its chain of global round trips, and nn_bwd3_kernel has one instead of
    // four (B = 1: 10.2 vs 12.2 us, B = 10: 10.1 vs 12.2 us).  Larger grids are throughput-bound and the compacted
    // lists of nn_bwd2_kernel cost fewer instructions and less shared memory per CTA (B = 50: 14.3 vs 16.3 us,
    // B = 512: 51 vs 65 us, B = 4096: 293 vs 436 us; profiles/r02_tune_bwd.txt).
    if (g_bwd_kernel == 3 || (g_bwd_kernel == 0 && (long long)grid <= (long long)sm_count())) {
      static std::atomic<unsigned> done3{0};
      if (!(done3.load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
        GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd3_kernel<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bwd3_smem_bytes(2048, kBwdStageMax)));
        GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd3_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bwd3_smem_bytes(1024, kBwdStageMax)));
        GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd3_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bwd3_smem_bytes(512, kBwdStageMax)));
        done3.fetch_or(1u << (dev & 31), std::memory_order_relaxed);
      }
      cudaLaunchConfig_t cfg3 = {};
      cfg3.gridDim = dim3(grid);
      cfg3.blockDim = dim3(kBwd3Threads);
      cfg3.stream = st;
      cudaLaunchAttribute attr3[1];
      attr3[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr3[0].val.programmaticStreamSerializationAllowed = 1;
      cfg3.attrs = attr3;
      cfg3.numAttrs = g_pdl ? 1 : 0;
      if (per <= 512) {
        cfg3.dynamicSmemBytes = bwd3_smem_bytes(512, lpad);
        GA_CUDA_TRY(cudaLaunchKernelEx(&cfg3, nn_bwd3_kernel<512>, a));
      } else if (per <= 1024) {
        cfg3.dynamicSmemBytes = bwd3_smem_bytes(1024, lpad);
        GA_CUDA_TRY(cudaLaunchKernelEx(&cfg3, nn_bwd3_kernel<1024>, a));
      } else {
        cfg3.dynamicSmemBytes = bwd3_smem_bytes(2048, lpad);
        GA_CUDA_TRY(cudaLaunchKernelEx(&cfg3, nn_bwd3_kernel<2048>, a));
      }
      GA_LAUNCH_CHECK("nn_bwd3_kernel");
      return GA_OK;
    }
    // Programmatic dependent launch: the grid may start while the preceding kernel of the stream (the
    // forward search, which triggers early) drains, and blocks in griddepcontrol.wait before it reads
    // anything that kernel wrote.  After a kernel that never triggers this is an ordinary launch.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kBwd2Threads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    if (per <= 512) {
      cfg.dynamicSmemBytes = bwd2_smem_bytes(512, lpad);
      GA_CUDA_TRY(cudaLaunchKernelEx(&cfg, nn_bwd2_kernel<512>, a));
    } else if (per <= 1024) {
      cfg.dynamicSmemBytes = bwd2_smem_bytes(1024, lpad);
      GA_CUDA_TRY(cudaLaunchKernelEx(&cfg, nn_bwd2_kernel<1024>, a));
    } else {
      cfg.dynamicSmemBytes = bwd2_smem_bytes(2048, lpad);
      GA_CUDA_TRY(cudaLaunchKernelEx(&cfg, nn_bwd2_kernel<2048>, a));
    }
    GA_LAUNCH_CHECK("nn_bwd2_kernel");
    return GA_OK;
  }
  int split = g_bwd_split;
  if (split < 0) split = (long long)b * 8 <= (long long)sm_count() ? 1 : 0;
  {
    // raise the opt-in shared-memory limit once per device to the largest size the kernels can ask for
    static std::atomic<unsigned> done_mask{0};
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask.load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      // the largest requests: staged partner at kBwdStageMax points; unstaged at 65536 points
      const size_t big2048 = bwd_smem_bytes(2048, 65536) > bwd_smem_bytes(2048, kBwdStageMax)
                                 ? bwd_smem_bytes(2048, 65536) : bwd_smem_bytes(2048, kBwdStageMax);
      const size_t big512 = bwd_smem_bytes(512, 65536) > bwd_smem_bytes(512, kBwdStageMax)
                                ? bwd_smem_bytes(512, 65536) : bwd_smem_bytes(512, kBwdStageMax);
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd_kernel<2048, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big2048));
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd_kernel<2048, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big2048));
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd_kernel<512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big512));
      GA_CUDA_TRY(cudaFuncSetAttribute(nn_bwd_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big512));
      done_mask.fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  const bool stage = lmax <= kBwdStageMax && g_bwd_stage != 0;
  if (split) {
    a.nparts = (lmax + 511) / 512 < 4 ? (lmax + 511) / 512 : 4;
    auto k = stage ? nn_bwd_kernel<512, true> : nn_bwd_kernel<512, false>;
    k<<<(unsigned)(2 * b * a.nparts), kBwdThreads, bwd_smem_bytes(512, lmax), st>>>(a);
  } else {
    a.nparts = 1;
    auto k = stage ? nn_bwd_kernel<2048, true> : nn_bwd_kernel<2048, false>;
    k<<<(unsigned)(2 * b), kBwdThreads, bwd_smem_bytes(2048, lmax), st>>>(a);
  }
  GA_LAUNCH_CHECK("nn_bwd_kernel");
  return GA_OK;
}

// Forward search + gradient for upstream gradients that are already final when the call is made (a
// Chamfer loss: d loss / d dist is a constant).  Both kernels are launched back to back by the library,
// so nothing else can write grad_dist* between them: gradient CTAs whose batch element is complete run
// to the end while the search is still draining its last wave, instead of parking behind the grid
// dependency (nn_bwd2_kernel).  Same results as ga_nn_distance_fwd followed by ga_nn_distance_bwd.
extern "C" int ga_nn_distance_fwd_bwd(int b, int n, int m, const float* xyz1, const float* xyz2,
                                      const float* grad_dist1, const float* grad_dist2, float* dist1, int* idx1,
                                      float* dist2, int* idx2, float* grad_xyz1, float* grad_xyz2, int mode,
                                      ga_stream_t stream) {
  ga::t_want_tickets = 1;
  int rc = ga_nn_distance_fwd(b, n, m, xyz1, xyz2, dist1, idx1, dist2, idx2, mode, stream);
  ga::t_want_tickets = 0;
  if (rc != GA_OK) return rc;
  ga::t_gd_final = 1;
  rc = ga_nn_distance_bwd(b, n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, stream);
  ga::t_gd_final = 0;
  return rc;
}
