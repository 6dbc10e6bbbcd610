// Chamfer forward with the filter scan on the 5th-generation tensor cores (tcgen05.mma, TMEM).
// Same contract and same bits as nn_fwd_kernel / nn_fwd_mma_kernel (tf_nndistance.cpp:21-43,
// tf_nndistance_g.cu:5-131): the tensor cores only decide WHICH 32-target tiles can hold the
// reference argmin; those tiles are then walked with the fp32 filter and the survivors evaluated
// in the reference arithmetic with the strict-< / lowest-index rule (refine_tiles, nn_mma.cuh).
//
// Operands.  The filter h(q,t) ~ |t|^2 - 2 q.t is the K=16 bf16 contraction of nn_mma.cuh (same 15
// products, hence the same error bound e2 and the same window 2^-14 s^2), in natural K order:
//   A row (query,  C = -2 q_c):  X1  X1  X2  Y1  Y1  Y2  Z1  Z1  | Z2 1  1  1  X2 Y2 Z2 0
//   B row (target, n = fl|t|^2): x1  x2  x1  y1  y2  y1  z1  z2  | z1 n1 n2 n3 x2 y2 z2 0
// Both operands live in shared memory in the canonical no-swizzle K-major layout (8 rows x 16 B
// core matrices, K chunks 128 B apart, 8-row groups 256 B apart); they are built by the CTA from
// the caller's fp32 xyz, so no TMA descriptor is involved.
//
// One tcgen05.mma (M=128 queries = TMEM lanes, N=256 targets = TMEM columns, K=16) produces
// 32768 filter values in ~128 clocks.  The 512 TMEM columns hold two accumulator buffers.  A
// dedicated warp issues the MMAs; 16 warps drain: a thread owns one query row and 64 columns, reads
// them with two tcgen05.ld.32x32b.x32 (32 consecutive columns = one refine tile) and folds each tile
// into a key (tile minimum with a 4-bit tile number in the low mantissa bits) with FMNMX3, keeping
// the three smallest keys.  mbarriers pair the roles: tcgen05.commit -> full[b] -> drain,
// drain -> empty[b] -> issuer; a buffer is handed back as soon as its values sit in registers.
// Measured with clock64 stamps (tools/umma_trace.py): an MMA is visible to the drain warps ~130
// cycles after its issue, so two buffers suffice; what limits the pipeline is the instruction count
// of the issuer loop (its warp gets one issue slot in five), hence the A rows are built by drain warps.
//
// Work decomposition: persistent, one CTA per SM.  All (batch, direction, 128-query M-tile) jobs
// are cut into equal contiguous shares; a share is walked in segments of one target cloud, which
// is staged once per segment (pair-SoA for the refine + the B operand).  After the scan of a
// segment every thread refines the queries of the segment (one query per thread and pass).
// Clouds of 257..2048 points (larger or smaller ones take nn_fwd_mma_kernel / nn_fwd_kernel).
#include <atomic>

#include "nn_mma.cuh"

namespace ga {

constexpr int kUmmaM = 128;        // queries per MMA (TMEM lanes)
constexpr int kUmmaN = 256;        // targets per MMA (TMEM columns of one accumulator buffer)
constexpr int kUmmaBufs = 2;       // accumulator buffers: 2 x 256 = the 512 TMEM columns
constexpr int kUmmaABufs = 4;      // ring of A operands: job jl+2 is built while job jl is drained
constexpr int kUmmaCH = 2048;      // targets of a segment
constexpr int kUmmaWarps = 16;
constexpr int kUmmaThreads = kUmmaWarps * 32;
constexpr int kUmmaMaxJobs = kUmmaCH / kUmmaM;  // M-tiles of one cloud
constexpr int kUmmaT = kMmaT;      // refine tile

// shared memory map (bytes)
constexpr size_t kUmmaOffTgt = 0;                                                  // pair-SoA + pipeline pad
constexpr size_t kUmmaOffRed = kUmmaOffTgt + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;  // red[32]
constexpr size_t kUmmaOffB = kUmmaOffRed + 128;                                    // B operand, 32 B per target
constexpr size_t kUmmaOffA = kUmmaOffB + (size_t)kUmmaCH * 32;                     // A operand ring, 4 x 128 rows
constexpr size_t kUmmaOffKeys = kUmmaOffA + kUmmaABufs * (size_t)kUmmaM * 32;               // [job][3][4][128] float
constexpr size_t kUmmaOffBar = kUmmaOffKeys + (size_t)kUmmaMaxJobs * 3 * 4 * kUmmaM * 4;  // full[4], empty[4], tmem base
constexpr size_t kUmmaSmem = kUmmaOffBar + 80;
static_assert(kUmmaOffB % 128 == 0 && kUmmaOffA % 128 == 0 && kUmmaOffBar % 8 == 0, "operand alignment");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Byte offset of operand row `row`, K chunk 0 (chunk 1 is 128 B further).
__device__ __forceinline__ uint32_t umma_row_off(int row) { return (uint32_t)((row >> 3) * 256 + (row & 7) * 16); }

// Shared-memory matrix descriptor: no swizzle, K-major, LBO (K chunks) 128 B, SBO (8-row groups) 256 B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
         (1ull << 46);
}
// Instruction descriptor: D fp32, A and B bf16, both K-major, N=256, M=128.
// (N >> 3 in bits 17-22, M >> 4 in bits 24-28)
constexpr uint32_t kUmmaIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kUmmaN >> 3) << 17) |
                                ((uint32_t)(kUmmaM >> 4) << 24);

__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kUmmaIdesc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Bounded spin: a completion that never comes (a malformed descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 22); spin++) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]),
        "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]),
        "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM: warp 0 allocates all 512 columns and publishes the base address through shared memory.
__device__ __forceinline__ uint32_t umma_tmem_alloc(uint32_t* slot, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(slot);
}
__device__ __forceinline__ void umma_tmem_free(uint32_t base, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

// A operand row of query (qx,qy,qz); `ok` false gives a zero row (rows past the end of the cloud).
__device__ __forceinline__ void umma_a_row(unsigned char* __restrict__ aop, int row, float qx, float qy, float qz,
                                           bool ok) {
  const float X = ok ? -2.0f * qx : 0.0f, Y = ok ? -2.0f * qy : 0.0f, Z = ok ? -2.0f * qz : 0.0f;
  const float Xr = X - bf16r(X), Yr = Y - bf16r(Y), Zr = Z - bf16r(Z);
  const float one = ok ? 1.0f : 0.0f;
  uint4* p = reinterpret_cast<uint4*>(aop + umma_row_off(row));
  //      slots 0,1          2,3              4,5              6,7
  p[0] = make_uint4(pack_bf16(X, X), pack_bf16(Xr, Y), pack_bf16(Y, Yr), pack_bf16(Z, Z));
  //      slots 8,9          10,11            12,13            14,15
  p[8] = make_uint4(pack_bf16(Zr, one), pack_bf16(one, one), pack_bf16(Xr, Yr), pack_bf16(Zr, 0.0f));
}

// B operand row of staged target `tau` (pair-SoA `tgt`); targets >= cn are padding (h = 1e38).
__device__ __forceinline__ void umma_b_row(unsigned char* __restrict__ bop, const float4* __restrict__ tgt, int tau,
                                           int cn) {
  const bool ok = tau < cn;
  const float* pu = reinterpret_cast<const float*>(tgt + 2 * (tau >> 1)) + (tau & 1);
  const float x = ok ? pu[0] : 0.0f, y = ok ? pu[2] : 0.0f, z = ok ? pu[4] : 0.0f;
  const float n = ok ? pu[6] : 1.0e38f;
  const float xr = x - bf16r(x), yr = y - bf16r(y), zr = z - bf16r(z);
  const float nr = n - bf16r(n);
  const float nr2 = nr - bf16r(nr);
  uint4* p = reinterpret_cast<uint4*>(bop + umma_row_off(tau));
  p[0] = make_uint4(pack_bf16(x, xr), pack_bf16(x, y), pack_bf16(yr, y), pack_bf16(z, zr));
  p[8] = make_uint4(pack_bf16(z, n), pack_bf16(nr, nr2), pack_bf16(xr, yr), pack_bf16(zr, 0.0f));
}

// Fold one 32-target tile (32 TMEM columns of this thread's row) into the three smallest keys.
// Four independent FMNMX3 chains: a dependent FMNMX3 costs far more than its 2 issue cycles, and a
// scheduler has only four drain warps to interleave.
__device__ __forceinline__ void umma_fold(const float (&v)[32], int local_tile, float& c1, float& c2, float& c3) {
  float m0 = fmin3(v[0], v[1], v[2]), m1 = fmin3(v[8], v[9], v[10]);
  float m2 = fmin3(v[16], v[17], v[18]), m3 = fmin3(v[24], v[25], v[26]);
  m0 = fmin3(m0, v[3], v[4]);
  m1 = fmin3(m1, v[11], v[12]);
  m2 = fmin3(m2, v[19], v[20]);
  m3 = fmin3(m3, v[27], v[28]);
  m0 = fmin3(m0, v[5], v[6]);
  m1 = fmin3(m1, v[13], v[14]);
  m2 = fmin3(m2, v[21], v[22]);
  m3 = fmin3(m3, v[29], v[30]);
  m0 = fmin3(m0, v[7], m1);
  m2 = fmin3(m2, v[23], m3);
  const float m = fmin3(fmin3(m0, v[15], v[31]), m2, m2);
  const float key = __int_as_float((__float_as_int(m) & ~15) | local_tile);
  c3 = fminf(c3, fmaxf(c2, key));
  c2 = fminf(c2, fmaxf(c1, key));
  c1 = fminf(c1, key);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// TRACE (development, ga_debug_umma_trace): CTA 0 records clock64() stamps of its pipeline events.
// trace[role * 4096 + 2 * i + {0, 1}], role 0 = issuer (empty wait done, MMA issued), role 1 = drain warp 0
// lane 0 (full wait start, full wait done), role 3 = phases of a
// segment (start, staged, scanned, refined).
#define UMMA_STAMP(role, idx, which)                                                           \
  do {                                                                                         \
    if (TRACE && blockIdx.x == 0 && (idx) < 2048) trace[(role) * 4096 + 2 * (idx) + (which)] = clock64(); \
  } while (0)

// 17 warps: one scheduler holds 5 of them, 16384 / (5 * 32) = 102 registers per thread at most
template <int MODE, bool TRACE = false>
__global__ void __launch_bounds__(kUmmaThreads + 32, 1) nn_fwd_umma_kernel(const FwdArgs a, const int mt1,
                                                                         const int mt2, const long long J,
                                                                         long long* __restrict__ trace) {
  constexpr int THREADS = kUmmaThreads + 32;  // 16 drain warps + the MMA issuer warp
  extern __shared__ __align__(128) unsigned char smem[];
  float4* tgt = reinterpret_cast<float4*>(smem + kUmmaOffTgt);
  float* red = reinterpret_cast<float*>(smem + kUmmaOffRed);
  unsigned char* bop = smem + kUmmaOffB;
  unsigned char* aop = smem + kUmmaOffA;
  float* keys = reinterpret_cast<float*>(smem + kUmmaOffKeys);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kUmmaOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kUmmaOffBar + 64);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool issuer = warp == kUmmaWarps;
  const int quad = warp & 3;         // TMEM lane quadrant this warp may read
  const int slot = (warp >> 2) & 3;  // 64-column quarter of a 256-column accumulator buffer
  const int row = quad * 32 + lane;
  // full[b]: accumulator buffer b written (tcgen05.commit); empty[b]: its values sit in the registers of
  // all 16 drain warps
  const uint32_t full_base = smem_u32(bars), empty_base = smem_u32(bars + kUmmaBufs);
  auto full = [&](unsigned b) { return full_base + 8u * b; };
  auto empty = [&](unsigned b) { return empty_base + 8u * b; };

  if (tid == 0) {
    for (int b = 0; b < kUmmaBufs; b++) {
      mbar_init(full(b), 1);
      mbar_init(empty(b), kUmmaWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = umma_tmem_alloc(tmem_slot, warp);  // contains a __syncthreads
  const uint32_t tmem_row = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * 64);
  // Steps are numbered g = 0, 1, 2, ... over the whole kernel (gbase = first step of the segment, the
  // same in every thread).  Step g uses buffer g & 1; its full barrier completes with parity
  // (g >> 1) & 1, and before the issuer overwrites the buffer it needs the empty barrier of step g - 2
  // (parity ((g >> 1) - 1) & 1).
  unsigned gbase = 0;
  int nseg = 0;

  const long long jpb = (long long)mt1 + mt2;
  const long long j0 = J * blockIdx.x / gridDim.x, j1 = J * (blockIdx.x + 1) / gridDim.x;
  long long j = j0;
  while (j < j1) {
    // ---- segment: M-tiles [j, pend) share one (batch element, direction) ------------------
    const int batch = (int)(j / jpb);
    const long long r = j - (long long)batch * jpb;
    const bool rev = r >= mt1;
    const long long sbeg = (long long)batch * jpb + (rev ? mt1 : 0);
    const long long send = sbeg + (rev ? mt2 : mt1);
    const long long pend = send < j1 ? send : j1;
    const int nj = (int)(pend - j);
    const int ml0 = (int)(j - sbeg);  // first M-tile of the segment within its cloud
    const int nq = rev ? a.m : a.n;
    const int nt = rev ? a.n : a.m;
    const float* qpts = (rev ? a.xyz2 : a.xyz1) + (size_t)batch * nq * 3;
    const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
    const int ntile = (nt + kUmmaT - 1) / kUmmaT;
    const int nblk = (nt + kUmmaN - 1) / kUmmaN;  // MMAs per job (>= 2)

    if (tid == 0) UMMA_STAMP(3, 2 * nseg, 0);
    const float bm = stage_targets<THREADS, kUmmaT>(tgt, red, tpts, 0, nt, ntile, tid);
    for (int tau = tid; tau < nblk * kUmmaN; tau += THREADS) umma_b_row(bop, tgt, tau, nt);
    // A rows of job jl (ring slot jl & 3): query (ml0 + jl) * 128 + rw
    auto a_row = [&](int jl, int rw, float qx, float qy, float qz) {
      umma_a_row(aop + (size_t)(jl & (kUmmaABufs - 1)) * kUmmaM * 32, rw, qx, qy, qz, (ml0 + jl) * kUmmaM + rw < nq);
    };
    auto a_query = [&](int jl, int rw) {  // clamped index of that query
      const int qi = (ml0 + jl) * kUmmaM + rw;
      return qi < nq ? qi : 0;
    };
    if (tid < 2 * kUmmaM && (tid >> 7) < nj) {  // the first two jobs; later ones are built by the drain warps
      const int qs = a_query(tid >> 7, tid & 127);
      a_row(tid >> 7, tid & 127, __ldg(qpts + (size_t)qs * 3), __ldg(qpts + (size_t)qs * 3 + 1),
            __ldg(qpts + (size_t)qs * 3 + 2));
    }
    proxy_fence();
    __syncthreads();
    if (tid == 0) UMMA_STAMP(3, 2 * nseg, 1);

    const int S = nj * nblk;
    if (issuer) {
      // ---- MMA issuer warp: nothing but wait / issue / commit (every instruction of this loop is on
      //      the critical path of the pipeline: the warp shares its scheduler with four drain warps)
      const uint32_t desc_hi = (uint32_t)(256 >> 4) | (1u << 14);  // SBO, descriptor version (bits 32.., 46)
      const uint32_t a_lo = ((smem_u32(aop) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      const uint32_t b_lo = ((smem_u32(bop) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      int jl = 0, t = 0;
      for (int s = 0; s < S; s++) {
        const unsigned g = gbase + s, b = g & 1;
        if (g >= kUmmaBufs) mbar_wait(empty(b), ((g >> 1) - 1) & 1);
        tc_fence_after();
        UMMA_STAMP(0, g, 0);
        const uint64_t adesc = ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)(jl & (kUmmaABufs - 1)) * (kUmmaM * 32 / 16));
        const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)t * (kUmmaN * 32 / 16));
        if (lane == 0) {
          umma_issue(tmem + b * kUmmaN, adesc, bdesc);
          umma_commit(full(b));
        }
        __syncwarp();
        UMMA_STAMP(0, g, 1);
        if (++t == nblk) {
          t = 0;
          jl++;
        }
      }
    } else {
      // ---- drain warps: thread = (query row, 64-column quarter).  Two 32-column tiles per step; the
      //      TMEM load of the next tile is in flight while the current one is folded, and the buffer is
      //      handed back as soon as both tiles are in registers.  Warps 0-3 (one thread per row) also
      //      build the A rows of job jl + 2: coordinates loaded at the job's first step, row written at
      //      its second; the issuer reads them only after it has seen the empty barrier of a later
      //      step, which every warp arrives at after this point in program order.
      float c1 = kMmaBig, c2 = kMmaBig, c3 = kMmaBig;
      int jl = 0, t = 0;
      float v[32], w[32];
      float nx = 0.0f, ny = 0.0f, nz = 0.0f;
      {
        const unsigned g = gbase;
        mbar_wait(full(g & 1), (g >> 1) & 1);
        tc_fence_after();
        tmem_ld32(tmem_row + (g & 1) * kUmmaN, v);
        tmem_ld_wait();
      }
      for (int s = 0; s < S; s++) {
        const unsigned g = gbase + s, b = g & 1;
        tmem_ld32(tmem_row + b * kUmmaN + 32, w);
        if (slot == 0 && jl + 2 < nj) {
          if (t == 0) {
            const int qs = a_query(jl + 2, row);
            nx = __ldg(qpts + (size_t)qs * 3);
            ny = __ldg(qpts + (size_t)qs * 3 + 1);
            nz = __ldg(qpts + (size_t)qs * 3 + 2);
          } else if (t == 1) {
            a_row(jl + 2, row, nx, ny, nz);
            proxy_fence();
          }
        }
        umma_fold(v, 2 * t, c1, c2, c3);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(empty(b));
        if (s + 1 < S) {
          if (lane == 0 && warp == 0) UMMA_STAMP(1, g + 1, 0);
          mbar_wait(full(b ^ 1), ((g + 1) >> 1) & 1);
          if (lane == 0 && warp == 0) UMMA_STAMP(1, g + 1, 1);
          tc_fence_after();
          tmem_ld32(tmem_row + (b ^ 1) * kUmmaN, v);
        }
        umma_fold(w, 2 * t + 1, c1, c2, c3);
        tmem_ld_wait();
        if (++t == nblk) {  // job finished: publish this thread's keys, reset
          float* kq = keys + (size_t)jl * (3 * 4 * kUmmaM) + slot * kUmmaM + row;
          kq[0] = c1;
          kq[4 * kUmmaM] = c2;
          kq[8 * kUmmaM] = c3;
          c1 = c2 = c3 = kMmaBig;
          t = 0;
          jl++;
        }
      }
    }
    __syncthreads();  // keys of the whole segment are published
    if (tid == 0) UMMA_STAMP(3, 2 * nseg + 1, 0);

    // ---- refine: one query per thread and pass ---------------------------------------------
    const float t0x = __ldg(tpts), t0y = __ldg(tpts + 1), t0z = __ldg(tpts + 2);
    float* odist = (rev ? a.dist2 : a.dist1) + (size_t)batch * nq;
    int* oidx = (rev ? a.idx2 : a.idx1) + (size_t)batch * nq;
    float* mdist = rev ? a.mdist2 : a.mdist1;
    int* midx = rev ? a.midx2 : a.midx1;
    for (int base = 0; base < nj * kUmmaM && !issuer; base += kUmmaThreads) {
      const int ql = base + tid;                  // query within the segment
      const int qi = ml0 * kUmmaM + ql;           // query within its cloud
      QueryState<1> qs;
      qs.valid[0] = ql < nj * kUmmaM && qi < nq;
      const int qsafe = qs.valid[0] ? qi : 0;
      qs.qx[0] = __ldg(qpts + (size_t)qsafe * 3);
      qs.qy[0] = __ldg(qpts + (size_t)qsafe * 3 + 1);
      qs.qz[0] = __ldg(qpts + (size_t)qsafe * 3 + 2);
      qs.qabs[0] = query_abs(qs.qx[0], qs.qy[0], qs.qz[0]);
      qs.ax2[0] = -2.0f * qs.qx[0];
      qs.ay2[0] = -2.0f * qs.qy[0];
      qs.az2[0] = -2.0f * qs.qz[0];
      qs.d0[0] = sqdist<MODE>(t0x, t0y, t0z, qs.qx[0], qs.qy[0], qs.qz[0]);
      qs.best[0] = __int_as_float(0x7f800000);
      qs.besti[0] = 0;
      qs.m1g[0] = __int_as_float(0x7f800000);

      int cnt[1] = {0}, ta[1] = {0}, tb[1] = {0};
      float thr[1] = {0.0f};
      if (qs.valid[0]) {
        const float* kq = keys + (size_t)(ql >> 7) * (3 * 4 * kUmmaM) + (ql & 127);
        float k[3][4];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int sl = 0; sl < 4; sl++) k[i][sl] = kq[(i * 4 + sl) * kUmmaM];
        const float mn = fminf(fminf(k[0][0], k[0][1]), fminf(k[0][2], k[0][3]));
        thr[0] = mn + mma_window(qs.qabs[0], bm);
        int c = 0;
#pragma unroll
        for (int sl = 0; sl < 4; sl++) {
#pragma unroll
          for (int i = 0; i < 2; i++) {
            if (!(k[i][sl] > thr[0])) {
              const int loc = __float_as_int(k[i][sl]) & 15;      // 2 * (MMA block) + half
              const int tile = (loc >> 1) * 8 + sl * 2 + (loc & 1);
              if (c == 0) ta[0] = tile;
              if (c == 1) tb[0] = tile;
              c++;
            }
          }
          if (!(k[2][sl] > thr[0])) c += 3;  // a third tile of one thread: exact scan
        }
        // a tile beyond the staged tiles can only come from padding under a non-finite window
        if ((c >= 1 && ta[0] >= ntile) || (c >= 2 && tb[0] >= ntile)) c = 3;
        cnt[0] = c;
      }
      refine_tiles<MODE, 1>(qs, tgt, 0, nt, ntile, cnt, ta, tb, thr);
      if (qs.valid[0]) {
        float d;
        int i;
        finish_query<1>(qs, 0, d, i);
        odist[qi] = d;
        oidx[qi] = i;
        if (mdist != nullptr) {
          mdist[(size_t)batch * nq + qi] = d;
          midx[(size_t)batch * nq + qi] = i;
        }
      }
    }
    if (tid == 0) UMMA_STAMP(3, 2 * nseg + 1, 1);
    nseg++;
    j = pend;
    gbase += (unsigned)S;
    // the next segment's stage_targets starts with a __syncthreads: tgt / keys are free by then
  }
  umma_tmem_free(tmem, warp);
}

// Debug / evidence: raw tcgen05 filter values h(q,t) of one cloud pair (n queries, m <= 2048
// targets), out[q*m + t]; same operand builders, descriptors and TMEM reads as the product kernel.
__global__ void __launch_bounds__(128) umma_filter_dump_kernel(int n, int m, const float* __restrict__ q,
                                                               const float* __restrict__ tp, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float4* tgt = reinterpret_cast<float4*>(smem + kUmmaOffTgt);
  float* red = reinterpret_cast<float*>(smem + kUmmaOffRed);
  unsigned char* bop = smem + kUmmaOffB;
  unsigned char* aop = smem + kUmmaOffA;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kUmmaOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kUmmaOffBar + 64);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t full0 = smem_u32(bars);
  if (tid == 0) {
    mbar_init(full0, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = umma_tmem_alloc(tmem_slot, warp);
  const int ntile = (m + kUmmaT - 1) / kUmmaT, nblk = (m + kUmmaN - 1) / kUmmaN;
  stage_targets<128, kUmmaT>(tgt, red, tp, 0, m, ntile, tid);
  for (int tau = tid; tau < nblk * kUmmaN; tau += 128) umma_b_row(bop, tgt, tau, m);
  const int qi = blockIdx.x * kUmmaM + tid;
  const int qs = qi < n ? qi : 0;
  umma_a_row(aop, tid, __ldg(q + (size_t)qs * 3), __ldg(q + (size_t)qs * 3 + 1), __ldg(q + (size_t)qs * 3 + 2), qi < n);
  proxy_fence();
  __syncthreads();
  uint32_t phase = 0;
  for (int t = 0; t < nblk; t++) {
    if (tid == 0) {
      tc_fence_after();
      umma_issue(tmem, umma_desc(smem_u32(aop)), umma_desc(smem_u32(bop + (size_t)t * kUmmaN * 32)));
      umma_commit(full0);
    }
    mbar_wait(full0, phase);
    phase ^= 1;
    tc_fence_after();
    for (int c = 0; c < kUmmaN; c += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; e++) {
        const int tau = t * kUmmaN + c + e;
        if (qi < n && tau < m) out[(size_t)qi * m + tau] = v[e];
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  umma_tmem_free(tmem, warp);
}

int g_umma_grid = 0;  // tuning hook (key 12): CTAs of the tcgen05 kernel (0 = one per SM)

bool fwd_umma_supported(int n, int m) { return n > kUmmaN && m > kUmmaN && n <= kUmmaCH && m <= kUmmaCH; }

static int launch_fwd_umma_impl(const FwdArgs& a, int mode, cudaStream_t st, long long* trace) {
  if (!fwd_umma_supported(a.n, a.m)) {
    set_error("nn_fwd_umma_kernel: clouds of %d..%d points", kUmmaN + 1, kUmmaCH);
    return GA_ERR_UNSUPPORTED;
  }
  const int mt1 = (a.n + kUmmaM - 1) / kUmmaM, mt2 = (a.m + kUmmaM - 1) / kUmmaM;
  const long long J = (long long)a.b * (mt1 + mt2);
  if (J <= 0) return GA_OK;
  auto k = trace != nullptr ? nn_fwd_umma_kernel<GA_MODE_CPU_EXACT, true>
                            : (mode == GA_MODE_CPU_EXACT ? nn_fwd_umma_kernel<GA_MODE_CPU_EXACT, false>
                                                         : nn_fwd_umma_kernel<GA_MODE_GPU_REF, false>);
  {
    static std::atomic<unsigned> done_mask[3];
    const int slot = trace != nullptr ? 2 : mode;
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[slot].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmem));
      done_mask[slot].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  long long grid = g_umma_grid > 0 ? g_umma_grid : sm_count();
  if (grid > J) grid = J;
  k<<<(unsigned)grid, kUmmaThreads + 32, kUmmaSmem, st>>>(a, mt1, mt2, J, trace);
  GA_LAUNCH_CHECK("nn_fwd_umma_kernel");
  return GA_OK;
}

int launch_fwd_umma(const FwdArgs& a, int mode, cudaStream_t st) { return launch_fwd_umma_impl(a, mode, st, nullptr); }

}  // namespace ga

extern "C" int ga_debug_umma_filter(int n, int m, const float* xyz1, const float* xyz2, float* out,
                                    ga_stream_t stream) {
  if (n <= 0 || m <= 0 || m > ga::kUmmaCH) {
    ga::set_error("ga_debug_umma_filter: need n > 0 and 0 < m <= %d", ga::kUmmaCH);
    return GA_ERR_INVALID_ARGUMENT;
  }
  cudaStream_t st = ga::as_stream(stream);
  auto k = ga::umma_filter_dump_kernel;
  GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ga::kUmmaSmem));
  k<<<(n + ga::kUmmaM - 1) / ga::kUmmaM, 128, ga::kUmmaSmem, st>>>(n, m, xyz1, xyz2, out);
  GA_LAUNCH_CHECK("umma_filter_dump_kernel");
  return GA_OK;
}

// Development: one traced run of the tcgen05 kernel (mode 0); `trace` = 4 * 4096 int64 on the device.
extern "C" int ga_debug_umma_trace(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                                   float* dist2, int* idx2, long long* trace, ga_stream_t stream) {
  ga::FwdArgs a;
  a.b = b; a.n = n; a.m = m;
  a.xyz1 = xyz1; a.xyz2 = xyz2;
  a.dist1 = dist1; a.idx1 = idx1; a.dist2 = dist2; a.idx2 = idx2;
  a.tiles1 = a.tiles2 = 0;
  a.mdist1 = a.mdist2 = nullptr;
  a.ticket = nullptr;
  a.call_id = 0;
  a.ticket_debug = 0;
  a.midx1 = a.midx2 = nullptr;
  return ga::launch_fwd_umma_impl(a, GA_MODE_CPU_EXACT, ga::as_stream(stream), trace);
}
