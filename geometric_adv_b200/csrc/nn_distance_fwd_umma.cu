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
// 32768 filter values in ~128 clocks; the SM's ALU pipe needs ~350 clocks to min-reduce them (one
// FMNMX3 per two values), so the tensor pipe is never the limiter and the whole design is about
// keeping the ALU pipe fed: see the kernel comment below.
// Clouds of 257..2048 points (larger or smaller ones take nn_fwd_mma_kernel / nn_fwd_kernel).
#include <atomic>

#include "nn_mma.cuh"

namespace ga {

constexpr int kUmmaM = 128;        // queries per MMA (TMEM lanes)
constexpr int kUmmaN = 256;        // targets per MMA (TMEM columns of one accumulator buffer)
constexpr int kUmmaCH = 2048;      // targets of a segment
constexpr int kUmmaT = kMmaT;      // refine tile

// shared memory map (bytes)
constexpr size_t kUmmaOffTgt = 0;                                                  // pair-SoA + pipeline pad
constexpr size_t kUmmaOffRed = kUmmaOffTgt + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;  // red[32]
constexpr size_t kUmmaOffB = kUmmaOffRed + 128;                                    // B operand, 32 B per target
constexpr size_t kUmmaOffA = kUmmaOffB + (size_t)kUmmaCH * 32;                     // A operand, 128 rows
constexpr size_t kUmmaOffBar = kUmmaOffA + (size_t)kUmmaM * 32;                    // full barrier, tmem base
constexpr size_t kUmmaSmem = kUmmaOffBar + 80;  // (layout of the dump kernel; the product kernel has its own, kTc*)
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
// 64 columns as two x32 loads in flight together (a single x64 load needs 82 registers: six warps per scheduler leave 80).
__device__ __forceinline__ void tmem_ld32x2(uint32_t taddr, float (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr + 0)
      : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(v[32]), "=f"(v[33]), "=f"(v[34]), "=f"(v[35]), "=f"(v[36]), "=f"(v[37]), "=f"(v[38]), "=f"(v[39]), "=f"(v[40]), "=f"(v[41]), "=f"(v[42]), "=f"(v[43]), "=f"(v[44]), "=f"(v[45]), "=f"(v[46]), "=f"(v[47]), "=f"(v[48]), "=f"(v[49]), "=f"(v[50]), "=f"(v[51]), "=f"(v[52]), "=f"(v[53]), "=f"(v[54]), "=f"(v[55]), "=f"(v[56]), "=f"(v[57]), "=f"(v[58]), "=f"(v[59]), "=f"(v[60]), "=f"(v[61]), "=f"(v[62]), "=f"(v[63])
      : "r"(taddr + 32)
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

// Fold one 32-target tile (32 TMEM columns of this thread's row) into the three smallest keys; the key is
// the tile minimum with the tile number (0..63) in its low 6 mantissa bits (mma_window_wide absorbs the
// perturbation).  Four independent FMNMX3 chains: a dependent FMNMX3 costs far more than its 2 issue cycles.
__device__ __forceinline__ void umma_fold(const float (&v)[32], int tile, float& c1, float& c2, float& c3) {
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
  const float key = __int_as_float((__float_as_int(m) & ~63) | tile);
  c3 = fminf(c3, fmaxf(c2, key));
  c2 = fminf(c2, fmaxf(c1, key));
  c1 = fminf(c1, key);
}

// Two adjacent tiles out of one 64-column load: eight independent FMNMX3 chains.
__device__ __forceinline__ float umma_tile_min(const float (&v)[64], int o) {
  float m0 = fmin3(v[o + 0], v[o + 1], v[o + 2]), m1 = fmin3(v[o + 8], v[o + 9], v[o + 10]);
  float m2 = fmin3(v[o + 16], v[o + 17], v[o + 18]), m3 = fmin3(v[o + 24], v[o + 25], v[o + 26]);
  m0 = fmin3(m0, v[o + 3], v[o + 4]);
  m1 = fmin3(m1, v[o + 11], v[o + 12]);
  m2 = fmin3(m2, v[o + 19], v[o + 20]);
  m3 = fmin3(m3, v[o + 27], v[o + 28]);
  m0 = fmin3(m0, v[o + 5], v[o + 6]);
  m1 = fmin3(m1, v[o + 13], v[o + 14]);
  m2 = fmin3(m2, v[o + 21], v[o + 22]);
  m3 = fmin3(m3, v[o + 29], v[o + 30]);
  m0 = fmin3(m0, v[o + 7], m1);
  m2 = fmin3(m2, v[o + 23], m3);
  return fmin3(fmin3(m0, v[o + 15], v[o + 31]), m2, m2);
}
__device__ __forceinline__ void umma_fold2(const float (&v)[64], int tile, float& c1, float& c2, float& c3) {
  const float ma = umma_tile_min(v, 0), mb = umma_tile_min(v, 32);
  const float ka = __int_as_float((__float_as_int(ma) & ~63) | tile);
  const float kb = __int_as_float((__float_as_int(mb) & ~63) | (tile + 1));
  // (c1 <= c2 <= c3) := the three smallest of {c1, c2, c3, ka, kb}
  const float lo = fminf(ka, kb), hi = fmaxf(ka, kb);
  const float x = fmaxf(c1, lo), y = fminf(c2, hi);
  c3 = fminf(fmaxf(x, y), c3);
  c2 = fminf(x, y);
  c1 = fminf(c1, lo);
}

// ---- the kernel ---------------------------------------------------------------------------------
// Persistent, one CTA of 24 warps per SM, warp-specialised; a job = 128 queries of one cloud against the whole
// other cloud; a CTA owns a contiguous range of jobs (J / grid: 10.8 at BASELINE config 2, 1.8 % quantisation).
//   scan    (warps 0-15)  the only warps that touch TMEM.  All sixteen work on the same job: warp w reads the TMEM
//                         lane quadrant w & 3 (= 32 query rows) and column quarter w >> 2 (two 32-target tiles) of
//                         every step's accumulator (128 queries x 256 targets), so a THREAD owns one query row and a
//                         quarter of the targets, and the three smallest tile keys of that quarter live in three
//                         registers for the whole job.  Per step and warp: 2 x (tcgen05.ld.32x32b.x32 -> 16 FMNMX3 ->
//                         key, 5 FMNMX).  tools/drainbench.cu: on a scheduler a tile costs its TMEM load (~22 clk of
//                         register-file writes) PLUS its fold (~24 clk of ALU), the two do not overlap; four scan
//                         warps per scheduler hide each other's load and dependency latencies, so the scheduler
//                         runs at that rate (8 tiles x ~47 clk per step).  After a job's last step each thread
//                         writes its keys to a 2-deep ring.
//   issuer  (warp 23)     one thread feeds the tensor core: step s goes to accumulator s & 1 as soon as all sixteen
//                         scan warps have read step s - 2 out of it (acc_empty), tcgen05.commit -> acc_full.
//   helper  (warps 16-19) builds the A operand of job k + 2 as soon as job k's keys have arrived (its MMAs are then
//                         complete), merges the four column quarters' keys and refines (refine_tiles, one query per
//                         thread), writes dist / idx.
//   stager  (warps 20-22) stages the target cloud of the next (batch, direction) into the free one of two slots
//                         (pair-SoA for the refine + B operand) while the other one is in use.  The CTA's first
//                         cloud is staged by all 24 warps before the roles split.
// Roles meet only through mbarriers (a_full, keys_full, keys_empty, slot_full, acc_full, acc_empty) and one counter
// per slot (refined jobs): the scan warps never wait for a global load, a refine or a staging pass.
constexpr int kTcThreads = 768;               // 24 warps, 80 registers each (six warps per scheduler)
constexpr int kTcScanWarps = 16;
constexpr int kTcStagers = 96;               // threads of the stager role (warps 20-22)
constexpr int kTcN = 256;                    // targets per MMA = columns of one accumulator; two accumulators
constexpr int kTcRing = 2;                   // A operands / key sets in flight
constexpr size_t kTcSlot = (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32 + (size_t)kUmmaCH * 32;  // one staged cloud
constexpr size_t kTcOffA = 2 * kTcSlot;                                        // [2] A operands, 128 rows x 32 B
constexpr size_t kTcOffKeys = kTcOffA + (size_t)kTcRing * kUmmaM * 32;         // [2][4][3][128] float
constexpr size_t kTcOffRed = kTcOffKeys + (size_t)kTcRing * 4 * 3 * kUmmaM * 4;  // warp maxima [24] + bm[2]
constexpr size_t kTcOffBar = kTcOffRed + 128;  // a_full[2] keys_full[2] keys_empty[2] slot_full[2] acc_full[2] acc_empty[2] | cnt[2] | tmem
constexpr size_t kTcSmem = kTcOffBar + 12 * 8 + 8 + 8;
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) |
                              ((uint32_t)(kUmmaM >> 4) << 24);
static_assert(kTcSlot % 128 == 0 && kTcOffA % 128 == 0 && kTcOffBar % 8 == 0, "operand alignment");
constexpr int kTcTrace = 1408;  // development trace, 32-bit clock stamps
static_assert(kTcSmem + kTcTrace * 4 <= 232448, "shared memory of one SM");

__device__ __forceinline__ void umma_issue_idesc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Wait for the phase with parity `parity`: a non-blocking test first (the usual case: long complete), then
// the suspending try_wait.  Bounded: a completion that never comes traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  if (!mbar_test(bar, parity)) mbar_wait(bar, parity);
}
// Spin on test_wait: for the single issuer thread, whose wake-up latency is on the accumulator's critical path.
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 22); spin++)
    if (mbar_test(bar, parity)) return;
  __trap();
}

struct TcJob {  // one 128-query job
  int batch, rev, ml, cs;  // batch element, direction, M-tile within its cloud, cloud sequence number within the CTA
};
__device__ __forceinline__ void tc_job_next(TcJob& jb, int mt1, int mt2) {
  if (++jb.ml == (jb.rev ? mt2 : mt1)) {
    jb.ml = 0;
    jb.cs++;
    jb.batch += jb.rev;
    jb.rev ^= 1;
  }
}

// Pair-SoA of cloud `tpts` (nt points) by threads tix of nth; returns this thread's max |coordinate|.
__device__ __forceinline__ float tc_stage_pairs(float4* __restrict__ tgt, const float* __restrict__ tpts, int nt,
                                                int npair, int tix, int nth) {
  const float kInf = __int_as_float(0x7f800000);
  float lmax = 0.0f;
  for (int p0 = tix; p0 < npair; p0 += nth * 4) {  // four pairs per thread in flight
    float c[4][6];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int p = p0 + u * nth;
#pragma unroll
      for (int e = 0; e < 6; e++) c[u][e] = (p < npair && 2 * p + (e >= 3) < nt) ? __ldg(tpts + (size_t)p * 6 + e) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int p = p0 + u * nth;
      if (p >= npair) break;
      float n0 = kInf, n1 = kInf;
      if (2 * p < nt) {
        n0 = fmaf(c[u][2], c[u][2], fmaf(c[u][1], c[u][1], c[u][0] * c[u][0]));
        lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(c[u][0]), fabsf(c[u][1])), fabsf(c[u][2])));
      }
      if (2 * p + 1 < nt) {
        n1 = fmaf(c[u][5], c[u][5], fmaf(c[u][4], c[u][4], c[u][3] * c[u][3]));
        lmax = fmaxf(lmax, fmaxf(fmaxf(fabsf(c[u][3]), fabsf(c[u][4])), fabsf(c[u][5])));
      }
      tgt[2 * p] = make_float4(c[u][0], c[u][3], c[u][1], c[u][4]);
      tgt[2 * p + 1] = make_float4(c[u][2], c[u][5], n0, n1);
    }
  }
  return lmax;
}

// DEV: development build with the trace and the timing switches `dbg` (bit 0: helper skips the refine, bit 1: scan
// skips the drain); the product build has neither in its loops.
template <int MODE, bool DEV>
__global__ void __launch_bounds__(kTcThreads, 1) nn_fwd_umma_kernel(const FwdArgs a, const int mt1, const int mt2,
                                                                   const long long J, const int dbg_in,
                                                                   long long* __restrict__ trace_in) {
  const int dbg = DEV ? (dbg_in & 255) : 0;
  long long* const trace = DEV ? trace_in : nullptr;
  constexpr int N = kTcN, NT = kTcN / 32;
  // development trace (ga_debug_umma_trace): CTA 0 stamps clock() into shared memory (a global store in front of a
  // releasing mbarrier arrive would put an L2 round trip on the path it measures) and copies them out at the end:
  //   [2 s], [2 s + 1]         issuer: before the acc_empty wait of step s / MMA of step s issued      (s < 128)
  //   [256 + 2 s], [.. + 1]    scan warp 0: acc_full of step s seen / accumulator handed back            (s < 128)
  //   [512 + 2 k], [.. + 1]    helper: keys of job k seen / job k refined and written                    (k < 64)
  //   [640 + 2 c], [.. + 1]    stager: staging of cloud c begins / ends                                   (c < 32)
  //   [704], [705]             TMEM allocated / roles begin
  //   [768 + 128 g + s]        scan warp 4 g: accumulator of step s handed back                           (s < 128)
  //   [1280 + s]               scan warp 0: step s folded                                                  (s < 128)
  //   [4096 + 4 c ..]          every CTA c, globaltimer ns: entry / TMEM allocated / roles begin / exit
  const bool tr = DEV && trace != nullptr && blockIdx.x == (unsigned)(dbg_in >> 8);  // dbg bits 8.. = the traced CTA
  if (trace != nullptr && threadIdx.x == 0) trace[4096 + 4 * blockIdx.x] = (long long)global_ns();
  asm volatile("griddepcontrol.launch_dependents;");
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* strace = reinterpret_cast<uint32_t*>(smem + kTcSmem);  // kTcTrace words, only with `trace`
  unsigned char* aops = smem + kTcOffA;
  float* keyring = reinterpret_cast<float*>(smem + kTcOffKeys);
  float* red = reinterpret_cast<float*>(smem + kTcOffRed);  // [0..23] warp maxima, [24..25] bm of the slots
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kTcOffBar);
  volatile int* done_cnt = reinterpret_cast<volatile int*>(smem + kTcOffBar + 12 * 8);  // [2] refined jobs per slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kTcOffBar + 12 * 8 + 8);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto keys_full = [&](int i) { return bar0 + 8u * (2 + i); };
  auto keys_empty = [&](int i) { return bar0 + 8u * (4 + i); };
  auto slot_full = [&](int i) { return bar0 + 8u * (6 + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (8 + i); };
  auto acc_empty = [&](int i) { return bar0 + 8u * (10 + i); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 0-3 scan (column quarters), 4 helper, 5 stager (three warps), 6 issuer
  const int role = warp < kTcScanWarps ? warp >> 2 : (warp < 20 ? 4 : (warp < 23 ? 5 : 6));
  const int row = (warp & 3) * 32 + lane;  // thread within its group of four warps = query row of a job
  if (tr)
    for (int i = threadIdx.x; i < kTcTrace; i += kTcThreads) strace[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 2; i++) {
      // one arrival per WARP everywhere (lane 0 behind a __syncwarp): 32 lanes arriving on one barrier word are 32
      // serialised shared-memory atomics
      mbar_init(a_full(i), 4);
      mbar_init(keys_full(i), kTcScanWarps);
      mbar_init(keys_empty(i), 4);
      mbar_init(slot_full(i), kTcStagers / 32);
      mbar_init(acc_full(i), 1);
      mbar_init(acc_empty(i), kTcScanWarps);  // one arrival per scan warp
    }
    done_cnt[0] = 0;
    done_cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = umma_tmem_alloc(tmem_slot, warp);  // contains a __syncthreads
  if (tr && tid == 0) strace[704] = (uint32_t)clock();
  if (trace != nullptr && tid == 0) trace[4096 + 4 * blockIdx.x + 1] = (long long)global_ns();

  const long long jpb = (long long)mt1 + mt2;
  const long long j0 = J * blockIdx.x / gridDim.x, j1 = J * (blockIdx.x + 1) / gridDim.x;
  const int nj = (int)(j1 - j0);
  TcJob first;
  first.batch = (int)(j0 / jpb);
  {
    const int r = (int)(j0 - (long long)first.batch * jpb);
    first.rev = r >= mt1;
    first.ml = first.rev ? r - mt1 : r;
    first.cs = 0;
  }
  const int cid0 = first.batch * 2 + first.rev;  // first cloud (batch * 2 + direction)

  // ---- the CTA's first cloud -> slot 0, by everybody -----------------------------------------------------------
  {
    const int nt = first.rev ? a.n : a.m;
    const float* tpts = (first.rev ? a.xyz1 : a.xyz2) + (size_t)first.batch * nt * 3;
    float4* tgt = reinterpret_cast<float4*>(smem);
    unsigned char* bop = smem + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;
    const int spj = (nt + N - 1) / N;
    float lmax = tc_stage_pairs(tgt, tpts, nt, spj * (N / 2), tid, kTcThreads);
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    if (tid == 0) {
      float m = red[0];
      for (int w = 1; w < kTcThreads / 32; w++) m = fmaxf(m, red[w]);
      red[24] = m;
    }
    for (int tau = tid; tau < spj * N; tau += kTcThreads) umma_b_row(bop, tgt, tau, nt);
    proxy_fence();
    __syncthreads();
  }
  if (tr && tid == 0) strace[705] = (uint32_t)clock();
  if (trace != nullptr && tid == 0) trace[4096 + 4 * blockIdx.x + 2] = (long long)global_ns();

  if (role == 5) {
    // ---- stager: cloud cs -> slot cs & 1, as soon as every job of cloud cs - 2 has been refined ----------
    const int tix = tid - 20 * 32;
    int last_cs;
    {
      const long long jl = j1 - 1;
      const int lb = (int)(jl / jpb);
      last_cs = lb * 2 + (((int)(jl - (long long)lb * jpb)) >= mt1 ? 1 : 0) - cid0;
    }
    if (lane == 0) mbar_arrive(slot_full(0));  // slot 0 was staged above
    for (int cs = 1; cs <= last_cs; cs++) {
      const int cid = cid0 + cs, batch = cid >> 1, rev = cid & 1, slot = cs & 1;
      const int nt = rev ? a.n : a.m;
      if (cs >= 2) {
        // jobs of cloud cs - 2 inside this CTA's range
        const int pc = cid - 2, pb = pc >> 1, pr = pc & 1;
        const long long cb = (long long)pb * jpb + (pr ? mt1 : 0), ce = cb + (pr ? mt2 : mt1);
        const int want = (int)((ce < j1 ? ce : j1) - (cb > j0 ? cb : j0));
        if (tix == 0) {
          while (done_cnt[slot] != want) __nanosleep(200);
          done_cnt[slot] = 0;
        }
        asm volatile("bar.sync 4, 96;" ::: "memory");
      }
      if (tr && tix == 0 && cs < 32) strace[640 + 2 * cs] = (uint32_t)clock();
      const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
      float4* tgt = reinterpret_cast<float4*>(smem + slot * kTcSlot);
      unsigned char* bop = smem + slot * kTcSlot + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;
      const int spj = (nt + N - 1) / N;
      float lmax = tc_stage_pairs(tgt, tpts, nt, spj * (N / 2), tix, kTcStagers);
      lmax = warp_max(lmax);
      if (lane == 0) red[warp] = lmax;
      asm volatile("bar.sync 4, 96;" ::: "memory");  // pair-SoA complete: the B rows are built from it
      if (tix == 0) red[24 + slot] = fmaxf(fmaxf(red[20], red[21]), red[22]);
      for (int tau = tix; tau < spj * N; tau += kTcStagers) umma_b_row(bop, tgt, tau, nt);
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_full(slot));
      if (tr && tix == 0 && cs < 32) strace[640 + 2 * cs + 1] = (uint32_t)clock();
    }
  } else if (role < 4) {
    // ---- scan --------------------------------------------------------------------------------------------------
    // Nothing but: wait acc_full, two tiles, arrive acc_empty.  The state of the NEXT step's barrier is sampled in
    // the middle of the drain, so that the usual case (its MMA finished long ago) costs no latency at the top of
    // the next step.
    const int g = role;
    // everything the step loop needs sits in a handful of registers (made opaque, or the compiler rebuilds the
    // shared-memory window base from a special register on every use)
    uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * (N / 4));
    uint32_t bfull = acc_full(0), bempty = acc_empty(0);
    asm volatile("" : "+r"(tmem_row), "+r"(bfull), "+r"(bempty));
    const int spj1 = (a.m + N - 1) / N, spj2 = (a.n + N - 1) / N;  // steps per job, direction 1 -> 2 and 2 -> 1
    unsigned ns = 0;  // steps drained so far: step s uses accumulator s & 1, barrier parity (s >> 1) & 1
    bool ready = false;
    TcJob jb = first;
    for (int k = 0; k < nj; k++) {
      const int spj = jb.rev ? spj2 : spj1;
      float c1 = kMmaBig, c2 = kMmaBig, c3 = kMmaBig;
      int tile = g * (NT / 4);
#pragma unroll 1
      for (int t = 0; t < spj; t++, tile += NT, ns++) {
        const uint32_t bo = (ns & 1) * 8;
        if (!ready) mbar_wait_fast(bfull + bo, (ns >> 1) & 1);
        tc_fence_after();
        if (DEV && tr && tid == 0 && ns < 128) strace[256 + 2 * ns] = (uint32_t)clock();
        if (DEV && (dbg & 2)) {  // timing experiment: no TMEM loads, no fold
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bempty + bo);
          ready = mbar_test(bfull + (bo ^ 8), ((ns + 1) >> 1) & 1);
        } else {
          float v[64];
          tmem_ld32x2(tmem_row + (ns & 1) * N, v);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bempty + bo);  // all values of this warp's quarter sit in registers
          if (DEV && tr && lane == 0 && (warp & 3) == 0 && ns < 128) {
            const uint32_t now = (uint32_t)clock();
            strace[768 + 128 * g + ns] = now;
            if (warp == 0) strace[256 + 2 * ns + 1] = now;
          }
          ready = mbar_test(bfull + (bo ^ 8), ((ns + 1) >> 1) & 1);  // consumed at the top of the next step
          umma_fold2(v, tile, c1, c2, c3);
        }
        if (DEV && tr && tid == 0 && ns < 128) strace[1280 + ns] = (uint32_t)clock();
      }
      // this column quarter's keys of job k -> ring slot k & 1 (free once the helper has read job k - 2's); every
      // address is an offset from the one shared-memory address the step loop holds
      if (DEV && (dbg & 8)) { tc_job_next(jb, mt1, mt2); continue; }
      const uint32_t kb = (uint32_t)(k & 1);
      if (k >= kTcRing) mbar_wait_fast(bfull - 8u * 8 + 8u * 4 + 8u * kb, ((k >> 1) - 1) & 1);  // keys_empty
      const uint32_t kaddr = bfull - (uint32_t)(kTcOffBar + 8 * 8 - kTcOffKeys) + (kb * 4 + g) * (3 * kUmmaM * 4) + row * 4;
      asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+512], %2;\n\tst.shared.f32 [%0+1024], %3;" ::"r"(kaddr), "f"(c1), "f"(c2), "f"(c3) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bfull - 8u * 8 + 8u * 2 + 8u * kb);  // keys_full
      tc_job_next(jb, mt1, mt2);
    }
  } else if (role == 6) {
    // ---- issuer: one thread feeds the tensor core.  Step s goes to accumulator s & 1 as soon as the sixteen scan
    // warps have read step s - 2 out of it; the first step of a job also needs the job's A operand and its cloud.
    if (lane == 0) {
      const uint32_t desc_hi = (uint32_t)(256 >> 4) | (1u << 14);  // SBO, descriptor version (bits 32.., 46)
      const uint32_t a_lo = ((smem_u32(aops) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      unsigned s_no = 0;
      TcJob jb = first;
      for (int k = 0; k < nj; k++) {
        const int nt = jb.rev ? a.n : a.m;
        const int spj = (nt + N - 1) / N;
        const int slot = jb.cs & 1;
        const unsigned char* bop = smem + slot * kTcSlot + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;
        const uint32_t b_lo = ((smem_u32(bop) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
        if (!(DEV && (dbg & 8) && k >= 2)) mbar_wait_fast(a_full(k & 1), (k >> 1) & 1);
        mbar_wait_fast(slot_full(slot), (jb.cs >> 1) & 1);
        for (int t = 0; t < spj; t++, s_no++) {
          if (tr && s_no < 128) strace[2 * s_no] = (uint32_t)clock();
          if (s_no >= 2) mbar_wait_poll(acc_empty(s_no & 1), ((s_no >> 1) - 1) & 1);  // one thread: spinning is cheap
          tc_fence_after();
          umma_issue_idesc(tmem + (s_no & 1) * N, ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)(k & 1) * (kUmmaM * 32 / 16)),
                           ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)t * (N * 32 / 16)), kTcIdesc);
          umma_commit(acc_full(s_no & 1));
          if (tr && s_no < 128) strace[2 * s_no + 1] = (uint32_t)clock();
        }
        tc_job_next(jb, mt1, mt2);
      }
    }
  } else {
    // ---- helper: A operands two jobs ahead, refine behind the scan -----------------------------------------------
    // No global load sits on its path: the query coordinates of job k + 2 are requested before the wait for job k's
    // keys and consumed after it (A operand row; kept in registers for the job's own refine two jobs later).
    auto query_of = [&](const TcJob& jq, float& x, float& y, float& z, bool& ok) {
      const int nq = jq.rev ? a.m : a.n;
      const float* qpts = (jq.rev ? a.xyz2 : a.xyz1) + (size_t)jq.batch * nq * 3;
      const int qi = jq.ml * kUmmaM + row;
      ok = qi < nq;
      const int qs = ok ? qi : 0;
      x = __ldg(qpts + (size_t)qs * 3);
      y = __ldg(qpts + (size_t)qs * 3 + 1);
      z = __ldg(qpts + (size_t)qs * 3 + 2);
    };
    auto build_a = [&](int k, float x, float y, float z, bool ok) {  // A operand row of this thread's query of job k
      umma_a_row(aops + (size_t)(k & 1) * kUmmaM * 32, row, x, y, z, ok);
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(k & 1));
    };
    TcJob jb = first, jn = first;  // job k, job k + 2
    float qx, qy, qz, q1x = 0.f, q1y = 0.f, q1z = 0.f;  // queries of job k and job k + 1
    bool qok, q1ok = false;
    query_of(jn, qx, qy, qz, qok);
    tc_job_next(jn, mt1, mt2);
    if (nj > 1) query_of(jn, q1x, q1y, q1z, q1ok);
    tc_job_next(jn, mt1, mt2);
    build_a(0, qx, qy, qz, qok);
    if (nj > 1) build_a(1, q1x, q1y, q1z, q1ok);
    for (int k = 0; k < ((DEV && (dbg & 8)) ? 0 : nj); k++) {
      const int nq = jb.rev ? a.m : a.n, nt = jb.rev ? a.n : a.m;
      const int qi = jb.ml * kUmmaM + row;
      const bool valid = qok;
      float nx = 0.f, ny = 0.f, nz = 0.f;
      bool nok = false;
      if (k + kTcRing < nj) query_of(jn, nx, ny, nz, nok);  // in flight across the wait below
      mbar_wait_fast(keys_full(k & 1), (k >> 1) & 1);
      if (tr && tid == 512 && k < 64) strace[512 + 2 * k] = (uint32_t)clock();
      const float* src = keyring + (size_t)(k & 1) * 4 * 3 * kUmmaM + row;
      float c1 = src[0], c2 = src[kUmmaM], c3 = src[2 * kUmmaM];
#pragma unroll
      for (int gq = 1; gq < 4; gq++)
        merge3(c1, c2, c3, src[(3 * gq) * kUmmaM], src[(3 * gq + 1) * kUmmaM], src[(3 * gq + 2) * kUmmaM]);
      __syncwarp();
      if (lane == 0) mbar_arrive(keys_empty(k & 1));
      // job k's MMAs are complete: its A operand may be rebuilt for job k + 2
      if (k + kTcRing < nj) build_a(k + kTcRing, nx, ny, nz, nok);
      const int slot = jb.cs & 1;
      mbar_wait_fast(slot_full(slot), (jb.cs >> 1) & 1);  // long complete (the scan needed it first): acquire only
      const float4* tgt = reinterpret_cast<const float4*>(smem + slot * kTcSlot);
      const float bm = *reinterpret_cast<volatile float*>(red + 24 + slot);
      const float4 tg0 = tgt[0], tg1 = tgt[1];  // target 0 of the cloud: {x0,x1,y0,y1} {z0,z1,n0,n1}
      const float t0x = tg0.x, t0y = tg0.z, t0z = tg1.x;
      // ---- refine: this thread's query against the one or two tiles inside the window -------------------
      QueryState<1> qs;
      qs.valid[0] = valid;
      qs.qx[0] = qx;
      qs.qy[0] = qy;
      qs.qz[0] = qz;
      qs.qabs[0] = query_abs(qx, qy, qz);
      qs.ax2[0] = -2.0f * qx;
      qs.ay2[0] = -2.0f * qy;
      qs.az2[0] = -2.0f * qz;
      qs.d0[0] = sqdist<MODE>(t0x, t0y, t0z, qx, qy, qz);
      qs.best[0] = __int_as_float(0x7f800000);
      qs.besti[0] = 0;
      qs.m1g[0] = __int_as_float(0x7f800000);
      const int ntile = (nt + kUmmaT - 1) / kUmmaT;
      int cnt[1] = {0}, ta[1] = {0}, tb[1] = {0};
      float thr[1] = {0.0f};
      if (valid) {
        thr[0] = c1 + mma_window_wide(qs.qabs[0], bm);
        int c = !(c1 > thr[0]) ? 1 : 0;
        c += !(c2 > thr[0]) ? 1 : 0;
        if (!(c3 > thr[0])) c = 3;
        ta[0] = __float_as_int(c1) & 63;
        tb[0] = __float_as_int(c2) & 63;
        // a tile beyond the staged tiles can only come from padding / sentinels under a non-finite window
        if ((c >= 1 && ta[0] >= ntile) || (c >= 2 && tb[0] >= ntile)) c = 3;
        cnt[0] = c;
      }
      if (!(dbg & 1)) refine_tiles<MODE, 1>(qs, tgt, 0, nt, ntile, cnt, ta, tb, thr);
      if (valid) {
        float d;
        int bi;
        finish_query<1>(qs, 0, d, bi);
        float* odist = (jb.rev ? a.dist2 : a.dist1) + (size_t)jb.batch * nq;
        int* oidx = (jb.rev ? a.idx2 : a.idx1) + (size_t)jb.batch * nq;
        odist[qi] = d;
        oidx[qi] = bi;
        float* mdist = jb.rev ? a.mdist2 : a.mdist1;
        if (mdist != nullptr) {
          int* midx = jb.rev ? a.midx2 : a.midx1;
          mdist[(size_t)jb.batch * nq + qi] = d;
          midx[(size_t)jb.batch * nq + qi] = bi;
        }
      }
      // this job no longer needs its cloud's slot
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (row == 0) atomicAdd(const_cast<int*>(done_cnt) + slot, 1);
      if (tr && tid == 512 && k < 64) strace[512 + 2 * k + 1] = (uint32_t)clock();
      qx = q1x; qy = q1y; qz = q1z; qok = q1ok;
      q1x = nx; q1y = ny; q1z = nz; q1ok = nok;
      tc_job_next(jb, mt1, mt2);
      tc_job_next(jn, mt1, mt2);
    }
  }
  umma_tmem_free(tmem, warp);
  if (trace != nullptr && tid == 0) trace[4096 + 4 * blockIdx.x + 3] = (long long)global_ns();
  if (tr) {
    __syncthreads();
    for (int i = tid; i < kTcTrace; i += kTcThreads) trace[i] = strace[i];
  }
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

int g_umma_grid = 0;    // tuning hook (key 12): CTAs of the tcgen05 kernel (0 = one per SM)
int g_umma_groups = 0;  // tuning hook (key 20): timing experiments (bit 0: helper skips the refine, bit 1: scan skips the drain)

bool fwd_umma_supported(int n, int m) { return n > 256 && m > 256 && n <= kUmmaCH && m <= kUmmaCH; }

static int launch_fwd_umma_impl(const FwdArgs& a, int mode, cudaStream_t st, long long* trace) {
  if (!fwd_umma_supported(a.n, a.m)) {
    set_error("nn_fwd_umma_kernel: clouds of %d..%d points", 257, kUmmaCH);
    return GA_ERR_UNSUPPORTED;
  }
  const int mt1 = (a.n + kUmmaM - 1) / kUmmaM, mt2 = (a.m + kUmmaM - 1) / kUmmaM;
  const long long J = (long long)a.b * (mt1 + mt2);
  if (J <= 0) return GA_OK;
  const bool dev = trace != nullptr || g_umma_groups != 0;
  auto k = mode == GA_MODE_CPU_EXACT ? (dev ? nn_fwd_umma_kernel<GA_MODE_CPU_EXACT, true> : nn_fwd_umma_kernel<GA_MODE_CPU_EXACT, false>)
                                     : (dev ? nn_fwd_umma_kernel<GA_MODE_GPU_REF, true> : nn_fwd_umma_kernel<GA_MODE_GPU_REF, false>);
  {
    static std::atomic<unsigned> done_mask[4];
    int device = 0;
    GA_CUDA_TRY(cudaGetDevice(&device));
    if (!(done_mask[mode * 2 + dev].load(std::memory_order_relaxed) & (1u << (device & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTcSmem + kTcTrace * 4)));
      done_mask[mode * 2 + dev].fetch_or(1u << (device & 31), std::memory_order_relaxed);
    }
  }
  long long grid = g_umma_grid > 0 ? g_umma_grid : sm_count();
  if (grid > J) grid = J;
  k<<<(unsigned)grid, kTcThreads, kTcSmem + (trace != nullptr ? kTcTrace * 4 : 0), st>>>(a, mt1, mt2, J, g_umma_groups, trace);
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


// Development: one traced run of the tcgen05 kernel (mode 0); `trace` = 8192 int64 on the device, zeroed by the caller.
extern "C" int ga_debug_umma_trace(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                                   float* dist2, int* idx2, long long* trace, ga_stream_t stream) {
  ga::FwdArgs a;
  a.b = b; a.n = n; a.m = m;
  a.xyz1 = xyz1; a.xyz2 = xyz2;
  a.dist1 = dist1; a.idx1 = idx1; a.dist2 = dist2; a.idx2 = idx2;
  a.tiles1 = a.tiles2 = 0;
  a.mdist1 = a.mdist2 = nullptr;
  a.midx1 = a.midx2 = nullptr;
  a.ticket = nullptr;
  a.call_id = 0;
  a.ticket_debug = 0;
  return ga::launch_fwd_umma_impl(a, GA_MODE_CPU_EXACT, ga::as_stream(stream), trace);
}
