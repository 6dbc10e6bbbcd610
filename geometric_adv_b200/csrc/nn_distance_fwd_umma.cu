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

// ---- the kernel ---------------------------------------------------------------------------------
// Persistent, one CTA of 16 warps per SM, warp-specialised:
//   scan    (warps 0-7)   the only warps that touch TMEM.  All eight work on the same 128-query job: warps w and
//                         w + 4 share a TMEM lane quadrant and each drains one half of the columns of a step's
//                         accumulator (128 queries x 256 targets), so a THREAD owns one query row and half of
//                         the targets, and a query's three smallest tile keys of that half live in three
//                         registers for the whole job.  Two accumulators: at the top of step s the leader
//                         issues the tcgen05.mma of step s + 1 (260 clk issue to visible, tools/mmalat.cu) into
//                         the buffer handed back during step s - 1 (acc_empty, one arrival per warp); every
//                         warp waits for step s (acc_full), 4 x (tcgen05.ld.32x32b.x32 -> 16 FMNMX3 -> key,
//                         5 FMNMX).  Two scan warps per scheduler keep its ALU pipe and its TMEM read path busy
//                         (tools/drainbench.cu: ~52 clk per tile and scheduler, load and fold do not overlap).
//                         After a job's last step each thread writes its keys to a 4-deep ring.
//   helper  (warps 8-11)  builds the A operand of a job four jobs ahead and, once both column halves' keys of a
//                         job have arrived, merges them and refines (refine_tiles, one query per thread).
//   stager  (warps 12-15) stages the target cloud of the next (batch, direction) into the free one of two
//                         slots (pair-SoA for the refine + B operand) while the other one is in use.
// Roles meet only through mbarriers (a_full, keys_full, keys_empty, slot_full, acc_full, acc_empty) and one
// counter per slot (refined jobs): the scan warps never wait for a global load, a refine or a staging pass,
// and the two halves of the scan are balanced by construction (same columns of every step).
// A CTA owns a contiguous range of whole jobs (J / grid: 10.8 at BASELINE config 2, 1.8 % quantisation).
constexpr int kTcThreads = 512;
constexpr int kTcN = 256;                    // targets per MMA = columns of one accumulator; two accumulators
constexpr int kTcRing = 4;                   // A operands / key sets in flight
constexpr size_t kTcSlot = (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32 + (size_t)kUmmaCH * 32;  // one staged cloud
constexpr size_t kTcOffA = 2 * kTcSlot;                                        // [4] A operands, 128 rows x 32 B
constexpr size_t kTcOffKeys = kTcOffA + (size_t)kTcRing * kUmmaM * 32;         // [4][2][3][128] float
constexpr size_t kTcOffRed = kTcOffKeys + (size_t)kTcRing * 2 * 3 * kUmmaM * 4;  // stager reduction [4] + bm[2]
constexpr size_t kTcOffBar = kTcOffRed + 32;  // a_full[4] keys_full[4] keys_empty[4] slot_full[2] acc_full[2] acc_empty[2] | cnt[2] | tmem
constexpr size_t kTcSmem = kTcOffBar + 18 * 8 + 8 + 8;
constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) |
                              ((uint32_t)(kUmmaM >> 4) << 24);
static_assert(kTcSlot % 128 == 0 && kTcOffA % 128 == 0 && kTcOffBar % 8 == 0, "operand alignment");
static_assert(kTcSmem <= 232448, "shared memory of one SM");

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
__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Wait for the phase with parity `parity`: a non-blocking test first (the usual case: long complete), then
// the suspending try_wait.  Bounded: a completion that never comes traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
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
  if (!done) mbar_wait(bar, parity);
}

// Poll with test_wait and a short sleep: for the waits on the scan's critical path (the suspending try_wait wakes a
// thread up late; a tight test_wait loop would take the issue slots of the group that is draining).
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 22); spin++) {
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
    if (done) return;
    __nanosleep(20);
  }
  __trap();
}

// Drain one accumulator (NT tiles of 32 columns) into the key triple.  One 32-register buffer: load, wait, fold.
// tools/drainbench.cu: on a scheduler a tile costs its TMEM load (22-30 clk of register-file writes) PLUS its fold
// (24 clk of ALU), whatever the warp count -- the two do not overlap -- and software double-buffering only adds
// 32 register moves per tile (72.8 vs 51.8 clk per tile and scheduler at two warps per scheduler).
// `handback` runs as soon as the last load has landed.
template <int NT, class F>
__device__ __forceinline__ void tc_drain(uint32_t tmem_row, int tile0, float& c1, float& c2, float& c3, F&& handback) {
#pragma unroll
  for (int e = 0; e < NT; e++) {
    float v[32];
    tmem_ld32(tmem_row + 32 * e, v);
    tmem_ld_wait();
    if (e + 1 == NT) handback();
    umma_fold(v, tile0 + e, c1, c2, c3);
  }
}

struct TcJob {  // one 128-query job
  int batch, rev, ml, cs;  // batch element, direction, M-tile within its cloud, cloud sequence number within the CTA
};

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) nn_fwd_umma_kernel(const FwdArgs a, const int mt1, const int mt2,
                                                                   const long long J) {
  constexpr int N = kTcN, NT = kTcN / 32;
  asm volatile("griddepcontrol.launch_dependents;");
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* aops = smem + kTcOffA;
  float* keyring = reinterpret_cast<float*>(smem + kTcOffKeys);
  float* red = reinterpret_cast<float*>(smem + kTcOffRed);  // [0..3] stager warps, [4..5] bm of the slots
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + kTcOffBar);
  volatile int* done_cnt = reinterpret_cast<volatile int*>(smem + kTcOffBar + 18 * 8);  // [2] refined jobs per slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kTcOffBar + 18 * 8 + 8);
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto keys_full = [&](int i) { return bar0 + 8u * (4 + i); };
  auto keys_empty = [&](int i) { return bar0 + 8u * (8 + i); };
  auto slot_full = [&](int i) { return bar0 + 8u * (12 + i); };
  auto acc_full = [&](int i) { return bar0 + 8u * (14 + i); };
  auto acc_empty = [&](int i) { return bar0 + 8u * (16 + i); };

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int role = warp >> 2;                // 0 / 1 scan groups, 2 helper, 3 stager
  const int row = (warp & 3) * 32 + lane;    // thread within its group = query row of a job
  if (tid == 0) {
    for (int i = 0; i < 4; i++) {
      mbar_init(a_full(i), 128);
      mbar_init(keys_full(i), 256);
      mbar_init(keys_empty(i), 128);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(slot_full(i), 128);
      mbar_init(acc_full(i), 1);
      mbar_init(acc_empty(i), 8);  // one arrival per scan warp
    }
    done_cnt[0] = 0;
    done_cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = umma_tmem_alloc(tmem_slot, warp);  // contains a __syncthreads

  const long long jpb = (long long)mt1 + mt2;
  const long long j0 = J * blockIdx.x / gridDim.x, j1 = J * (blockIdx.x + 1) / gridDim.x;
  const int nj = (int)(j1 - j0);
  const int cid0 = (int)(j0 / jpb) * 2 + ((j0 % jpb) >= mt1 ? 1 : 0);  // first cloud (batch * 2 + direction)
  auto job = [&](int k) {
    TcJob jb;
    const long long jg = j0 + k;
    jb.batch = (int)(jg / jpb);
    const int r = (int)(jg - (long long)jb.batch * jpb);
    jb.rev = r >= mt1;
    jb.ml = jb.rev ? r - mt1 : r;
    jb.cs = jb.batch * 2 + jb.rev - cid0;
    return jb;
  };

  if (role == 3) {
    // ---- stager: cloud cs -> slot cs & 1, as soon as every job of cloud cs - 2 has been refined ----------
    const TcJob last = job(nj - 1);
    for (int cs = 0; cs <= last.cs; cs++) {
      const int cid = cid0 + cs, batch = cid >> 1, rev = cid & 1, slot = cs & 1;
      const int nt = rev ? a.n : a.m;
      if (cs >= 2) {
        // jobs of cloud cs - 2 inside this CTA's range
        const int pc = cid - 2, pb = pc >> 1, pr = pc & 1;
        const long long cb = (long long)pb * jpb + (pr ? mt1 : 0), ce = cb + (pr ? mt2 : mt1);
        const int want = (int)((ce < j1 ? ce : j1) - (cb > j0 ? cb : j0));
        if (row == 0) {
          while (done_cnt[slot] != want) __nanosleep(100);
          done_cnt[slot] = 0;
        }
        group_barrier(4);
      }
      const float* tpts = (rev ? a.xyz1 : a.xyz2) + (size_t)batch * nt * 3;
      float4* tgt = reinterpret_cast<float4*>(smem + slot * kTcSlot);
      unsigned char* bop = smem + slot * kTcSlot + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;
      const float kInf = __int_as_float(0x7f800000);
      const int spj = (nt + N - 1) / N;
      const int npair = spj * (N / 2);
      float lmax = 0.0f;
      for (int p0 = row; p0 < npair; p0 += 128 * 4) {  // four pairs per thread in flight
        float c[4][6];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int p = p0 + u * 128;
#pragma unroll
          for (int e = 0; e < 6; e++) c[u][e] = (p < npair && 2 * p + (e >= 3) < nt) ? __ldg(tpts + (size_t)p * 6 + e) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int p = p0 + u * 128;
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
      lmax = warp_max(lmax);
      if (lane == 0) red[warp & 3] = lmax;
      group_barrier(4);  // pair-SoA complete: the B rows are built from it
      if (row == 0) red[4 + slot] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
      for (int tau = row; tau < spj * N; tau += 128) umma_b_row(bop, tgt, tau, nt);
      proxy_fence();
      mbar_arrive(slot_full(slot));
    }
  } else if (role < 2) {
    // ---- scan: warps w and w + 4 share a TMEM lane quadrant; group g = warp >> 2 drains column half g (four of the
    // eight 32-target tiles) of every step's accumulator.  Two 256-column accumulators: the MMA of step s + 1 is
    // issued at the top of step s into the buffer that was handed back during step s - 1.
    const int g = role;
    const bool leader = tid == 0;
    const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * (N / 2));
    const uint32_t desc_hi = (uint32_t)(256 >> 4) | (1u << 14);  // SBO, descriptor version (bits 32.., 46)
    const uint32_t a_lo = ((smem_u32(aops) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
    auto issue_step = [&](unsigned s_no, int slot, int ab, int t) {  // leader only; step number s_no -> buffer s_no & 1
      const unsigned char* bop = smem + slot * kTcSlot + (size_t)kUmmaCH * 16 + (size_t)kPipeU * 32;
      const uint32_t b_lo = ((smem_u32(bop) & 0x3ffffu) >> 4) | ((uint32_t)(128 >> 4) << 16);
      // the buffer's previous contents (step s_no - 2) have been read by all eight warps
      if (s_no >= 2) mbar_wait_fast(acc_empty(s_no & 1), ((s_no >> 1) - 1) & 1);
      tc_fence_after();
      umma_issue_idesc(tmem + (s_no & 1) * N, ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)ab * (kUmmaM * 32 / 16)),
                       ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)t * (N * 32 / 16)), kTcIdesc);
      umma_commit(acc_full(s_no & 1));
    };
    auto job_ready = [&](int k, const TcJob& jb) {  // what the first MMA of job k needs: its A operand and its cloud
      mbar_wait_fast(a_full(k & 3), (k >> 2) & 1);
      mbar_wait_fast(slot_full(jb.cs & 1), (jb.cs >> 1) & 1);
    };
    unsigned ns = 0;  // steps drained so far: step s uses accumulator s & 1, barrier parity (s >> 1) & 1
    TcJob cur = job(0);
    if (leader) {
      job_ready(0, cur);
      issue_step(0, cur.cs & 1, 0, 0);
    }
    for (int k = 0; k < nj; k++) {
      const int nt = cur.rev ? a.n : a.m;
      const int spj = (nt + N - 1) / N;
      const bool more = k + 1 < nj;
      TcJob nxt = cur;
      if (more) nxt = job(k + 1);
      float c1 = kMmaBig, c2 = kMmaBig, c3 = kMmaBig;
      for (int t = 0; t < spj; t++) {
        if (leader) {  // one step ahead
          if (t + 1 < spj) {
            issue_step(ns + 1, cur.cs & 1, k & 3, t + 1);
          } else if (more) {
            job_ready(k + 1, nxt);
            issue_step(ns + 1, nxt.cs & 1, (k + 1) & 3, 0);
          }
        }
        mbar_wait_fast(acc_full(ns & 1), (ns >> 1) & 1);
        tc_fence_after();
        tc_drain<NT / 2>(tmem_row + (ns & 1) * N, t * NT + g * (NT / 2), c1, c2, c3, [&] {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(ns & 1));
        });
        ns++;
      }
      // this column half's keys of job k -> ring slot k & 3 (free once the helper has read job k - 4's)
      if (k >= kTcRing) mbar_wait_fast(keys_empty(k & 3), ((k >> 2) - 1) & 1);
      float* dst = keyring + ((size_t)(k & 3) * 2 + g) * 3 * kUmmaM + row;
      dst[0] = c1;
      dst[kUmmaM] = c2;
      dst[2 * kUmmaM] = c3;
      mbar_arrive(keys_full(k & 3));
      cur = nxt;
    }
  } else {
    // ---- helper: A operands four jobs ahead, refine behind the scan -------------------------------------------
    auto build_a = [&](int k) {  // A operand row of this thread's query of job k
      const TcJob jb = job(k);
      const int nq = jb.rev ? a.m : a.n;
      const float* qpts = (jb.rev ? a.xyz2 : a.xyz1) + (size_t)jb.batch * nq * 3;
      const int qi = jb.ml * kUmmaM + row;
      const bool ok = qi < nq;
      const int qs = ok ? qi : 0;
      umma_a_row(aops + (size_t)(k & 3) * kUmmaM * 32, row, __ldg(qpts + (size_t)qs * 3), __ldg(qpts + (size_t)qs * 3 + 1),
                 __ldg(qpts + (size_t)qs * 3 + 2), ok);
      proxy_fence();
      mbar_arrive(a_full(k & 3));
    };
    for (int k = 0; k < kTcRing && k < nj; k++) build_a(k);
    for (int k = 0; k < nj; k++) {
      const TcJob jb = job(k);
      const int nq = jb.rev ? a.m : a.n, nt = jb.rev ? a.n : a.m;
      const float* qpts = (jb.rev ? a.xyz2 : a.xyz1) + (size_t)jb.batch * nq * 3;
      const float* tpts = (jb.rev ? a.xyz1 : a.xyz2) + (size_t)jb.batch * nt * 3;
      const int qi = jb.ml * kUmmaM + row;
      const bool valid = qi < nq;
      const int qsafe = valid ? qi : 0;
      const float qx = __ldg(qpts + (size_t)qsafe * 3), qy = __ldg(qpts + (size_t)qsafe * 3 + 1),
                  qz = __ldg(qpts + (size_t)qsafe * 3 + 2);
      const float t0x = __ldg(tpts), t0y = __ldg(tpts + 1), t0z = __ldg(tpts + 2);
      mbar_wait_fast(keys_full(k & 3), (k >> 2) & 1);
      const float* src = keyring + (size_t)(k & 3) * 2 * 3 * kUmmaM + row;
      float c1 = src[0], c2 = src[kUmmaM], c3 = src[2 * kUmmaM];
      merge3(c1, c2, c3, src[3 * kUmmaM], src[4 * kUmmaM], src[5 * kUmmaM]);
      mbar_arrive(keys_empty(k & 3));
      // job k's MMAs are complete: its A operand may be rebuilt for job k + 4
      if (k + kTcRing < nj) build_a(k + kTcRing);
      // ---- refine: this thread's query against the one or two tiles inside the window -------------------
      const int slot = jb.cs & 1;
      mbar_wait_fast(slot_full(slot), (jb.cs >> 1) & 1);  // long complete (the scan needed it first): acquire only
      const float4* tgt = reinterpret_cast<const float4*>(smem + slot * kTcSlot);
      const float bm = *reinterpret_cast<volatile float*>(red + 4 + slot);
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
      refine_tiles<MODE, 1>(qs, tgt, 0, nt, ntile, cnt, ta, tb, thr);
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
      group_barrier(2);
      if (row == 0) atomicAdd(const_cast<int*>(done_cnt) + slot, 1);
    }
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

int g_umma_grid = 0;    // tuning hook (key 12): CTAs of the tcgen05 kernel (0 = one per SM)
int g_umma_groups = 0;  // tuning hook (key 20): warp groups per CTA (0 = auto, 2 or 4)

bool fwd_umma_supported(int n, int m) { return n > 256 && m > 256 && n <= kUmmaCH && m <= kUmmaCH; }

int launch_fwd_umma(const FwdArgs& a, int mode, cudaStream_t st) {
  if (!fwd_umma_supported(a.n, a.m)) {
    set_error("nn_fwd_umma_kernel: clouds of %d..%d points", 257, kUmmaCH);
    return GA_ERR_UNSUPPORTED;
  }
  const int mt1 = (a.n + kUmmaM - 1) / kUmmaM, mt2 = (a.m + kUmmaM - 1) / kUmmaM;
  const long long J = (long long)a.b * (mt1 + mt2);
  if (J <= 0) return GA_OK;
  auto k = mode == GA_MODE_CPU_EXACT ? nn_fwd_umma_kernel<GA_MODE_CPU_EXACT> : nn_fwd_umma_kernel<GA_MODE_GPU_REF>;
  {
    static std::atomic<unsigned> done_mask[2];
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    if (!(done_mask[mode].load(std::memory_order_relaxed) & (1u << (dev & 31)))) {
      GA_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
      done_mask[mode].fetch_or(1u << (dev & 31), std::memory_order_relaxed);
    }
  }
  long long grid = g_umma_grid > 0 ? g_umma_grid : sm_count();
  if (grid > J) grid = J;
  k<<<(unsigned)grid, kTcThreads, kTcSmem, st>>>(a, mt1, mt2, J);
  GA_LAUNCH_CHECK("nn_fwd_umma_kernel");
  return GA_OK;
}

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

