// HOST-pointer entry points: the drop-in for the reference's CPU-tensor kernels
// (REGISTER_KERNEL_BUILDER(... DEVICE_CPU ...), tf_nndistance.cpp:83,166).  The
// caller's buffers are copied to a per-thread device arena, the device entry
// points run on a per-thread stream, results are copied straight back into the
// caller's memory.  Pinned caller memory gives true async DMA; pageable memory
// works too (the driver stages it).
#include "ga_common.cuh"

namespace ga {

struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  int dev = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // second lane for chunked copy/compute overlap
  // captured pipeline (ga_nn_distance_fwd_bwd_host): copy-out lane, one kernel lane per chunk, fork/join events
  cudaStream_t out_lane = nullptr;
  cudaStream_t k_lane[8] = {};
  cudaEvent_t ev[48] = {};
  unsigned long long generation = 0;  // bumped when `base` moves: cached graphs hold device addresses
  // streamed ingest (issue_pipeline_streamed): per-group arrival flags on the device, the pinned word the flag
  // copies read, and the mapped word a CTA raises when it gave up waiting
  int* d_ready = nullptr;
  int* h_one = nullptr;
  int* h_abort = nullptr;
  int* d_abort = nullptr;
  // deliberately no destructor: at process teardown the CUDA context may already be gone
};
static thread_local Arena t_arena;

static int arena_reserve(size_t bytes, char** base, cudaStream_t* st) {
  int dev = 0;
  GA_CUDA_TRY(cudaGetDevice(&dev));
  Arena& A = t_arena;
  if (A.dev != dev) {
    if (A.base) cudaFree(A.base);
    if (A.stream) cudaStreamDestroy(A.stream);
    if (A.stream2) cudaStreamDestroy(A.stream2);
    if (A.out_lane) cudaStreamDestroy(A.out_lane);
    A.out_lane = nullptr;
    for (auto& k : A.k_lane) {
      if (k) cudaStreamDestroy(k);
      k = nullptr;
    }
    for (auto& e : A.ev) {
      if (e) cudaEventDestroy(e);
      e = nullptr;
    }
    if (A.d_ready) cudaFree(A.d_ready);
    if (A.h_one) cudaFreeHost(A.h_one);
    if (A.h_abort) cudaFreeHost(A.h_abort);
    A.d_ready = nullptr;
    A.h_one = A.h_abort = A.d_abort = nullptr;
    A.generation++;
    A.base = nullptr;
    A.cap = 0;
    A.stream = nullptr;
    A.stream2 = nullptr;
    A.dev = dev;
  }
  if (!A.stream) GA_CUDA_TRY(cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking));
  if (!A.stream2) GA_CUDA_TRY(cudaStreamCreateWithFlags(&A.stream2, cudaStreamNonBlocking));
  if (bytes > A.cap) {
    if (A.base) {
      GA_CUDA_TRY(cudaStreamSynchronize(A.stream));
      GA_CUDA_TRY(cudaStreamSynchronize(A.stream2));
      GA_CUDA_TRY(cudaFree(A.base));
      A.base = nullptr;
      A.cap = 0;
    }
    size_t want = bytes + (bytes >> 2) + (1u << 20);
    GA_CUDA_TRY(cudaMalloc(&A.base, want));
    A.cap = want;
    A.generation++;
  }
  *base = A.base;
  *st = A.stream;
  return GA_OK;
}

struct Carver {
  char* p;
  size_t off = 0;
  explicit Carver(char* base) : p(base) {}
  template <class T>
  T* take(size_t count) {
    T* r = reinterpret_cast<T*>(p + off);
    off += (count * sizeof(T) + 255) & ~(size_t)255;
    return r;
  }
};
static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

int nn_distance_fwd_mirrored(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                             float* dist2, int* idx2, float* mdist1, int* midx1, float* mdist2, int* midx2,
                             int mode, ga_stream_t stream);  // nn_distance_fwd.cu

// ---- experimental zero-copy path for pinned caller memory ---------------------------------
// With pinned (page-locked, UVA-mapped) host buffers the GPU can read inputs and write results
// itself: ONE ingest kernel pulls all inputs over PCIe, the forward kernel mirrors dist/idx to
// the host while it computes, and the backward kernel writes the gradients straight to the
// host.  It replaces 4 + 6 cudaMemcpyAsync calls, but SM-driven PCIe traffic turned out slower
// than the copy engines on this platform (numbers at the call site); kept behind a tuning key.
struct IngestArgs {
  const void* src[4];
  void* dst[4];
  size_t bytes[4];
  int nseg;
};

__global__ void __launch_bounds__(256) ingest_kernel(const IngestArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  for (int sgm = 0; sgm < a.nseg; sgm++) {
    const size_t bytes = a.bytes[sgm];
    const bool al16 = (((uintptr_t)a.src[sgm] | (uintptr_t)a.dst[sgm]) & 15) == 0;
    if (al16) {
      const int4* src = reinterpret_cast<const int4*>(a.src[sgm]);
      int4* dst = reinterpret_cast<int4*>(a.dst[sgm]);
      const size_t nv = bytes / 16;
      size_t i = tid;
      for (; i + 3 * nthreads < nv; i += 4 * nthreads) {  // four PCIe reads in flight per thread
        const int4 v0 = src[i], v1 = src[i + nthreads], v2 = src[i + 2 * nthreads], v3 = src[i + 3 * nthreads];
        dst[i] = v0;
        dst[i + nthreads] = v1;
        dst[i + 2 * nthreads] = v2;
        dst[i + 3 * nthreads] = v3;
      }
      for (; i < nv; i += nthreads) dst[i] = src[i];
      const int* s4 = reinterpret_cast<const int*>(a.src[sgm]);
      int* d4 = reinterpret_cast<int*>(a.dst[sgm]);
      for (size_t w = nv * 4 + tid; w < bytes / 4; w += nthreads) d4[w] = s4[w];
    } else {
      const int* s4 = reinterpret_cast<const int*>(a.src[sgm]);
      int* d4 = reinterpret_cast<int*>(a.dst[sgm]);
      for (size_t w = tid; w < bytes / 4; w += nthreads) d4[w] = s4[w];
    }
  }
}

// Streamed ingest by the SMs (issue_pipeline_pulled): a few CTAs pull the clouds of one batch element after the other
// out of the caller's pinned buffers (16-byte loads over PCIe, all of an element's loads of a thread in flight at
// once) and raise that element's arrival flag; the forward kernel, launched behind this one as a programmatic
// dependent (it starts as soon as every CTA here has passed the trigger below, and never waits for the grid), has
// its CTAs wait per batch element (wait_ready, nn_search.cuh).  No copy node, no per-group overhead: the search of
// element e runs under the transfer of the elements behind it.  v1 / v2: 16-byte vectors per cloud of a batch element.
struct PullArgs {
  const int4* src1;
  const int4* src2;
  int4* dst1;
  int4* dst2;
  int* ready;
  int b, v1, v2;
  int mute;  // test hook (key 29): the flag of the last batch element is never raised
};

__global__ void __launch_bounds__(256) ingest_stream_kernel(const PullArgs a) {
  asm volatile("griddepcontrol.launch_dependents;");
  constexpr int U = 6;  // vectors in flight per thread and round (256 threads x 6 x 16 B = one 2048-point cloud)
  for (int e = blockIdx.x; e < a.b; e += gridDim.x) {
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
      const int nv = c ? a.v2 : a.v1;
      const int4* src = (c ? a.src2 : a.src1) + (size_t)e * nv;
      int4* dst = (c ? a.dst2 : a.dst1) + (size_t)e * nv;
      for (int i0 = threadIdx.x; i0 < nv; i0 += 256 * U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
          if (i0 + u * 256 < nv) v[u] = src[i0 + u * 256];
#pragma unroll
        for (int u = 0; u < U; u++)
          if (i0 + u * 256 < nv) dst[i0 + u * 256] = v[u];
      }
    }
    __syncthreads();  // every thread's stores are issued ...
    if (threadIdx.x == 0 && !(a.mute && e == a.b - 1)) {
      __threadfence();  // ... and ordered before the flag (cumulative over the barrier)
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.ready + e), "r"(1) : "memory");
    }
  }
}

int g_host_chunks = 0;  // tuning hook (key 3): force the chunk count of the copy path
int g_host_path = 0;  // tuning hook (ga_set_tuning key 2): 0 auto, 1 force copies, 2 force zero-copy

// true if `p` (a host pointer handed to a *_host entry point) can be dereferenced by the device
static bool device_can_touch(const void* p) {
  if (p == nullptr) return false;
  if (g_host_path == 1) return false;
  if (g_host_path == 2) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost && at.devicePointer == p;
}

#define GA_TRY(expr)             \
  do {                           \
    int _rc = (expr);            \
    if (_rc != GA_OK) return _rc; \
  } while (0)


// ---- graph replay of the host fwd+bwd step -------------------------------------------------
// The direct path below costs ~24 driver calls per step (10 copies and 2 kernels per chunk), about
// as long on the CPU as the whole step takes on the GPU.  A training loop hands in the SAME pinned
// buffers every step (src/adv_ae.py feeds fixed placeholders), so the second time a (shape,
// pointers) key is seen the step is captured once as a CUDA graph and replayed from then on with
// one cudaGraphLaunch.  The graph is laid out as a three-stage pipeline over chunks of the batch: H2D copies chained on one lane (chunk i's inputs arrive
// before chunk i+1's start), each chunk's kernels on its own lane (free to overlap the tail of
// the previous chunk), D2H copies chained on a third lane: dist/idx right after the forward,
// gradients after the backward.  Pageable buffers never take this path.
int g_host_graph = 0;         // tuning hook (key 10): 0 auto, 1 never replay, 2 capture on first sight
int g_host_graph_chunks = 0;  // tuning hook (key 11): chunks of the captured pipeline (0 = auto)
int g_host_graph_epoch = 0;   // bumped by ga_set_tuning(10 | 11 | 17): cached graphs of older epochs are dropped
int g_host_pull_mute = 0;     // test hook (key 29): 1 = the ingest kernel withholds one arrival flag (exercises the give-up path)
int g_host_pull = 0;          // tuning hook (key 27): clouds pulled by the SMs behind arrival flags: 0 auto, -1 off, n = ingest CTAs
int g_host_stream = 0;        // tuning hook (key 26): streamed ingest of the replayed step: 0 off, n = arrival groups
int g_host_graph_mirror = 0;  // tuning hook (key 17): 0 auto, 1 = always copy dist/idx, 2 = always let the forward
                              // kernel write them straight to the pinned host buffers inside the replayed graph

struct HostGraph {
  cudaGraphExec_t exec = nullptr;
  int dev = -1, b = 0, n = 0, m = 0, mode = 0, nchunk = 0, launches = 0, seen = 0;
  bool failed = false;
  bool streamed = false;   // exec is the streamed pipeline
  bool no_stream = false;  // a streamed replay gave up waiting once: this key stays on the chunked pipeline
  const void* ptr[10] = {};
  unsigned long long generation = 0, stamp = 0;
  int epoch = 0;
};
static thread_local HostGraph t_graphs[4];
static thread_local int t_last_streamed = 0;  // ga_debug_host_streamed
static thread_local unsigned long long t_graph_clock = 0;
long long launch_count_now();  // core.cu

static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

struct FwdBwdBufs {
  const float *xyz1, *xyz2, *gd1, *gd2;             // host in
  float *dist1, *dist2, *gx1, *gx2;                 // host out
  int *idx1, *idx2;
  float *d_x1, *d_x2, *d_g1, *d_g2, *d_d1, *d_d2, *d_o1, *d_o2;  // arena
  int *d_i1, *d_i2;
  bool mirror;  // dist/idx host buffers are device-accessible: the forward kernel mirrors its outputs there
};

constexpr int kMaxReadyGroups = 32;
constexpr int kMaxReadyElems = 16384;  // arrival flags of the pulled pipeline: one per batch element

// Buffers of the streamed pipeline; called outside any capture.
static int stream_buffers() {
  Arena& A = t_arena;
  if (!A.d_ready) GA_CUDA_TRY(cudaMalloc(&A.d_ready, sizeof(int) * kMaxReadyElems));
  if (!A.h_one) {
    GA_CUDA_TRY(cudaHostAlloc(&A.h_one, sizeof(int), cudaHostAllocDefault));
    *A.h_one = 1;
  }
  if (!A.h_abort) {
    GA_CUDA_TRY(cudaHostAlloc(&A.h_abort, sizeof(int), cudaHostAllocMapped));
    *A.h_abort = 0;
    GA_CUDA_TRY(cudaHostGetDevicePointer(&A.d_abort, A.h_abort, 0));
  }
  return GA_OK;
}

// The step with streamed ingest (wait_ready, nn_search.cuh): ONE forward launch for the whole batch, started at
// the same time as the first H2D copy.  The clouds arrive in `groups` groups of batch elements, each followed by a
// 4-byte copy that raises the group's flag; the CTAs of a batch element wait for it.  The search of group g runs
// under the copy of group g+1, and no launch is split (a split launch pays the kernel's wave steps per piece:
// two halves of B=50 cost 2 x 41 us instead of 56).  The upstream gradients follow the clouds on the same lane;
// the rest is issue_pipeline with one chunk.
static int issue_pipeline_streamed(const FwdBwdBufs& f, int b, int n, int m, int mode, int groups) {
  Arena& A = t_arena;
  cudaStream_t sin = A.stream, sout = A.out_lane, sk = A.k_lane[0];
  cudaEvent_t fork = A.ev[0], join = A.ev[1], ev_g = A.ev[10], ev_f = A.ev[18], ev_b = A.ev[26];
  const int per = (b + groups - 1) / groups;
  groups = (b + per - 1) / per;
  GA_CUDA_TRY(cudaMemsetAsync(A.d_ready, 0, sizeof(int) * groups, sin));
  GA_CUDA_TRY(cudaEventRecord(fork, sin));
  GA_CUDA_TRY(cudaStreamWaitEvent(sout, fork, 0));
  GA_CUDA_TRY(cudaStreamWaitEvent(sk, fork, 0));
  const size_t e1 = (size_t)b * n, e2 = (size_t)b * m;

  t_ready_arm.flags = A.d_ready;
  t_ready_arm.per = per;
  t_ready_arm.abort_word = A.d_abort;
  const int rc = f.mirror ? nn_distance_fwd_mirrored(b, n, m, f.d_x1, f.d_x2, f.d_d1, f.d_i1, f.d_d2, f.d_i2, f.dist1,
                                                     f.idx1, f.dist2, f.idx2, mode, (ga_stream_t)sk)
                          : ga_nn_distance_fwd(b, n, m, f.d_x1, f.d_x2, f.d_d1, f.d_i1, f.d_d2, f.d_i2, mode,
                                               (ga_stream_t)sk);
  t_ready_arm.flags = nullptr;
  GA_TRY(rc);
  GA_CUDA_TRY(cudaEventRecord(ev_f, sk));

  for (int g = 0; g < groups; g++) {
    const int b0 = g * per, b1 = b0 + per < b ? b0 + per : b;
    const size_t o1 = (size_t)b0 * n, o2 = (size_t)b0 * m, c1 = (size_t)(b1 - b0) * n, c2 = (size_t)(b1 - b0) * m;
    GA_CUDA_TRY(cudaMemcpyAsync(f.d_x1 + o1 * 3, f.xyz1 + o1 * 3, c1 * 12, cudaMemcpyHostToDevice, sin));
    GA_CUDA_TRY(cudaMemcpyAsync(f.d_x2 + o2 * 3, f.xyz2 + o2 * 3, c2 * 12, cudaMemcpyHostToDevice, sin));
    GA_CUDA_TRY(cudaMemcpyAsync(A.d_ready + g, A.h_one, sizeof(int), cudaMemcpyHostToDevice, sin));
  }
  GA_CUDA_TRY(cudaMemcpyAsync(f.d_g1, f.gd1, e1 * 4, cudaMemcpyHostToDevice, sin));
  GA_CUDA_TRY(cudaMemcpyAsync(f.d_g2, f.gd2, e2 * 4, cudaMemcpyHostToDevice, sin));
  GA_CUDA_TRY(cudaEventRecord(ev_g, sin));

  GA_CUDA_TRY(cudaStreamWaitEvent(sk, ev_g, 0));
  GA_TRY(ga_nn_distance_bwd(b, n, m, f.d_x1, f.d_x2, f.d_g1, f.d_i1, f.d_g2, f.d_i2, f.d_o1, f.d_o2, (ga_stream_t)sk));
  GA_CUDA_TRY(cudaEventRecord(ev_b, sk));

  if (!f.mirror && (f.dist1 || f.idx1 || f.dist2 || f.idx2)) {
    GA_CUDA_TRY(cudaStreamWaitEvent(sout, ev_f, 0));
    if (f.dist1) GA_CUDA_TRY(cudaMemcpyAsync(f.dist1, f.d_d1, e1 * 4, cudaMemcpyDeviceToHost, sout));
    if (f.idx1) GA_CUDA_TRY(cudaMemcpyAsync(f.idx1, f.d_i1, e1 * 4, cudaMemcpyDeviceToHost, sout));
    if (f.dist2) GA_CUDA_TRY(cudaMemcpyAsync(f.dist2, f.d_d2, e2 * 4, cudaMemcpyDeviceToHost, sout));
    if (f.idx2) GA_CUDA_TRY(cudaMemcpyAsync(f.idx2, f.d_i2, e2 * 4, cudaMemcpyDeviceToHost, sout));
  }
  GA_CUDA_TRY(cudaStreamWaitEvent(sout, ev_b, 0));
  GA_CUDA_TRY(cudaMemcpyAsync(f.gx1, f.d_o1, e1 * 12, cudaMemcpyDeviceToHost, sout));
  GA_CUDA_TRY(cudaMemcpyAsync(f.gx2, f.d_o2, e2 * 12, cudaMemcpyDeviceToHost, sout));
  GA_CUDA_TRY(cudaEventRecord(join, sout));
  GA_CUDA_TRY(cudaStreamWaitEvent(sin, join, 0));
  return GA_OK;
}

static int pipeline_lanes(int nchunk) {
  Arena& A = t_arena;
  if (!A.out_lane) GA_CUDA_TRY(cudaStreamCreateWithFlags(&A.out_lane, cudaStreamNonBlocking));
  for (int i = 0; i < nchunk; i++)
    if (!A.k_lane[i]) GA_CUDA_TRY(cudaStreamCreateWithFlags(&A.k_lane[i], cudaStreamNonBlocking));
  for (auto& e : A.ev)
    if (!e) GA_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return GA_OK;
}

// Issue the pipeline on (A.stream = copy-in and origin, k lanes, out lane).  Used under capture.
static int issue_pipeline(const FwdBwdBufs& f, int b, int n, int m, int mode, int nchunk) {
  Arena& A = t_arena;
  cudaStream_t sin = A.stream, sout = A.out_lane;
  cudaEvent_t fork = A.ev[0], join = A.ev[1];
  cudaEvent_t* ev_x = A.ev + 2;    // xyz of chunk ch on the device
  cudaEvent_t* ev_g = A.ev + 10;   // upstream gradients of chunk ch on the device
  cudaEvent_t* ev_f = A.ev + 18;   // forward of chunk ch done
  cudaEvent_t* ev_b = A.ev + 26;   // backward of chunk ch done
  GA_CUDA_TRY(cudaEventRecord(fork, sin));
  GA_CUDA_TRY(cudaStreamWaitEvent(sout, fork, 0));
  for (int ch = 0; ch < nchunk; ch++) {
    const int b0 = (int)((long long)b * ch / nchunk), b1 = (int)((long long)b * (ch + 1) / nchunk);
    const int bc = b1 - b0;
    cudaStream_t sk = A.k_lane[ch];
    const size_t o1 = (size_t)b0 * n, o2 = (size_t)b0 * m, c1 = (size_t)bc * n, c2 = (size_t)bc * m;
    GA_CUDA_TRY(cudaMemcpyAsync(f.d_x1 + o1 * 3, f.xyz1 + o1 * 3, c1 * 12, cudaMemcpyHostToDevice, sin));
    GA_CUDA_TRY(cudaMemcpyAsync(f.d_x2 + o2 * 3, f.xyz2 + o2 * 3, c2 * 12, cudaMemcpyHostToDevice, sin));
    GA_CUDA_TRY(cudaEventRecord(ev_x[ch], sin));
    GA_CUDA_TRY(cudaMemcpyAsync(f.d_g1 + o1, f.gd1 + o1, c1 * 4, cudaMemcpyHostToDevice, sin));
    GA_CUDA_TRY(cudaMemcpyAsync(f.d_g2 + o2, f.gd2 + o2, c2 * 4, cudaMemcpyHostToDevice, sin));
    GA_CUDA_TRY(cudaEventRecord(ev_g[ch], sin));

    GA_CUDA_TRY(cudaStreamWaitEvent(sk, ev_x[ch], 0));
    if (f.mirror) {
      // results stream over PCIe while the search runs (1.6 MB of posted writes spread over the kernel)
      GA_TRY(nn_distance_fwd_mirrored(bc, n, m, f.d_x1 + o1 * 3, f.d_x2 + o2 * 3, f.d_d1 + o1, f.d_i1 + o1, f.d_d2 + o2,
                                      f.d_i2 + o2, f.dist1 + o1, f.idx1 + o1, f.dist2 + o2, f.idx2 + o2, mode,
                                      (ga_stream_t)sk));
    } else {
      GA_TRY(ga_nn_distance_fwd(bc, n, m, f.d_x1 + o1 * 3, f.d_x2 + o2 * 3, f.d_d1 + o1, f.d_i1 + o1, f.d_d2 + o2,
                                f.d_i2 + o2, mode, (ga_stream_t)sk));
    }
    GA_CUDA_TRY(cudaEventRecord(ev_f[ch], sk));
    GA_CUDA_TRY(cudaStreamWaitEvent(sk, ev_g[ch], 0));
    GA_TRY(ga_nn_distance_bwd(bc, n, m, f.d_x1 + o1 * 3, f.d_x2 + o2 * 3, f.d_g1 + o1, f.d_i1 + o1, f.d_g2 + o2,
                              f.d_i2 + o2, f.d_o1 + o1 * 3, f.d_o2 + o2 * 3, (ga_stream_t)sk));
    GA_CUDA_TRY(cudaEventRecord(ev_b[ch], sk));

    if (!f.mirror && (f.dist1 || f.idx1 || f.dist2 || f.idx2)) {  // a NULL output is not copied back
      GA_CUDA_TRY(cudaStreamWaitEvent(sout, ev_f[ch], 0));
      if (f.dist1) GA_CUDA_TRY(cudaMemcpyAsync(f.dist1 + o1, f.d_d1 + o1, c1 * 4, cudaMemcpyDeviceToHost, sout));
      if (f.idx1) GA_CUDA_TRY(cudaMemcpyAsync(f.idx1 + o1, f.d_i1 + o1, c1 * 4, cudaMemcpyDeviceToHost, sout));
      if (f.dist2) GA_CUDA_TRY(cudaMemcpyAsync(f.dist2 + o2, f.d_d2 + o2, c2 * 4, cudaMemcpyDeviceToHost, sout));
      if (f.idx2) GA_CUDA_TRY(cudaMemcpyAsync(f.idx2 + o2, f.d_i2 + o2, c2 * 4, cudaMemcpyDeviceToHost, sout));
    }
    GA_CUDA_TRY(cudaStreamWaitEvent(sout, ev_b[ch], 0));
    GA_CUDA_TRY(cudaMemcpyAsync(f.gx1 + o1 * 3, f.d_o1 + o1 * 3, c1 * 12, cudaMemcpyDeviceToHost, sout));
    GA_CUDA_TRY(cudaMemcpyAsync(f.gx2 + o2 * 3, f.d_o2 + o2 * 3, c2 * 12, cudaMemcpyDeviceToHost, sout));
  }
  // every kernel lane ends in ev_b[ch], which the out lane waited for: joining the out lane joins all
  GA_CUDA_TRY(cudaEventRecord(join, sout));
  GA_CUDA_TRY(cudaStreamWaitEvent(sin, join, 0));
  return GA_OK;
}

// The step with the clouds pulled by the SMs (ingest_stream_kernel): kernel lane = ingest -> forward (programmatic
// dependent, waits per batch element) -> backward; the upstream gradients come by DMA on the copy-in lane meanwhile;
// gradients (and dist/idx unless mirrored) leave on the out lane as in issue_pipeline.  ctas = ingest CTAs.
static int issue_pipeline_pulled(const FwdBwdBufs& f, int b, int n, int m, int mode, int ctas) {
  Arena& A = t_arena;
  cudaStream_t sin = A.stream, sout = A.out_lane, sk = A.k_lane[0];
  cudaEvent_t fork = A.ev[0], join = A.ev[1], ev_g = A.ev[10], ev_f = A.ev[18], ev_b = A.ev[26];
  const size_t e1 = (size_t)b * n, e2 = (size_t)b * m;
  GA_CUDA_TRY(cudaMemsetAsync(A.d_ready, 0, sizeof(int) * b, sin));
  GA_CUDA_TRY(cudaEventRecord(fork, sin));
  GA_CUDA_TRY(cudaStreamWaitEvent(sout, fork, 0));
  GA_CUDA_TRY(cudaStreamWaitEvent(sk, fork, 0));

  PullArgs pa;
  pa.src1 = reinterpret_cast<const int4*>(f.xyz1);
  pa.src2 = reinterpret_cast<const int4*>(f.xyz2);
  pa.dst1 = reinterpret_cast<int4*>(f.d_x1);
  pa.dst2 = reinterpret_cast<int4*>(f.d_x2);
  pa.ready = A.d_ready;
  pa.b = b;
  pa.v1 = n * 12 / 16;
  pa.v2 = m * 12 / 16;
  pa.mute = g_host_pull_mute;
  ingest_stream_kernel<<<ctas < b ? ctas : b, 256, 0, sk>>>(pa);
  GA_LAUNCH_CHECK("ingest_stream_kernel");

  t_ready_arm.flags = A.d_ready;
  t_ready_arm.per = 1;
  t_ready_arm.abort_word = A.d_abort;
  t_ready_arm.pdl = 1;
  const int rc = f.mirror ? nn_distance_fwd_mirrored(b, n, m, f.d_x1, f.d_x2, f.d_d1, f.d_i1, f.d_d2, f.d_i2, f.dist1,
                                                     f.idx1, f.dist2, f.idx2, mode, (ga_stream_t)sk)
                          : ga_nn_distance_fwd(b, n, m, f.d_x1, f.d_x2, f.d_d1, f.d_i1, f.d_d2, f.d_i2, mode,
                                               (ga_stream_t)sk);
  t_ready_arm.flags = nullptr;
  t_ready_arm.pdl = 0;
  GA_TRY(rc);
  GA_CUDA_TRY(cudaEventRecord(ev_f, sk));

  GA_CUDA_TRY(cudaMemcpyAsync(f.d_g1, f.gd1, e1 * 4, cudaMemcpyHostToDevice, sin));
  GA_CUDA_TRY(cudaMemcpyAsync(f.d_g2, f.gd2, e2 * 4, cudaMemcpyHostToDevice, sin));
  GA_CUDA_TRY(cudaEventRecord(ev_g, sin));

  GA_CUDA_TRY(cudaStreamWaitEvent(sk, ev_g, 0));
  GA_TRY(ga_nn_distance_bwd(b, n, m, f.d_x1, f.d_x2, f.d_g1, f.d_i1, f.d_g2, f.d_i2, f.d_o1, f.d_o2, (ga_stream_t)sk));
  GA_CUDA_TRY(cudaEventRecord(ev_b, sk));

  if (!f.mirror && (f.dist1 || f.idx1 || f.dist2 || f.idx2)) {
    GA_CUDA_TRY(cudaStreamWaitEvent(sout, ev_f, 0));
    if (f.dist1) GA_CUDA_TRY(cudaMemcpyAsync(f.dist1, f.d_d1, e1 * 4, cudaMemcpyDeviceToHost, sout));
    if (f.idx1) GA_CUDA_TRY(cudaMemcpyAsync(f.idx1, f.d_i1, e1 * 4, cudaMemcpyDeviceToHost, sout));
    if (f.dist2) GA_CUDA_TRY(cudaMemcpyAsync(f.dist2, f.d_d2, e2 * 4, cudaMemcpyDeviceToHost, sout));
    if (f.idx2) GA_CUDA_TRY(cudaMemcpyAsync(f.idx2, f.d_i2, e2 * 4, cudaMemcpyDeviceToHost, sout));
  }
  GA_CUDA_TRY(cudaStreamWaitEvent(sout, ev_b, 0));
  GA_CUDA_TRY(cudaMemcpyAsync(f.gx1, f.d_o1, e1 * 12, cudaMemcpyDeviceToHost, sout));
  GA_CUDA_TRY(cudaMemcpyAsync(f.gx2, f.d_o2, e2 * 12, cudaMemcpyDeviceToHost, sout));
  GA_CUDA_TRY(cudaEventRecord(join, sout));
  GA_CUDA_TRY(cudaStreamWaitEvent(sin, join, 0));
  return GA_OK;
}

// groups > 0: the streamed pipeline with that many arrival groups (nchunk is 1 then); groups < 0: the pulled pipeline
// with -groups ingest CTAs.
static int capture_pipeline(HostGraph& G, const FwdBwdBufs& f, int b, int n, int m, int mode, int nchunk,
                            int groups = 0) {
  GA_TRY(pipeline_lanes(nchunk));
  if (groups != 0) GA_TRY(stream_buffers());
  cudaStream_t s0 = t_arena.stream;
  const long long l0 = launch_count_now();
  GA_CUDA_TRY(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
  const int rc = groups > 0   ? issue_pipeline_streamed(f, b, n, m, mode, groups)
                 : groups < 0 ? issue_pipeline_pulled(f, b, n, m, mode, -groups)
                              : issue_pipeline(f, b, n, m, mode, nchunk);
  t_ready_arm.flags = nullptr;
  t_ready_arm.pdl = 0;
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(s0, &graph);
  if (rc != GA_OK || e != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc != GA_OK) return rc;
    return cuda_fail(e != cudaSuccess ? e : cudaErrorUnknown, "cudaStreamEndCapture (host pipeline)");
  }
  const cudaError_t ei = cudaGraphInstantiate(&G.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) {
    G.exec = nullptr;
    cudaGetLastError();
    return cuda_fail(ei, "cudaGraphInstantiate (host pipeline)");
  }
  G.launches = (int)(launch_count_now() - l0);
  count_launch(-G.launches);  // counted at capture, but nothing ran yet: replays count below
  G.nchunk = nchunk;
  G.streamed = groups != 0;
  return GA_OK;
}

// Returns the cache slot for this call (never null): a hit, or the least recently used slot re-keyed.
static HostGraph& graph_slot(int dev, int b, int n, int m, int mode, const void* const (&ptr)[10]) {
  Arena& A = t_arena;
  HostGraph* lru = &t_graphs[0];
  for (auto& G : t_graphs) {
    bool same = G.seen > 0 && G.dev == dev && G.b == b && G.n == n && G.m == m && G.mode == mode &&
                G.generation == A.generation && G.epoch == g_host_graph_epoch;
    for (int i = 0; same && i < 10; i++) same = G.ptr[i] == ptr[i];
    if (same) {
      G.stamp = ++t_graph_clock;
      return G;
    }
    if (G.stamp < lru->stamp) lru = &G;
  }
  if (lru->exec) cudaGraphExecDestroy(lru->exec);
  *lru = HostGraph();
  lru->dev = dev; lru->b = b; lru->n = n; lru->m = m; lru->mode = mode;
  for (int i = 0; i < 10; i++) lru->ptr[i] = ptr[i];
  lru->generation = A.generation;
  lru->epoch = g_host_graph_epoch;
  lru->stamp = ++t_graph_clock;
  return *lru;
}

}  // namespace ga

using namespace ga;

extern "C" {

int ga_nn_distance_fwd_host(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                            float* dist2, int* idx2, int mode) {
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_nn_distance_fwd_host: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const size_t e1 = (size_t)b * n, e2 = (size_t)b * m;
  if (e1 + e2 == 0) return GA_OK;
  char* base;
  cudaStream_t st;
  GA_TRY(arena_reserve(padded(e1 * 12) + padded(e2 * 12) + 2 * padded(e1 * 4) + 2 * padded(e2 * 4), &base, &st));
  Carver c(base);
  float* d_x1 = c.take<float>(e1 * 3);
  float* d_x2 = c.take<float>(e2 * 3);
  float* d_d1 = c.take<float>(e1);
  int* d_i1 = c.take<int>(e1);
  float* d_d2 = c.take<float>(e2);
  int* d_i2 = c.take<int>(e2);
  if (e1) GA_CUDA_TRY(cudaMemcpyAsync(d_x1, xyz1, e1 * 12, cudaMemcpyHostToDevice, st));
  if (e2) GA_CUDA_TRY(cudaMemcpyAsync(d_x2, xyz2, e2 * 12, cudaMemcpyHostToDevice, st));
  GA_TRY(ga_nn_distance_fwd(b, n, m, d_x1, d_x2, d_d1, d_i1, d_d2, d_i2, mode, (ga_stream_t)st));
  if (e1) {
    GA_CUDA_TRY(cudaMemcpyAsync(dist1, d_d1, e1 * 4, cudaMemcpyDeviceToHost, st));
    GA_CUDA_TRY(cudaMemcpyAsync(idx1, d_i1, e1 * 4, cudaMemcpyDeviceToHost, st));
  }
  if (e2) {
    GA_CUDA_TRY(cudaMemcpyAsync(dist2, d_d2, e2 * 4, cudaMemcpyDeviceToHost, st));
    GA_CUDA_TRY(cudaMemcpyAsync(idx2, d_i2, e2 * 4, cudaMemcpyDeviceToHost, st));
  }
  GA_CUDA_TRY(cudaStreamSynchronize(st));
  return GA_OK;
}

int ga_nn_distance_bwd_host(int b, int n, int m, const float* xyz1, const float* xyz2, const float* grad_dist1,
                            const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1,
                            float* grad_xyz2) {
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_nn_distance_bwd_host: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const size_t e1 = (size_t)b * n, e2 = (size_t)b * m;
  if (e1 + e2 == 0) return GA_OK;
  char* base;
  cudaStream_t st;
  GA_TRY(arena_reserve(2 * padded(e1 * 12) + 2 * padded(e2 * 12) + 2 * padded(e1 * 4) + 2 * padded(e2 * 4), &base,
                       &st));
  Carver c(base);
  float* d_x1 = c.take<float>(e1 * 3);
  float* d_x2 = c.take<float>(e2 * 3);
  float* d_g1 = c.take<float>(e1);
  int* d_i1 = c.take<int>(e1);
  float* d_g2 = c.take<float>(e2);
  int* d_i2 = c.take<int>(e2);
  float* d_o1 = c.take<float>(e1 * 3);
  float* d_o2 = c.take<float>(e2 * 3);
  if (e1) {
    GA_CUDA_TRY(cudaMemcpyAsync(d_x1, xyz1, e1 * 12, cudaMemcpyHostToDevice, st));
    GA_CUDA_TRY(cudaMemcpyAsync(d_g1, grad_dist1, e1 * 4, cudaMemcpyHostToDevice, st));
    GA_CUDA_TRY(cudaMemcpyAsync(d_i1, idx1, e1 * 4, cudaMemcpyHostToDevice, st));
  }
  if (e2) {
    GA_CUDA_TRY(cudaMemcpyAsync(d_x2, xyz2, e2 * 12, cudaMemcpyHostToDevice, st));
    GA_CUDA_TRY(cudaMemcpyAsync(d_g2, grad_dist2, e2 * 4, cudaMemcpyHostToDevice, st));
    GA_CUDA_TRY(cudaMemcpyAsync(d_i2, idx2, e2 * 4, cudaMemcpyHostToDevice, st));
  }
  GA_TRY(ga_nn_distance_bwd(b, n, m, d_x1, d_x2, d_g1, d_i1, d_g2, d_i2, d_o1, d_o2, (ga_stream_t)st));
  if (e1) GA_CUDA_TRY(cudaMemcpyAsync(grad_xyz1, d_o1, e1 * 12, cudaMemcpyDeviceToHost, st));
  if (e2) GA_CUDA_TRY(cudaMemcpyAsync(grad_xyz2, d_o2, e2 * 12, cudaMemcpyDeviceToHost, st));
  GA_CUDA_TRY(cudaStreamSynchronize(st));
  return GA_OK;
}

int ga_nn_distance_fwd_bwd_host(int b, int n, int m, const float* xyz1, const float* xyz2,
                                const float* grad_dist1, const float* grad_dist2, float* dist1, int* idx1,
                                float* dist2, int* idx2, float* grad_xyz1, float* grad_xyz2, int mode) {
  if (b < 0 || n < 0 || m < 0) {
    set_error("ga_nn_distance_fwd_bwd_host: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const size_t e1 = (size_t)b * n, e2 = (size_t)b * m;
  if (e1 + e2 == 0) return GA_OK;
  char* base;
  cudaStream_t st;
  GA_TRY(arena_reserve(2 * padded(e1 * 12) + 2 * padded(e2 * 12) + 3 * padded(e1 * 4) + 3 * padded(e2 * 4), &base,
                       &st));
  Carver c(base);
  float* d_x1 = c.take<float>(e1 * 3);
  float* d_x2 = c.take<float>(e2 * 3);
  float* d_g1 = c.take<float>(e1);
  float* d_g2 = c.take<float>(e2);
  float* d_d1 = c.take<float>(e1);
  int* d_i1 = c.take<int>(e1);
  float* d_d2 = c.take<float>(e2);
  int* d_i2 = c.take<int>(e2);
  float* d_o1 = c.take<float>(e1 * 3);
  float* d_o2 = c.take<float>(e2 * 3);
  // Opt-in (ga_set_tuning(2, 2)): every caller buffer is pinned and UVA-mapped.  Measured on the
  // box it is NOT faster than the copy engines (ingest kernel 36 GB/s vs 43 GB/s DMA; gradients
  // written over PCIe by the SMs 30 GB/s vs 45 GB/s DMA; whole step 284 us vs 266 us), so the
  // default stays with cudaMemcpyAsync.
  if (g_host_path == 2 && n > 0 && m > 0 && dist1 && idx1 && dist2 && idx2 && device_can_touch(xyz1) && device_can_touch(xyz2) && device_can_touch(grad_dist1) &&
      device_can_touch(grad_dist2) && device_can_touch(dist1) && device_can_touch(idx1) &&
      device_can_touch(dist2) && device_can_touch(idx2) && device_can_touch(grad_xyz1) &&
      device_can_touch(grad_xyz2)) {
    IngestArgs ia;
    ia.nseg = 4;
    ia.src[0] = xyz1; ia.dst[0] = d_x1; ia.bytes[0] = e1 * 12;
    ia.src[1] = xyz2; ia.dst[1] = d_x2; ia.bytes[1] = e2 * 12;
    ia.src[2] = grad_dist1; ia.dst[2] = d_g1; ia.bytes[2] = e1 * 4;
    ia.src[3] = grad_dist2; ia.dst[3] = d_g2; ia.bytes[3] = e2 * 4;
    ingest_kernel<<<sm_count() * 4, 256, 0, st>>>(ia);
    GA_LAUNCH_CHECK("ingest_kernel");
    GA_TRY(nn_distance_fwd_mirrored(b, n, m, d_x1, d_x2, d_d1, d_i1, d_d2, d_i2, dist1, idx1, dist2, idx2, mode,
                                    (ga_stream_t)st));
    GA_TRY(ga_nn_distance_bwd(b, n, m, d_x1, d_x2, d_g1, d_i1, d_g2, d_i2, grad_xyz1, grad_xyz2, (ga_stream_t)st));
    GA_CUDA_TRY(cudaStreamSynchronize(st));
    return GA_OK;
  }
  // Same pinned buffers as before: replay the captured pipeline (see HostGraph above).
  if (g_host_graph != 1 && n > 0 && m > 0 && b >= 1) {
    const void* const key[10] = {xyz1, xyz2, grad_dist1, grad_dist2, dist1, idx1, dist2, idx2, grad_xyz1, grad_xyz2};
    int dev = 0;
    GA_CUDA_TRY(cudaGetDevice(&dev));
    HostGraph& G = graph_slot(dev, b, n, m, mode, key);
    G.seen++;
    // Pinned status is re-checked on EVERY call (10 attribute look-ups, well under a microsecond each): a
    // caller may free a pinned buffer and get pageable memory back at the same address, and the captured
    // copy nodes would then DMA from unpinned pages.  Outputs 4..7 (dist/idx) may be NULL = not wanted.
    bool all_pinned = true;
    for (int i = 0; all_pinned && i < 10; i++) all_pinned = (key[i] == nullptr && i >= 4 && i < 8) || is_pinned(key[i]);
    if (!all_pinned && G.exec != nullptr) {
      cudaGraphExecDestroy(G.exec);
      G.exec = nullptr;
    }
    if (all_pinned && !G.failed && G.exec == nullptr && (G.seen >= 2 || g_host_graph == 2)) {
      {
        // Measured on the B200 box (profiles/r01_tune_e2e.json, B x 2048 x 2048): every extra chunk
        // costs ~25 us of dependency hops between copy and compute nodes, so the overlap only pays
        // once: B=10 one chunk 93 us (direct path 113), B=50 two chunks 210 us (direct 239, one chunk
        // 219, three 224, eight 350), B=200 two chunks 575 us (direct 611).
        const size_t traffic = (e1 + e2) * 36;
        int nc = g_host_graph_chunks > 0 ? g_host_graph_chunks : (traffic >= ((size_t)4 << 20) ? 2 : 1);
        nc = nc < 1 ? 1 : (nc > 8 ? 8 : nc);
        if (nc > b) nc = b;
        // The forward kernel writes dist/idx straight into the pinned host buffers while it searches, which saves
        // four copy nodes per chunk and keeps the D2H engine free for the gradients.  Since the kernel stores whole
        // 128-byte lines (mma_write) the posted PCIe writes no longer slow it down at any size (profiles/
        // r02_tune_e2e.txt: B=10 83 vs 96 us with copies, B=50 183 vs 213, B=200 527 vs 578; with the old
        // 32-byte pieces B=50 was 212 vs 212 and B=200 lost).
        const bool mirror = (g_host_graph_mirror == 2 || g_host_graph_mirror == 0) &&
                            dist1 && idx1 && dist2 && idx2 && device_can_touch(dist1) && device_can_touch(idx1) &&
                            device_can_touch(dist2) && device_can_touch(idx2);
        FwdBwdBufs f = {xyz1, xyz2, grad_dist1, grad_dist2, dist1, dist2, grad_xyz1, grad_xyz2, idx1, idx2,
                        d_x1, d_x2, d_g1, d_g2, d_d1, d_d2, d_o1, d_o2, d_i1, d_i2, mirror};
        // Streamed ingest (issue_pipeline_streamed), opt-in (key 26 = number of arrival groups).  It works, and it
        // does not pay on this platform: every copy node of a captured chain costs ~4 us before its first byte
        // moves, and a group needs three (profiles/r02_tune_e2e.txt, B=50, dist/idx mirrored: one group 185 us,
        // two 205, four 227, eight 297, against 183 for the two-chunk pipeline above).
        int groups = 0;
        if (g_host_stream > 0 && !G.no_stream && fwd_ready_supported(b, n, m)) {
          groups = g_host_stream;
          if (groups > kMaxReadyGroups) groups = kMaxReadyGroups;
          if (groups > b) groups = b;
        }
        // Pulled ingest (issue_pipeline_pulled): the clouds come over PCIe by the loads of a few CTAs instead of
        // copy nodes, one arrival flag per batch element, the search runs under the transfer.  Measured
        // (profiles/r02_tune_e2e.txt, dist/idx mirrored): B=50 174 us with 8 ingest CTAs (2: 225, 4: 187, 12: 177,
        // 32: 173, 50: 186) against 183-200 for the two-chunk pipeline; B=200 512 against 529.  Default from 4 MB
        // of traffic (below that the tcgen05 kernel serves the step and the copies are short); key 27 = -1 turns
        // it off, n > 0 sets the CTAs.  Needs device-readable, 16-byte aligned clouds.
        if (g_host_pull >= 0 && groups == 0 && !G.no_stream && b <= kMaxReadyElems && fwd_ready_supported(b, n, m) &&
            (g_host_pull > 0 || (traffic >= ((size_t)4 << 20) && traffic <= ((size_t)1 << 30))) &&
            device_can_touch(xyz1) && device_can_touch(xyz2) && ((uintptr_t)xyz1 & 15) == 0 && ((uintptr_t)xyz2 & 15) == 0)
          groups = -(g_host_pull > 0 ? g_host_pull : 8);
        if (capture_pipeline(G, f, b, n, m, mode, groups != 0 ? 1 : nc, groups) != GA_OK) G.failed = true;  // direct path
      }
    }
    t_last_streamed = 0;
    if (G.exec != nullptr) {
      GA_CUDA_TRY(cudaGraphLaunch(G.exec, st));
      count_launch(G.launches);
      GA_CUDA_TRY(cudaStreamSynchronize(st));
      t_last_streamed = G.streamed ? 1 : 0;
      if (!(G.streamed && *reinterpret_cast<volatile int*>(t_arena.h_abort) != 0)) return GA_OK;
      // A CTA gave up waiting for its clouds (wait_ready): the step's results are incomplete.  Drop the graph, keep
      // this key off the streamed pipeline, and redo the step on the direct path below.
      *reinterpret_cast<volatile int*>(t_arena.h_abort) = 0;
      cudaGraphExecDestroy(G.exec);
      G.exec = nullptr;
      G.streamed = false;
      G.no_stream = true;
      t_last_streamed = -1;
    }
  }
  // Batch elements are independent: split the batch into chunks that alternate between two
  // streams, so the H2D copy of chunk i+1 and the D2H copy of chunk i-1 run under the kernels
  // of chunk i (the copy engines are full duplex).  Needs pinned caller memory to overlap.
  // Measured on the B200 box (PCIe Gen5 x16, ~4 us per copy call, 50 GB/s), B=50 N=M=2048
  // (8 MB moved): 1 chunk 266 us, 2 chunks 245 us, 3 chunks 285 us, 4 chunks 307 us -- the extra
  // driver calls eat the overlap quickly, so: two chunks from 4 MB, four only for big batches.
  const size_t bytes = (e1 + e2) * 40;
  int nchunk = (bytes >= ((size_t)64 << 20) && b >= 4) ? 4 : ((bytes >= ((size_t)4 << 20) && b >= 2) ? 2 : 1);
  if (g_host_chunks > 0) nchunk = g_host_chunks < b ? g_host_chunks : b;
  cudaStream_t lanes[2] = {st, t_arena.stream2};
  for (int ch = 0; ch < nchunk; ch++) {
    const int b0 = (int)((long long)b * ch / nchunk), b1 = (int)((long long)b * (ch + 1) / nchunk);
    const int bc = b1 - b0;
    if (bc == 0) continue;
    cudaStream_t s = lanes[ch & 1];
    const size_t o1 = (size_t)b0 * n, o2 = (size_t)b0 * m, c1 = (size_t)bc * n, c2 = (size_t)bc * m;
    if (c1) {
      GA_CUDA_TRY(cudaMemcpyAsync(d_x1 + o1 * 3, xyz1 + o1 * 3, c1 * 12, cudaMemcpyHostToDevice, s));
      GA_CUDA_TRY(cudaMemcpyAsync(d_g1 + o1, grad_dist1 + o1, c1 * 4, cudaMemcpyHostToDevice, s));
    }
    if (c2) {
      GA_CUDA_TRY(cudaMemcpyAsync(d_x2 + o2 * 3, xyz2 + o2 * 3, c2 * 12, cudaMemcpyHostToDevice, s));
      GA_CUDA_TRY(cudaMemcpyAsync(d_g2 + o2, grad_dist2 + o2, c2 * 4, cudaMemcpyHostToDevice, s));
    }
    GA_TRY(ga_nn_distance_fwd(bc, n, m, d_x1 + o1 * 3, d_x2 + o2 * 3, d_d1 + o1, d_i1 + o1, d_d2 + o2, d_i2 + o2,
                              mode, (ga_stream_t)s));
    GA_TRY(ga_nn_distance_bwd(bc, n, m, d_x1 + o1 * 3, d_x2 + o2 * 3, d_g1 + o1, d_i1 + o1, d_g2 + o2, d_i2 + o2,
                              d_o1 + o1 * 3, d_o2 + o2 * 3, (ga_stream_t)s));
    if (c1) {
      if (dist1) GA_CUDA_TRY(cudaMemcpyAsync(dist1 + o1, d_d1 + o1, c1 * 4, cudaMemcpyDeviceToHost, s));
      if (idx1) GA_CUDA_TRY(cudaMemcpyAsync(idx1 + o1, d_i1 + o1, c1 * 4, cudaMemcpyDeviceToHost, s));
      GA_CUDA_TRY(cudaMemcpyAsync(grad_xyz1 + o1 * 3, d_o1 + o1 * 3, c1 * 12, cudaMemcpyDeviceToHost, s));
    }
    if (c2) {
      if (dist2) GA_CUDA_TRY(cudaMemcpyAsync(dist2 + o2, d_d2 + o2, c2 * 4, cudaMemcpyDeviceToHost, s));
      if (idx2) GA_CUDA_TRY(cudaMemcpyAsync(idx2 + o2, d_i2 + o2, c2 * 4, cudaMemcpyDeviceToHost, s));
      GA_CUDA_TRY(cudaMemcpyAsync(grad_xyz2 + o2 * 3, d_o2 + o2 * 3, c2 * 12, cudaMemcpyDeviceToHost, s));
    }
  }
  GA_CUDA_TRY(cudaStreamSynchronize(lanes[0]));
  if (nchunk > 1) GA_CUDA_TRY(cudaStreamSynchronize(lanes[1]));
  return GA_OK;
}

int ga_knn_host(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx) {
  if (b < 0 || n < 0 || m < 0 || k <= 0) {
    set_error(k <= 0 ? "SelectionSort expects positive k" : "ga_knn_host: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const size_t e1 = (size_t)b * n, e2 = (size_t)b * m, eo = (size_t)b * m * k;
  if (e2 == 0) return GA_OK;
  char* base;
  cudaStream_t st;
  GA_TRY(arena_reserve(padded(e1 * 12) + padded(e2 * 12) + 2 * padded(eo * 4), &base, &st));
  Carver c(base);
  float* d_x1 = c.take<float>(e1 * 3);
  float* d_x2 = c.take<float>(e2 * 3);
  float* d_v = c.take<float>(eo);
  int* d_i = c.take<int>(eo);
  if (e1) GA_CUDA_TRY(cudaMemcpyAsync(d_x1, xyz1, e1 * 12, cudaMemcpyHostToDevice, st));
  GA_CUDA_TRY(cudaMemcpyAsync(d_x2, xyz2, e2 * 12, cudaMemcpyHostToDevice, st));
  GA_TRY(ga_knn(b, n, m, k, d_x1, d_x2, d_v, d_i, (ga_stream_t)st));
  GA_CUDA_TRY(cudaMemcpyAsync(val, d_v, eo * 4, cudaMemcpyDeviceToHost, st));
  GA_CUDA_TRY(cudaMemcpyAsync(idx, d_i, eo * 4, cudaMemcpyDeviceToHost, st));
  GA_CUDA_TRY(cudaStreamSynchronize(st));
  return GA_OK;
}

int ga_knn_dists_host(int b, int n, int k, const float* pc, float* out) {
  if (b < 0 || n < 0 || k <= 0) {
    set_error(k <= 0 ? "SelectionSort expects positive k" : "ga_knn_dists_host: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  const size_t e1 = (size_t)b * n, eo = (size_t)b * n * k;
  if (e1 == 0) return GA_OK;
  char* base;
  cudaStream_t st;
  GA_TRY(arena_reserve(padded(e1 * 12) + padded(eo * 4), &base, &st));
  Carver c(base);
  float* d_x = c.take<float>(e1 * 3);
  float* d_o = c.take<float>(eo);
  GA_CUDA_TRY(cudaMemcpyAsync(d_x, pc, e1 * 12, cudaMemcpyHostToDevice, st));
  GA_TRY(ga_knn_dists(b, n, k, d_x, d_o, (ga_stream_t)st));
  GA_CUDA_TRY(cudaMemcpyAsync(out, d_o, eo * 4, cudaMemcpyDeviceToHost, st));
  GA_CUDA_TRY(cudaStreamSynchronize(st));
  return GA_OK;
}


// 1 if the calling thread's last ga_nn_distance_fwd_bwd_host replayed the streamed pipeline, -1 if that replay gave
// up waiting and the step was redone on the direct path, 0 otherwise (tests, tools/tune_e2e.py)
int ga_debug_host_streamed(void) { return t_last_streamed; }

// development hooks (tools/diag_e2e.py): the two zero-copy building blocks on a caller stream
int ga_debug_ingest(const void* src, void* dst, size_t bytes, ga_stream_t stream) {
  IngestArgs ia;
  ia.nseg = 1;
  ia.src[0] = src; ia.dst[0] = dst; ia.bytes[0] = bytes;
  ingest_kernel<<<sm_count() * 4, 256, 0, as_stream(stream)>>>(ia);
  GA_LAUNCH_CHECK("ingest_kernel");
  return GA_OK;
}
int ga_debug_fwd_mirrored(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                          float* dist2, int* idx2, float* mdist1, int* midx1, float* mdist2, int* midx2,
                          ga_stream_t stream) {
  return nn_distance_fwd_mirrored(b, n, m, xyz1, xyz2, dist1, idx1, dist2, idx2, mdist1, midx1, mdist2, midx2, 0,
                                  stream);
}

}  // extern "C"
