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
  if (g_host_path == 2 && n > 0 && m > 0 && device_can_touch(xyz1) && device_can_touch(xyz2) && device_can_touch(grad_dist1) &&
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
      GA_CUDA_TRY(cudaMemcpyAsync(dist1 + o1, d_d1 + o1, c1 * 4, cudaMemcpyDeviceToHost, s));
      GA_CUDA_TRY(cudaMemcpyAsync(idx1 + o1, d_i1 + o1, c1 * 4, cudaMemcpyDeviceToHost, s));
      GA_CUDA_TRY(cudaMemcpyAsync(grad_xyz1 + o1 * 3, d_o1 + o1 * 3, c1 * 12, cudaMemcpyDeviceToHost, s));
    }
    if (c2) {
      GA_CUDA_TRY(cudaMemcpyAsync(dist2 + o2, d_d2 + o2, c2 * 4, cudaMemcpyDeviceToHost, s));
      GA_CUDA_TRY(cudaMemcpyAsync(idx2 + o2, d_i2 + o2, c2 * 4, cudaMemcpyDeviceToHost, s));
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
