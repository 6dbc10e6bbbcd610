// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Builds the reference's own CPU kernels into oracle/_ref/libga_ref.so so the
// C restatement in oracle/ga_oracle.c (and, through it, the CUDA path) can be
// checked against the real thing.  The reference translation unit is compiled
// where it lies (the Makefile passes its path as GA_REF_NNDISTANCE_CPP); no
// reference source is copied into this repository.
//
//   NnDistanceOp::Compute      external/structural_losses/tf_nndistance.cpp:45-83
//   NnDistanceGradOp::Compute  external/structural_losses/tf_nndistance.cpp:84-166
//
// The `_mt` entry points split the batch over std::thread workers in THIS file
// (the reference is single threaded); they are the "all host cores" CPU baseline.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include GA_REF_NNDISTANCE_CPP  // the unmodified reference file

#include <atomic>
#include <thread>
#include <functional>
#include <vector>

// The reference file declares these two GPU launchers (tf_nndistance.cpp:168,208)
// and defines them in its .cu file, which is not part of the CPU oracle.
void NmDistanceKernelLauncher(int, int, const float*, int, const float*, float*, int*, float*, int*) {
  fprintf(stderr, "ga_ref: GPU launcher is not available in the CPU oracle\n");
  abort();
}
void NmDistanceGradKernelLauncher(int, int, const float*, int, const float*, const float*, const int*,
                                  const float*, const int*, float*, float*) {
  fprintf(stderr, "ga_ref: GPU launcher is not available in the CPU oracle\n");
  abort();
}

namespace {
using tensorflow::OpKernelContext;
using tensorflow::Tensor;
using tensorflow::TensorShape;

char g_last_error[512] = "";

int finish(const OpKernelContext& ctx) {
  if (ctx.status.ok()) return 0;
  snprintf(g_last_error, sizeof(g_last_error), "%s", ctx.status.msg_.c_str());
  return -1;
}

TensorShape shape_of(int rank, const long long* dims) {
  switch (rank) {
    case 0: return TensorShape{};
    case 1: return TensorShape{dims[0]};
    case 2: return TensorShape{dims[0], dims[1]};
    case 3: return TensorShape{dims[0], dims[1], dims[2]};
    default: return TensorShape{dims[0], dims[1], dims[2], dims[3]};
  }
}
}  // namespace

extern "C" {

const char* ga_ref_last_error() { return g_last_error; }

// Shapes are passed explicitly so the reference's own rank / last-dim / batch
// checks (tf_nndistance.cpp:51-58) can be exercised with bad shapes too.
int ga_ref_nn_distance_shaped(const float* xyz1, int rank1, const long long* dims1, const float* xyz2,
                              int rank2, const long long* dims2, float* dist1, int* idx1, float* dist2,
                              int* idx2, long long cap1, long long cap2) {
  tensorflow::OpKernelConstruction c;
  NnDistanceOp op(&c);
  OpKernelContext ctx;
  ctx.inputs = {Tensor((void*)xyz1, shape_of(rank1, dims1)), Tensor((void*)xyz2, shape_of(rank2, dims2))};
  ctx.out_ptr = {dist1, idx1, dist2, idx2};
  ctx.out_capacity = {cap1, cap1, cap2, cap2};
  op.Compute(&ctx);
  return finish(ctx);
}

int ga_ref_nn_distance(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                       float* dist2, int* idx2) {
  long long d1[3] = {b, n, 3}, d2[3] = {b, m, 3};
  return ga_ref_nn_distance_shaped(xyz1, 3, d1, xyz2, 3, d2, dist1, idx1, dist2, idx2, (long long)b * n,
                                   (long long)b * m);
}

int ga_ref_nn_distance_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* gd1,
                            const int* idx1, const float* gd2, const int* idx2, float* gxyz1, float* gxyz2) {
  tensorflow::OpKernelConstruction c;
  NnDistanceGradOp op(&c);
  OpKernelContext ctx;
  ctx.inputs = {Tensor((void*)xyz1, TensorShape{b, n, 3}), Tensor((void*)xyz2, TensorShape{b, m, 3}),
                Tensor((void*)gd1, TensorShape{b, n}),     Tensor((void*)idx1, TensorShape{b, n}),
                Tensor((void*)gd2, TensorShape{b, m}),     Tensor((void*)idx2, TensorShape{b, m})};
  ctx.out_ptr = {gxyz1, gxyz2};
  ctx.out_capacity = {(long long)b * n * 3, (long long)b * m * 3};
  op.Compute(&ctx);
  return finish(ctx);
}

// All-cores variants: one reference Compute() per batch element, batch elements
// handed out to worker threads.  Results are identical to the single call because
// batch elements are independent in the reference loops.
int ga_ref_max_threads() {
  unsigned h = std::thread::hardware_concurrency();
  return h ? (int)h : 1;
}

static int run_batch_parallel(int b, int threads, const std::function<int(int)>& one) {
  if (threads <= 0) threads = ga_ref_max_threads();
  if (threads > b) threads = b > 0 ? b : 1;
  std::atomic<int> next(0), rc(0);
  auto worker = [&]() {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= b) break;
      if (one(i) != 0) rc.store(-1);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return rc.load();
}

int ga_ref_nn_distance_mt(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                          float* dist2, int* idx2, int threads) {
  return run_batch_parallel(b, threads, [&](int i) {
    return ga_ref_nn_distance(1, n, m, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3, dist1 + (size_t)i * n,
                              idx1 + (size_t)i * n, dist2 + (size_t)i * m, idx2 + (size_t)i * m);
  });
}

int ga_ref_nn_distance_grad_mt(int b, int n, int m, const float* xyz1, const float* xyz2, const float* gd1,
                               const int* idx1, const float* gd2, const int* idx2, float* gxyz1, float* gxyz2,
                               int threads) {
  return run_batch_parallel(b, threads, [&](int i) {
    return ga_ref_nn_distance_grad(1, n, m, xyz1 + (size_t)i * n * 3, xyz2 + (size_t)i * m * 3,
                                   gd1 + (size_t)i * n, idx1 + (size_t)i * n, gd2 + (size_t)i * m,
                                   idx2 + (size_t)i * m, gxyz1 + (size_t)i * n * 3, gxyz2 + (size_t)i * m * 3);
  });
}

}  // extern "C"
