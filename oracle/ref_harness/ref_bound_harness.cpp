// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// The drop-in boundary, for real: the reference's tf_nndistance.cpp compiled UNMODIFIED (where it lies) together
// with the binding file a maintainer adds (geometric_adv_b200/bindings/tf_nndistance_b200.cpp, which defines the two
// launchers tf_nndistance.cpp declares at :168 and :208) and linked against libga_b200.so.  The entry points below
// run the reference's own GPU OpKernels
//   NnDistanceGpuOp::Compute      external/structural_losses/tf_nndistance.cpp:169-206
//   NnDistanceGradGpuOp::Compute  external/structural_losses/tf_nndistance.cpp:209-252
// through the TF header stub with DEVICE pointers: shape checks, allocate_output and the launcher call are the
// reference's code, the kernels behind the launchers are this repo's.
#include <cstdio>

#include GA_REF_NNDISTANCE_CPP  // the unmodified reference file

namespace {
using tensorflow::OpKernelContext;
using tensorflow::Tensor;
using tensorflow::TensorShape;
char g_err[512] = "";
int finish(const OpKernelContext& ctx) {
  if (ctx.status.ok()) return 0;
  snprintf(g_err, sizeof(g_err), "%s", ctx.status.msg_.c_str());
  return -1;
}
TensorShape shape_of(int rank, const long long* d) {
  switch (rank) {
    case 0: return TensorShape{};
    case 1: return TensorShape{d[0]};
    case 2: return TensorShape{d[0], d[1]};
    case 3: return TensorShape{d[0], d[1], d[2]};
    default: return TensorShape{d[0], d[1], d[2], d[3]};
  }
}
}  // namespace

extern "C" {

const char* ga_bound_last_error() { return g_err; }

// device pointers; shapes passed as they are so that the reference's own OP_REQUIRES run on them
int ga_bound_nn_distance_gpu(const float* xyz1, int rank1, const long long* dims1, const float* xyz2, int rank2,
                             const long long* dims2, float* dist1, int* idx1, float* dist2, int* idx2, long long cap1,
                             long long cap2) {
  tensorflow::OpKernelConstruction c;
  NnDistanceGpuOp op(&c);
  OpKernelContext ctx;
  ctx.inputs = {Tensor((void*)xyz1, shape_of(rank1, dims1)), Tensor((void*)xyz2, shape_of(rank2, dims2))};
  ctx.out_ptr = {dist1, idx1, dist2, idx2};
  ctx.out_capacity = {cap1, cap1, cap2, cap2};
  op.Compute(&ctx);
  return finish(ctx);
}

int ga_bound_nn_distance_grad_gpu(int b, int n, int m, const float* xyz1, const float* xyz2, const float* gd1,
                                  const int* idx1, const float* gd2, const int* idx2, float* gxyz1, float* gxyz2) {
  tensorflow::OpKernelConstruction c;
  NnDistanceGradGpuOp op(&c);
  OpKernelContext ctx;
  ctx.inputs = {Tensor((void*)xyz1, TensorShape{b, n, 3}), Tensor((void*)xyz2, TensorShape{b, m, 3}),
                Tensor((void*)gd1, TensorShape{b, n}),     Tensor((void*)idx1, TensorShape{b, n}),
                Tensor((void*)gd2, TensorShape{b, m}),     Tensor((void*)idx2, TensorShape{b, m})};
  ctx.out_ptr = {gxyz1, gxyz2};
  ctx.out_capacity = {(long long)b * n * 3, (long long)b * m * 3};
  op.Compute(&ctx);
  return finish(ctx);
}

}  // extern "C"
