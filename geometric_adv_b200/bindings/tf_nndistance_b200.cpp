// tf_nndistance_b200.cpp -- the file a maintainer of geometric_adv adds next to
// external/structural_losses/tf_nndistance.cpp, INSTEAD of tf_nndistance_g.cu.
//
// The reference's GPU ops (NnDistanceGpuOp / NnDistanceGradGpuOp, tf_nndistance.cpp:169-252) end in two
// raw-pointer launchers that tf_nndistance.cpp only declares (:168, :208) and tf_nndistance_g.cu defines
// (:128-131, :152-157).  This file defines them on top of the C ABI of libga_b200.so; tf_nndistance.cpp itself
// is compiled unmodified:
//
//   g++ -std=c++11 -shared -fPIC tf_nndistance.cpp tf_nndistance_b200.cpp -o tf_nndistance_so.so \
//       -I$TF_INC -I<repo>/include -L<repo>/geometric_adv_b200 -lga_b200 -Wl,-rpath,<repo>/geometric_adv_b200
//
// The test harness builds exactly this (against its stub of the two TF headers) and tests/test_bound_gpu.py drives
// the reference's own Compute() methods through it on the B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ga_b200.h"

namespace {
// GA_MODE_GPU_REF reproduces the bits of the CUDA kernel this file replaces (fma contraction as nvcc compiles
// tf_nndistance_g.cu); GA_MODE_CPU_EXACT those of the CPU op NnDistanceOp.  GA_B200_MODE=cpu_exact selects the latter.
int g_mode = -1;
int mode() {
  if (g_mode < 0) {
    const char* e = std::getenv("GA_B200_MODE");
    g_mode = (e != nullptr && std::strcmp(e, "cpu_exact") == 0) ? GA_MODE_CPU_EXACT : GA_MODE_GPU_REF;
  }
  return g_mode;
}
// TensorFlow 1.13 custom ops launch on the legacy default stream (no stream argument anywhere in the reference's
// <<< >>>); pass context->eigen_gpu_device().stream() through here to run on TF's compute stream instead.
ga_stream_t g_stream = nullptr;
void report(int rc, const char* what) {
  if (rc != GA_OK) std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, ga_last_error());  // chamfer3D.cu:145-151 printf()s too
}
}  // namespace

extern "C" void ga_b200_tf_set_mode(int m) { g_mode = m; }
extern "C" void ga_b200_tf_set_stream(void* s) { g_stream = static_cast<ga_stream_t>(s); }

void NmDistanceKernelLauncher(int b, int n, const float* xyz, int m, const float* xyz2, float* result, int* result_i,
                              float* result2, int* result2_i) {
  report(ga_nn_distance_fwd(b, n, m, xyz, xyz2, result, result_i, result2, result2_i, mode(), g_stream),
         "NmDistanceKernelLauncher");
}

void NmDistanceGradKernelLauncher(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                                  const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1,
                                  float* grad_xyz2) {
  report(ga_nn_distance_bwd(b, n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, g_stream),
         "NmDistanceGradKernelLauncher");
}
