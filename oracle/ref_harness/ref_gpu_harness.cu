// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Compiles the reference's own CUDA kernels, UNMODIFIED and from where they lie
// under /root/reference (paths passed by oracle/Makefile), for sm_100a into
// oracle/_ref/libga_ref_gpu.so.  This is "the kernel to beat on the same box"
//  and the bit-level pin for mode 1 (the FMA contraction nvcc
// gives the reference kernel).  Launches go to the legacy default stream, as in
// the reference.
//
//   NmDistanceKernelLauncher      external/structural_losses/tf_nndistance_g.cu:128-131
//   NmDistanceGradKernelLauncher  external/structural_losses/tf_nndistance_g.cu:152-157
//   selectionSortLauncher         external/grouping/tf_grouping_g.cu:129-132
//   groupPointLauncher            external/grouping/tf_grouping_g.cu:133-136
#include <cuda_runtime.h>
#define GOOGLE_CUDA 1
#include GA_REF_NNDISTANCE_CU
#include GA_REF_GROUPING_CU

extern "C" {
int ga_refgpu_nn_distance(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                          float* dist2, int* idx2) {
  NmDistanceKernelLauncher(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2);
  return (int)cudaGetLastError();
}
int ga_refgpu_nn_distance_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* gd1,
                               const int* idx1, const float* gd2, const int* idx2, float* gxyz1, float* gxyz2) {
  NmDistanceGradKernelLauncher(b, n, xyz1, m, xyz2, gd1, idx1, gd2, idx2, gxyz1, gxyz2);
  return (int)cudaGetLastError();
}
int ga_refgpu_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out) {
  selectionSortLauncher(b, n, m, k, dist, outi, out);
  return (int)cudaGetLastError();
}
int ga_refgpu_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx,
                          float* out) {
  groupPointLauncher(b, n, c, m, nsample, points, idx, out);
  return (int)cudaGetLastError();
}
}
