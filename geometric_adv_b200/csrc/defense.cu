// Off-surface defense epilogue on the GPU: split every cloud into the points
// whose score (mean distance to the nearest neighbours, run_defense_surface.py:187-191) exceeds a
// threshold and the rest.  Restates get_outlier_pc_inlier_pc (src/adversary_utils.py:149-178), a
// numpy loop over clouds: stable (index-ascending) compaction of both parts; a part that is
// neither empty nor the whole cloud is padded with its own last point (duplicates do not change
// the latent vector under global max pooling); empty parts stay zero.  A NaN score belongs to
// neither part (both comparisons are false), exactly as in numpy.
#include "ga_common.cuh"

namespace ga {

struct SplitArgs {
  int b, n;
  const float* pc;     // (b,n,3)
  const float* score;  // (b,n)
  float thresh;
  float* outlier_pc;   // (b,n,3)
  int* outlier_idx;    // (b,n)
  int* outlier_num;    // (b)
  float* inlier_pc;    // (b,n,3)
};

constexpr int kSplitThreads = 256;

__global__ void __launch_bounds__(kSplitThreads) split_by_threshold_kernel(const SplitArgs a) {
  __shared__ int wsum[2][kSplitThreads / 32];
  __shared__ int base[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cl = blockIdx.x, n = a.n;
  const float* pc = a.pc + (size_t)cl * n * 3;
  const float* sc = a.score + (size_t)cl * n;
  float* opc = a.outlier_pc + (size_t)cl * n * 3;
  float* ipc = a.inlier_pc + (size_t)cl * n * 3;
  int* oidx = a.outlier_idx + (size_t)cl * n;
  if (tid == 0) base[0] = base[1] = 0;
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  for (int c0 = 0; c0 < n; c0 += kSplitThreads) {
    const int i = c0 + tid;
    const bool in = i < n;
    const float s = in ? __ldg(sc + i) : 0.f;
    const bool is_out = in && s > a.thresh;
    const bool is_in = in && s <= a.thresh;
    const unsigned mo = __ballot_sync(0xffffffffu, is_out), mi = __ballot_sync(0xffffffffu, is_in);
    if (lane == 0) {
      wsum[0][warp] = __popc(mo);
      wsum[1][warp] = __popc(mi);
    }
    __syncthreads();
    int offo = base[0], offi = base[1];
    for (int w = 0; w < warp; w++) {
      offo += wsum[0][w];
      offi += wsum[1][w];
    }
    float x = 0.f, y = 0.f, z = 0.f;
    if (in) {
      x = __ldg(pc + (size_t)i * 3);
      y = __ldg(pc + (size_t)i * 3 + 1);
      z = __ldg(pc + (size_t)i * 3 + 2);
    }
    if (is_out) {
      const int p = offo + __popc(mo & lt);
      opc[(size_t)p * 3] = x; opc[(size_t)p * 3 + 1] = y; opc[(size_t)p * 3 + 2] = z;
      oidx[p] = i;
    }
    if (is_in) {
      const int p = offi + __popc(mi & lt);
      ipc[(size_t)p * 3] = x; ipc[(size_t)p * 3 + 1] = y; ipc[(size_t)p * 3 + 2] = z;
    }
    __syncthreads();
    if (tid == 0) {
      int so = 0, si = 0;
      for (int w = 0; w < kSplitThreads / 32; w++) {
        so += wsum[0][w];
        si += wsum[1][w];
      }
      base[0] += so;
      base[1] += si;
    }
    __syncthreads();
  }
  const int no = base[0], ni = base[1];
  if (tid == 0) a.outlier_num[cl] = no;
  // tails: index list zero-filled; point lists padded with the part's last point (or zeros if empty)
  for (int p = no + tid; p < n; p += kSplitThreads) oidx[p] = 0;
  {
    float lx = 0.f, ly = 0.f, lz = 0.f;
    if (no > 0) { lx = opc[(size_t)(no - 1) * 3]; ly = opc[(size_t)(no - 1) * 3 + 1]; lz = opc[(size_t)(no - 1) * 3 + 2]; }
    for (int p = no + tid; p < n; p += kSplitThreads) { opc[(size_t)p * 3] = lx; opc[(size_t)p * 3 + 1] = ly; opc[(size_t)p * 3 + 2] = lz; }
  }
  {
    float lx = 0.f, ly = 0.f, lz = 0.f;
    if (ni > 0) { lx = ipc[(size_t)(ni - 1) * 3]; ly = ipc[(size_t)(ni - 1) * 3 + 1]; lz = ipc[(size_t)(ni - 1) * 3 + 2]; }
    for (int p = ni + tid; p < n; p += kSplitThreads) { ipc[(size_t)p * 3] = lx; ipc[(size_t)p * 3 + 1] = ly; ipc[(size_t)p * 3 + 2] = lz; }
  }
}

}  // namespace ga

extern "C" int ga_split_by_threshold(int b, int n, const float* pc, const float* score, float thresh,
                                     float* outlier_pc, int* outlier_idx, int* outlier_num, float* inlier_pc,
                                     ga_stream_t stream) {
  using namespace ga;
  if (b < 0 || n < 0) {
    set_error("ga_split_by_threshold: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (b == 0) return GA_OK;
  if (n == 0) return cudaMemsetAsync(outlier_num, 0, sizeof(int) * (size_t)b, as_stream(stream)) == cudaSuccess
                         ? GA_OK
                         : cuda_fail(cudaGetLastError(), "cudaMemsetAsync");
  SplitArgs a;
  a.b = b; a.n = n; a.pc = pc; a.score = score; a.thresh = thresh;
  a.outlier_pc = outlier_pc; a.outlier_idx = outlier_idx; a.outlier_num = outlier_num; a.inlier_pc = inlier_pc;
  split_by_threshold_kernel<<<b, kSplitThreads, 0, as_stream(stream)>>>(a);
  GA_LAUNCH_CHECK("split_by_threshold_kernel");
  return GA_OK;
}
