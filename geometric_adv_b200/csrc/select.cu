// Second half of attacker/prepare_indices_for_attack.py on the GPU: sort_dist_mat (:167-180).
//
// For every row of the (rows, s) Chamfer matrix block and every target class c (columns
// [slice[c], slice[c+1])), nn_idx[row, slice[c] + r] = class-local index of the r-th nearest
// instance, int16 -- the array the attack reads its targets from (src/adversary_utils.py:51-63:
// the first num_pc_for_target entries of a row's class block, skipping entry 0 for the shape's own
// class).  The reference calls np.argsort (quicksort: the order of exact ties is unspecified);
// this kernel is the stable variant (ties by ascending index), NaN last, -0 == +0, exactly what
// np.argsort(kind="stable") gives.
//
// One CTA per (row, class): the class block of the row is staged in shared memory as
// order-preserving integer keys and every thread ranks its own elements by counting the keys
// that precede them (broadcast reads; blocks are a few hundred columns, so O(L^2) is cheaper
// than a sorting network and needs no padding).  Also sums the matrix with its transpose first
// when asked (CD = D + D^T from the directed terms), so the sort can run right behind the all-gather.
#include "ga_common.cuh"

namespace ga {

constexpr int kSortThreads = 256;
constexpr int kSortMaxClass = 12288;  // keys of one class block in 48 KB of shared memory

__device__ __forceinline__ unsigned order_key(float v) {
  if (v != v) return 0xffffffffu;          // NaN sorts last
  v += 0.0f;                               // -0 -> +0
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(kSortThreads) sort_dist_mat_kernel(int s, int rows, const float* __restrict__ dm,
                                                                    int nclass, const int* __restrict__ slice,
                                                                    short* __restrict__ nn_idx) {
  extern __shared__ unsigned keys[];
  const int row = blockIdx.x, c = blockIdx.y, tid = threadIdx.x;
  const int c0 = slice[c], c1 = slice[c + 1];
  const int len = c1 - c0;
  const float* src = dm + (size_t)row * s + c0;
  for (int j = tid; j < len; j += kSortThreads) keys[j] = order_key(__ldg(src + j));
  __syncthreads();
  short* dst = nn_idx + (size_t)row * s + c0;
  for (int j = tid; j < len; j += kSortThreads) {
    const unsigned kj = keys[j];
    int rank = 0;
    int i = 0;
    for (; i + 4 <= len; i += 4) {  // broadcast reads, four compares per round
      const uint4 k4 = *reinterpret_cast<const uint4*>(keys + i);
      rank += (k4.x < kj || (k4.x == kj && i < j)) ? 1 : 0;
      rank += (k4.y < kj || (k4.y == kj && i + 1 < j)) ? 1 : 0;
      rank += (k4.z < kj || (k4.z == kj && i + 2 < j)) ? 1 : 0;
      rank += (k4.w < kj || (k4.w == kj && i + 3 < j)) ? 1 : 0;
    }
    for (; i < len; i++) rank += (keys[i] < kj || (keys[i] == kj && i < j)) ? 1 : 0;
    dst[rank] = (short)j;
  }
}

// out[r, j] = d[row0 + r, j] + d[j, row0 + r]  (d is the full (s,s) matrix of directed terms)
__global__ void __launch_bounds__(256) symmetrize_rows_kernel(int s, int row0, int rows, const float* __restrict__ d,
                                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int jb = blockIdx.x * 32, rb = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {  // tile[k][tx] = d[jb + k, row0 + rb + tx]  (coalesced along the row of d)
    const int j = jb + k, r = rb + tx;
    tile[k][tx] = (j < s && r < rows) ? __ldg(d + (size_t)j * s + row0 + r) : 0.0f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int r = rb + k, j = jb + tx;
    if (r < rows && j < s) out[(size_t)r * s + j] = __ldg(d + (size_t)(row0 + r) * s + j) + tile[tx][k];
  }
}

}  // namespace ga

using namespace ga;

extern "C" {

int ga_sort_dist_mat(int s, int rows, const float* dist_rows, int nclass, const int* slice_idx_dev, int max_class,
                     short* nn_idx, ga_stream_t stream) {
  if (s < 0 || rows < 0 || nclass < 0 || max_class < 0) {
    set_error("ga_sort_dist_mat: negative size");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (s > 32767 + 1 || max_class > 32768) {
    set_error("ga_sort_dist_mat: class-local indices are int16 (prepare_indices_for_attack.py:168), %d columns", max_class);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (max_class > kSortMaxClass) {
    set_error("ga_sort_dist_mat: classes of more than %d shapes are not supported (%d)", kSortMaxClass, max_class);
    return GA_ERR_UNSUPPORTED;
  }
  if (rows == 0 || nclass == 0 || s == 0) return GA_OK;
  if (dist_rows == nullptr || slice_idx_dev == nullptr || nn_idx == nullptr) {
    set_error("ga_sort_dist_mat: null pointer");
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (nclass > 65535) {
    set_error("ga_sort_dist_mat: too many classes (%d)", nclass);
    return GA_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)((max_class + 3) & ~3) * sizeof(unsigned);
  sort_dist_mat_kernel<<<dim3((unsigned)rows, (unsigned)nclass), kSortThreads, smem, as_stream(stream)>>>(
      s, rows, dist_rows, nclass, slice_idx_dev, nn_idx);
  GA_LAUNCH_CHECK("sort_dist_mat_kernel");
  return GA_OK;
}

int ga_symmetrize_rows(int s, int row0, int rows, const float* directed, float* out, ga_stream_t stream) {
  if (s < 0 || row0 < 0 || rows < 0 || row0 + rows > s) {
    set_error("ga_symmetrize_rows: bad sizes (s=%d row0=%d rows=%d)", s, row0, rows);
    return GA_ERR_INVALID_ARGUMENT;
  }
  if (rows == 0 || s == 0) return GA_OK;
  if (directed == nullptr || out == nullptr) {
    set_error("ga_symmetrize_rows: null pointer");
    return GA_ERR_INVALID_ARGUMENT;
  }
  symmetrize_rows_kernel<<<dim3((unsigned)((s + 31) / 32), (unsigned)((rows + 31) / 32)), 256, 0, as_stream(stream)>>>(
      s, row0, rows, directed, out);
  GA_LAUNCH_CHECK("symmetrize_rows_kernel");
  return GA_OK;
}

}  // extern "C"
