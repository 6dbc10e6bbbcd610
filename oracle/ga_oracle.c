/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * CPU restatement ("oracle") of the geometric_adv Chamfer / kNN hot path, in
 * plain C, single threaded, written from the reference's algorithm.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (geometric_adv_b200/)
 * never does.  Citations are path:line below /root/reference.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 *   - against the reference's own CPU kernels compiled unmodified
 *     (oracle/_ref/libga_ref.so, built by oracle/Makefile from
 *     external/structural_losses/tf_nndistance.cpp), when that library is
 *     present, and against fixtures generated from it (tests/golden/);
 *   - against the known answer printed by the reference's
 *     external/grouping/test/selection_sort.cpp (tests/golden/);
 *   - against the numpy statement of the reference's kNN fallback path
 *     (defender/get_knn_dists_per_point.py:125-137, src/general_utils.py:94-106).
 * Not pinned (no reference artefact exists): the TensorFlow-GPU arithmetic
 * of knn_point's distance tensor (tf_grouping.py:66-68, tensorflow-gpu
 * 1.13.2, not vendored).  We restate it as ((dx*dx+dy*dy)+dz*dz), unfused,
 * which is also what the reference's numpy path squares before its sqrt.
 *
 * Build: gcc -O2 -ffp-contract=off (no -march): every product and sum below
 * is a separate fp32 rounding, as in the reference's g++ -O2 build
 * (tf_nndistance_compile.sh:9).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- nn_distance ------------------------------------------------------- */

/* One direction.  tf_nndistance.cpp:21-43 (nnsearch): target minus query,
 * ((x*x)+(y*y))+(z*z) in fp32, first strict minimum wins; m==0 leaves 0/0.
 * mode 1 uses the contraction the reference's CUDA kernel
 * (tf_nndistance_g.cu:24-27) gets from nvcc: fma(z,z,fma(x,x,y*y)). */
static void oracle_nnsearch(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx,
                            int mode) {
  for (int i = 0; i < b; i++) {
    for (int j = 0; j < n; j++) {
      const float* q = xyz1 + ((size_t)i * n + j) * 3;
      float best = 0.0f;
      int besti = 0;
      for (int k = 0; k < m; k++) {
        const float* t = xyz2 + ((size_t)i * m + k) * 3;
        float x = t[0] - q[0];
        float y = t[1] - q[1];
        float z = t[2] - q[2];
        float d;
        if (mode == 0) {
          float xx = x * x, yy = y * y, zz = z * z;
          float s = xx + yy;
          d = s + zz;
        } else {
          d = fmaf(z, z, fmaf(x, x, y * y));
        }
        if (k == 0 || d < best) {
          best = d;
          besti = k;
        }
      }
      dist[(size_t)i * n + j] = best;
      idx[(size_t)i * n + j] = besti;
    }
  }
}

/* tf_nndistance.cpp:79-80: the search is run twice with the roles swapped. */
void ga_oracle_nn_distance(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1,
                           float* dist2, int* idx2, int mode) {
  oracle_nnsearch(b, n, m, xyz1, xyz2, dist1, idx1, mode);
  oracle_nnsearch(b, m, n, xyz2, xyz1, dist2, idx2, mode);
}

/* tf_nndistance.cpp:122-163: zero, then loop 1 over cloud 1, loop 2 over
 * cloud 2, per batch element; g*(x1-x2) is rounded before it is accumulated. */
void ga_oracle_nn_distance_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* gd1,
                                const int* idx1, const float* gd2, const int* idx2, float* gxyz1, float* gxyz2) {
  memset(gxyz1, 0, sizeof(float) * (size_t)b * n * 3);
  memset(gxyz2, 0, sizeof(float) * (size_t)b * m * 3);
  for (int i = 0; i < b; i++) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    float* o1 = gxyz1 + (size_t)i * n * 3;
    float* o2 = gxyz2 + (size_t)i * m * 3;
    for (int j = 0; j < n; j++) {
      int j2 = idx1[(size_t)i * n + j];
      float g = gd1[(size_t)i * n + j] * 2;
      for (int c = 0; c < 3; c++) {
        float t = g * (p1[j * 3 + c] - p2[j2 * 3 + c]);
        o1[j * 3 + c] += t;
        o2[j2 * 3 + c] -= t;
      }
    }
    for (int j = 0; j < m; j++) {
      int j2 = idx2[(size_t)i * m + j];
      float g = gd2[(size_t)i * m + j] * 2;
      for (int c = 0; c < 3; c++) {
        float t = g * (p2[j * 3 + c] - p1[j2 * 3 + c]);
        o2[j * 3 + c] += t;
        o1[j2 * 3 + c] -= t;
      }
    }
  }
}

/* Per-cloud Chamfer scalar as the scripts reduce it:
 * mean(dist1,1)+mean(dist2,1) (src/adv_ae.py:120-121,
 * attacker/prepare_indices_for_attack.py:113-114).  The sums are taken in
 * index order in fp32; TensorFlow's reduce_mean order is unpinned, so callers
 * compare this one with a tolerance (1e-6 relative). */
void ga_oracle_chamfer_per_cloud(int b, int n, int m, const float* dist1, const float* dist2, float* out) {
  for (int i = 0; i < b; i++) {
    float s1 = 0.0f, s2 = 0.0f;
    for (int j = 0; j < n; j++) s1 += dist1[(size_t)i * n + j];
    for (int j = 0; j < m; j++) s2 += dist2[(size_t)i * m + j];
    out[i] = s1 / (float)n + s2 / (float)m;
  }
}

/* ---- grouping ----------------------------------------------------------- */

/* tf_grouping_g.cu:83-123 == grouping/test/selection_sort.cpp:20-63: copy the
 * (b,m,n) matrix, identity indices, then k rounds of "first minimum of [s,n)
 * by strict <, swap into slot s".  Full (b,m,n) outputs like the op. */
void ga_oracle_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out) {
  for (size_t r = 0; r < (size_t)b * m; r++) {
    const float* src = dist + r * n;
    float* v = out + r * n;
    int* ix = outi + r * n;
    for (int s = 0; s < n; s++) {
      v[s] = src[s];
      ix[s] = s;
    }
    for (int s = 0; s < k && s < n; s++) {
      int mn = s;
      for (int t = s + 1; t < n; t++)
        if (v[t] < v[mn]) mn = t;
      if (mn != s) {
        float tv = v[mn];
        v[mn] = v[s];
        v[s] = tv;
        int ti = ix[mn];
        ix[mn] = ix[s];
        ix[s] = ti;
      }
    }
  }
}

/* Squared distance of knn_point's matrix entry: tf_grouping.py:66-68,
 * dataset (xyz1) minus query (xyz2), squared, summed over the 3 coordinates. */
static float oracle_sqdist(const float* p, const float* q) {
  float x = p[0] - q[0], y = p[1] - q[1], z = p[2] - q[2];
  float xx = x * x, yy = y * y, zz = z * z;
  float s = xx + yy;
  return s + zz;
}

/* knn_point(k, xyz1, xyz2): tf_grouping.py:48-75.  xyz1 (b,n,3) is the data
 * set, xyz2 (b,m,3) the queries; returns the first k columns of the selection
 * sort of row j = [sqdist(xyz1[i], xyz2[j]) for i<n]: val (b,m,k), idx (b,m,k). */
void ga_oracle_knn_point(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx) {
  float* row = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  float* v = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int* ix = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < b; i++) {
    for (int j = 0; j < m; j++) {
      const float* q = xyz2 + ((size_t)i * m + j) * 3;
      for (int t = 0; t < n; t++) row[t] = oracle_sqdist(xyz1 + ((size_t)i * n + t) * 3, q);
      ga_oracle_selection_sort(1, n, 1, k, row, ix, v);
      for (int s = 0; s < k; s++) {
        val[((size_t)i * m + j) * k + s] = v[s];
        idx[((size_t)i * m + j) * k + s] = ix[s];
      }
    }
  }
  free(row);
  free(v);
  free(ix);
}

/* tf_grouping_g.cu:40-57: out[b,j,s,:] = points[b, idx[b,j,s], :]. */
void ga_oracle_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx,
                           float* out) {
  for (int i = 0; i < b; i++)
    for (int j = 0; j < m; j++)
      for (int s = 0; s < nsample; s++) {
        int ii = idx[((size_t)i * m + j) * nsample + s];
        for (int l = 0; l < c; l++)
          out[(((size_t)i * m + j) * nsample + s) * c + l] = points[((size_t)i * n + ii) * c + l];
      }
}

/* defender/get_knn_dists_per_point.py:78-81: knn_point(k+1, pc, pc), drop the
 * first neighbour, group, subtract the centre, sqrt(sum(delta^2)) -> (b,n,k).
 * Same values as the numpy path (:125-137): sort of the norms, column 0 dropped. */
void ga_oracle_knn_dists(int b, int n, int k, const float* pc, float* out) {
  int kk = k + 1;
  float* val = (float*)malloc(sizeof(float) * (size_t)b * n * kk);
  int* idx = (int*)malloc(sizeof(int) * (size_t)b * n * kk);
  ga_oracle_knn_point(b, n, n, kk, pc, pc, val, idx);
  for (int i = 0; i < b; i++)
    for (int j = 0; j < n; j++) {
      const float* ctr = pc + ((size_t)i * n + j) * 3;
      for (int s = 0; s < k; s++) {
        int ii = idx[((size_t)i * n + j) * kk + s + 1];
        out[((size_t)i * n + j) * k + s] = sqrtf(oracle_sqdist(pc + ((size_t)i * n + ii) * 3, ctr));
      }
    }
  free(val);
  free(idx);
}

/* src/adversary_utils.py:149-178 (get_outlier_pc_inlier_pc): per cloud, points whose score exceeds
 * the threshold vs the rest, index order kept; a part that is neither empty nor the whole cloud
 * is padded with its last point; a NaN score falls into neither part. */
void ga_oracle_split_by_threshold(int b, int n, const float* pc, const float* score, float thresh,
                                  float* outlier_pc, int* outlier_idx, int* outlier_num, float* inlier_pc) {
  memset(outlier_pc, 0, sizeof(float) * (size_t)b * n * 3);
  memset(inlier_pc, 0, sizeof(float) * (size_t)b * n * 3);
  memset(outlier_idx, 0, sizeof(int) * (size_t)b * n);
  for (int l = 0; l < b; l++) {
    const float* p = pc + (size_t)l * n * 3;
    const float* s = score + (size_t)l * n;
    float* op = outlier_pc + (size_t)l * n * 3;
    float* ip = inlier_pc + (size_t)l * n * 3;
    int no = 0, ni = 0;
    for (int i = 0; i < n; i++) {
      if (s[i] > thresh) {
        memcpy(op + (size_t)no * 3, p + (size_t)i * 3, 3 * sizeof(float));
        outlier_idx[(size_t)l * n + no] = i;
        no++;
      }
      if (s[i] <= thresh) {
        memcpy(ip + (size_t)ni * 3, p + (size_t)i * 3, 3 * sizeof(float));
        ni++;
      }
    }
    outlier_num[l] = no;
    if (no > 0 && no < n)
      for (int i = no; i < n; i++) memcpy(op + (size_t)i * 3, op + (size_t)(no - 1) * 3, 3 * sizeof(float));
    if (ni > 0 && ni < n)
      for (int i = ni; i < n; i++) memcpy(ip + (size_t)i * 3, ip + (size_t)(ni - 1) * 3, 3 * sizeof(float));
  }
}
