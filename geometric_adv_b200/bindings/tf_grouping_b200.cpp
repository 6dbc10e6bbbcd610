// tf_grouping_b200.cpp -- added next to external/grouping/tf_grouping.cpp INSTEAD of tf_grouping_g.cu: the two
// launchers the scripts of geometric_adv reach (selectionSortLauncher, tf_grouping.cpp:108 / tf_grouping_g.cu:129-132;
// groupPointLauncher, tf_grouping.cpp:142 / tf_grouping_g.cu:133-136) on top of the C ABI of libga_b200.so.
// queryBallPointLauncher and groupPointGradLauncher are not called by any script of geometric_adv
// and keep their definitions from tf_grouping_g.cu if they are wanted.
//
// knn_point (tf_grouping.py:48-75) still builds the dense (b,m,n) tensor in front of SelectionSort; to drop it,
// register one more op whose Compute calls ga_knn(b, n, m, k, xyz1, xyz2, val, idx, stream).
#include <cstdio>

#include "ga_b200.h"

void selectionSortLauncher(int b, int n, int m, int k, const float* dist, int* outi, float* out) {
  const int rc = ga_selection_sort(b, n, m, k, dist, outi, out, nullptr);
  if (rc != GA_OK) std::fprintf(stderr, "selectionSortLauncher failed (%d): %s\n", rc, ga_last_error());
}

void groupPointLauncher(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out) {
  const int rc = ga_group_point(b, n, c, m, nsample, points, idx, out, nullptr);
  if (rc != GA_OK) std::fprintf(stderr, "groupPointLauncher failed (%d): %s\n", rc, ga_last_error());
}
