// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the map-based dynamic-object removal of the local-map update (SURVEY.md 8f "next" #2, kernel part):
//   map_scan_feature_pts_distance_removal   src/include/subMap.h:1063-1098
//   call site update_local_map              :886-905 (only the dynamic-class cloud is filtered), thresholds :842-845
// A feature point survives when it is outside the centre disc (x^2 + y^2 > center_radius^2) or when its squared
// distance d2 to the nearest map point satisfies (d2 > near^2 && d2 < dyn_min^2) || d2 > dyn_max^2; survivors keep
// their order (:1078-1092).  Clouds of <= 10 points are left untouched (:1069-1070).  The 1-NN is pcl::search::KdTree
// (FLANN, exact) -> orc_kdtree.
#include "orc_api.h"
#include <cstring>

extern "C" int32_t orc_map_distance_filter(const float* feat4, int32_t n, const float* map4, int32_t m, float center_radius,
                                           float dyn_min, float dyn_max, float near_thre, uint8_t* keep) {
  if (n <= 10) { for (int i = 0; i < n; i++) keep[i] = 1; return n; }
  void* tree = m > 0 ? orc_kdtree_build(map4, m) : nullptr;
  int kept = 0;
  for (int i = 0; i < n; i++) {
    const float* p = feat4 + 4 * (size_t)i;
    bool k;
    if (p[0] * p[0] + p[1] * p[1] > center_radius * center_radius) k = true;
    else {
      int idx; float d2 = 0.f;
      if (!tree || orc_kdtree_knn(tree, p, 1, &idx, &d2) < 1) k = true;     // empty map: nothing to compare with (reference would index an empty result: UB)
      else k = (d2 > near_thre * near_thre && d2 < dyn_min * dyn_min) || d2 > dyn_max * dyn_max;
    }
    keep[i] = k ? 1 : 0; kept += k;
  }
  if (tree) orc_kdtree_free(tree);
  return kept;
}
