// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of pcl::VoxelGrid<PointXYZI>::applyFilter as the reference uses it
// (downSizeFilterCorner/Surf: odomEstimationNode.cpp:110-111, :196-201, :272-277; leaf sizes
// mappingCornerLeafSize 0.2 / mappingSurfLeafSize 0.4, config/params.yaml:132-133).
// PCL is a third-party dependency absent from /root/reference (inferred PCL 1.8.1, unpinned);
// its published algorithm (pcl/filters/impl/voxel_grid.hpp) is restated here:
//   min/max over the cloud -> min_b = floor(min * inv_leaf), div_b = max_b - min_b + 1;
//   ijk = (int)(floor(p * inv_leaf) - (float)min_b); idx = ijk . (1, div_x, div_x * div_y);
//   sort by idx; one output point per occupied voxel in ascending idx = centroid of all fields
//   (x, y, z, intensity; fp32 accumulation, then division by the count).
// Deviation (documented): PCL sorts with std::sort on idx only (unstable), so the fp32 summation order
// inside a voxel is unspecified upstream; here (and on the GPU) points of a voxel are summed in
// ascending input index.  If dx*dy*dz overflows int32 PCL warns and returns the input unchanged.
#include "orc_api.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <numeric>
#include <vector>

extern "C" int32_t orc_voxel_grid(const float* pts4, int32_t n, float leaf, float* out4, int32_t cap) {
  if (n <= 0) return 0;
  const float inv = 1.0f / leaf;   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) { float v = pts4[4 * (size_t)i + d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v); }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {   // "Leaf size is too small": output = input
    int m = std::min(n, cap);
    memcpy(out4, pts4, sizeof(float) * 4 * (size_t)m);
    return n;
  }
  int minb[3], maxb[3], divb[3];
  for (int d = 0; d < 3; d++) { minb[d] = (int)std::floor(mn[d] * inv); maxb[d] = (int)std::floor(mx[d] * inv); divb[d] = maxb[d] - minb[d] + 1; }
  const int mul[3] = {1, divb[0], divb[0] * divb[1]};
  std::vector<uint32_t> idx(n);
  for (int i = 0; i < n; i++) {
    const float* p = pts4 + 4 * (size_t)i;
    int ijk0 = (int)(std::floor(p[0] * inv) - (float)minb[0]);
    int ijk1 = (int)(std::floor(p[1] * inv) - (float)minb[1]);
    int ijk2 = (int)(std::floor(p[2] * inv) - (float)minb[2]);
    idx[i] = (uint32_t)(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]);
  }
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return idx[a] < idx[b]; });
  int m = 0;
  for (int i = 0; i < n;) {
    int j = i;
    float s[4] = {0, 0, 0, 0};
    while (j < n && idx[order[j]] == idx[order[i]]) {
      const float* p = pts4 + 4 * (size_t)order[j];
      s[0] += p[0]; s[1] += p[1]; s[2] += p[2]; s[3] += p[3];
      j++;
    }
    const float cnt = (float)(j - i);
    if (m < cap) { out4[4 * (size_t)m] = s[0] / cnt; out4[4 * (size_t)m + 1] = s[1] / cnt; out4[4 * (size_t)m + 2] = s[2] / cnt; out4[4 * (size_t)m + 3] = s[3] / cnt; }
    m++;
    i = j;
  }
  return m;
}
