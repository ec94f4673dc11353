// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the LOAM feature extraction front end:
//   projectPointCloud    src/core/laserProcessing.cpp:467-510   (F1; deskew is out of scope = identity)
//   cloudExtraction      :515-539  (F2)
//   calculateSmoothness  :544-563  (F3)
//   markOccludedPoints   :568-605  (F4)
//   extractFeatures      :610-713  (F5)
// (src/core/featureExtraction.cpp:125-366 is a byte-identical duplicate.)
//
// Documented resolutions of reference quirks:
//  Q3  std::sort covers [sp, ep) but the pick loops run over [sp, ep]            -> reproduced.
//  Q5  cloudSmoothness is clear()ed then indexed; entries outside [5, M-5) are stale from the
//      previous frame (index 4 of ring 0 is the only one ever read)           -> engine is
//      stateless: entry i is {curvature[i] (0 outside the stencil range), i}.
//  --  std::sort is unstable on equal curvatures                              -> key = (value, index).
//  --  pointColInd[ind + l] is read at index -1 when point 4 of ring 0 is picked (UB) -> indices
//      outside [0, M) end the neighbour walk.
#include "orc_api.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

extern "C" {

// F1 + F2.  pts4: n x {x,y,z,intensity}; ring: n.  Outputs (capacity n_scan*horizon):
// src_index (index of the input point kept in each extracted slot), col_ind, range,
// start_ring/end_ring (n_scan each).  Returns M = number of extracted points.
int32_t orc_project_scan(const float* pts4, const uint16_t* ring, int32_t n, const orc_feat_params* prm,
                         int32_t* src_index, int32_t* col_ind, float* range, int32_t* start_ring, int32_t* end_ring) {
  const int NS = prm->n_scan, H = prm->horizon;
  std::vector<float> rangeMat((size_t)NS * H, FLT_MAX);
  std::vector<int32_t> srcMat((size_t)NS * H, -1);
  const float ang_res_x = 360.0 / float(H);   // static float ang_res_x (:493)
  for (int i = 0; i < n; i++) {
    const float x = pts4[4 * (size_t)i], y = pts4[4 * (size_t)i + 1], z = pts4[4 * (size_t)i + 2];
    float r = std::sqrt(x * x + y * y + z * z);   // pointDistance, common.h
    if (r < prm->min_range || r > prm->max_range) continue;
    int rowIdn = ring[i];
    if (rowIdn < 0 || rowIdn >= NS) continue;
    if (rowIdn % prm->downsample_rate != 0) continue;
    // atan2f taken as the correctly rounded float of the double routine (see DESIGN.md numerics)
    float horizonAngle = (float)((double)((float)std::atan2((double)x, (double)y) * 180) / M_PI);
    int columnIdn = (int)(-std::round((horizonAngle - 90.0) / ang_res_x) + H / 2);
    if (columnIdn >= H) columnIdn -= H;
    if (columnIdn < 0 || columnIdn >= H) continue;
    if (rangeMat[(size_t)rowIdn * H + columnIdn] != FLT_MAX) continue;   // first hit wins (:499)
    rangeMat[(size_t)rowIdn * H + columnIdn] = r;
    srcMat[(size_t)rowIdn * H + columnIdn] = i;
  }
  int count = 0;
  for (int i = 0; i < NS; i++) {
    start_ring[i] = count - 1 + 5;
    for (int j = 0; j < H; j++)
      if (rangeMat[(size_t)i * H + j] != FLT_MAX) {
        col_ind[count] = j; range[count] = rangeMat[(size_t)i * H + j]; src_index[count] = srcMat[(size_t)i * H + j];
        ++count;
      }
    end_ring[i] = count - 1 - 5;
  }
  return count;
}

// F3-F5.  Index lists refer to the extracted cloud; order = reference push order.
// label_out (M, optional): 1 corner, -1 flat, 0 other.  curvature_out (M, optional).
void orc_extract_features(const float* range, const int32_t* col_ind, int32_t M,
                          const int32_t* start_ring, const int32_t* end_ring, const orc_feat_params* prm,
                          int32_t* corner_idx, int32_t* n_corner, int32_t* sharp_idx, int32_t* n_sharp,
                          int32_t* flat_idx, int32_t* n_flat, int32_t* surf_idx, int32_t* n_surf,
                          float* curvature_out, int32_t* label_out) {
  std::vector<float> curv(M, 0.f);
  std::vector<int> picked(M, 0), label(M, 0);
  struct Sm { float value; int ind; };
  std::vector<Sm> sm(M);
  for (int i = 0; i < M; i++) { sm[i].value = 0.f; sm[i].ind = i; }
  // calculateSmoothness (:544-563)
  for (int i = 5; i < M - 5; i++) {
    float d = range[i - 5] + range[i - 4] + range[i - 3] + range[i - 2] + range[i - 1] - range[i] * 10 +
              range[i + 1] + range[i + 2] + range[i + 3] + range[i + 4] + range[i + 5];
    curv[i] = d * d;
    sm[i].value = curv[i]; sm[i].ind = i;
  }
  // markOccludedPoints (:568-605)
  for (int i = 5; i < M - 6; ++i) {
    float depth1 = range[i], depth2 = range[i + 1];
    int columnDiff = std::abs(int(col_ind[i + 1] - col_ind[i]));
    if (columnDiff < 10) {
      if (depth1 - depth2 > 0.3) { for (int k = -5; k <= 0; k++) picked[i + k] = 1; }
      else if (depth2 - depth1 > 0.3) { for (int k = 1; k <= 6; k++) picked[i + k] = 1; }
    }
    float diff1 = std::abs(float(range[i - 1] - range[i]));
    float diff2 = std::abs(float(range[i + 1] - range[i]));
    if (diff1 > 0.02 * range[i] && diff2 > 0.02 * range[i]) picked[i] = 1;
  }
  // extractFeatures (:610-713)
  int nc = 0, nsh = 0, nfl = 0, nsf = 0;
  auto colAt = [&](int i, bool& ok) { ok = (i >= 0 && i < M); return ok ? col_ind[i] : 0; };
  auto mark_neighbours = [&](int ind) {
    for (int l = 1; l <= 5; l++) {
      bool a, b; int c1 = colAt(ind + l, a), c0 = colAt(ind + l - 1, b);
      if (!a || !b) break;
      if (std::abs(c1 - c0) > 10) break;
      picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
      bool a, b; int c1 = colAt(ind + l, a), c0 = colAt(ind + l + 1, b);
      if (!a || !b) break;
      if (std::abs(c1 - c0) > 10) break;
      picked[ind + l] = 1;
    }
  };
  for (int i = 0; i < prm->n_scan; i++) {
    for (int j = 0; j < 6; j++) {
      int sp = (start_ring[i] * (6 - j) + end_ring[i] * j) / 6;
      int ep = (start_ring[i] * (5 - j) + end_ring[i] * (j + 1)) / 6 - 1;
      if (sp >= ep) continue;
      std::sort(sm.begin() + sp, sm.begin() + ep, [](const Sm& a, const Sm& b) {
        return a.value < b.value || (a.value == b.value && a.ind < b.ind);
      });
      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = sm[k].ind;
        if (picked[ind] == 0 && curv[ind] > prm->edge_thr) {
          largestPickedNum++;
          if (largestPickedNum <= 20) {
            label[ind] = 1;
            corner_idx[nc++] = ind;
            if (largestPickedNum <= 4) sharp_idx[nsh++] = ind;
          } else break;
          picked[ind] = 1;
          mark_neighbours(ind);
        }
      }
      largestPickedNum = 0;
      for (int k = sp; k <= ep; k++) {
        int ind = sm[k].ind;
        if (picked[ind] == 0 && curv[ind] < prm->surf_thr) {
          largestPickedNum++;
          label[ind] = -1;
          picked[ind] = 1;
          if (largestPickedNum <= 10) flat_idx[nfl++] = ind;
          mark_neighbours(ind);
        }
      }
      for (int k = sp; k <= ep; k++)
        if (label[k] <= 0) surf_idx[nsf++] = k;
    }
  }
  *n_corner = nc; *n_sharp = nsh; *n_flat = nfl; *n_surf = nsf;
  if (curvature_out) memcpy(curvature_out, curv.data(), sizeof(float) * M);
  if (label_out) for (int i = 0; i < M; i++) label_out[i] = label[i];
}

}  // extern "C"
