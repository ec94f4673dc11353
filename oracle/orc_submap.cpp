// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the local-map / submap assembly around scan-to-map variants B / C (SURVEY.md 8f "next" #2, row T4):
//   localMap_t / submap_t class clouds      src/include/subMap.h:435-777 (submap_dynamic, _pole, _ground, _building, _outlier)
//   SubMapManager::insert_local_map         subMap.h:957-1055   transformPointCloud of the key frame's class clouds by its pose,
//                                                              map-based dynamic removal of the DYNAMIC class only (:1001-1017),
//                                                              append_feature (:742-753), bounding box of all five clouds (:1046-1050)
//   CloudUtility::get_cloud_bbx_cpt, transform_bbx, get_intersection_bbx   subMap.h:131-228
//   SubMapOdometryNode::extractSlidingCloud src/node/subMapOptmizationNode.cpp:1369-1432: sensor box (+-70, +-70, -10..20) moved by
//                                                              the current pose, intersected with the map box (pad 2), every class
//                                                              voxel-filtered IN PLACE (0.1 / 0.05 / 0.4 / 0.2 / 0.6), box-filtered IN
//                                                              PLACE (strict inequalities, subMap.h:1125-1150), corner map = pole,
//                                                              surf map = ground + building + dynamic
// Map-side labels are not restated: F9-F14 read the label of the QUERY point only (subMapOptmizationNode.cpp:1671, :1795).
#include "orc_api.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
struct Submap { std::vector<float> cls[5]; };
inline int cnt(const Submap& S, int c) { return (int)(S.cls[c].size() / 4); }
}  // namespace

extern "C" {

void* orc_submap_create() { return new Submap(); }
void orc_submap_free(void* h) { delete (Submap*)h; }
void orc_submap_clear(void* h) { for (auto& c : ((Submap*)h)->cls) c.clear(); }

// insert_local_map.  pts[c] = xyzi of class c (0 dynamic, 1 pole, 2 ground, 3 building, 4 outlier), pose6 = [roll, pitch, yaw, x, y, z].
// Outputs: counts[5] after the append, bound[6] = min xyz, max xyz of all five clouds.
void orc_submap_insert(void* h, const float* const* pts, const int32_t* n, const float* pose6, int32_t dynrem_on, int32_t max_num_pts,
                       float center_radius, float dist_min, float dist_max, float near_dist, int32_t* counts, double* bound) {
  Submap& S = *(Submap*)h;
  float T[12]; orc_pose_to_affine(pose6, T);
  int feature_point_num = 0;
  for (int c = 0; c < 5; c++) feature_point_num += cnt(S, c);
  dist_max = std::max(dist_max, (float)((double)dist_min + 0.1));                     // :979
  for (int c = 0; c < 5; c++) {
    std::vector<float> moved(4 * (size_t)n[c]);
    for (int i = 0; i < n[c]; i++) {                                                    // transformPointCloud (common.cpp:134-160)
      const float* p = pts[c] + 4 * (size_t)i; float* q = &moved[4 * (size_t)i];
      q[0] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3];
      q[1] = T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7];
      q[2] = T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11];
      q[3] = p[3];
    }
    if (c == 0 && dynrem_on && feature_point_num > max_num_pts / 5 && n[c] > 0) {       // :980-985: the dynamic class against the map's dynamic cloud
      std::vector<uint8_t> keep((size_t)n[c]);
      orc_map_distance_filter(moved.data(), n[c], S.cls[0].data(), cnt(S, 0), center_radius, dist_min, dist_max, near_dist, keep.data());
      std::vector<float> kept;
      for (int i = 0; i < n[c]; i++) if (keep[i]) kept.insert(kept.end(), &moved[4 * (size_t)i], &moved[4 * (size_t)i] + 4);
      moved.swap(kept);
    }
    S.cls[c].insert(S.cls[c].end(), moved.begin(), moved.end());
  }
  double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};     // get_cloud_bbx over merge_feature_points
  for (int c = 0; c < 5; c++) {
    counts[c] = cnt(S, c);
    for (int i = 0; i < counts[c]; i++) for (int d = 0; d < 3; d++) {
      const double v = S.cls[c][4 * (size_t)i + d];
      if (mn[d] > v) mn[d] = v;
      if (mx[d] < v) mx[d] = v;
    }
  }
  for (int d = 0; d < 3; d++) { bound[d] = mn[d]; bound[3 + d] = mx[d]; }
}

// extractSlidingCloud.  map_bound = localMap->bound (from the last insert).  Returns the corner / surf registration map
// (capacity = current point counts) and leaves the voxel- and box-filtered class clouds in the submap.
void orc_submap_extract(void* h, const float* cur_pose6, const float* leaf5, const double* map_bound, float* corner_out, int32_t* nc,
                        float* surf_out, int32_t* ns, int32_t* counts) {
  Submap& S = *(Submap*)h;
  float T[12]; orc_pose_to_affine(cur_pose6, T);
  // cur_bbx (+-70, +-70, -10 .. 20), its centre, transform_bbx (float matrix entries times double coordinates)
  const double bmin[3] = {-70.0, -70.0, -10.0}, bmax[3] = {70.0, 70.0, 20.0};
  double cp[3], cpo[3];
  for (int d = 0; d < 3; d++) cp[d] = 0.5 * (bmin[d] + bmax[d]);
  for (int d = 0; d < 3; d++) cpo[d] = (double)T[4 * d] * cp[0] + (double)T[4 * d + 1] * cp[1] + (double)T[4 * d + 2] * cp[2] + (double)T[4 * d + 3];
  double lo[3], hi[3];
  for (int d = 0; d < 3; d++) { hi[d] = bmax[d] - cp[d] + cpo[d]; lo[d] = bmin[d] - cp[d] + cpo[d]; }
  const float pad = 2.0f;
  for (int d = 0; d < 3; d++) { lo[d] = std::max(lo[d], map_bound[d]) - pad; hi[d] = std::min(hi[d], map_bound[3 + d]) + pad; }   // get_intersection_bbx
  for (int c = 0; c < 5; c++) {
    const int n = cnt(S, c);
    if (n > 0) {                                                                         // voxel_downsample_pcl: empty clouds are left alone
      std::vector<float> out(4 * (size_t)n);
      const int m = orc_voxel_grid(S.cls[c].data(), n, leaf5[c], out.data(), n);
      out.resize(4 * (size_t)m); S.cls[c].swap(out);
    }
    std::vector<float> kept;                                                             // bbx_filter (strict inequalities)
    for (int i = 0; i < cnt(S, c); i++) {
      const float* p = &S.cls[c][4 * (size_t)i];
      if ((double)p[0] > lo[0] && (double)p[0] < hi[0] && (double)p[1] > lo[1] && (double)p[1] < hi[1] && (double)p[2] > lo[2] && (double)p[2] < hi[2])
        kept.insert(kept.end(), p, p + 4);
    }
    S.cls[c].swap(kept);
    counts[c] = cnt(S, c);
  }
  *nc = cnt(S, 1);
  if (*nc) memcpy(corner_out, S.cls[1].data(), sizeof(float) * S.cls[1].size());
  size_t o = 0;
  for (int c : {2, 3, 0}) { if (!S.cls[c].empty()) memcpy(surf_out + o, S.cls[c].data(), sizeof(float) * S.cls[c].size()); o += S.cls[c].size(); }
  *ns = (int32_t)(o / 4);
}

// SubMapOdometryNode::detectLoopClosureForSubMap (subMapOptmizationNode.cpp:2739-2916).  Pose arithmetic: Eigen::Affine3f
// products / inverse in fp32, natural order (same resolution as the streaming odometry flow, lis_slam_b200/stream.py).
namespace {
void T16_of(const float* pose6, float* T) { float t12[12]; orc_pose_to_affine(pose6, t12); memcpy(T, t12, sizeof(t12)); T[12] = T[13] = T[14] = 0.f; T[15] = 1.f; }
float cof(const float* T, int i, int j) { const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3; return T[i1 * 4 + j1] * T[i2 * 4 + j2] - T[i1 * 4 + j2] * T[i2 * 4 + j1]; }
void inv16(const float* T, float* R) {
  const float c0 = cof(T, 0, 0), c1 = cof(T, 1, 0), c2 = cof(T, 2, 0);
  const float det = (c0 * T[0] + c1 * T[4]) + c2 * T[8];
  const float invdet = 1.f / det;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i * 4 + j] = cof(T, j, i) * invdet;
  for (int i = 0; i < 3; i++) R[i * 4 + 3] = ((-R[i * 4 + 0]) * T[3] + (-R[i * 4 + 1]) * T[7]) + (-R[i * 4 + 2]) * T[11];
  R[12] = R[13] = R[14] = 0.f; R[15] = 1.f;
}
void mul16(const float* A, const float* B, float* Cm) {
  float r[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float a = 0.f; for (int k = 0; k < 4; k++) a += A[i * 4 + k] * B[k * 4 + j]; r[i * 4 + j] = a; }
  memcpy(Cm, r, sizeof(r));
}
}  // namespace

// cand_pose: per candidate [use_epsc, prekey_pose6 (6), epsc_T (16), submap_pose6 (6)] = 29 floats; submaps: P handles.
// out: found, best, best_score, correction[16], key2pre[16], t_correct[16], constraint6; per-candidate fitness / converged / iters.
int32_t orc_loop_verify(const float* key4, int32_t n, const float* key_pose6, const float* key_rel_pose6, int32_t P, void* const* submaps,
                        const float* cand_pose, float fitness_threshold, const orc_icp_params* prm, int32_t* best, double* best_score,
                        float* correction16, float* key2pre16, float* t_correct16, float* constraint6, double* fitness_out, int32_t* conv_out) {
  *best = -1; *best_score = DBL_MAX;
  std::vector<float> K((size_t)std::max(P, 1) * 16);
  std::vector<orc_icp_result> res((size_t)std::max(P, 1));
  for (int i = 0; i < P; i++) {
    const float* c = cand_pose + 29 * (size_t)i;
    float* T = &K[16 * (size_t)i];
    if (c[0] != 0.f) { float A[16]; T16_of(c + 1, A); mul16(A, c + 7, T); }
    else { float A[16], Ai[16], B[16]; T16_of(c + 23, A); T16_of(key_pose6, B); inv16(A, Ai); mul16(Ai, B, T); }
    const Submap& S = *(const Submap*)submaps[i];
    std::vector<float> tgt;
    for (int cl = 0; cl < 4; cl++) tgt.insert(tgt.end(), S.cls[cl].begin(), S.cls[cl].end());
    std::vector<float> src(4 * (size_t)n);
    for (int k = 0; k < n; k++) {
      const float* p = key4 + 4 * (size_t)k; float* q = &src[4 * (size_t)k];
      q[0] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3];
      q[1] = T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7];
      q[2] = T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11];
      q[3] = p[3];
    }
    orc_icp(src.data(), n, tgt.data(), (int)(tgt.size() / 4), prm, &res[i]);
    fitness_out[i] = res[i].fitness; conv_out[i] = res[i].converged;
    if (!res[i].converged || res[i].fitness > *best_score) continue;
    *best_score = res[i].fitness; *best = i;
  }
  if (*best < 0) return 0;
  memcpy(correction16, res[*best].T, sizeof(float) * 16);
  memcpy(key2pre16, &K[16 * (size_t)*best], sizeof(float) * 16);
  if (*best_score > (double)fitness_threshold) return 0;
  float R[16], Ri[16], M[16];
  T16_of(key_rel_pose6, R); inv16(R, Ri);
  mul16(correction16, key2pre16, M); mul16(M, Ri, t_correct16);
  constraint6[0] = t_correct16[3]; constraint6[1] = t_correct16[7]; constraint6[2] = t_correct16[11];
  constraint6[3] = (float)std::atan2((double)t_correct16[9], (double)t_correct16[10]);
  constraint6[4] = (float)std::asin((double)-t_correct16[8]);
  constraint6[5] = (float)std::atan2((double)t_correct16[4], (double)t_correct16[0]);
  return 1;
}

int32_t orc_submap_get(void* h, int32_t c, float* out, int32_t cap) {
  Submap& S = *(Submap*)h;
  const int n = std::min(cnt(S, c), cap);
  if (n) memcpy(out, S.cls[c].data(), sizeof(float) * 4 * (size_t)n);
  return cnt(S, c);
}

}  // extern "C"
