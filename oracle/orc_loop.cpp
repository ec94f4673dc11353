// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the EPSC loop detector (rows F17 / F18 of SURVEY.md §8a):
//   EPSCGeneration::project        src/core/epscGeneration.cpp:84-120   360-sector {count, last x, last y, last label}
//   EPSCGeneration::globalICP      :258-401 (the Affine3f overload)     sector-count shift search + 2-D PCL ICP
//   EPSCGeneration::loopDetection  :663-992                             travel-distance gate, per-candidate re-description
//   constants                      src/include/epscGeneration.h:9-16 (SKIP_NEIBOUR_DISTANCE 20, INFLATION_COVARIANCE 0.01,
//                                  DISTANCE_THRESHOLD 0.75); descriptor flags config/params.yaml:22-28 (FEPSC only by default)
// Descriptor kinds restated: EPSC, SEPSC, FEPSC and the pose fallback (UsingPoseFlag); ISC / SC / SSC are not part of the
// hot path named by BASELINE.json and are not restated.
//
// Third-party semantics (absent from /root/reference, restated from the published algorithm):
//   * pcl::IterativeClosestPoint with DEFAULT parameters (epscGeneration.cpp:321-325): max_iterations 10,
//     transformation_epsilon 0, euclidean_fitness_epsilon -DBL_MAX, max correspondence distance sqrt(DBL_MAX)
//     -> orc_icp() with {1e18, 10, 0, -DBL_MAX} (1-NN over everything, Umeyama fit, stop at 10 iterations or
//     |mse - prev| < 1e-12).
//   * pcl::getTranslationAndEulerAngles: roll = atan2(R21, R22), pitch = asin(-R20), yaw = atan2(R10, R00).
//   * pcl::transformPointCloud(Affine3f): ((R00 x + R01 y) + R02 z) + tx, fp32, left to right.
//   * Eigen::AngleAxisf(a, UnitZ).toRotationMatrix(): [[c, -s, 0], [s, c, 0], [0, 0, (1 - c) + c]].
// Resolutions (identical here and on the GPU): sin / cos / atan2 / asin of a float are the correctly rounded float of
// the double routine (DESIGN.md numerics); Affine products accumulate k = 0..3 left to right in fp32.
// Quirk Q8 (reference UB, NOT reproducible): globalICP reads column j + i - 360 without a second wrap, which indexes
// past the 360-column row when yaw_diff >= 332 deg (:279-282); the restatement wraps modulo 360.
#include "orc_api.h"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

const int SECT = 360;
const double PI = 3.14159265358979323846;

inline float f_sin(float a) { return (float)std::sin((double)a); }
inline float f_cos(float a) { return (float)std::cos((double)a); }
inline float f_atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float f_asin(float a) { return (float)std::asin((double)a); }

struct Proj { float v[SECT][4]; };   // count, last x, last y, last label

struct Loop {
  uint8_t using_map[256];
  int use_epsc, use_sepsc, use_fepsc, use_pose;
  std::vector<Proj> proj;
  std::vector<std::vector<uint8_t>> epsc, sepsc, fepsc;
  std::vector<double> travel;
  std::vector<double> px, py;
  std::vector<float> yaw;
};

// EPSCGeneration::project (:84-120)
void project(const float* sem4, const uint16_t* label, int n, Proj& out) {
  memset(&out, 0, sizeof(out));
  const float step = (float)(2. * PI / 360.f);
  for (int i = 0; i < n; i++) {
    const unsigned l = label[i];
    if (!(l == 13 || l == 14 || l == 16 || l == 18 || l == 19)) continue;
    const float x = sem4[4 * (size_t)i], y = sem4[4 * (size_t)i + 1];
    const float distance = std::sqrt(x * x + y * y);
    if ((double)distance < 1e-2) continue;
    const float angle = (float)(PI + (double)f_atan2(y, x));
    const int sector_id = (int)std::floor(angle / step);
    if (sector_id >= SECT || sector_id < 0) continue;
    out.v[sector_id][0] += 1.f;
    out.v[sector_id][1] = x;
    out.v[sector_id][2] = y;
    out.v[sector_id][3] = (float)l;
  }
}

void matmul4(const float A[16], const float B[16], float C[16]) {
  float r[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; r[i * 4 + j] = s; }
  memcpy(C, r, sizeof(r));
}

void rot_z(float angle, float T[16]) {   // Identity.rotate(AngleAxisf(angle, UnitZ))
  const float c = f_cos(angle), s = f_sin(angle);
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.f : 0.f;
  T[0] = c; T[1] = -s; T[4] = s; T[5] = c; T[10] = (1.f - c) + c;
}

// EPSCGeneration::globalICP(ssc_dis1, ssc_dis2, yaw_diff) (:258-401) -> trans * trans1 (row-major 4x4)
void global_icp(const Proj& d1, const Proj& d2, float yaw_diff, float T[16]) {
  const float step = (float)(2. * PI / 360.f);
  double similarity = 100000;
  float angle = yaw_diff;
  if ((double)angle >= 2. * PI) angle = (float)((double)angle - 2. * PI);
  if (angle < 0) angle = (float)((double)angle + 2. * PI);
  const int tmp_id = (int)std::floor(angle / step);
  for (int i = tmp_id - 30; i < tmp_id + 30; ++i) {
    float dis_count = 0;
    for (int j = 0; j < SECT; ++j) {
      int new_col = ((j + i) % SECT + SECT) % SECT;   // Q8: reference wraps once only
      dis_count += std::fabs(d1.v[j][0] - d2.v[new_col][0]);
    }
    if ((double)dis_count < similarity) { similarity = dis_count; angle = (float)i; }
  }
  angle = angle * step;
  const float cs = f_cos(angle), sn = f_sin(angle);
  std::vector<float> c1, c2;
  for (int i = 0; i < SECT; ++i) {
    if (d1.v[i][3] > 0) { c1.push_back(d1.v[i][1]); c1.push_back(d1.v[i][2]); c1.push_back(0.f); c1.push_back(0.f); }
    if (d2.v[i][3] > 0) {
      const float tpx = d2.v[i][1] * cs - d2.v[i][2] * sn;
      const float tpy = d2.v[i][1] * sn + d2.v[i][2] * cs;
      c2.push_back(tpx); c2.push_back(tpy); c2.push_back(0.f); c2.push_back(0.f);
    }
  }
  orc_icp_params ip; ip.max_corr_dist = 1e18f; ip.max_iters = 10; ip.trans_eps = 0.0; ip.fitness_eps = -DBL_MAX;
  orc_icp_result ir;
  orc_icp(c2.data(), (int)(c2.size() / 4), c1.data(), (int)(c1.size() / 4), &ip, &ir);   // source = cloud2, target = cloud1
  float T1[16];
  rot_z(angle, T1);
  matmul4(ir.T, T1, T);
}

inline void euler_of(const float T[16], float& x, float& y, float& yaw) { x = T[3]; y = T[7]; yaw = f_atan2(T[4], T[0]); }

void transform_cloud(const float* in4, int n, const float T[16], std::vector<float>& out) {
  out.resize(4 * (size_t)n);
  for (int i = 0; i < n; i++) {
    const float* p = in4 + 4 * (size_t)i;
    out[4 * (size_t)i + 0] = ((T[0] * p[0] + T[1] * p[1]) + T[2] * p[2]) + T[3];
    out[4 * (size_t)i + 1] = ((T[4] * p[0] + T[5] * p[1]) + T[6] * p[2]) + T[7];
    out[4 * (size_t)i + 2] = ((T[8] * p[0] + T[9] * p[1]) + T[10] * p[2]) + T[11];
    out[4 * (size_t)i + 3] = p[3];
  }
}

void planar_transform(float dx, float dy, float angle, float T[16]) {   // translation << dx, dy, 0 ; rotate(AngleAxisf(angle, Z))
  rot_z(angle, T);
  T[3] = dx; T[7] = dy; T[11] = 0.f;
}

}  // namespace

extern "C" {

void* orc_loop_create(const uint8_t* using_map, int32_t use_epsc, int32_t use_sepsc, int32_t use_fepsc, int32_t use_pose) {
  Loop* L = new Loop();
  memcpy(L->using_map, using_map, 256);
  L->use_epsc = use_epsc; L->use_sepsc = use_sepsc; L->use_fepsc = use_fepsc; L->use_pose = use_pose;
  return L;
}
void orc_loop_free(void* h) { delete (Loop*)h; }

// EPSCGeneration::loopDetection (:663-992).  odom = row-major 4x4.  Outputs (capacity 4 each): kind (0 EPSC, 1 SEPSC,
// 2 FEPSC, 3 POSE) in the reference push order, matched frame id, score (pose kind: the position distance) and the 4x4
// matched_frame_transform.  Returns the number of matches; *current_id = current_frame_id, *n_cand = gated candidates.
int32_t orc_loop_detect(void* h, const float* corner4, int32_t nc, const float* surf4, int32_t ns,
                        const float* sem4, const uint16_t* sem_label, int32_t nsem, const float* odom,
                        int32_t* current_id, int32_t* n_cand, int32_t* kinds, int32_t* ids, double* scores, float* T16s) {
  Loop& L = *(Loop*)h;
  const float x_t = odom[3], y_t = odom[7];
  const float yaw_t = f_atan2(odom[4], odom[0]);
  const double cx = x_t, cy = y_t;
  if (L.travel.empty()) L.travel.push_back(0);
  else {
    const double ex = L.px.back() - cx, ey = L.py.back() - cy;
    L.travel.push_back(L.travel.back() + std::sqrt(ex * ex + ey * ey + 0.0));
  }
  const int cur = (int)L.px.size();
  *current_id = cur;
  int best_id[3] = {-1, -1, -1};
  double best_score[3] = {0, 0, 0};
  float best_T[3][16];
  double min_distance = 1000000; int best_pose = -1; float best_pose_T[16];
  Proj cur_dis;
  project(sem4, sem_label, nsem, cur_dis);
  int ncand = 0;
  std::vector<float> tc, ts, tsem;
  std::vector<uint8_t> e(1600), se(1600), fe(1600);
  for (int i = 0; i < cur; i++) {
    const double delta_travel = L.travel.back() - L.travel[i];
    // posArr.back() is the PREVIOUS frame (the current pose is pushed after the loop, :899)
    const double qx = L.px[i] - L.px.back(), qy = L.py[i] - L.py.back();
    const double pos_distance = std::sqrt(qx * qx + qy * qy + 0.0);
    if (!(delta_travel > 20.0 && pos_distance < delta_travel * 0.01)) continue;
    ncand++;
    const float yaw_diff = yaw_t - L.yaw[i];
    float T[16];
    global_icp(L.proj[i], cur_dis, yaw_diff, T);
    float diff_x, diff_y, angle;
    euler_of(T, diff_x, diff_y, angle);
    transform_cloud(sem4, nsem, T, tsem);
    transform_cloud(corner4, nc, T, tc);
    transform_cloud(surf4, ns, T, ts);
    orc_epsc_describe(tc.data(), nc, ts.data(), ns, tsem.data(), sem_label, nsem, L.using_map, e.data(), se.data(), fe.data());
    const double sector_step = 2 * PI / 80;
    for (int kind = 0; kind < 3; kind++) {
      if (!(kind == 0 ? L.use_epsc : kind == 1 ? L.use_sepsc : L.use_fepsc)) continue;
      const std::vector<uint8_t>& hist = kind == 0 ? L.epsc[i] : kind == 1 ? L.sepsc[i] : L.fepsc[i];
      const std::vector<uint8_t>& now = kind == 0 ? e : kind == 1 ? se : fe;
      int32_t shift, sad;
      const double score = orc_epsc_distance(hist.data(), now.data(), &shift, &sad);
      if (score > 0.75 && score > best_score[kind]) {
        best_score[kind] = score; best_id[kind] = i;
        // EPSC rotates by yaw_diff + shift (:816-829), SEPSC by the ICP yaw + shift (:837-850), FEPSC by the ICP yaw (:857-869)
        double a = kind == 0 ? (double)yaw_diff : (double)angle;
        if (kind != 2 && sad >= 0) a = a + shift * sector_step;
        planar_transform(diff_x, diff_y, (float)a, best_T[kind]);
      }
    }
    if (L.use_pose && pos_distance < min_distance) { min_distance = pos_distance; best_pose = i; memcpy(best_pose_T, T, sizeof(T)); }
  }
  *n_cand = ncand;
  L.px.push_back(cx); L.py.push_back(cy); L.yaw.push_back(yaw_t); L.proj.push_back(cur_dis);
  orc_epsc_describe(corner4, nc, surf4, ns, sem4, sem_label, nsem, L.using_map, e.data(), se.data(), fe.data());
  L.epsc.push_back(e); L.sepsc.push_back(se); L.fepsc.push_back(fe);
  int m = 0;
  for (int kind = 0; kind < 3; kind++) {
    if (!(kind == 0 ? L.use_epsc : kind == 1 ? L.use_sepsc : L.use_fepsc) || best_id[kind] < 0) continue;
    kinds[m] = kind; ids[m] = best_id[kind]; scores[m] = best_score[kind]; memcpy(T16s + 16 * m, best_T[kind], sizeof(float) * 16); m++;
  }
  if (L.use_pose && best_pose >= 0) { kinds[m] = 3; ids[m] = best_pose; scores[m] = min_distance; memcpy(T16s + 16 * m, best_pose_T, sizeof(float) * 16); m++; }
  return m;
}

// project() / globalICP() exposed for unit tests
void orc_loop_project(const float* sem4, const uint16_t* label, int32_t n, float* out1440) {
  Proj p; project(sem4, label, n, p); memcpy(out1440, p.v, sizeof(p.v));
}
void orc_loop_global_icp(const float* proj1, const float* proj2, float yaw_diff, float* T16) {
  Proj a, b; memcpy(a.v, proj1, sizeof(a.v)); memcpy(b.v, proj2, sizeof(b.v));
  global_icp(a, b, yaw_diff, T16);
}

}  // extern "C"
