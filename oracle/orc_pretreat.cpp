// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the sweep pre-treatment in front of the feature extractor (SURVEY.md 8f "next" #3):
//   ring / time synthesis     src/node/laserPretreatmentNode.cpp:60-230 (dup src/core/laserPretreatment.cpp:20-160):
//                             removeNaNFromPointCloud, removeClosedPointCloud (:244-272), scanID from the elevation angle
//                             (N_SCAN 16 / 32 / 64, :95-126), relTime from the azimuth with the sequential halfPassed state
//                             (:128-141), point.time = scanPeriod (double 0.1, :14) * relTime
//   constant-velocity de-skew DistortionAdjust::AdjustCloud / UpdateMatrix src/core/distortionAdjust.cpp:419-479: every point
//                             but the first, real_time = time - scan_period / 2, R = (AngleAxis z * AngleAxis y * AngleAxis x)
//                             of angular_rate * real_time, p' = R p + velocity * real_time
// Third-party semantics restated (Eigen 3.3, fp32): AngleAxis -> Quaternion (half-angle sin / cos), quaternion product,
// Quaternion -> AngleAxis (2 atan2(|v|, |w|), axis = v / +-|v|), AngleAxis::toRotationMatrix (Rodrigues).  Eigen's vectorised
// operation order is unpinnable; the scalar formulas of its generic path are used, identically on the GPU (csrc/pretreat.cuh).
// atan / atan2 / sin / cos of floats follow the repo-wide resolution: correctly rounded float of the double routine.
#include "orc_api.h"
#include <cmath>
#include <cstring>
#include <vector>

namespace {
const double PI = 3.14159265358979323846;
inline float f_atan2(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float f_atan(float a) { return (float)std::atan((double)a); }
inline float f_sin(float a) { return (float)std::sin((double)a); }
inline float f_cos(float a) { return (float)std::cos((double)a); }

int scan_id(float x, float y, float z, int n_scan) {
  const float angle = (float)((double)f_atan(z / std::sqrt(x * x + y * y)) * 180 / PI);
  int id;
  if (n_scan == 16) { id = (int)((double)((angle + 15) / 2) + 0.5); if (id > n_scan - 1 || id < 0) return -1; return id; }
  if (n_scan == 32) { id = (int)(((double)angle + 92.0 / 3.0) * 3.0 / 4.0); if (id > n_scan - 1 || id < 0) return -1; return id; }
  if (n_scan == 64) {
    if ((double)angle >= -8.83) id = (int)((double)(2 - angle) * 3.0 + 0.5);
    else id = n_scan / 2 + (int)((-8.83 - (double)angle) * 2.0 + 0.5);
    if ((double)angle > 2 || (double)angle < -24.33 || id > 50 || id < 0) return -1;
    return id;
  }
  return -1;
}

struct Quat { float w, x, y, z; };
inline Quat q_axis(float angle, int axis) {          // Quaternionf(AngleAxisf(angle, Unit axis))
  const float ha = 0.5f * angle;
  Quat q{f_cos(ha), 0.f, 0.f, 0.f};
  const float s = f_sin(ha);
  (axis == 0 ? q.x : axis == 1 ? q.y : q.z) = s * 1.f;
  return q;
}
inline Quat q_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
void update_matrix(const float w[3], float real_time, float R[9]) {      // DistortionAdjust::UpdateMatrix
  const float ax = w[0] * real_time, ay = w[1] * real_time, az = w[2] * real_time;
  const Quat q = q_mul(q_mul(q_axis(az, 2), q_axis(ay, 1)), q_axis(ax, 0));
  float n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  float angle, ux, uy, uz;
  if (n != 0.f) {
    angle = 2.f * f_atan2(n, std::fabs(q.w));
    if (q.w < 0.f) n = -n;
    ux = q.x / n; uy = q.y / n; uz = q.z / n;
  } else { angle = 0.f; ux = 1.f; uy = 0.f; uz = 0.f; }
  const float s = f_sin(angle), c = f_cos(angle);
  const float sx = s * ux, sy = s * uy, sz = s * uz;
  const float cx = (1.f - c) * ux, cy = (1.f - c) * uy, cz = (1.f - c) * uz;
  float t;
  t = cx * uy; R[1] = t - sz; R[3] = t + sz;
  t = cx * uz; R[2] = t + sy; R[6] = t - sy;
  t = cy * uz; R[5] = t - sx; R[7] = t + sx;
  R[0] = cx * ux + c; R[4] = cy * uy + c; R[8] = cz * uz + c;
}
}  // namespace

extern "C" {

// laserCloudInfoHandler pre-treatment.  Returns the number of output points; out4 / ring_out / time_out have capacity n.
int32_t orc_pretreat(const float* pts4, int32_t n, int32_t n_scan, double scan_period, float min_range, float max_range,
                     float* out4, uint16_t* ring_out, float* time_out) {
  std::vector<int> kept;
  for (int i = 0; i < n; i++) {
    const float x = pts4[4 * (size_t)i], y = pts4[4 * (size_t)i + 1], z = pts4[4 * (size_t)i + 2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;                 // removeNaNFromPointCloud
    const float r2 = x * x + y * y + z * z;
    if (r2 < min_range * min_range) continue;
    if (r2 > max_range * max_range) continue;
    kept.push_back(i);
  }
  const int m = (int)kept.size();
  if (m == 0) return 0;
  const float* p0 = pts4 + 4 * (size_t)kept[0]; const float* pl = pts4 + 4 * (size_t)kept[m - 1];
  const float startOri = -f_atan2(p0[1], p0[0]);
  float endOri = (float)((double)-f_atan2(pl[1], pl[0]) + 2 * PI);
  if ((double)(endOri - startOri) > 3 * PI) endOri = (float)((double)endOri - 2 * PI);
  else if ((double)(endOri - startOri) < PI) endOri = (float)((double)endOri + 2 * PI);
  bool halfPassed = false;
  int o = 0;
  for (int k = 0; k < m; k++) {
    const float* p = pts4 + 4 * (size_t)kept[k];
    const int id = scan_id(p[0], p[1], p[2], n_scan);
    if (id < 0) continue;
    float ori = -f_atan2(p[1], p[0]);
    if (!halfPassed) {
      if ((double)ori < (double)startOri - PI / 2) ori = (float)((double)ori + 2 * PI);
      else if ((double)ori > (double)startOri + PI * 3 / 2) ori = (float)((double)ori - 2 * PI);
      if ((double)(ori - startOri) > PI) halfPassed = true;
    } else {
      ori = (float)((double)ori + 2 * PI);
      if ((double)ori < (double)endOri - PI * 3 / 2) ori = (float)((double)ori + 2 * PI);
      else if ((double)ori > (double)endOri + PI / 2) ori = (float)((double)ori - 2 * PI);
    }
    const float relTime = (ori - startOri) / (endOri - startOri);
    memcpy(out4 + 4 * (size_t)o, p, 16);
    ring_out[o] = (uint16_t)id;
    time_out[o] = (float)(scan_period * (double)relTime);
    o++;
  }
  return o;
}

// DistortionAdjust::AdjustCloud: n - 1 output points (the first input point is skipped, distortionAdjust.cpp:436)
int32_t orc_deskew_cv(const float* pts4, const float* time, int32_t n, float scan_period, const float* lin_vel3, const float* ang_vel3, float* out4) {
  int o = 0;
  for (int i = 1; i < n; i++) {
    const float* p = pts4 + 4 * (size_t)i;
    const float real_time = (float)((double)time[i] - (double)scan_period / 2.0);
    float R[9]; update_matrix(ang_vel3, real_time, R);
    float* q = out4 + 4 * (size_t)o;
    for (int r = 0; r < 3; r++) {
      const float rot = (R[3 * r] * p[0] + R[3 * r + 1] * p[1]) + R[3 * r + 2] * p[2];
      q[r] = rot + lin_vel3[r] * real_time;
    }
    q[3] = p[3];
    o++;
  }
  return o;
}

}  // extern "C"
