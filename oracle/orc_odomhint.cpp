// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
// cloud_info hints of the odometry front end:
//   * transformUpdate (src/node/odomEstimationNode.cpp:976-1006): roll / pitch pulled toward the IMU attitude by a
//     quaternion slerp with weight imuRPYWeight, then the roll / pitch / z clamps (constraintTransformation);
//   the branches of updateInitialGuess (:297-419) are plain Affine3f products and live in the Python flow
//   (lis_slam_b200/stream.py OdometryStream._update_initial_guess) that tests drive with this oracle as its backend.
// tf (ROS geometry, LinearMath/Quaternion.h + Matrix3x3.h) is a third-party dependency that is absent from
// /root/reference; its published algorithm is restated here in double precision (tfScalar = double):
// setRPY, angleShortestPath, slerp, Matrix3x3::setRotation, getEulerYPR (solution 1).
#include <cmath>
#include <cstring>
#include "orc_api.h"

namespace {
struct Q { double x, y, z, w; };

Q set_rpy(double roll, double pitch, double yaw) {
  double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
  double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
  Q q;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  q.w = cr * cp * cy + sr * sp * sy;
  return q;
}
double dot(const Q& a, const Q& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
double angle_shortest_path(const Q& a, const Q& b) {
  double s = std::sqrt(dot(a, a) * dot(b, b));
  double d = dot(a, b);
  if (d < 0) return std::acos(-d / s) * 2.0;   // dot(-q) = -dot(q)
  return std::acos(d / s) * 2.0;
}
Q slerp(const Q& a, const Q& b, double t) {
  double theta = angle_shortest_path(a, b) / 2.0;
  if (theta != 0.0) {
    double d = 1.0 / std::sin(theta);
    double s0 = std::sin((1.0 - t) * theta);
    double s1 = std::sin(t * theta);
    Q r;
    if (dot(a, b) < 0) {
      r.x = (a.x * s0 + -b.x * s1) * d; r.y = (a.y * s0 + -b.y * s1) * d; r.z = (a.z * s0 + -b.z * s1) * d; r.w = (a.w * s0 + -b.w * s1) * d;
    } else {
      r.x = (a.x * s0 + b.x * s1) * d; r.y = (a.y * s0 + b.y * s1) * d; r.z = (a.z * s0 + b.z * s1) * d; r.w = (a.w * s0 + b.w * s1) * d;
    }
    return r;
  }
  return a;
}
// tf::Matrix3x3(q).getRPY(roll, pitch, yaw)
void get_rpy(const Q& q, double* roll, double* pitch, double* yaw) {
  double d = dot(q, q), s = 2.0 / d;
  double xs = q.x * s, ys = q.y * s, zs = q.z * s;
  double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
  double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
  double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
  double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  if (std::fabs(m20) >= 1) {
    *yaw = 0;
    double delta = std::atan2(m21, m22);
    if (m20 < 0) { *pitch = M_PI / 2.0; *roll = delta; }
    else { *pitch = -M_PI / 2.0; *roll = delta; }
  } else {
    *pitch = -std::asin(m20);
    double c = std::cos(*pitch);
    *roll = std::atan2(m21 / c, m22 / c);
    *yaw = std::atan2(m10 / c, m00 / c);
  }
}
float clampf(float v, float lim) {   // constraintTransformation (src/core/common.cpp:286-292); lim <= 0 = disabled (orc_lm_params)
  if (!(lim > 0.f)) return v;
  if (v < -lim) v = -lim;
  if (v > lim) v = lim;
  return v;
}
}  // namespace

extern "C" void orc_transform_update(float pose6[6], int32_t imu_available, float imu_roll, float imu_pitch, float imu_rpy_weight,
                                     float rot_tol, float z_tol) {
  if (imu_available) {
    if (std::abs(imu_pitch) < 1.4) {
      double w = imu_rpy_weight, r, p, y;
      Q tq = set_rpy(pose6[0], 0, 0), iq = set_rpy(imu_roll, 0, 0);
      get_rpy(slerp(tq, iq, w), &r, &p, &y);
      pose6[0] = (float)r;
      tq = set_rpy(0, pose6[1], 0); iq = set_rpy(0, imu_pitch, 0);
      get_rpy(slerp(tq, iq, w), &r, &p, &y);
      pose6[1] = (float)p;
    }
  }
  pose6[0] = clampf(pose6[0], rot_tol);
  pose6[1] = clampf(pose6[1], rot_tol);
  pose6[5] = clampf(pose6[5], z_tol);
}
