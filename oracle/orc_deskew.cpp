// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the per-point motion de-skew of the range-image projection (SURVEY.md 8f "next" #3):
//   LaserProcessing::findRotation   src/core/laserProcessing.cpp:368-400   IMU rotation table lookup / interpolation
//   LaserProcessing::findPosition   :405-422                               always zero (positional de-skew is commented out)
//   LaserProcessing::deskewPoint    :427-462                               p' = (T_first^-1 * T(p.time)) p
//   call site                       :501 (inside projectPointCloud: range and column come from the ORIGINAL point,
//                                   only the stored coordinates are de-skewed; the first point that passes the
//                                   filters and owns a cell fixes transStartInverse, :439-443)
// Inputs: the IMU rotation table imuTime / imuRotX,Y,Z that imuDeskewInfo integrates (:213-262; ROS glue, out of
// scope), timeScanCur and the per-point relative time (PointXYZIRT::time).
// Third-party semantics restated (absent from /root/reference): pcl::getTransformation(0,0,0,rx,ry,rz) = the fp32
// closed form of Rz(rz) Ry(ry) Rx(rx); Eigen::Affine3f::inverse() = 3x3 cofactor inverse (adjugate / det, det = (c00
// m00 + c10 m10) + c20 m20); Affine * Affine = coefficient-wise 3x3 product, left to right.  sin / cos of a float are
// the correctly rounded float of the double routine (DESIGN.md numerics).
#include "orc_api.h"
#include <cmath>
#include <cstring>

namespace {

inline float f_sin(float a) { return (float)std::sin((double)a); }
inline float f_cos(float a) { return (float)std::cos((double)a); }

void rot_of(float roll, float pitch, float yaw, float R[9]) {   // pcl::getTransformation, linear part
  const float A = f_cos(yaw), B = f_sin(yaw), C = f_cos(pitch), D = f_sin(pitch), E = f_cos(roll), F = f_sin(roll);
  const float DE = D * E, DF = D * F;
  R[0] = A * C; R[1] = A * DF - B * E; R[2] = B * F + A * DE;
  R[3] = B * C; R[4] = A * E + B * DF; R[5] = B * DE - A * F;
  R[6] = -D;    R[7] = C * F;          R[8] = C * E;
}

inline float cof(const float m[9], int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}

void inverse3(const float m[9], float inv[9]) {
  const float c0 = cof(m, 0, 0), c1 = cof(m, 1, 0), c2 = cof(m, 2, 0);
  const float det = (c0 * m[0] + c1 * m[3]) + c2 * m[6];
  const float invdet = 1.f / det;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) inv[i * 3 + j] = cof(m, j, i) * invdet;
}

// findRotation (:368-400)
void find_rotation(double pointTime, const double* imuTime, const double* imuRot, int imuPointerCur, float r[3]) {
  int front = 0;
  while (front < imuPointerCur) { if (pointTime < imuTime[front]) break; ++front; }
  if (pointTime > imuTime[front] || front == 0) {
    for (int a = 0; a < 3; a++) r[a] = (float)imuRot[3 * front + a];
  } else {
    const int back = front - 1;
    const double ratioFront = (pointTime - imuTime[back]) / (imuTime[front] - imuTime[back]);
    const double ratioBack = (imuTime[front] - pointTime) / (imuTime[front] - imuTime[back]);
    for (int a = 0; a < 3; a++) r[a] = (float)(imuRot[3 * front + a] * ratioFront + imuRot[3 * back + a] * ratioBack);
  }
}

}  // namespace

extern "C" {

// De-skews the M extracted points (src_index = owner of every extracted slot, from orc_project_scan).  n_imu =
// imuPointerCur + 1 table entries; n_imu <= 0 = de-skew disabled (points pass through).  out4: M x float4.
void orc_deskew(const float* pts4, const float* time, const int32_t* src_index, int32_t M,
                const double* imu_time, const double* imu_rot3, int32_t n_imu, double time_scan_cur, float* out4) {
  if (n_imu <= 0) { for (int i = 0; i < M; i++) memcpy(out4 + 4 * (size_t)i, pts4 + 4 * (size_t)src_index[i], 16); return; }
  int first = -1;
  for (int i = 0; i < M; i++) if (first < 0 || src_index[i] < first) first = src_index[i];   // first point processed (:439)
  float r[3], Rs[9], Sinv[9];
  find_rotation(time_scan_cur + (double)time[first], imu_time, imu_rot3, n_imu - 1, r);
  rot_of(r[0], r[1], r[2], Rs);
  inverse3(Rs, Sinv);
  for (int i = 0; i < M; i++) {
    const int s = src_index[i];
    const float* p = pts4 + 4 * (size_t)s;
    float Rc[9], Bt[9];
    find_rotation(time_scan_cur + (double)time[s], imu_time, imu_rot3, n_imu - 1, r);
    rot_of(r[0], r[1], r[2], Rc);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) Bt[a * 3 + b] = (Sinv[a * 3] * Rc[b] + Sinv[a * 3 + 1] * Rc[3 + b]) + Sinv[a * 3 + 2] * Rc[6 + b];
    float* o = out4 + 4 * (size_t)i;
    o[0] = Bt[0] * p[0] + Bt[1] * p[1] + Bt[2] * p[2] + 0.f;
    o[1] = Bt[3] * p[0] + Bt[4] * p[1] + Bt[5] * p[2] + 0.f;
    o[2] = Bt[6] * p[0] + Bt[7] * p[1] + Bt[8] * p[2] + 0.f;
    o[3] = p[3];
  }
}

}  // extern "C"
