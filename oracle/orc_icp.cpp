// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the loop-closure verification ICP:
//   detectLoopClosureForSubMap   src/node/subMapOptmizationNode.cpp:2739-2916
//   pcl::IterativeClosestPoint with setMaxCorrespondenceDistance(10), setMaximumIterations(30),
//   setTransformationEpsilon(1e-4), setEuclideanFitnessEpsilon(1e-4), setRANSACIterations(0) (:2763-2769),
//   source pre-transformed by the initial guess (:2822-2824), align() with identity guess (:2831),
//   getFitnessScore() (:2835), accept iff converged and score <= historyKeyframeFitnessScore 0.5 (:2855).
// PCL is a third-party dependency absent from /root/reference (inferred PCL 1.8.1, unpinned); restated from its
// published algorithm (registration/impl/icp.hpp, default_convergence_criteria.hpp,
// correspondence_estimation.hpp, transformation_estimation_svd.hpp):
//   loop { 1-NN correspondences with d^2 <= max^2 ; < 3 => not converged ; rigid fit of (cur -> target) by
//          Umeyama/SVD ; cur = T * cur ; final = T * final ; ++iter ; converged = criteria() }
//   criteria: iter >= max -> true ; cos(angle) >= 1 - eps_T && |t|^2 <= eps_T -> true ;
//             |mse - prev| < 1e-12 -> true ; |mse - prev| / prev < eps_fit -> true ; prev = mse.
//   fitness = mean squared 1-NN distance of the source transformed by the final transformation.
// Deviation (documented): PCL accumulates the Umeyama means/covariance in fp32 in Eigen's GEMM order
// (unpinnable); here and on the GPU the 17 sums are accumulated in fp64 and the 3x3 rotation is solved in fp64,
// then rounded to fp32 like PCL's Matrix4f.
#include "orc_api.h"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>



namespace {

// symmetric Jacobi (cyclic, fp64) for n <= 4: eigenvalues W, eigenvectors in the COLUMNS of V
void jacobi_sym(double* A, int n, double* W, double* V) {
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = i == j;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (std::fabs(A[p * n + q]) < 1e-300) continue;
      double theta = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
      double t = (theta >= 0 ? 1 : -1) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
      double c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk; }
      for (int k = 0; k < n; k++) { double vkp = V[k * n + p], vkq = V[k * n + q]; V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < n; i++) W[i] = A[i * n + i];
}

// optimal proper rotation R (dst ~ R src) from the cross-covariance H = sum (dst - mean_d)(src - mean_s)^T:
// Horn's closed form (largest eigenvector of the 4x4 N matrix == Umeyama/SVD with the det fix).
void best_rotation(const double H[9], double R[9]) {
  // M = sum src * dst^T = H^T
  const double Sxx = H[0], Sxy = H[3], Sxz = H[6], Syx = H[1], Syy = H[4], Syz = H[7], Szx = H[2], Szy = H[5], Szz = H[8];
  double N[16] = {Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx,
                  Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz,
                  Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy,
                  Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz};
  double W[4], V[16];
  jacobi_sym(N, 4, W, V);
  int b = 0; for (int i = 1; i < 4; i++) if (W[i] > W[b]) b = i;
  double q0 = V[0 * 4 + b], q1 = V[1 * 4 + b], q2 = V[2 * 4 + b], q3 = V[3 * 4 + b];
  double nq = std::sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 /= nq; q1 /= nq; q2 /= nq; q3 /= nq;
  R[0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

inline void xform(const float T[16], const float* p, float* q) {   // Eigen: linear()*v (column order) + translation
  q[0] = (T[0] * p[0] + T[1] * p[1]) + T[2] * p[2] + T[3];
  q[1] = (T[4] * p[0] + T[5] * p[1]) + T[6] * p[2] + T[7];
  q[2] = (T[8] * p[0] + T[9] * p[1]) + T[10] * p[2] + T[11];
}
inline void matmul4(const float A[16], const float B[16], float C[16]) {
  float r[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0; for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j]; r[i * 4 + j] = s; }
  memcpy(C, r, sizeof(r));
}

}  // namespace

extern "C" int orc_icp(const float* src4, int32_t ns, const float* tgt4, int32_t nt, const orc_icp_params* prm, orc_icp_result* res) {
  memset(res, 0, sizeof(*res));
  for (int i = 0; i < 16; i++) res->T[i] = (i % 5 == 0) ? 1.f : 0.f;
  res->fitness = DBL_MAX;
  if (ns <= 0 || nt <= 0) return 0;
  void* tree = orc_kdtree_build(tgt4, nt);
  std::vector<float> cur(3 * (size_t)ns);
  for (int i = 0; i < ns; i++) for (int d = 0; d < 3; d++) cur[3 * (size_t)i + d] = src4[4 * (size_t)i + d];
  const double max2 = (double)prm->max_corr_dist * (double)prm->max_corr_dist;
  const double rot_thr = 1.0 - prm->trans_eps, trans_thr = prm->trans_eps;
  double prev_mse = DBL_MAX;
  int iters = 0; bool converged = false;
  float final_T[16]; memcpy(final_T, res->T, sizeof(final_T));
  for (;;) {
    double sp[3] = {0, 0, 0}, sq[3] = {0, 0, 0}, spq[9] = {0}, sd = 0; long long n = 0;
    for (int i = 0; i < ns; i++) {
      int idx; float d2;
      if (orc_kdtree_knn(tree, &cur[3 * (size_t)i], 1, &idx, &d2) < 1) continue;
      if ((double)d2 > max2) continue;
      const float* p = &cur[3 * (size_t)i]; const float* q = tgt4 + 4 * (size_t)idx;
      for (int a = 0; a < 3; a++) { sp[a] += p[a]; sq[a] += q[a]; for (int b = 0; b < 3; b++) spq[a * 3 + b] += (double)q[a] * (double)p[b]; }
      sd += d2; n++;
    }
    res->n_corr_last = (int32_t)n;
    if (n < 3) { converged = false; break; }
    double mp[3], mq[3], H[9], R[9];
    for (int a = 0; a < 3; a++) { mp[a] = sp[a] / n; mq[a] = sq[a] / n; }
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) H[a * 3 + b] = spq[a * 3 + b] / n - mq[a] * mp[b];   // dst x src^T
    best_rotation(H, R);
    float T[16] = {0};
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) T[a * 4 + b] = (float)R[a * 3 + b];
      T[a * 4 + 3] = (float)(mq[a] - (R[a * 3] * mp[0] + R[a * 3 + 1] * mp[1] + R[a * 3 + 2] * mp[2]));
    }
    T[15] = 1.f;
    for (int i = 0; i < ns; i++) { float q[3]; xform(T, &cur[3 * (size_t)i], q); cur[3 * (size_t)i] = q[0]; cur[3 * (size_t)i + 1] = q[1]; cur[3 * (size_t)i + 2] = q[2]; }
    matmul4(T, final_T, final_T);
    iters++;
    // DefaultConvergenceCriteria::hasConverged
    if (iters >= prm->max_iters) { converged = true; break; }
    const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1);
    const double tr2 = (double)T[3] * T[3] + (double)T[7] * T[7] + (double)T[11] * T[11];
    if (cos_angle >= rot_thr && tr2 <= trans_thr) { converged = true; break; }
    const double mse = sd / n;
    if (std::fabs(mse - prev_mse) < 1e-12) { converged = true; break; }
    if (std::fabs(mse - prev_mse) / prev_mse < prm->fitness_eps) { converged = true; break; }
    prev_mse = mse;
  }
  res->iters = iters; res->converged = converged ? 1 : 0;
  memcpy(res->T, final_T, sizeof(final_T));
  // getFitnessScore(): source transformed by the final transformation, unbounded 1-NN
  double fs = 0; long long nr = 0;
  for (int i = 0; i < ns; i++) {
    float q[3]; xform(final_T, src4 + 4 * (size_t)i, q);
    int idx; float d2;
    if (orc_kdtree_knn(tree, q, 1, &idx, &d2) < 1) continue;
    fs += d2; nr++;
  }
  res->fitness = nr > 0 ? fs / nr : DBL_MAX;
  orc_kdtree_free(tree);
  return res->converged;
}
