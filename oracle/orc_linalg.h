// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under lis_slam_b200/ may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use the oracle.
//
// PARITY UNPINNED: the reference (QingzhiWang/LIS-SLAM) ships no tests or golden
// vectors for this path and cannot be built here (ROS/PCL/OpenCV-C++/Eigen absent).
// The small dense routines below restate the *published algorithms* of the
// third-party routines the reference calls, and are pinned in tests/ against the
// same routines of the Python OpenCV wheel (cv2 4.13: cv2.eigen, cv2.solve(QR),
// cv2.invert(LU)) which is present in this image.
//
//   cv::eigen            <- odomEstimationNode.cpp:690, :928   (symmetric Jacobi)
//   cv::solve DECOMP_QR  <- odomEstimationNode.cpp:921         (Householder QR)
//   cv::Mat::inv (LU)    <- odomEstimationNode.cpp:945
//   Eigen colPivHouseholderQr().solve  <- odomEstimationNode.cpp:783
#pragma once
#include <cmath>
#include <cfloat>
#include <utility>

namespace orc {

// OpenCV's own hypot helper used by its Jacobi (core/src/lapack.cpp).
template <typename T> static inline T cv_hypot(T a, T b) {
  a = std::abs(a); b = std::abs(b);
  if (a > b) { b /= a; return a * std::sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * std::sqrt(1 + a * a); }
  return 0;
}

// Symmetric Jacobi eigen-decomposition, OpenCV JacobiImpl_ semantics:
// A (n x n, row-major, destroyed), eigenvalues W descending, eigenvectors are the
// ROWS of V.  n <= 6.
template <typename T> static void jacobi_eigen(T* A, int n, T* W, T* V) {
  const T eps = std::numeric_limits<T>::epsilon();
  int i, j, k, m;
  for (i = 0; i < n; i++) { for (j = 0; j < n; j++) V[i * n + j] = 0; V[i * n + i] = 1; }
  int indR[8], indC[8];
  T mv = 0;
  for (k = 0; k < n; k++) {
    W[k] = A[(n + 1) * k];
    if (k < n - 1) {
      for (m = k + 1, mv = std::abs(A[n * k + m]), i = k + 2; i < n; i++) {
        T val = std::abs(A[n * k + i]);
        if (mv < val) mv = val, m = i;
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = std::abs(A[k]), i = 1; i < k; i++) {
        T val = std::abs(A[n * i + k]);
        if (mv < val) mv = val, m = i;
      }
      indC[k] = m;
    }
  }
  int maxIters = n * n * 30;
  if (n > 1) for (int iters = 0; iters < maxIters; iters++) {
    for (k = 0, mv = std::abs(A[indR[0]]), i = 1; i < n - 1; i++) {
      T val = std::abs(A[n * i + indR[i]]);
      if (mv < val) mv = val, k = i;
    }
    int l = indR[k];
    for (i = 1; i < n; i++) {
      T val = std::abs(A[n * indC[i] + i]);
      if (mv < val) mv = val, k = indC[i], l = i;
    }
    T p = A[n * k + l];
    if (std::abs(p) <= eps) break;
    T y = (T)((W[l] - W[k]) * 0.5);
    T t = std::abs(y) + cv_hypot(p, y);
    T s = cv_hypot(p, t);
    T c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) s = -s, t = -t;
    A[n * k + l] = 0;
    W[k] -= t;
    W[l] += t;
    T a0, b0;
#define ORC_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
    for (i = 0; i < k; i++) ORC_ROT(A[n * i + k], A[n * i + l]);
    for (i = k + 1; i < l; i++) ORC_ROT(A[n * k + i], A[n * i + l]);
    for (i = l + 1; i < n; i++) ORC_ROT(A[n * k + i], A[n * l + i]);
    for (i = 0; i < n; i++) ORC_ROT(V[n * k + i], V[n * l + i]);
#undef ORC_ROT
    for (j = 0; j < 2; j++) {
      int idx = j == 0 ? k : l;
      if (idx < n - 1) {
        for (m = idx + 1, mv = std::abs(A[n * idx + m]), i = idx + 2; i < n; i++) {
          T val = std::abs(A[n * idx + i]);
          if (mv < val) mv = val, m = i;
        }
        indR[idx] = m;
      }
      if (idx > 0) {
        for (m = 0, mv = std::abs(A[idx]), i = 1; i < idx; i++) {
          T val = std::abs(A[n * i + idx]);
          if (mv < val) mv = val, m = i;
        }
        indC[idx] = m;
      }
    }
  }
  for (k = 0; k < n - 1; k++) {
    m = k;
    for (i = k + 1; i < n; i++) if (W[m] < W[i]) m = i;
    if (k != m) {
      std::swap(W[m], W[k]);
      for (i = 0; i < n; i++) std::swap(V[n * m + i], V[n * k + i]);
    }
  }
}

// Householder QR solve of a square n x n system, OpenCV hal::QR32f semantics
// (A row-major destroyed, b overwritten by the solution). Returns 0 if singular.
template <typename T> static int qr_solve(T* A, int n, T* b) {
  const int m = n;
  const T eps = std::numeric_limits<T>::epsilon() * 10;  // hal::QR32f: FLT_EPSILON*10
  T vl[8], hF[8];
  for (int l = 0; l < n; l++) {
    int vlSize = m - l;
    T vlNorm = 0;
    for (int i = 0; i < vlSize; i++) { vl[i] = A[(l + i) * n + l]; vlNorm += vl[i] * vl[i]; }
    T tmpV = vl[0];
    vl[0] = vl[0] + (vl[0] >= 0 ? 1 : -1) * std::sqrt(vlNorm);
    vlNorm = std::sqrt(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
    for (int i = 0; i < vlSize; i++) vl[i] /= vlNorm;
    for (int j = l; j < n; j++) {
      T v_lA = 0;
      for (int i = l; i < m; i++) v_lA += vl[i - l] * A[i * n + j];
      for (int i = l; i < m; i++) A[i * n + j] -= 2 * vl[i - l] * v_lA;
    }
    hF[l] = vl[0] * vl[0];
    for (int i = 1; i < vlSize; i++) A[(l + i) * n + l] = vl[i] / vl[0];
  }
  for (int l = 0; l < n; l++) {
    for (int j = 0; j < l; j++) vl[j] = 0;
    vl[l] = 1;
    for (int j = l + 1; j < m; j++) vl[j] = A[j * n + l];
    T v_lB = 0;
    for (int i = l; i < m; i++) v_lB += vl[i] * b[i];
    for (int i = l; i < m; i++) b[i] -= 2 * vl[i] * v_lB * hF[l];
  }
  for (int i = n - 1; i >= 0; i--) {
    for (int j = n - 1; j > i; j--) b[i] -= b[j] * A[i * n + j];
    if (std::abs(A[i * n + i]) < eps) return 0;
    b[i] /= A[i * n + i];
  }
  return 1;
}

// LU with partial pivoting, OpenCV hal::LU32f semantics: solves A X = B in place
// (B is n x nb row-major).  Returns 0 if singular.
template <typename T> static int lu_solve(T* A, int m, T* b, int nb) {
  const T eps = std::numeric_limits<T>::epsilon() * 10;
  int i, j, k, p = 1;
  for (i = 0; i < m; i++) {
    k = i;
    for (j = i + 1; j < m; j++)
      if (std::abs(A[j * m + i]) > std::abs(A[k * m + i])) k = j;
    if (std::abs(A[k * m + i]) < eps) return 0;
    if (k != i) {
      for (j = i; j < m; j++) std::swap(A[i * m + j], A[k * m + j]);
      for (j = 0; j < nb; j++) std::swap(b[i * nb + j], b[k * nb + j]);
      p = -p;
    }
    T d = -1 / A[i * m + i];
    for (j = i + 1; j < m; j++) {
      T alpha = A[j * m + i] * d;
      for (k = i + 1; k < m; k++) A[j * m + k] += alpha * A[i * m + k];
      for (k = 0; k < nb; k++) b[j * nb + k] += alpha * b[i * nb + k];
    }
  }
  for (i = m - 1; i >= 0; i--)
    for (j = 0; j < nb; j++) {
      T s = b[i * nb + j];
      for (k = i + 1; k < m; k++) s -= A[i * m + k] * b[k * nb + j];
      b[i * nb + j] = s / A[i * m + i];
    }
  return p;
}

// Least squares of a 5x3 system  A x = rhs  by column-pivoting Householder QR
// (Eigen::ColPivHouseholderQR semantics: pivot = largest remaining column norm,
// Householder makeHouseholderInPlace sign convention, rank from pivot threshold
// eps*min(rows,cols)*|maxpivot|).  A is 5x3 row-major, destroyed.
static inline void colpiv_qr_solve_5x3(float* A, const float* rhs, float* x) {
  const int R = 5, C = 3;
  float c[5];
  for (int i = 0; i < R; i++) c[i] = rhs[i];
  int perm[3] = {0, 1, 2};
  float hcoef[3];
  float maxpivot = 0.f;
  int nonzero = C;
  float colsq[3];
  for (int j = 0; j < C; j++) {
    float s = 0.f;
    for (int i = 0; i < R; i++) s += A[i * C + j] * A[i * C + j];
    colsq[j] = s;
  }
  for (int k = 0; k < C; k++) {
    // recompute the exact remaining squared norms (Eigen 3.2 behaviour; 3.3's
    // down-dating picks the same pivot away from ties)
    int big = k; float bigv = -1.f;
    for (int j = k; j < C; j++) {
      float s = 0.f;
      for (int i = k; i < R; i++) s += A[i * C + j] * A[i * C + j];
      colsq[j] = s;
      if (s > bigv) { bigv = s; big = j; }
    }
    if (nonzero == C && bigv <= 0.f) nonzero = k;
    if (big != k) {
      for (int i = 0; i < R; i++) std::swap(A[i * C + k], A[i * C + big]);
      std::swap(perm[k], perm[big]);
    }
    // Householder on column k, rows k..R-1
    float c0 = A[k * C + k];
    float tailsq = 0.f;
    for (int i = k + 1; i < R; i++) tailsq += A[i * C + k] * A[i * C + k];
    float beta, tau;
    if (tailsq <= FLT_MIN) {
      tau = 0.f; beta = c0;
      for (int i = k + 1; i < R; i++) A[i * C + k] = 0.f;
    } else {
      beta = std::sqrt(c0 * c0 + tailsq);
      if (c0 >= 0.f) beta = -beta;
      for (int i = k + 1; i < R; i++) A[i * C + k] = A[i * C + k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k * C + k] = beta;
    hcoef[k] = tau;
    if (std::abs(beta) > maxpivot) maxpivot = std::abs(beta);
    // apply H = I - tau v v^T (v = [1, essential]) to the trailing columns
    for (int j = k + 1; j < C; j++) {
      float tmp = 0.f;  // Eigen applyHouseholderOnTheLeft: essential^T*bottom, then += top
      for (int i = k + 1; i < R; i++) tmp += A[i * C + k] * A[i * C + j];
      tmp += A[k * C + j];
      A[k * C + j] -= tau * tmp;
      for (int i = k + 1; i < R; i++) A[i * C + j] -= tau * A[i * C + k] * tmp;
    }
  }
  // rank
  const float premult = std::abs(maxpivot) * (FLT_EPSILON * (float)C);
  int rank = 0;
  for (int i = 0; i < nonzero; i++) rank += (std::abs(A[i * C + i]) > premult) ? 1 : 0;
  // c = Q^T rhs
  for (int k = 0; k < C; k++) {
    float tmp = 0.f;
    for (int i = k + 1; i < R; i++) tmp += A[i * C + k] * c[i];
    tmp += c[k];
    c[k] -= hcoef[k] * tmp;
    for (int i = k + 1; i < R; i++) c[i] -= hcoef[k] * A[i * C + k] * tmp;
  }
  // back substitution on the rank x rank upper triangle
  float y[3] = {0.f, 0.f, 0.f};
  for (int i = rank - 1; i >= 0; i--) {
    float s = c[i];
    for (int j = i + 1; j < rank; j++) s -= A[i * C + j] * y[j];
    y[i] = s / A[i * C + i];
  }
  x[0] = x[1] = x[2] = 0.f;
  for (int i = 0; i < C; i++) x[perm[i]] = y[i];
}

}  // namespace orc
