// Small dense fp32 routines evaluated per point / per registration on the device.
// They follow the published algorithms of the third-party routines the reference
// calls so that the GPU path rounds like the CPU path (compiled with --fmad=false):
//   cv::eigen (symmetric Jacobi)      odomEstimationNode.cpp:690, :928
//   cv::solve(DECOMP_QR) (Householder) :921      cv::Mat::inv (LU) :945
//   Eigen colPivHouseholderQr().solve  :783
#pragma once
#include <cfloat>
#include <cuda_runtime.h>
#include <cmath>
#ifndef LISREG_HD
#define LISREG_HD __host__ __device__
#endif

namespace lisreg {

LISREG_HD __forceinline__ float cv_hypotf(float a, float b) {
  a = fabsf(a); b = fabsf(b);
  if (a > b) { b /= a; return a * sqrtf(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrtf(1 + a * a); }
  return 0.f;
}

LISREG_HD __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }

// Symmetric Jacobi: eigenvalues descending in W, eigenvectors in the ROWS of V.
// A, W, V and the index scratch indR/indC (N ints each) are caller-provided: device callers pass
// SHARED memory (data-dependent indexing of per-thread local arrays is avoided on the device).
template <int N>
LISREG_HD void jacobi_eigen(float* A, float* W, float* V, int* indR, int* indC) {
  const float eps = FLT_EPSILON;
  int i, j, k, m;
  for (i = 0; i < N; i++) { for (j = 0; j < N; j++) V[i * N + j] = 0.f; V[i * N + i] = 1.f; }
  for (i = 0; i < N; i++) { indR[i] = 0; indC[i] = 0; }
  float mv = 0.f;
  for (k = 0; k < N; k++) {
    W[k] = A[(N + 1) * k];
    if (k < N - 1) {
      for (m = k + 1, mv = fabsf(A[N * k + m]), i = k + 2; i < N; i++) {
        float val = fabsf(A[N * k + i]);
        if (mv < val) mv = val, m = i;
      }
      indR[k] = m;
    }
    if (k > 0) {
      for (m = 0, mv = fabsf(A[k]), i = 1; i < k; i++) {
        float val = fabsf(A[N * i + k]);
        if (mv < val) mv = val, m = i;
      }
      indC[k] = m;
    }
  }
  const int maxIters = N * N * 30;
  for (int iters = 0; iters < maxIters; iters++) {
    for (k = 0, mv = fabsf(A[indR[0]]), i = 1; i < N - 1; i++) {
      float val = fabsf(A[N * i + indR[i]]);
      if (mv < val) mv = val, k = i;
    }
    int l = indR[k];
    for (i = 1; i < N; i++) {
      float val = fabsf(A[N * indC[i] + i]);
      if (mv < val) mv = val, k = indC[i], l = i;
    }
    float p = A[N * k + l];
    if (fabsf(p) <= eps) break;
    float y = (float)((double)(W[l] - W[k]) * 0.5);
    float t = fabsf(y) + cv_hypotf(p, y);
    float s = cv_hypotf(p, t);
    float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) s = -s, t = -t;
    A[N * k + l] = 0.f;
    W[k] -= t;
    W[l] += t;
    float a0, b0;
#define LISREG_ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
    for (i = 0; i < k; i++) LISREG_ROT(A[N * i + k], A[N * i + l]);
    for (i = k + 1; i < l; i++) LISREG_ROT(A[N * k + i], A[N * i + l]);
    for (i = l + 1; i < N; i++) LISREG_ROT(A[N * k + i], A[N * l + i]);
    for (i = 0; i < N; i++) LISREG_ROT(V[N * k + i], V[N * l + i]);
#undef LISREG_ROT
    for (j = 0; j < 2; j++) {
      int idx = j == 0 ? k : l;
      if (idx < N - 1) {
        for (m = idx + 1, mv = fabsf(A[N * idx + m]), i = idx + 2; i < N; i++) {
          float val = fabsf(A[N * idx + i]);
          if (mv < val) mv = val, m = i;
        }
        indR[idx] = m;
      }
      if (idx > 0) {
        for (m = 0, mv = fabsf(A[idx]), i = 1; i < idx; i++) {
          float val = fabsf(A[N * i + idx]);
          if (mv < val) mv = val, m = i;
        }
        indC[idx] = m;
      }
    }
  }
  for (k = 0; k < N - 1; k++) {
    m = k;
    for (i = k + 1; i < N; i++) if (W[m] < W[i]) m = i;
    if (k != m) {
      swapf(W[m], W[k]);
      for (i = 0; i < N; i++) swapf(V[N * m + i], V[N * k + i]);
    }
  }
}

// 3x3 specialisation of the same algorithm with every operand in registers: the pivot pair
// (k,l) in {(0,1),(0,2),(1,2)} selects one of three statically-indexed code paths.  Same operation
// order as jacobi_eigen<3>, hence bit-identical results (checked in tests).
LISREG_HD __forceinline__ void jacobi_eigen3(float a00, float a01, float a02, float a11, float a12, float a22,
                                             float (&W)[3], float (&V)[9]) {
  const float eps = FLT_EPSILON;
  float w0 = a00, w1 = a11, w2 = a22;
  float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
  // indR[0] in {1,2}; indR[1] == 2; indC[1] == 0; indC[2] in {0,1}
  int r0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
  int c2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
  for (int iters = 0; iters < 270; iters++) {
    // pivot search
    int k = 0;
    float mv = fabsf(r0 == 1 ? a01 : a02);
    { float val = fabsf(a12); if (mv < val) mv = val, k = 1; }
    int l = (k == 0) ? r0 : 2;
    { float val = fabsf(a01); if (mv < val) mv = val, k = 0, l = 1; }
    { float val = fabsf(c2 == 0 ? a02 : a12); if (mv < val) mv = val, k = c2, l = 2; }
    const int kl = k * 3 + l;   // 1, 2 or 5
    const float p = kl == 1 ? a01 : (kl == 2 ? a02 : a12);
    if (fabsf(p) <= eps) break;
    const float wk = k == 0 ? w0 : w1, wl = l == 1 ? w1 : w2;
    const float y = (float)((double)(wl - wk) * 0.5);
    float t = fabsf(y) + cv_hypotf(p, y);
    float s = cv_hypotf(p, t);
    const float c = t / s;
    s = p / s; t = (p / t) * p;
    if (y < 0) s = -s, t = -t;
#define LISREG_ROT3(v0, v1) { const float a0_ = v0, b0_ = v1; v0 = a0_ * c - b0_ * s; v1 = a0_ * s + b0_ * c; }
    if (kl == 1) {          // (0,1)
      a01 = 0.f; w0 -= t; w1 += t;
      LISREG_ROT3(a02, a12);
      LISREG_ROT3(v00, v10); LISREG_ROT3(v01, v11); LISREG_ROT3(v02, v12);
    } else if (kl == 2) {   // (0,2)
      a02 = 0.f; w0 -= t; w2 += t;
      LISREG_ROT3(a01, a12);
      LISREG_ROT3(v00, v20); LISREG_ROT3(v01, v21); LISREG_ROT3(v02, v22);
    } else {                // (1,2)
      a12 = 0.f; w1 -= t; w2 += t;
      LISREG_ROT3(a01, a02);
      LISREG_ROT3(v10, v20); LISREG_ROT3(v11, v21); LISREG_ROT3(v12, v22);
    }
#undef LISREG_ROT3
    // index updates for idx = k then idx = l (only indR[0] and indC[2] can change)
    if (k == 0) r0 = (fabsf(a01) < fabsf(a02)) ? 2 : 1;
    if (l == 2) c2 = (fabsf(a02) < fabsf(a12)) ? 1 : 0;
  }
  // sort descending (selection sort with row swaps, as the reference routine)
  W[0] = w0; W[1] = w1; W[2] = w2;
  V[0] = v00; V[1] = v01; V[2] = v02; V[3] = v10; V[4] = v11; V[5] = v12; V[6] = v20; V[7] = v21; V[8] = v22;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    int m = k;
#pragma unroll
    for (int i = k + 1; i < 3; i++) if (W[m] < W[i]) m = i;
#pragma unroll
    for (int mm = k + 1; mm < 3; mm++) {
      if (mm == m) {
        swapf(W[mm], W[k]);
#pragma unroll
        for (int i = 0; i < 3; i++) swapf(V[3 * mm + i], V[3 * k + i]);
      }
    }
  }
}

// Householder QR solve, square N x N (A destroyed, b -> x). Returns 0 if singular.  Every loop has a compile-time trip
// count and is unrolled: called on local arrays the whole factorisation stays in registers (lm_solve_tail).
template <int N>
LISREG_HD __forceinline__ int qr_solve(float* A, float* b) {
  const float eps = FLT_EPSILON * 10;
  float vl[N], hF[N];
  #pragma unroll
  for (int l = 0; l < N; l++) {
    int vlSize = N - l;
    float vlNorm = 0.f;
    #pragma unroll
    for (int i = 0; i < vlSize; i++) { vl[i] = A[(l + i) * N + l]; vlNorm += vl[i] * vl[i]; }
    float tmpV = vl[0];
    vl[0] = vl[0] + (vl[0] >= 0.f ? 1.f : -1.f) * sqrtf(vlNorm);
    vlNorm = sqrtf(vlNorm + vl[0] * vl[0] - tmpV * tmpV);
    #pragma unroll
    for (int i = 0; i < vlSize; i++) vl[i] /= vlNorm;
    #pragma unroll
    for (int j = l; j < N; j++) {
      float v_lA = 0.f;
      #pragma unroll
      for (int i = l; i < N; i++) v_lA += vl[i - l] * A[i * N + j];
      #pragma unroll
      for (int i = l; i < N; i++) A[i * N + j] -= 2 * vl[i - l] * v_lA;
    }
    hF[l] = vl[0] * vl[0];
    #pragma unroll
    for (int i = 1; i < vlSize; i++) A[(l + i) * N + l] = vl[i] / vl[0];
  }
  #pragma unroll
  for (int l = 0; l < N; l++) {
    #pragma unroll
    for (int j = 0; j < l; j++) vl[j] = 0.f;
    vl[l] = 1.f;
    #pragma unroll
    for (int j = l + 1; j < N; j++) vl[j] = A[j * N + l];
    float v_lB = 0.f;
    #pragma unroll
    for (int i = l; i < N; i++) v_lB += vl[i] * b[i];
    #pragma unroll
    for (int i = l; i < N; i++) b[i] -= 2 * vl[i] * v_lB * hF[l];
  }
  #pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    #pragma unroll
    for (int j = N - 1; j > i; j--) b[i] -= b[j] * A[i * N + j];
    if (fabsf(A[i * N + i]) < eps) return 0;
    b[i] /= A[i * N + i];
  }
  return 1;
}

// LU with partial pivoting: solves A X = B in place (B is N x NB row-major).
template <int N, int NB>
LISREG_HD int lu_solve(float* A, float* b) {
  const float eps = FLT_EPSILON * 10;
  int i, j, k;
  for (i = 0; i < N; i++) {
    k = i;
    for (j = i + 1; j < N; j++)
      if (fabsf(A[j * N + i]) > fabsf(A[k * N + i])) k = j;
    if (fabsf(A[k * N + i]) < eps) return 0;
    if (k != i) {
      for (j = i; j < N; j++) swapf(A[i * N + j], A[k * N + j]);
      for (j = 0; j < NB; j++) swapf(b[i * NB + j], b[k * NB + j]);
    }
    float d = -1 / A[i * N + i];
    for (j = i + 1; j < N; j++) {
      float alpha = A[j * N + i] * d;
      for (k = i + 1; k < N; k++) A[j * N + k] += alpha * A[i * N + k];
      for (k = 0; k < NB; k++) b[j * NB + k] += alpha * b[i * NB + k];
    }
  }
  for (i = N - 1; i >= 0; i--)
    for (j = 0; j < NB; j++) {
      float s = b[i * NB + j];
      for (k = i + 1; k < N; k++) s -= A[i * N + k] * b[k * NB + j];
      b[i * NB + j] = s / A[i * N + i];
    }
  return 1;
}

// 5x3 least squares A x = rhs by column-pivoting Householder QR. A row-major, destroyed.
LISREG_HD __forceinline__ void colpiv_qr_solve_5x3(float* A, const float* rhs, float* x) {
  const int R = 5, C = 3;
  float c[5];
#pragma unroll
  for (int i = 0; i < R; i++) c[i] = rhs[i];
  int perm[3] = {0, 1, 2};
  float hcoef[3];
  float maxpivot = 0.f;
  int nonzero = C;
#pragma unroll
  for (int k = 0; k < C; k++) {
    int big = k; float bigv = -1.f;
#pragma unroll
    for (int j = k; j < C; j++) {
      float s = 0.f;
#pragma unroll
      for (int i = k; i < R; i++) s += A[i * C + j] * A[i * C + j];
      if (s > bigv) { bigv = s; big = j; }
    }
    if (nonzero == C && bigv <= 0.f) nonzero = k;
    if (big != k) {
#pragma unroll
      for (int j = k + 1; j < C; j++)
        if (j == big) {
#pragma unroll
          for (int i = 0; i < R; i++) swapf(A[i * C + k], A[i * C + j]);
          int t = perm[k]; perm[k] = perm[j]; perm[j] = t;
        }
    }
    float c0 = A[k * C + k];
    float tailsq = 0.f;
#pragma unroll
    for (int i = k + 1; i < R; i++) tailsq += A[i * C + k] * A[i * C + k];
    float beta, tau;
    if (tailsq <= FLT_MIN) {
      tau = 0.f; beta = c0;
#pragma unroll
      for (int i = k + 1; i < R; i++) A[i * C + k] = 0.f;
    } else {
      beta = sqrtf(c0 * c0 + tailsq);
      if (c0 >= 0.f) beta = -beta;
#pragma unroll
      for (int i = k + 1; i < R; i++) A[i * C + k] = A[i * C + k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k * C + k] = beta;
    hcoef[k] = tau;
    if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
#pragma unroll
    for (int j = k + 1; j < C; j++) {
      float tmp = 0.f;
#pragma unroll
      for (int i = k + 1; i < R; i++) tmp += A[i * C + k] * A[i * C + j];
      tmp += A[k * C + j];
      A[k * C + j] -= tau * tmp;
#pragma unroll
      for (int i = k + 1; i < R; i++) A[i * C + j] -= tau * A[i * C + k] * tmp;
    }
  }
  const float premult = fabsf(maxpivot) * (FLT_EPSILON * (float)C);
  int rank = 0;
#pragma unroll
  for (int i = 0; i < C; i++) rank += (i < nonzero && fabsf(A[i * C + i]) > premult) ? 1 : 0;
#pragma unroll
  for (int k = 0; k < C; k++) {
    float tmp = 0.f;
#pragma unroll
    for (int i = k + 1; i < R; i++) tmp += A[i * C + k] * c[i];
    tmp += c[k];
    c[k] -= hcoef[k] * tmp;
#pragma unroll
    for (int i = k + 1; i < R; i++) c[i] -= hcoef[k] * A[i * C + k] * tmp;
  }
  float y[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = C - 1; i >= 0; i--) {
    if (i < rank) {
      float s = c[i];
#pragma unroll
      for (int j = i + 1; j < C; j++) if (j < rank) s -= A[i * C + j] * y[j];
      y[i] = s / A[i * C + i];
    }
  }
  x[0] = x[1] = x[2] = 0.f;
#pragma unroll
  for (int i = 0; i < C; i++) {
    if (perm[i] == 0) x[0] = y[i]; else if (perm[i] == 1) x[1] = y[i]; else x[2] = y[i];
  }
}

}  // namespace lisreg
