// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
//
// CPU restatement of the LIS-SLAM scan-to-map "LM" inner loop, sequential
// semantics (the reference's OpenMP pragmas are inert as built, SURVEY.md header):
//   scan2SubMapOptimization   src/node/odomEstimationNode.cpp:596-626
//   cornerOptimization        :633-747      surfOptimization  :749-827
//   combineOptimizationCoeffs :829-850      LMOptimization    :852-974
//   transformUpdate (clamps)  :1001-1003    pointAssociateToMap :243-258
//   trans2Affine3f            src/core/common.cpp:55-58 (pcl::getTransformation)
//   variants B/C deltas       src/node/subMapOptmizationNode.cpp:1509-2001, 4485-4976
// Third-party semantics restated: FLANN KDTreeSingleIndex exact 5-NN, cv::eigen,
// cv::solve(QR), cv::Mat::inv, cv GEMM (double accumulate), Eigen colPivHouseholderQr.
#include "orc_api.h"
#include "orc_linalg.h"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ------------------------------------------------------------------ kd-tree
// Exact k-NN, L2 in the FLANN functor's op order ((dx*dx)+dy*dy)+dz*dz, results
// sorted ascending by (distance, index).  Leaf size 15 as PCL's KdTreeFLANN.
struct KdTree {
  struct Node { int left, right; int begin, end; int dim; float split; };
  std::vector<Node> nodes;
  std::vector<int> perm;      // permutation of point indices
  std::vector<float> pts;     // xyz packed in tree order (3 floats)
  int n = 0;

  int build_rec(const float* p4, int b, int e) {
    Node nd; nd.left = nd.right = -1; nd.begin = b; nd.end = e; nd.dim = 0; nd.split = 0;
    int id = (int)nodes.size();
    nodes.push_back(nd);
    if (e - b <= 15) return id;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = b; i < e; i++)
      for (int d = 0; d < 3; d++) {
        float v = p4[4 * perm[i] + d];
        mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
      }
    int dim = 0; float best = mx[0] - mn[0];
    for (int d = 1; d < 3; d++) if (mx[d] - mn[d] > best) { best = mx[d] - mn[d]; dim = d; }
    if (!(best > 0.f)) return id;  // all identical: keep as (big) leaf
    int mid = (b + e) / 2;
    std::nth_element(perm.begin() + b, perm.begin() + mid, perm.begin() + e,
                     [&](int a, int c) { return p4[4 * a + dim] < p4[4 * c + dim]; });
    float split = p4[4 * perm[mid] + dim];
    int l = build_rec(p4, b, mid);
    int r = build_rec(p4, mid, e);
    nodes[id].left = l; nodes[id].right = r; nodes[id].dim = dim; nodes[id].split = split;
    return id;
  }
  void build(const float* p4, int n_) {
    n = n_;
    perm.resize(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    nodes.clear(); nodes.reserve(n / 4 + 16);
    if (n > 0) build_rec(p4, 0, n);
    pts.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++)
      for (int d = 0; d < 3; d++) pts[3 * (size_t)i + d] = p4[4 * perm[i] + d];
  }
  struct Best { float d[8]; int i[8]; int k, cnt; };
  static inline void insert(Best& b, float d, int idx) {
    if (b.cnt == b.k) {
      if (d > b.d[b.k - 1] || (d == b.d[b.k - 1] && idx > b.i[b.k - 1])) return;
    }
    int pos = b.cnt < b.k ? b.cnt : b.k - 1;
    while (pos > 0 && (b.d[pos - 1] > d || (b.d[pos - 1] == d && b.i[pos - 1] > idx))) {
      b.d[pos] = b.d[pos - 1]; b.i[pos] = b.i[pos - 1]; pos--;
    }
    b.d[pos] = d; b.i[pos] = idx;
    if (b.cnt < b.k) b.cnt++;
  }
  void search_rec(int id, const float* q, Best& b) const {
    const Node& nd = nodes[id];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; i++) {
        const float* p = &pts[3 * (size_t)i];
        float d0 = q[0] - p[0], d1 = q[1] - p[1], d2 = q[2] - p[2];
        float d = d0 * d0; d = d + d1 * d1; d = d + d2 * d2;
        insert(b, d, perm[i]);
      }
      return;
    }
    float diff = q[nd.dim] - nd.split;
    int nearc = diff < 0 ? nd.left : nd.right, farc = diff < 0 ? nd.right : nd.left;
    search_rec(nearc, q, b);
    float dd = diff * diff;
    if (b.cnt < b.k || dd <= b.d[b.k - 1]) search_rec(farc, q, b);
  }
  int knn(const float* q, int k, int* idx, float* sqd) const {
    Best b; b.k = k; b.cnt = 0;
    if (n > 0) search_rec(0, q, b);
    for (int j = 0; j < b.cnt; j++) { idx[j] = b.i[j]; sqd[j] = b.d[j]; }
    return b.cnt;
  }
};

// ------------------------------------------------------------------ pose
// sinf/cosf taken as the correctly rounded float of the double routine, so that the CPU
// oracle and the GPU path (which evaluates sin/cos in fp64) round identically; this
// differs from a given libm's sinf by at most the last ulp, on rare arguments.
static inline float sinf_cr(float x) { return (float)std::sin((double)x); }
static inline float cosf_cr(float x) { return (float)std::cos((double)x); }
// pcl::getTransformation(x,y,z,roll,pitch,yaw) closed form, fp32 (common.cpp:55-58).
static void pose_to_affine(const float t[6], float T[12]) {
  float roll = t[0], pitch = t[1], yaw = t[2];
  float A = cosf_cr(yaw), B = sinf_cr(yaw), C = cosf_cr(pitch), D = sinf_cr(pitch);
  float E = cosf_cr(roll), F = sinf_cr(roll), DE = D * E, DF = D * F;
  T[0] = A * C;  T[1] = A * DF - B * E;  T[2] = B * F + A * DE;   T[3] = t[3];
  T[4] = B * C;  T[5] = A * E + B * DF;  T[6] = B * DE - A * F;   T[7] = t[4];
  T[8] = -D;     T[9] = C * F;           T[10] = C * E;           T[11] = t[5];
}
// pointAssociateToMap, odomEstimationNode.cpp:243-258
static inline void associate(const float T[12], const float* p, float* q) {
  q[0] = T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3];
  q[1] = T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7];
  q[2] = T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11];
}

// ------------------------------------------------------------------ coefficients
// cornerOptimization body after the kNN, odomEstimationNode.cpp:657-742.
// nb = 5 neighbours xyz (sorted by distance). Returns 1 if accepted (s > 0.1).
static int corner_coeff(const float* q, const float* nb, float* coeff) {
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; j++) { cx += nb[3 * j]; cy += nb[3 * j + 1]; cz += nb[3 * j + 2]; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
  for (int j = 0; j < 5; j++) {
    float ax = nb[3 * j] - cx, ay = nb[3 * j + 1] - cy, az = nb[3 * j + 2] - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az; a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  float A[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33}, W[3], V[9];
  jacobi_eigen<float>(A, 3, W, V);
  if (!(W[0] > 3 * W[1])) return 0;
  float x0 = q[0], y0 = q[1], z0 = q[2];
  float x1 = cx + 0.1 * V[0], y1 = cy + 0.1 * V[1], z1 = cz + 0.1 * V[2];  // double literal math
  float x2 = cx - 0.1 * V[0], y2 = cy - 0.1 * V[1], z2 = cz - 0.1 * V[2];
  float a012 = std::sqrt(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                         ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                         ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
  float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
              (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
  float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) -
               (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
               (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  float ld2 = a012 / l12;
  float s = 1 - 0.9 * std::fabs(ld2);  // double literal math, narrowed
  coeff[0] = la; coeff[1] = lb; coeff[2] = lc; coeff[3] = ld2; coeff[4] = s;  // caller applies w*s
  return s > 0.1 ? 1 : 0;
}

// surfOptimization body after the kNN, odomEstimationNode.cpp:776-821.
static int surf_coeff(const float* q, const float* nb, float* coeff) {
  float A0[15], B0[5] = {-1, -1, -1, -1, -1}, X0[3];
  for (int j = 0; j < 5; j++) { A0[3 * j] = nb[3 * j]; A0[3 * j + 1] = nb[3 * j + 1]; A0[3 * j + 2] = nb[3 * j + 2]; }
  colpiv_qr_solve_5x3(A0, B0, X0);
  float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
  float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  for (int j = 0; j < 5; j++)
    if (std::fabs(pa * nb[3 * j] + pb * nb[3 * j + 1] + pc * nb[3 * j + 2] + pd) > 0.2) return 0;
  float pd2 = pa * q[0] + pb * q[1] + pc * q[2] + pd;
  float s = 1 - 0.9 * std::fabs(pd2) / std::sqrt(std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]));
  coeff[0] = pa; coeff[1] = pb; coeff[2] = pc; coeff[3] = pd2; coeff[4] = s;
  return s > 0.1 ? 1 : 0;
}

// coeff = w*s*(dir, dist): variant A has w == 1 ((1*s)*la == s*la bit-exactly);
// variants B/C: w = 2.0 - LabelSorce[label] (double literal, narrowed),
// subMapOptmizationNode.cpp:1669-1676, :1793-1800.
static inline void apply_weight(const float raw[5], float w, float out[4]) {
  out[0] = w * raw[4] * raw[0]; out[1] = w * raw[4] * raw[1];
  out[2] = w * raw[4] * raw[2]; out[3] = w * raw[4] * raw[3];
}
static inline float label_weight(const orc_lm_params* prm, const uint16_t* labels, int i) {
  if (!prm->use_label_weight) return 1.0f;
  unsigned l = labels ? labels[i] : 0;
  float sc = l < ORC_LUT_SIZE ? prm->label_score[l] : 0.f;  // std::map::operator[] default 0
  float w = 2.0 - sc;
  return w;
}

}  // namespace orc

using namespace orc;

extern "C" {

void* orc_kdtree_build(const float* pts4, int32_t n) {
  KdTree* t = new KdTree();
  t->build(pts4, n);
  return t;
}
void orc_kdtree_free(void* t) { delete (KdTree*)t; }
int32_t orc_kdtree_knn(const void* t, const float* q3, int32_t k, int32_t* idx, float* sqd) {
  return ((const KdTree*)t)->knn(q3, k, idx, sqd);
}
void orc_knn_batch(const void* t, const float* q4, int32_t nq, int32_t k, int32_t* idx, float* sqd,
                   int32_t n_threads) {
  const KdTree* kt = (const KdTree*)t;
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
  for (int i = 0; i < nq; i++) {
    int li[8]; float ld[8];
    int c = kt->knn(q4 + 4 * (size_t)i, k, li, ld);
    for (int j = 0; j < k; j++) {
      idx[(size_t)i * k + j] = j < c ? li[j] : -1;
      sqd[(size_t)i * k + j] = j < c ? ld[j] : FLT_MAX;
    }
  }
}

int orc_corner_coeff(const float q[3], const float nb[15], float coeff[4]) {
  float c5[5]; int r = corner_coeff(q, nb, c5); apply_weight(c5, 1.0f, coeff); return r;
}
int orc_surf_coeff(const float q[3], const float nb[15], float coeff[4]) {
  float c5[5]; int r = surf_coeff(q, nb, c5); apply_weight(c5, 1.0f, coeff); return r;
}
void orc_pose_to_affine(const float pose6[6], float T12[12]) { pose_to_affine(pose6, T12); }

void orc_jacobi_eigen_f32(const float* A, int32_t n, float* W, float* V) {
  float a[36]; memcpy(a, A, sizeof(float) * n * n);
  jacobi_eigen<float>(a, n, W, V);
}
int orc_qr_solve_f32(const float* A, int32_t n, const float* b, float* x) {
  float a[36]; memcpy(a, A, sizeof(float) * n * n); memcpy(x, b, sizeof(float) * n);
  return qr_solve<float>(a, n, x);
}
int orc_lu_inv_f32(const float* A, int32_t n, float* Ainv) {
  float a[36]; memcpy(a, A, sizeof(float) * n * n);
  for (int i = 0; i < n * n; i++) Ainv[i] = 0; for (int i = 0; i < n; i++) Ainv[i * n + i] = 1;
  return lu_solve<float>(a, n, Ainv, n);
}
void orc_plane_fit_5x3(const float* A15, float* x3) {
  float a[15], b[5] = {-1, -1, -1, -1, -1}; memcpy(a, A15, sizeof(a));
  colpiv_qr_solve_5x3(a, b, x3);
}

int orc_scan2map(const float* corner, const uint16_t* clabel, int32_t nc,
                 const float* surf, const uint16_t* slabel, int32_t ns,
                 const float* map_corner, int32_t mc, const float* map_surf, int32_t ms,
                 float pose[6], const orc_lm_params* prm, orc_lm_result* res, orc_lm_iter* log) {
  using clk = std::chrono::steady_clock;
  memset(res, 0, sizeof(*res));
  res->is_degenerate = prm->degenerate_in;
  res->deltaR = 100; res->deltaT = 100;
  // guard, odomEstimationNode.cpp:598
  if (!(nc > prm->edge_min_valid && ns > prm->surf_min_valid)) { res->status = 1; return 1; }
  const int nthr = prm->n_threads > 0 ? prm->n_threads : 1;
  auto t0 = clk::now();
  KdTree kc, ks;  // :602-603 (rebuilt on every call, as the reference does)
  kc.build(map_corner, mc);
  ks.build(map_surf, ms);
  auto t1 = clk::now();
  res->ms_build = std::chrono::duration<double, std::milli>(t1 - t0).count();

  std::vector<float> coeffC((size_t)nc * 4), coeffS((size_t)ns * 4);
  std::vector<uint8_t> flagC(nc), flagS(ns);
  std::vector<float> selP, selC;
  bool isDegenerate = prm->degenerate_in != 0;
  bool any_small = false;
  int iter = 0;
  for (; iter < prm->max_iters; iter++) {
    float T[12];
    pose_to_affine(pose, T);  // updatePointAssociateToSubMap :628-631
    // cornerOptimization :633-747
#pragma omp parallel for num_threads(nthr) schedule(dynamic, 64)
    for (int i = 0; i < nc; i++) {
      flagC[i] = 0;
      float q[3]; associate(T, corner + 4 * (size_t)i, q);
      int idx[5]; float sqd[5];
      int c = kc.knn(q, 5, idx, sqd);
      if (c == 5 && sqd[4] < prm->sqdist_gate) {  // Q6: maps with <5 pts => no correspondence
        float nb[15];
        for (int j = 0; j < 5; j++) for (int d = 0; d < 3; d++) nb[3 * j + d] = map_corner[4 * (size_t)idx[j] + d];
        float c5[5];
        if (corner_coeff(q, nb, c5)) {
          apply_weight(c5, label_weight(prm, clabel, i), &coeffC[4 * (size_t)i]);
          flagC[i] = 1;
        }
      }
    }
    // surfOptimization :749-827
#pragma omp parallel for num_threads(nthr) schedule(dynamic, 64)
    for (int i = 0; i < ns; i++) {
      flagS[i] = 0;
      float q[3]; associate(T, surf + 4 * (size_t)i, q);
      int idx[5]; float sqd[5];
      int c = ks.knn(q, 5, idx, sqd);
      if (c == 5 && sqd[4] < prm->sqdist_gate) {
        float nb[15];
        for (int j = 0; j < 5; j++) for (int d = 0; d < 3; d++) nb[3 * j + d] = map_surf[4 * (size_t)idx[j] + d];
        float c5[5];
        if (surf_coeff(q, nb, c5)) { apply_weight(c5, label_weight(prm, slabel, i), &coeffS[4 * (size_t)i]); flagS[i] = 1; }
      }
    }
    // combineOptimizationCoeffs :829-850
    selP.clear(); selC.clear();
    int nCs = 0, nSs = 0;
    for (int i = 0; i < nc; i++) if (flagC[i]) {
      selP.insert(selP.end(), corner + 4 * (size_t)i, corner + 4 * (size_t)i + 3);
      selC.insert(selC.end(), &coeffC[4 * (size_t)i], &coeffC[4 * (size_t)i] + 4); nCs++;
    }
    for (int i = 0; i < ns; i++) if (flagS[i]) {
      selP.insert(selP.end(), surf + 4 * (size_t)i, surf + 4 * (size_t)i + 3);
      selC.insert(selC.end(), &coeffS[4 * (size_t)i], &coeffS[4 * (size_t)i] + 4); nSs++;
    }
    const int nSel = nCs + nSs;
    res->n_sel_last = nSel;
    res->iters = iter + 1;
    orc_lm_iter* L = log ? &log[iter] : nullptr;
    if (L) { memset(L, 0, sizeof(*L)); L->n_sel = nSel; L->n_corner_sel = nCs; L->n_surf_sel = nSs; memcpy(L->pose, pose, 24); }
    // LMOptimization :852-974
    if (nSel < prm->min_sel) { any_small = true; continue; }
    float srx = sinf_cr(pose[1]), crx = cosf_cr(pose[1]);
    float sry = sinf_cr(pose[2]), cry = cosf_cr(pose[2]);
    float srz = sinf_cr(pose[0]), crz = cosf_cr(pose[0]);
    double AtA_d[36] = {0}, AtB_d[6] = {0};
    for (int i = 0; i < nSel; i++) {
      float px = selP[3 * (size_t)i + 1], py = selP[3 * (size_t)i + 2], pz = selP[3 * (size_t)i + 0];  // lidar -> camera
      float cx = selC[4 * (size_t)i + 1], cy = selC[4 * (size_t)i + 2], cz = selC[4 * (size_t)i + 0];
      float ci = selC[4 * (size_t)i + 3];
      float arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx +
                  (-srx * srz * px - crz * srx * py - crx * pz) * cy +
                  (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz;
      float ary = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx +
                  ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz;
      float arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx +
                  (crx * crz * px - crx * srz * py) * cy +
                  ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz;
      float row[6] = {arz, arx, ary, cz, cx, cy};
      float b = -ci;
      // cv GEMM on CV_32F accumulates in double (products of two floats are exact in double)
      for (int r = 0; r < 6; r++) {
        for (int c = r; c < 6; c++) AtA_d[r * 6 + c] += (double)row[r] * (double)row[c];
        AtB_d[r] += (double)row[r] * (double)b;
      }
    }
    float AtA[36], AtB[6], X[6];
    for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) AtA[r * 6 + c] = AtA[c * 6 + r] = (float)AtA_d[r * 6 + c];
    for (int r = 0; r < 6; r++) AtB[r] = (float)AtB_d[r];
    {
      float a[36]; memcpy(a, AtA, sizeof(a)); memcpy(X, AtB, sizeof(X));
      if (!qr_solve<float>(a, 6, X)) for (int r = 0; r < 6; r++) X[r] = 0.f;  // cv::solve leaves dst = 0 when singular
    }
    float matP[36] = {0};  // Q1: a LOCAL all-zero matP on iterations >= 1 (:880)
    if (iter == 0) {
      float a[36], E[6], V[36], V2[36];
      memcpy(a, AtA, sizeof(a));
      jacobi_eigen<float>(a, 6, E, V);
      memcpy(V2, V, sizeof(V2));
      isDegenerate = false;
      for (int i = 5; i >= 0; i--) {
        if (E[i] < prm->degenerate_eig) { for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0; isDegenerate = true; }
        else break;
      }
      // matP = matV.inv() * matV2 :945  (LU inverse, then GEMM with double accumulate)
      float Vc[36], Vinv[36]; memcpy(Vc, V, sizeof(Vc));
      for (int i = 0; i < 36; i++) Vinv[i] = 0; for (int i = 0; i < 6; i++) Vinv[i * 6 + i] = 1;
      if (!lu_solve<float>(Vc, 6, Vinv, 6)) for (int i = 0; i < 36; i++) Vinv[i] = 0;
      for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
        double s = 0; for (int k = 0; k < 6; k++) s += (double)Vinv[r * 6 + k] * (double)V2[k * 6 + c];
        matP[r * 6 + c] = (float)s;
      }
    }
    if (isDegenerate) {
      float X2[6]; memcpy(X2, X, sizeof(X2));
      for (int r = 0; r < 6; r++) { double s = 0; for (int k = 0; k < 6; k++) s += (double)matP[r * 6 + k] * (double)X2[k]; X[r] = (float)s; }
    }
    for (int r = 0; r < 6; r++) pose[r] += X[r];
    const float r2d = 57.29578f;  // pcl::rad2deg(float alpha) = alpha * 57.29578f
    float dR = std::sqrt(std::pow((double)(X[0] * r2d), 2) + std::pow((double)(X[1] * r2d), 2) + std::pow((double)(X[2] * r2d), 2));
    float dT = std::sqrt(std::pow((double)(X[3] * 100), 2) + std::pow((double)(X[4] * 100), 2) + std::pow((double)(X[5] * 100), 2));
    res->deltaR = dR; res->deltaT = dT;
    if (L) { memcpy(L->AtA, AtA, sizeof(AtA)); memcpy(L->AtB, AtB, sizeof(AtB)); memcpy(L->X, X, sizeof(X));
             memcpy(L->pose, pose, 24); L->solved = 1; L->deltaR = dR; L->deltaT = dT; }
    res->converged = (dR < prm->conv_rot_deg && dT < prm->conv_trans_cm) ? 1 : 0;
    if (res->converged && prm->early_exit) break;
  }
  res->is_degenerate = isDegenerate ? 1 : 0;
  // transformUpdate clamps :1001-1003 (IMU slerp lives in the host adapter)
  if (prm->rot_tolerance > 0) {
    pose[0] = std::min(std::max(pose[0], -prm->rot_tolerance), prm->rot_tolerance);
    pose[1] = std::min(std::max(pose[1], -prm->rot_tolerance), prm->rot_tolerance);
  }
  if (prm->z_tolerance > 0) pose[5] = std::min(std::max(pose[5], -prm->z_tolerance), prm->z_tolerance);
  res->ms_iters = std::chrono::duration<double, std::milli>(clk::now() - t1).count();
  res->status = any_small ? 2 : 0;
  return res->status;
}

}  // extern "C"
