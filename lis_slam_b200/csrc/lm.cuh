// Fused scan-to-map iteration kernel:
//   transform -> gated exact 5-NN -> {point-to-line | point-to-plane} coefficient ->
//   Jacobian row -> 27-term J^T J / J^T r reduction -> (last block) 6x6 solve + pose update.
// One launch = one Gauss-Newton iteration of EVERY registration of a batch.
//
// Reference: cornerOptimization odomEstimationNode.cpp:633-747, surfOptimization :749-827,
// combineOptimizationCoeffs :829-850 (disappears: row order does not affect A^T A),
// LMOptimization :852-974, transformUpdate clamps :1001-1003; variants B/C
// subMapOptmizationNode.cpp:1509-2001, :4485-4976.  Per-point arithmetic keeps the
// reference's fp32 expression order (translation unit is compiled with --fmad=false);
// the 27 sums are accumulated in fp64 like OpenCV's float GEMM.
#pragma once
#include "grid.cuh"
#include "smallmat.cuh"
#include "../../include/lisreg.h"

namespace lisreg {

struct MapDev { GridDev corner, surf; };

struct RegDesc {
  const float4* corner; const float4* surf;
  const uint16_t* clabel; const uint16_t* slabel;
  int nc, ns, map_slot, pad;
  const int* nc_ptr; const int* ns_ptr;   // optional device-resident counts (frame pipeline); override nc/ns
};
__device__ __forceinline__ int reg_nc(const RegDesc& d) { return d.nc_ptr ? *d.nc_ptr : d.nc; }
__device__ __forceinline__ int reg_ns(const RegDesc& d) { return d.ns_ptr ? *d.ns_ptr : d.ns; }

struct RegState {
  float pose[6];
  float T[12];
  float trig[6];   // srx crx sry cry srz crz  (rx<-pitch, ry<-yaw, rz<-roll, :862-867)
  int iter, done, converged, degenerate, any_small, n_sel_last, status;
  float deltaR, deltaT;
  int nc, ns;
};

struct LmParamsDev {
  int max_iters, early_exit;
  float gate, conv_rot, conv_trans;
  int edge_min, surf_min, min_sel;
  float degenerate_eig;
  int use_w;
  float rot_tol, z_tol;
  int degenerate_in;
  float label_score[LISREG_LUT_SIZE];
};

constexpr int LM_THREADS = 128;
constexpr int LM_NSUM = 32;   // 21 AtA (upper) + 6 AtB + nCorner + nSurf + pad

LISREG_HD __forceinline__ float sinf_cr(float x) { return (float)sin((double)x); }
LISREG_HD __forceinline__ float cosf_cr(float x) { return (float)cos((double)x); }

// pcl::getTransformation closed form (common.cpp:55-58) + LOAM trig set (:862-867)
LISREG_HD inline void state_refresh(RegState& s) {
  float roll = s.pose[0], pitch = s.pose[1], yaw = s.pose[2];
  float A = cosf_cr(yaw), B = sinf_cr(yaw), C = cosf_cr(pitch), D = sinf_cr(pitch);
  float E = cosf_cr(roll), F = sinf_cr(roll), DE = D * E, DF = D * F;
  s.T[0] = A * C;  s.T[1] = A * DF - B * E;  s.T[2] = B * F + A * DE;  s.T[3] = s.pose[3];
  s.T[4] = B * C;  s.T[5] = A * E + B * DF;  s.T[6] = B * DE - A * F;  s.T[7] = s.pose[4];
  s.T[8] = -D;     s.T[9] = C * F;           s.T[10] = C * E;          s.T[11] = s.pose[5];
  s.trig[0] = D; s.trig[1] = C;   // srx crx <- pitch
  s.trig[2] = B; s.trig[3] = A;   // sry cry <- yaw
  s.trig[4] = F; s.trig[5] = E;   // srz crz <- roll
}

// The coefficient of a correspondence splits into a part that depends on the 5 neighbours ONLY (the fitted line /
// plane and its validity test) and a part that depends on the query position.  The first part is what costs (3x3
// Jacobi eigen-decomposition, 5x3 column-pivoting QR) and it does not change while a query keeps its neighbours, so
// k_lm_resid caches it per query (GeomCache) and k_knn_* invalidate it whenever the neighbour list changes.
struct GeomCache { float4 a, b; };   // corner: a = {x1,y1,z1,x2}, b = {y2,z2,-,state}; surf: a = {pa,pb,pc,pd}, b = {-,-,-,state}
constexpr float GEOM_INVALID = 0.f, GEOM_OK = 1.f, GEOM_REJECTED = 2.f;   // state (b.w)

// cornerOptimization, neighbour part (:657-702): centroid, covariance, cv::eigen, line test, the two line points
LISREG_HD __forceinline__ bool corner_geom(const float4 (&nb)[5], float (&ln)[6]) {
  float cx = 0, cy = 0, cz = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) { cx += nb[j].x; cy += nb[j].y; cz += nb[j].z; }
  cx /= 5; cy /= 5; cz /= 5;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    float ax = nb[j].x - cx, ay = nb[j].y - cy, az = nb[j].z - cz;
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az; a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  float W[3], V[9];
  jacobi_eigen3(a11, a12, a13, a22, a23, a33, W, V);
  if (!(W[0] > 3 * W[1])) return false;
  ln[0] = (float)((double)cx + 0.1 * (double)V[0]); ln[1] = (float)((double)cy + 0.1 * (double)V[1]); ln[2] = (float)((double)cz + 0.1 * (double)V[2]);
  ln[3] = (float)((double)cx - 0.1 * (double)V[0]); ln[4] = (float)((double)cy - 0.1 * (double)V[1]); ln[5] = (float)((double)cz - 0.1 * (double)V[2]);
  return true;
}
// cornerOptimization, query part (:704-742). raw = {la,lb,lc,ld2,s}
LISREG_HD __forceinline__ bool corner_query(float x0, float y0, float z0, const float (&ln)[6], float (&raw)[5]) {
  const float x1 = ln[0], y1 = ln[1], z1 = ln[2], x2 = ln[3], y2 = ln[4], z2 = ln[5];
  float a012 = sqrtf(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                     ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                     ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
  float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
              (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
  float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) -
               (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
               (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
  float ld2 = a012 / l12;
  float s = (float)(1.0 - 0.9 * (double)fabsf(ld2));
  raw[0] = la; raw[1] = lb; raw[2] = lc; raw[3] = ld2; raw[4] = s;
  return (double)s > 0.1;
}
// cornerOptimization body after the kNN (:657-742). nb = 5 neighbours. raw = {la,lb,lc,ld2,s}
LISREG_HD __forceinline__ bool corner_coeff(float x0, float y0, float z0, const float4 (&nb)[5], float (&raw)[5]) {
  float ln[6];
  if (!corner_geom(nb, ln)) return false;
  return corner_query(x0, y0, z0, ln, raw);
}

// surfOptimization, neighbour part (:776-806): plane fit A0 x = -1 (Eigen colPivHouseholderQr), normalisation, 5-point test
LISREG_HD __forceinline__ bool surf_geom(const float4 (&nb)[5], float (&pl)[4]) {
  float A0[15], X0[3];
  const float B0[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
#pragma unroll
  for (int j = 0; j < 5; j++) { A0[3 * j] = nb[j].x; A0[3 * j + 1] = nb[j].y; A0[3 * j + 2] = nb[j].z; }
  colpiv_qr_solve_5x3(A0, B0, X0);
  float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
  float ps = sqrtf(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  bool valid = true;
#pragma unroll
  for (int j = 0; j < 5; j++)
    if ((double)fabsf(pa * nb[j].x + pb * nb[j].y + pc * nb[j].z + pd) > 0.2) valid = false;
  pl[0] = pa; pl[1] = pb; pl[2] = pc; pl[3] = pd;
  return valid;
}
// surfOptimization, query part (:808-821). raw = {pa,pb,pc,pd2,s}
LISREG_HD __forceinline__ bool surf_query(float x0, float y0, float z0, const float (&pl)[4], float (&raw)[5]) {
  const float pa = pl[0], pb = pl[1], pc = pl[2], pd = pl[3];
  float pd2 = pa * x0 + pb * y0 + pc * z0 + pd;
  float s = (float)(1.0 - 0.9 * (double)fabsf(pd2) / (double)sqrtf(sqrtf(x0 * x0 + y0 * y0 + z0 * z0)));
  raw[0] = pa; raw[1] = pb; raw[2] = pc; raw[3] = pd2; raw[4] = s;
  return (double)s > 0.1;
}
// surfOptimization body after the kNN (:776-821). raw = {pa,pb,pc,pd2,s}
LISREG_HD __forceinline__ bool surf_coeff(float x0, float y0, float z0, const float4 (&nb)[5], float (&raw)[5]) {
  float pl[4];
  if (!surf_geom(nb, pl)) return false;
  return surf_query(x0, y0, z0, pl, raw);
}

// scratch of the 6x6 solve; lives in SHARED memory on the device (one per solving warp)
struct SolveScratch {
  float AtA[36], a[36], V[36], V2[36], Vinv[36], matP[36];
  float AtB[6], X[6], X2[6], E[6];
  int indR[6], indC[6];
};

// LMOptimization tail (:869-973) run by one thread once all tiles of a registration are in.
LISREG_HD inline void lm_solve_tail(RegState& st, const LmParamsDev& prm, const double* sums, lisreg_lm_iter* log, SolveScratch& sc) {
  const int nC = (int)sums[27], nS = (int)sums[28], nSel = nC + nS;
  const int iter = st.iter;
  st.n_sel_last = nSel;
  if (log) {
    for (int i = 0; i < 36; i++) log->AtA[i] = 0.f;
    for (int i = 0; i < 6; i++) { log->AtB[i] = 0.f; log->X[i] = 0.f; log->pose[i] = st.pose[i]; }
    log->n_sel = nSel; log->n_corner_sel = nC; log->n_surf_sel = nS; log->solved = 0; log->deltaR = 0.f; log->deltaT = 0.f;
  }
  bool converged = false;
  if (nSel < prm.min_sel) {
    st.any_small = 1;
  } else {
    int q = 0;
    for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) { sc.AtA[r * 6 + c] = sc.AtA[c * 6 + r] = (float)sums[q]; q++; }
    for (int r = 0; r < 6; r++) sc.AtB[r] = (float)sums[21 + r];
    {   // the factorisation runs on register copies (shared-memory operands made it a chain of dependent LDS / STS)
      float qa[36], qx[6];
#pragma unroll
      for (int i = 0; i < 36; i++) qa[i] = sc.AtA[i];
#pragma unroll
      for (int i = 0; i < 6; i++) qx[i] = sc.AtB[i];
      const int ok = qr_solve<6>(qa, qx);
#pragma unroll
      for (int i = 0; i < 6; i++) sc.X[i] = ok ? qx[i] : 0.f;
    }
    for (int i = 0; i < 36; i++) sc.matP[i] = 0.f;   // Q1: local all-zero matP on iterations >= 1
    if (iter == 0) {
      for (int i = 0; i < 36; i++) sc.a[i] = sc.AtA[i];
      jacobi_eigen<6>(sc.a, sc.E, sc.V, sc.indR, sc.indC);
      for (int i = 0; i < 36; i++) sc.V2[i] = sc.V[i];
      st.degenerate = 0;
      for (int i = 5; i >= 0; i--) {
        if (sc.E[i] < prm.degenerate_eig) { for (int j = 0; j < 6; j++) sc.V2[i * 6 + j] = 0.f; st.degenerate = 1; }
        else break;
      }
      if (st.degenerate) {   // matP = matV.inv() * matV2 (:945); only consumed when degenerate
        for (int i = 0; i < 36; i++) { sc.a[i] = sc.V[i]; sc.Vinv[i] = 0.f; }
        for (int i = 0; i < 6; i++) sc.Vinv[i * 6 + i] = 1.f;
        if (!lu_solve<6, 6>(sc.a, sc.Vinv)) for (int i = 0; i < 36; i++) sc.Vinv[i] = 0.f;
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
          double s = 0; for (int k = 0; k < 6; k++) s += (double)sc.Vinv[r * 6 + k] * (double)sc.V2[k * 6 + c];
          sc.matP[r * 6 + c] = (float)s;
        }
      }
    }
    if (st.degenerate) {
      for (int i = 0; i < 6; i++) sc.X2[i] = sc.X[i];
      for (int r = 0; r < 6; r++) { double s = 0; for (int k = 0; k < 6; k++) s += (double)sc.matP[r * 6 + k] * (double)sc.X2[k]; sc.X[r] = (float)s; }
    }
    for (int r = 0; r < 6; r++) st.pose[r] += sc.X[r];
    const float r2d = 57.29578f;
    double r0 = (double)(sc.X[0] * r2d), r1 = (double)(sc.X[1] * r2d), r2 = (double)(sc.X[2] * r2d);
    double t0 = (double)(sc.X[3] * 100), t1 = (double)(sc.X[4] * 100), t2 = (double)(sc.X[5] * 100);
    float dR = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    float dT = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2);
    st.deltaR = dR; st.deltaT = dT;
    converged = ((double)dR < (double)prm.conv_rot) && ((double)dT < (double)prm.conv_trans);
    st.converged = converged ? 1 : 0;
    if (log) {
      for (int i = 0; i < 36; i++) log->AtA[i] = sc.AtA[i];
      for (int i = 0; i < 6; i++) { log->AtB[i] = sc.AtB[i]; log->X[i] = sc.X[i]; log->pose[i] = st.pose[i]; }
      log->solved = 1; log->deltaR = dR; log->deltaT = dT;
    }
  }
  st.iter = iter + 1;
  if ((converged && prm.early_exit) || st.iter >= prm.max_iters) {
    st.done = 1;
    if (prm.rot_tol > 0.f) {   // transformUpdate clamps (:1001-1003)
      st.pose[0] = fminf(fmaxf(st.pose[0], -prm.rot_tol), prm.rot_tol);
      st.pose[1] = fminf(fmaxf(st.pose[1], -prm.rot_tol), prm.rot_tol);
    }
    if (prm.z_tol > 0.f) st.pose[5] = fminf(fmaxf(st.pose[5], -prm.z_tol), prm.z_tol);
    st.status = st.any_small ? LISREG_FEW_CORRESPONDENCES : LISREG_OK;
  } else {
    state_refresh(st);
  }
}

__global__ void k_lm_init(const RegDesc* __restrict__ descs, RegState* __restrict__ states, const float* __restrict__ pose_in,
                          LmParamsDev prm, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  RegState s;
  for (int i = 0; i < 6; i++) s.pose[i] = pose_in[6 * b + i];
  s.iter = 0; s.done = 0; s.converged = 0; s.degenerate = prm.degenerate_in; s.any_small = 0; s.n_sel_last = 0; s.status = 0;
  s.deltaR = 100.f; s.deltaT = 100.f;
  const RegDesc d = descs[b];
  s.nc = reg_nc(d); s.ns = reg_ns(d);
  if (!(s.nc > prm.edge_min && s.ns > prm.surf_min)) { s.done = 1; s.status = LISREG_NOT_ENOUGH_FEATURES; }   // :598
  state_refresh(s);
  states[b] = s;
}

__global__ void k_lm_finish(const RegState* __restrict__ states, float* __restrict__ pose_out, lisreg_lm_result* __restrict__ res, int B) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const RegState s = states[b];
  lisreg_lm_result r;
  r.status = s.status; r.iters = s.iter; r.converged = s.converged; r.is_degenerate = s.degenerate;
  r.n_sel_last = s.n_sel_last; r.deltaR = s.deltaR; r.deltaT = s.deltaT;
  r.n_corner = s.nc; r.n_surf = s.ns;
  for (int i = 0; i < 6; i++) { r.pose[i] = s.pose[i]; pose_out[6 * b + i] = s.pose[i]; }
  res[b] = r;
}

// Row/column of the 27 accumulated products: lanes 0..20 = upper triangle of A^T A (row-major),
// lanes 21..26 = A^T b (column 6 of the stored row is b).
__device__ __forceinline__ void lm_pair_of_lane(int lane, int& r, int& c) {
  if (lane >= 21) { r = lane - 21; c = 6; return; }
  // row starts of the row-major upper triangle: 0, 6, 11, 15, 18, 20
  const int rr = (lane >= 6) + (lane >= 11) + (lane >= 15) + (lane >= 18) + (lane >= 20);
  const int base = rr == 0 ? 0 : rr == 1 ? 6 : rr == 2 ? 11 : rr == 3 ? 15 : rr == 4 ? 18 : 20;
  r = rr; c = rr + (lane - base);
}

constexpr int LM_MAX_TILE = 512;
constexpr int LM_ROW_STRIDE = LM_MAX_TILE + 1;   // shared-memory stride of a Jacobian row (doubles): odd => conflict-free column reads

// ---------------------------------------------------------------------------------------------------------
// One Gauss-Newton iteration = the 5-NN stage (k_knn_check -> k_knn_scan -> k_knn_shell: few registers, latency
// bound -> many resident warps) + k_lm_resid (coefficients + 27-term reduction: many registers) + k_lm_solve.
// The 5 neighbour positions of every query cross through global memory (20 B per query).
//
// The 5-NN stage exploits the temporal coherence of the iteration: between two iterations a query moves by
// centimetres or less, so its 5 nearest neighbours almost never change.  Every real search also records the
// SAFE RADIUS s of the query = a lower bound on the distance from the searched position q_ref to every map
// point that is NOT one of the 5 neighbours (the 6th best candidate visited, and the bound of everything not
// visited).  At the next iteration the query sits at q with |q - q_ref| = delta: by the triangle inequality
// every non-neighbour is farther than s - delta, so if (s - delta)^2 exceeds the largest of the 5 re-evaluated
// neighbour distances the old set IS the exact 5-NN set of q - proved, not assumed - and only its order is
// refreshed (5 gathers instead of ~25 candidates).  Queries that fail the test are searched again.
// Results are bit-identical to searching every query from scratch (LISREG_KNN_NOSKIP=1 does exactly that;
// tests/test_lm_parity.py compares the two).
//
// Three kernels, each with dense warps; the work of the next one is compacted into a global list (one slot id per
// query, appended with one warp-aggregated atomic per 32 queries; results do not depend on the list order):
//   k_knn_check  every query: coalesced state loads, 5 gathers, the proof; failures -> scan list
//   k_knn_scan   listed queries: flattened 3x3x3 block scan - the nine row streaks of a query are staged in shared
//                memory and walked as ONE candidate list, so a warp iterates max(total) times instead of
//                sum(max per row) times - with a two-candidates-ahead prefetch; queries whose 5th neighbour may
//                lie outside the block -> shell list
//   k_knn_search<SHELL>  deferred queries: flattened scan of the row streaks inside the ball that can still hold a
//                neighbour (knn6_wide_flat; sequential knn_ball_walk as the fallback for tiny cells); also gives
//                rejected queries a bound so that they are not searched again either
// ---------------------------------------------------------------------------------------------------------
struct KnnState { float x, y, z, s; };   // searched position q_ref and safe radius: > 0 accepted (bound of every non-neighbour),
                                         // < 0 rejected (-bound of the 5th-nearest distance), 0 = none: search again
constexpr float KNN_PAD = 0.1f;          // proof margin (m) the ball walk of sparse queries leaves around its result

#define KNN_INF __int_as_float(0x7f800000)   // +inf

// Sorted insert into the running top-6 (distances + positions).
// TIE = false (the hot path): candidates arrive in ascending position order, so on equal distance the resident entry
// - smaller position - stays ahead and ONE float compare per stage suffices.  The order of bit-equal distances that the
// oracle uses is (d^2, ORIGINAL index) though (grid.cuh knn_key_less); the searches therefore check their final list
// for equal neighbouring distances (knn6_has_tie) and, only then, run again with TIE = true, which fetches the original
// indices on equality.  Real clouds almost never tie, so the common path pays five compares per query.
template <bool TIE>
__device__ __forceinline__ void knn6_insert_mono(const float4* __restrict__ pts, float (&bd)[6], unsigned (&bp)[6], float d, unsigned p) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    bool keep = bd[j] <= d;
    if (TIE) { if (bd[j] == d && d < __int_as_float(0x7f800000)) keep = knn_orig(pts, bp[j]) < knn_orig(pts, p); }   // (a displaced +inf sentinel bubbles past its peers)
    const float lo_d = keep ? bd[j] : d; const unsigned lo_p = keep ? bp[j] : p;
    d = keep ? d : bd[j]; p = keep ? p : bp[j];
    bd[j] = lo_d; bp[j] = lo_p;
  }
}
// a tie that can change the 5-NN set or its order shows up as two equal neighbouring distances in the final top-6
__device__ __forceinline__ bool knn6_has_tie(const float (&bd)[6]) {
  bool t = false;
#pragma unroll
  for (int j = 0; j < 5; j++) t |= (bd[j] == bd[j + 1]) && (bd[j] < __int_as_float(0x7f800000));
  return t;
}

__device__ __forceinline__ float knn_dist2(float qx, float qy, float qz, float4 m) {
  const float dx = qx - m.x, dy = qy - m.y, dz = qz - m.z;
  float d = dx * dx; d = d + dy * dy; d = d + dz * dz;   // FLANN L2 functor op order, no FMA
  return d;
}

// walks the nr candidate ranges staged in this thread's column of the shared range table (stride LM_THREADS) as ONE
// list.  Software pipeline: c0 is processed while c1 and c2 are in flight.  adv() steps the cursor (p, e, k) to the
// next candidate of the flattened list; past the end it keeps returning the last valid position (harmless
// re-load) and `left` counts what is really there.
template <bool TIE>
__device__ __forceinline__ void knn6_scan_ranges(const float4* __restrict__ pts, const uint2* rng, int nr, float qx, float qy, float qz,
                                                 float (&bd)[6], unsigned (&bp)[6]) {
  if (nr <= 0) return;
  int left = 0;
  for (int k = 0; k < nr; k++) { const uint2 r = rng[k * LM_THREADS]; left += (int)(r.y - r.x); }
  uint2 r0 = rng[0];
  unsigned p = r0.x, e = r0.y; int k = 1;
  auto adv = [&]() {
    unsigned np = p + 1;
    if (np == e && k < nr) { const uint2 r = rng[k * LM_THREADS]; k++; np = r.x; e = r.y; }
    else if (np == e) { np = p; e = p + 1; }     // stay on the last candidate
    p = np;
  };
  unsigned p0 = p; float4 c0 = __ldg(&pts[p]); adv();
  unsigned p1 = p; float4 c1 = __ldg(&pts[p]); adv();
  while (left > 0) {
    const unsigned p2 = p; const float4 c2 = __ldg(&pts[p]); adv();
    const float d = knn_dist2(qx, qy, qz, c0);
    if (TIE ? d <= bd[5] : d < bd[5]) knn6_insert_mono<TIE>(pts, bd, bp, d, p0);
    c0 = c1; p0 = p1; c1 = c2; p1 = p2;
    left--;
  }
}

// flattened 3x3x3 block scan; rng = this thread's column of the shared range table (stride LM_THREADS, 10 rows:
// up to nine non-empty row streaks and a terminator).  Returns true when the outer shells are needed; lb = lower
// bound of everything outside the block.
__device__ __forceinline__ bool knn6_block_flat(const GridDev& g, float qx, float qy, float qz, float gate,
                                                float (&bd)[6], unsigned (&bp)[6], uint2* rng, float& lb) {
#pragma unroll
  for (int j = 0; j < 6; j++) { bd[j] = KNN_INF; bp[j] = 0xffffffffu; }
  lb = KNN_INF;
  if (g.n <= 0) return false;
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
  const float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
  const float minf = fminf(fminf(fminf(rx, 1.f - rx), fminf(ry, 1.f - ry)), fminf(rz, 1.f - rz));
  // the x extent is the same for the nine rows; a row is ONE contiguous streak of the sorted point array
  const int xa = cx - 1 > 0 ? cx - 1 : 0, xb = cx + 1 < g.nx - 1 ? cx + 1 : g.nx - 1;
  int nr = 0;
  if (xa <= xb) {
    const uint32_t* __restrict__ cs = g.cell_start;
#pragma unroll
    for (int r = 0; r < 9; r++) {
      const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
      if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
        const int rowbase = (z * g.ny + y) * g.nx;
        const uint32_t b = __ldg(&cs[rowbase + xa]), e = __ldg(&cs[rowbase + xb + 1]);
        if (e > b) { rng[nr * LM_THREADS] = make_uint2(b, e); nr++; }
      }
    }
  }
  knn6_scan_ranges<false>(g.pts, rng, nr, qx, qy, qz, bd, bp);
  if (knn6_has_tie(bd)) {                                        // rare: order the equal distances by original index
#pragma unroll
    for (int j = 0; j < 6; j++) { bd[j] = KNN_INF; bp[j] = 0xffffffffu; }
    knn6_scan_ranges<true>(g.pts, rng, nr, qx, qy, qz, bd, bp);
  }
  lb = knn_block_lb(g, minf);
  const float lb2 = lb * lb;
  return !(bd[4] < lb2 || lb2 >= gate);
}

// Deferred queries (the 5th neighbour may lie outside the 3x3x3 block): flattened scan of every row streak that
// intersects the ball of radius Rn = min(sqrt(d5 of the block), sqrt(gate)) + KNN_PAD around the query - rows pruned
// by their distance in y / z, streaks clipped in x - over the (2T+1)^2 rows of the Chebyshev-T block that contains the
// ball.  Exact (the 5-NN of the block bound the true 5th distance; beyond the gate nothing matters) and every map
// point NOT visited is farther than Rn, which the caller records as the bound.  KNN_WIDE_T = largest T staged in
// shared memory (returns false beyond: the caller falls back to the sequential ball walk).
constexpr int KNN_WIDE_T = 3;
constexpr int KNN_WIDE_ROWS = 48;   // staged row streaks per query (48 KB of static shared memory per 128 threads); the 49 rows of T = 3
                                    // never all intersect the ball (its corner rows are farther than any admissible radius)
__device__ __forceinline__ bool knn6_wide_flat(const GridDev& g, float qx, float qy, float qz, float gate, float d5_block,
                                               float (&bd)[6], unsigned (&bp)[6], uint2* rng, float& lbu) {
#pragma unroll
  for (int j = 0; j < 6; j++) { bd[j] = KNN_INF; bp[j] = 0xffffffffu; }
  lbu = KNN_INF;
  if (g.n <= 0) return true;
  const float eps = 1e-3f;
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
  const float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
  const float minf = fminf(fminf(fminf(rx, 1.f - rx), fminf(ry, 1.f - ry)), fminf(rz, 1.f - rz));
  const float Rn = sqrtf(fminf(d5_block, gate)) + KNN_PAD;
  // smallest T whose outside is farther than Rn: ((T + minf - eps) h >= Rn)
  int T = (int)ceilf(Rn * g.inv_h - minf + eps);
  if (T < 1) T = 1;
  if (T > KNN_WIDE_T) return false;
  int nr = 0;
  const uint32_t* __restrict__ cs = g.cell_start;
  for (int z = cz - T; z <= cz + T; z++) {
    if (z < 0 || z >= g.nz) continue;
    float dz = z < cz ? fz - (float)(z + 1) : (z > cz ? (float)z - fz : 0.f);
    dz = dz - eps > 0.f ? dz - eps : 0.f;
    for (int y = cy - T; y <= cy + T; y++) {
      if (y < 0 || y >= g.ny) continue;
      float dy = y < cy ? fy - (float)(y + 1) : (y > cy ? (float)y - fy : 0.f);
      dy = dy - eps > 0.f ? dy - eps : 0.f;
      const float rem = Rn * Rn - (dy * dy + dz * dz) * g.h * g.h;
      if (rem < 0.f) continue;                                   // the whole row is out of reach
      const float rad = sqrtf(rem) * g.inv_h + eps;              // reach along x, in cells
      int xa = (int)floorf(fx - rad), xb = (int)floorf(fx + rad);
      xa = xa > cx - T ? xa : cx - T; xb = xb < cx + T ? xb : cx + T;
      xa = xa > 0 ? xa : 0; xb = xb < g.nx - 1 ? xb : g.nx - 1;
      if (xa > xb) continue;
      const int rowbase = (z * g.ny + y) * g.nx;
      const uint32_t b = __ldg(&cs[rowbase + xa]), e = __ldg(&cs[rowbase + xb + 1]);
      if (e > b) { if (nr == KNN_WIDE_ROWS) return false; rng[nr * LM_THREADS] = make_uint2(b, e); nr++; }
    }
  }
  knn6_scan_ranges<false>(g.pts, rng, nr, qx, qy, qz, bd, bp);
  if (knn6_has_tie(bd)) {
#pragma unroll
    for (int j = 0; j < 6; j++) { bd[j] = KNN_INF; bp[j] = 0xffffffffu; }
    knn6_scan_ranges<true>(g.pts, rng, nr, qx, qy, qz, bd, bp);
  }
  lbu = Rn;
  return true;
}

// writes the result of a real search: neighbour positions (or -1 = rejected) and the refreshed state
__device__ __forceinline__ void knn_commit(int* __restrict__ tnbr, KnnState* __restrict__ tstate, GeomCache* __restrict__ tgeom, int tile_pts, int l,
                                           float qx, float qy, float qz, float gate,
                                           float d5, float d6, float lbu, const unsigned (&pos)[5]) {
  KnnState st; st.x = qx; st.y = qy; st.z = qz; st.s = 0.f;
  tgeom[l].b.w = GEOM_INVALID;                 // new neighbour list: the cached line / plane is stale
  if (d5 < gate) {
#pragma unroll
    for (int j = 0; j < 5; j++) tnbr[j * tile_pts + l] = (int)pos[j];
    // every non-neighbour is farther than min(6th best visited, bound of the unvisited); the factor absorbs the
    // fp32 rounding of the distance evaluation and of the square root
    st.s = fminf(sqrtf(d6), lbu) * 0.99999f;
  } else {
    // rejected (5th neighbour outside the gate): -s = lower bound of the 5th-nearest distance, so that the next
    // iteration can prove "still rejected" without searching
    tnbr[l] = -1;
    st.s = -(fminf(sqrtf(d5), lbu) * 0.99999f);
  }
  tstate[l] = st;
}

// appends the slots of the lanes with `need` to a global list: one atomic per warp
__device__ __forceinline__ void knn_list_append(bool need, unsigned slot, unsigned* __restrict__ list, int* __restrict__ counter) {
  const unsigned m = __ballot_sync(0xffffffffu, need);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (need) list[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

// slot id of a query = position in the [B][max_tiles][tile_pts] slot space (tile_pts = 1 << tile_shift)
struct KnnSlot { int b, tile, l, q; };
__device__ __forceinline__ KnnSlot knn_slot_decode(unsigned slot, int max_tiles, int tile_shift) {
  KnnSlot s;
  const unsigned bt = slot >> tile_shift;
  s.l = (int)(slot & ((1u << tile_shift) - 1u));
  s.b = (int)(bt / (unsigned)max_tiles); s.tile = (int)(bt - (unsigned)s.b * (unsigned)max_tiles);
  s.q = (s.tile << tile_shift) + s.l;
  return s;
}

// ---- k_knn_check: grid (tiles, B), one thread per query ----
__global__ void __launch_bounds__(LM_THREADS)
k_knn_check(const RegDesc* __restrict__ descs, const RegState* __restrict__ states, const MapDev* __restrict__ maps,
            float gate, int* __restrict__ nbr, KnnState* __restrict__ kstate, GeomCache* __restrict__ geom, unsigned* __restrict__ scan_list,
            int* __restrict__ counter, int max_tiles, int tile_shift, int use_state) {
  const int b = blockIdx.y, tid = threadIdx.x;
  const int tile_pts = 1 << tile_shift;
  __shared__ RegDesc sd;
  __shared__ float sT[12];
  __shared__ int sdone;
  if (tid == 0) { sd = descs[b]; sd.nc = states[b].nc; sd.ns = states[b].ns; sdone = states[b].done; }
  if (tid < 12) sT[tid] = states[b].T[tid];
  __syncthreads();
  if (sdone) return;
  const int n = sd.nc + sd.ns;
  const int ntiles = (n + tile_pts - 1) >> tile_shift;
  const MapDev& mp = maps[sd.map_slot];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int q0 = tile << tile_shift;
    const int qn = min(n - q0, tile_pts);
    const size_t bt = (size_t)b * max_tiles + tile;
    int* tnbr = nbr + bt * 5 * tile_pts;         // [5][tile_pts]
    KnnState* tstate = kstate + bt * tile_pts;   // [tile_pts]
    GeomCache* tgeom = geom + bt * tile_pts;
    for (int l = tid; l < tile_pts; l += LM_THREADS) {   // tile_pts is a multiple of LM_THREADS: warps stay whole
      bool need = false;
      if (l < qn) {
        need = true;
        if (use_state) {
          // state, query and neighbour ids are independent loads: issue them together (most queries take the proof path)
          const float4 ref = *reinterpret_cast<const float4*>(&tstate[l]);
          const int q = q0 + l;
          const bool is_corner = q < sd.nc;
          const float4 p = is_corner ? __ldg(&sd.corner[q]) : __ldg(&sd.surf[q - sd.nc]);
          int pos[5];
#pragma unroll
          for (int j = 0; j < 5; j++) pos[j] = tnbr[j * tile_pts + l];
          if (ref.w != 0.f) {
            const float x0 = sT[0] * p.x + sT[1] * p.y + sT[2] * p.z + sT[3];
            const float y0 = sT[4] * p.x + sT[5] * p.y + sT[6] * p.z + sT[7];
            const float z0 = sT[8] * p.x + sT[9] * p.y + sT[10] * p.z + sT[11];
            const float mx = x0 - ref.x, my = y0 - ref.y, mz = z0 - ref.z;
            const float delta = sqrtf(mx * mx + my * my + mz * mz) * 1.00001f + 1e-6f;   // upper bound of the move
            const float r = fabsf(ref.w) - delta;
            if (r > 0.f && ref.w < 0.f) {
              // rejected last time: the 5th-nearest point was farther than |s|, so it is still farther than r
              if (r * r * 0.99999f > gate) need = false;          // still rejected; nbr[l] is already -1
            } else if (r > 0.f) {
              const float4* __restrict__ pts = is_corner ? mp.corner.pts : mp.surf.pts;
              knn_key key[5];
              float bmax = 0.f;
#pragma unroll
              for (int j = 0; j < 5; j++) {
                const float d = knn_dist2(x0, y0, z0, __ldg(&pts[pos[j]]));
                bmax = fmaxf(bmax, d);
                key[j] = ((knn_key)(unsigned)__float_as_int(d) << 32) | (knn_key)(unsigned)pos[j];
              }
              if (r * r * 0.99999f > bmax) {     // no other map point can be as close as the farthest neighbour
                need = false;
                if (bmax < gate) {
                  // refresh the (d^2, original index) order: 9-comparator network on (d^2, position) keys, then - only when
                  // two of the five distances are bit-equal - an insertion sort with the full order
#define LISREG_CSWAP(i, j) { const knn_key a_ = key[i], b_ = key[j]; const bool sw_ = b_ < a_; key[i] = sw_ ? b_ : a_; key[j] = sw_ ? a_ : b_; }
                  LISREG_CSWAP(0, 1) LISREG_CSWAP(3, 4) LISREG_CSWAP(2, 4) LISREG_CSWAP(2, 3) LISREG_CSWAP(1, 4)
                  LISREG_CSWAP(0, 3) LISREG_CSWAP(0, 2) LISREG_CSWAP(1, 3) LISREG_CSWAP(1, 2)
#undef LISREG_CSWAP
                  bool tie = false;
#pragma unroll
                  for (int j = 0; j < 4; j++) tie |= (unsigned)(key[j] >> 32) == (unsigned)(key[j + 1] >> 32);
                  if (tie) {
#pragma unroll
                    for (int j = 1; j < 5; j++)
#pragma unroll
                      for (int k = j; k > 0; k--)
                        if (knn_key_less(pts, key[k], key[k - 1])) { const knn_key t_ = key[k]; key[k] = key[k - 1]; key[k - 1] = t_; }
                  }
                  bool changed = false;
#pragma unroll
                  for (int j = 0; j < 5; j++) changed |= knn_key_pos(key[j]) != pos[j];
                  if (changed) {
#pragma unroll
                    for (int j = 0; j < 5; j++) tnbr[j * tile_pts + l] = knn_key_pos(key[j]);
                    tgeom[l].b.w = GEOM_INVALID;   // same set, new order: the fits round differently
                  }
                } else {
                  // the 5th neighbour left the gate: rejected now.  Its distance (>= sqrt(gate), and the neighbours are
                  // the exact 5-NN) is the bound of the 5th-nearest distance from HERE
                  tnbr[l] = -1;
                  KnnState ns; ns.x = x0; ns.y = y0; ns.z = z0; ns.s = -(sqrtf(bmax) * 0.99999f);
                  tstate[l] = ns;
                }
              }
            }
          }
        }
      }
      knn_list_append(need, (unsigned)(bt * tile_pts + l), scan_list, counter);
    }
  }
}

// ---- k_knn_scan / k_knn_shell: grid-stride over a slot list, one thread per listed query ----
#ifndef LM_KNN_MIN_BLOCKS
#define LM_KNN_MIN_BLOCKS 8
#endif
template <bool SHELL>
__global__ void __launch_bounds__(LM_THREADS, LM_KNN_MIN_BLOCKS)
k_knn_search(const RegDesc* __restrict__ descs, const RegState* __restrict__ states, const MapDev* __restrict__ maps,
             float gate, int* __restrict__ nbr, KnnState* __restrict__ kstate, GeomCache* __restrict__ geom, const unsigned* __restrict__ list,
             const int* __restrict__ counter, unsigned* __restrict__ shell_list, int* __restrict__ shell_counter,
             int min_list, int max_tiles, int tile_shift) {
  __shared__ uint2 s_rng[(SHELL ? KNN_WIDE_ROWS : 9) * LM_THREADS];
  const int tile_pts = 1 << tile_shift;
  const int total = *counter;
  if (total < min_list) return;             // short lists are searched by k_knn_coop
  const int lane = threadIdx.x & 31;
  for (int base = (blockIdx.x * LM_THREADS + threadIdx.x) - lane; base < total; base += gridDim.x * LM_THREADS) {
    const int i = base + lane;
    bool need_shell = false;
    unsigned slot = 0u;
    if (i < total) {
      slot = list[i];
      const KnnSlot ks = knn_slot_decode(slot, max_tiles, tile_shift);
      const RegDesc* __restrict__ d = &descs[ks.b];
      const RegState* __restrict__ st = &states[ks.b];
      const int nc = st->nc;
      const bool is_corner = ks.q < nc;
      const float4 p = is_corner ? __ldg(&d->corner[ks.q]) : __ldg(&d->surf[ks.q - nc]);
      // pointAssociateToMap (:243-258)
      const float x0 = st->T[0] * p.x + st->T[1] * p.y + st->T[2] * p.z + st->T[3];
      const float y0 = st->T[4] * p.x + st->T[5] * p.y + st->T[6] * p.z + st->T[7];
      const float z0 = st->T[8] * p.x + st->T[9] * p.y + st->T[10] * p.z + st->T[11];
      const MapDev& mp = maps[d->map_slot];
      const GridDev& g = is_corner ? mp.corner : mp.surf;
      const size_t bt = (size_t)ks.b * max_tiles + ks.tile;
      int* tnbr = nbr + bt * 5 * tile_pts;
      KnnState* tstate = kstate + bt * tile_pts;
      GeomCache* tgeom = geom + bt * tile_pts;
      if (SHELL) {
        float bd[6]; unsigned bp[6]; float lbu;
        const float d5_block = tstate[ks.l].s;      // left by the block scan that deferred this query
        if (knn6_wide_flat(g, x0, y0, z0, gate, d5_block, bd, bp, s_rng + threadIdx.x, lbu)) {
          const unsigned pos[5] = {bp[0], bp[1], bp[2], bp[3], bp[4]};
          knn_commit(tnbr, tstate, tgeom, tile_pts, ks.l, x0, y0, z0, gate, bd[4], bd[5], lbu, pos);
        } else {                                     // ball too wide for the staged rows (tiny cells): sequential walk
          knn_key best[6];
          const float lb2 = knn_ball_walk<6, 4>(g, x0, y0, z0, gate, KNN_PAD, best);
          const unsigned pos[5] = {(unsigned)knn_key_pos(best[0]), (unsigned)knn_key_pos(best[1]), (unsigned)knn_key_pos(best[2]),
                                   (unsigned)knn_key_pos(best[3]), (unsigned)knn_key_pos(best[4])};
          knn_commit(tnbr, tstate, tgeom, tile_pts, ks.l, x0, y0, z0, gate, knn_key_d(best[4]), knn_key_d(best[5]), lb2, pos);
        }
      } else {
        float bd[6]; unsigned bp[6]; float lb;
        need_shell = knn6_block_flat(g, x0, y0, z0, gate, bd, bp, s_rng + threadIdx.x, lb);
        if (!need_shell) {
          const unsigned pos[5] = {bp[0], bp[1], bp[2], bp[3], bp[4]};
          knn_commit(tnbr, tstate, tgeom, tile_pts, ks.l, x0, y0, z0, gate, bd[4], bd[5], lb, pos);
        } else {
          tstate[ks.l].s = bd[4];                    // 5th best of the block: bounds the radius of the wide scan
        }
      }
    }
    if (!SHELL) knn_list_append(need_shell, slot, shell_list, shell_counter);
  }
}

// ---- k_knn_coop: WARP per query, for SHORT scan lists (late iterations: a few hundred queries whose proof failed).
// A thread-per-query search of a short list is pure latency (one thread walks ~50-100 dependent candidates while the
// GPU idles); here the 32 lanes split the row streaks of the ball of radius sqrt(gate) + KNN_PAD (same pruning / clipping
// as knn6_wide_flat), each keeps a private top-6 (positions ascend inside a lane), and six warp arg-min rounds on
// (d^2, position) keys merge them - exact, complete (no deferral), ~10x shorter dependent chain.
// Runs only when the list is shorter than `max_list`; k_knn_search<false> runs only when it is not.
__global__ void __launch_bounds__(LM_THREADS)
k_knn_coop(const RegDesc* __restrict__ descs, const RegState* __restrict__ states, const MapDev* __restrict__ maps,
           float gate, int* __restrict__ nbr, KnnState* __restrict__ kstate, GeomCache* __restrict__ geom, const unsigned* __restrict__ list,
           const int* __restrict__ counter, int max_list, int max_tiles, int tile_shift) {
  const int total = *counter;
  if (total >= max_list) return;
  const int tile_pts = 1 << tile_shift;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * LM_THREADS + threadIdx.x) >> 5, nwarps = (gridDim.x * LM_THREADS) >> 5;
  const unsigned FULL = 0xffffffffu;
  for (int i = warp; i < total; i += nwarps) {
    const unsigned slot = list[i];
    const KnnSlot ks = knn_slot_decode(slot, max_tiles, tile_shift);
    const RegDesc* __restrict__ d = &descs[ks.b];
    const RegState* __restrict__ st = &states[ks.b];
    const int nc = st->nc;
    const bool is_corner = ks.q < nc;
    const float4 p = is_corner ? __ldg(&d->corner[ks.q]) : __ldg(&d->surf[ks.q - nc]);
    const float x0 = st->T[0] * p.x + st->T[1] * p.y + st->T[2] * p.z + st->T[3];
    const float y0 = st->T[4] * p.x + st->T[5] * p.y + st->T[6] * p.z + st->T[7];
    const float z0 = st->T[8] * p.x + st->T[9] * p.y + st->T[10] * p.z + st->T[11];
    const MapDev& mp = maps[d->map_slot];
    const GridDev& g = is_corner ? mp.corner : mp.surf;
    const size_t bt = (size_t)ks.b * max_tiles + ks.tile;
    int* tnbr = nbr + bt * 5 * tile_pts;
    KnnState* tstate = kstate + bt * tile_pts;
    GeomCache* tgeom = geom + bt * tile_pts;
    float bd[6]; unsigned bp[6];
#pragma unroll
    for (int j = 0; j < 6; j++) { bd[j] = KNN_INF; bp[j] = 0xffffffffu; }
    float lbu = KNN_INF;
    bool fallback = false;
    if (g.n > 0) {
      const float eps = 1e-3f;
      const float fx = (x0 - g.ox) * g.inv_h, fy = (y0 - g.oy) * g.inv_h, fz = (z0 - g.oz) * g.inv_h;
      const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
      const float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
      const float minf = fminf(fminf(fminf(rx, 1.f - rx), fminf(ry, 1.f - ry)), fminf(rz, 1.f - rz));
      const float Rn = sqrtf(gate) + KNN_PAD;
      int T = (int)ceilf(Rn * g.inv_h - minf + eps);
      if (T < 1) T = 1;
      const int side = 2 * T + 1, nrows = side * side;
      if (nrows > 64) fallback = true;                         // tiny cells: lane 0 walks the ball sequentially
      else {
        // rows r = lane and lane + 32 (ascending (z, y) order inside a lane => ascending positions)
#pragma unroll 1
        for (int rr = 0; rr < 2; rr++) {
          const int r = lane + 32 * rr;
          if (r >= nrows) break;
          const int z = cz - T + r / side, y = cy - T + r % side;
          if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) continue;
          float dz = z < cz ? fz - (float)(z + 1) : (z > cz ? (float)z - fz : 0.f);
          dz = dz - eps > 0.f ? dz - eps : 0.f;
          float dy = y < cy ? fy - (float)(y + 1) : (y > cy ? (float)y - fy : 0.f);
          dy = dy - eps > 0.f ? dy - eps : 0.f;
          const float rem = Rn * Rn - (dy * dy + dz * dz) * g.h * g.h;
          if (rem < 0.f) continue;
          const float rad = sqrtf(rem) * g.inv_h + eps;
          int xa = (int)floorf(fx - rad), xb = (int)floorf(fx + rad);
          xa = xa > cx - T ? xa : cx - T; xb = xb < cx + T ? xb : cx + T;
          xa = xa > 0 ? xa : 0; xb = xb < g.nx - 1 ? xb : g.nx - 1;
          if (xa > xb) continue;
          const int rowbase = (z * g.ny + y) * g.nx;
          const uint32_t b = __ldg(&g.cell_start[rowbase + xa]), e = __ldg(&g.cell_start[rowbase + xb + 1]);
          for (uint32_t c = b; c < e; c++) {
            const float dd = knn_dist2(x0, y0, z0, __ldg(&g.pts[c]));
            if (dd <= bd[5]) knn6_insert_mono<true>(g.pts, bd, bp, dd, c);     // short lists only: always tie-aware
          }
        }
        lbu = Rn;
      }
    }
    unsigned pos[5] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    float d5 = KNN_INF, d6 = KNN_INF;
    if (fallback) {
      if (lane == 0) {
        knn_key best[6];
        lbu = knn_ball_walk<6, 4>(g, x0, y0, z0, gate, KNN_PAD, best);
#pragma unroll
        for (int j = 0; j < 5; j++) pos[j] = (unsigned)knn_key_pos(best[j]);
        d5 = knn_key_d(best[4]); d6 = knn_key_d(best[5]);
      }
    } else {
      // merge the 32 private ascending lists: six rounds of warp arg-min on (d^2 bits, position)
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const unsigned hd = __float_as_uint(bd[0]);                       // d^2 >= 0: bit order == value order; +inf = empty
        const unsigned md = __reduce_min_sync(FULL, hd);
        const bool cand = hd == md && bp[0] != 0xffffffffu;
        unsigned mpos;
        if (__popc(__ballot_sync(FULL, cand)) <= 1) mpos = __reduce_min_sync(FULL, cand ? bp[0] : 0xffffffffu);
        else {                                                            // bit-equal heads: the smaller ORIGINAL index wins
          const unsigned o = cand ? (unsigned)knn_orig(g.pts, bp[0]) : 0xffffffffu;
          const unsigned mo = __reduce_min_sync(FULL, o);
          mpos = __reduce_min_sync(FULL, (cand && o == mo) ? bp[0] : 0xffffffffu);
        }
        if (cand && bp[0] == mpos) {                                      // the owning lane pops its head
#pragma unroll
          for (int j = 0; j < 5; j++) { bd[j] = bd[j + 1]; bp[j] = bp[j + 1]; }
          bd[5] = KNN_INF; bp[5] = 0xffffffffu;
        }
        if (k < 5) pos[k] = mpos;
        if (k == 4) d5 = __uint_as_float(md);
        if (k == 5) d6 = __uint_as_float(md);
      }
    }
    if (lane == 0) knn_commit(tnbr, tstate, tgeom, tile_pts, ks.l, x0, y0, z0, gate, d5, d6, lbu, pos);
  }
}

#ifndef LM_RESID_MIN_BLOCKS
#define LM_RESID_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(LM_THREADS, LM_RESID_MIN_BLOCKS)
k_lm_resid(const RegDesc* __restrict__ descs, const RegState* __restrict__ states, const MapDev* __restrict__ maps,
           LmParamsDev prm, const int* __restrict__ nbr, GeomCache* __restrict__ geom, double* __restrict__ partials, int max_tiles, int tile_pts) {
  const int b = blockIdx.y, tid = threadIdx.x;
  __shared__ RegDesc sd;
  __shared__ float sT[12], sTrig[6];
  __shared__ int sdone;
  // Jacobian rows as doubles (converted once per element instead of once per product), row stride LM_ROW_STRIDE: the 27
  // lanes of the reduction read up to 7 different rows at the same column, the odd stride puts them in different banks
  __shared__ double s_row[7 * LM_ROW_STRIDE];
  __shared__ double swarp[LM_THREADS / 32][LM_NSUM];
  if (tid == 0) { sd = descs[b]; sd.nc = states[b].nc; sd.ns = states[b].ns; sdone = states[b].done; }
  if (tid < 12) sT[tid] = states[b].T[tid];
  if (tid >= 32 && tid < 38) sTrig[tid - 32] = states[b].trig[tid - 32];
  __syncthreads();
  if (sdone) return;
  const int n = sd.nc + sd.ns;
  const int ntiles = (n + tile_pts - 1) / tile_pts;
  const MapDev& mp = maps[sd.map_slot];
  const float srx = sTrig[0], crx = sTrig[1], sry = sTrig[2], cry = sTrig[3], srz = sTrig[4], crz = sTrig[5];
  const int lane = tid & 31, wid = tid >> 5;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int q0 = tile * tile_pts;
    const int qn = min(n - q0, tile_pts);
    const int* tnbr = nbr + ((size_t)b * max_tiles + tile) * 5 * tile_pts;
    GeomCache* tgeom = geom + ((size_t)b * max_tiles + tile) * tile_pts;
    int cntC = 0, cntS = 0;
    for (int l = tid; l < tile_pts; l += LM_THREADS) {
      float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int pos0 = l < qn ? tnbr[l] : -1;
      if (pos0 >= 0) {
        const int q = q0 + l;
        const bool is_corner = q < sd.nc;
        const float4 p = is_corner ? __ldg(&sd.corner[q]) : __ldg(&sd.surf[q - sd.nc]);
        GeomCache gc = tgeom[l];
        const float x0 = sT[0] * p.x + sT[1] * p.y + sT[2] * p.z + sT[3];
        const float y0 = sT[4] * p.x + sT[5] * p.y + sT[6] * p.z + sT[7];
        const float z0 = sT[8] * p.x + sT[9] * p.y + sT[10] * p.z + sT[11];
        if (gc.b.w == GEOM_INVALID) {
          // the neighbour list changed since the fit was cached: gather, fit the line / plane, cache it
          const float4* __restrict__ pts = is_corner ? mp.corner.pts : mp.surf.pts;
          float4 nb[5];
          nb[0] = __ldg(&pts[pos0]);
#pragma unroll
          for (int j = 1; j < 5; j++) nb[j] = __ldg(&pts[tnbr[j * tile_pts + l]]);
          if (is_corner) {
            float ln[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const bool okg = corner_geom(nb, ln);
            gc.a = make_float4(ln[0], ln[1], ln[2], ln[3]); gc.b = make_float4(ln[4], ln[5], 0.f, okg ? GEOM_OK : GEOM_REJECTED);
          } else {
            float pl[4];
            const bool okg = surf_geom(nb, pl);
            gc.a = make_float4(pl[0], pl[1], pl[2], pl[3]); gc.b = make_float4(0.f, 0.f, 0.f, okg ? GEOM_OK : GEOM_REJECTED);
          }
          tgeom[l] = gc;
        }
        float raw[5];
        bool ok = false;
        if (gc.b.w == GEOM_OK) {
          if (is_corner) { const float ln[6] = {gc.a.x, gc.a.y, gc.a.z, gc.a.w, gc.b.x, gc.b.y}; ok = corner_query(x0, y0, z0, ln, raw); }
          else { const float pl[4] = {gc.a.x, gc.a.y, gc.a.z, gc.a.w}; ok = surf_query(x0, y0, z0, pl, raw); }
        }
        if (ok) {
          float w = 1.0f;
          if (prm.use_w) {
            const uint16_t* lab = is_corner ? sd.clabel : sd.slabel;
            unsigned lb = lab ? lab[is_corner ? q : q - sd.nc] : 0u;
            float sc = lb < LISREG_LUT_SIZE ? prm.label_score[lb] : 0.f;
            w = (float)(2.0 - (double)sc);
          }
          const float ws = w * raw[4];
          const float c_x = ws * raw[0], c_y = ws * raw[1], c_z = ws * raw[2], c_i = ws * raw[3];
          if (is_corner) cntC++; else cntS++;
          const float px = p.y, py = p.z, pz = p.x;
          const float cx = c_y, cy = c_z, cz = c_x;
          const float arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx +
                            (-srx * srz * px - crz * srx * py - crx * pz) * cy +
                            (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz;
          const float ary = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx +
                            ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz;
          const float arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx +
                            (crx * crz * px - crx * srz * py) * cy +
                            ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz;
          row[0] = arz; row[1] = arx; row[2] = ary; row[3] = cz; row[4] = cx; row[5] = cy; row[6] = -c_i;
        }
      }
#pragma unroll
      for (int k = 0; k < 7; k++) s_row[k * LM_ROW_STRIDE + l] = (double)row[k];
    }
    __syncthreads();
    {
      int r, c;
      lm_pair_of_lane(lane < 27 ? lane : 0, r, c);
      const int per_warp = tile_pts / (LM_THREADS / 32);
      const double* __restrict__ ra = s_row + r * LM_ROW_STRIDE + wid * per_warp;
      const double* __restrict__ rc = s_row + c * LM_ROW_STRIDE + wid * per_warp;
      double v = 0.0;
      for (int i = 0; i < per_warp; i++) v += ra[i] * rc[i];
      if (lane < 27) swarp[wid][lane] = v;
      int cc = cntC, s2 = cntS;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { cc += __shfl_down_sync(0xffffffffu, cc, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
      if (lane == 0) { swarp[wid][27] = (double)cc; swarp[wid][28] = (double)s2; }
    }
    __syncthreads();
    double* mypart = partials + ((size_t)b * max_tiles + tile) * LM_NSUM;
    if (tid < 29) {
      double v = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < LM_THREADS / 32; w2++) v += swarp[w2][tid];
      mypart[tid] = v;
    }
    __syncthreads();
  }
}

// LMOptimization tail: one warp per registration sums the tile partials in tile order (fixed
// order => bit-reproducible) and lane 0 runs the 6x6 solve / degeneracy / pose update.
// Kept out of k_lm_resid so that the hot kernel has no large stack frame or cold code (measured in round 2: letting the
// last block of k_lm_resid run the solve costs 0.3 ms per 256-frame step - the 544-byte frame of the solve lands in the
// hot kernel - and saves nothing on the single-frame path).
constexpr int LM_SOLVE_THREADS = 128;
constexpr int LM_SOLVE_CHUNK = 32;       // tiles staged per round
// One block per registration: the tile partials are staged in shared memory by all threads (coalesced, every load in
// flight at once - a lane that walks the tiles itself waits one L2 round trip per tile, 20 us for a 7 k-point frame), then 29
// lanes add them up in tile order (fixed order => bit-reproducible) and lane 0 runs the 6x6 solve.
__global__ void __launch_bounds__(LM_SOLVE_THREADS)
k_lm_solve(const RegDesc* __restrict__ descs, RegState* __restrict__ states, LmParamsDev prm,
           const double* __restrict__ partials, lisreg_lm_iter* __restrict__ logs, int max_tiles, int tile_pts, int B) {
  __shared__ double s_part[LM_SOLVE_CHUNK * LM_NSUM];
  __shared__ double stot[LM_NSUM];
  __shared__ SolveScratch ssc;
  const int b = blockIdx.x;
  if (b >= B) return;
  if (states[b].done) return;
  const int n = states[b].nc + states[b].ns;
  const int ntiles = (n + tile_pts - 1) / tile_pts;
  const double* base = partials + (size_t)b * max_tiles * LM_NSUM;
  double v = 0.0;
  for (int t0 = 0; t0 < ntiles; t0 += LM_SOLVE_CHUNK) {
    const int cnt = min(LM_SOLVE_CHUNK, ntiles - t0) * LM_NSUM;
    for (int i = threadIdx.x; i < cnt; i += LM_SOLVE_THREADS) s_part[i] = base[(size_t)t0 * LM_NSUM + i];
    __syncthreads();
    if (threadIdx.x < 29) for (int i = threadIdx.x; i < cnt; i += LM_NSUM) v += s_part[i];
    __syncthreads();
  }
  if (threadIdx.x < 29) stot[threadIdx.x] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    RegState st = states[b];
    lisreg_lm_iter* lg = logs ? &logs[(size_t)b * LISREG_MAX_ITERS + st.iter] : nullptr;
    lm_solve_tail(st, prm, stot, lg, ssc);
    states[b] = st;
  }
}

// self-test of the small dense routines on the device (lisreg_selftest_smallmat)
__global__ void k_selftest_smallmat(const float* __restrict__ A36, const float* __restrict__ b6, float* __restrict__ out) {
  __shared__ SolveScratch sc;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < 36; i++) sc.a[i] = A36[i];
  jacobi_eigen<6>(sc.a, sc.E, sc.V, sc.indR, sc.indC);
  for (int i = 0; i < 6; i++) out[i] = sc.E[i];
  for (int i = 0; i < 36; i++) out[6 + i] = sc.V[i];
  for (int i = 0; i < 36; i++) sc.a[i] = A36[i];
  for (int i = 0; i < 6; i++) sc.X[i] = b6[i];
  int ok = qr_solve<6>(sc.a, sc.X);
  for (int i = 0; i < 6; i++) out[42 + i] = sc.X[i];
  out[48] = (float)ok;
  for (int i = 0; i < 36; i++) { sc.a[i] = A36[i]; sc.Vinv[i] = 0.f; }
  for (int i = 0; i < 6; i++) sc.Vinv[i * 6 + i] = 1.f;
  ok = lu_solve<6, 6>(sc.a, sc.Vinv);
  for (int i = 0; i < 36; i++) out[49 + i] = sc.Vinv[i];
  out[85] = (float)ok;
  // register-only 3x3 Jacobi on the leading 3x3 block
  float W3[3], V3[9];
  jacobi_eigen3(A36[0], A36[1], A36[2], A36[7], A36[8], A36[14], W3, V3);
  for (int i = 0; i < 3; i++) out[86 + i] = W3[i];
  for (int i = 0; i < 9; i++) out[89 + i] = V3[i];
}

// map-based dynamic-object removal (map_scan_feature_pts_distance_removal, subMap.h:1063-1098): a feature point
// survives outside the centre disc, or when its squared distance d2 to the nearest map point satisfies
// (d2 > near^2 && d2 < dyn_min^2) || d2 > dyn_max^2.  The 1-NN search is gated just above the largest finite
// threshold: a miss means d2 exceeds every threshold, which decides the test without knowing d2.
__global__ void k_map_distance_filter(GridDev g, const float4* __restrict__ feat, int n, float center_r2, float near2, float dmin2,
                                      float dmax2, float gate, unsigned char* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(&feat[i]);
  bool k;
  if (p.x * p.x + p.y * p.y > center_r2) k = true;
  else if (g.n <= 0) k = true;
  else {
    knn_key best[1];
    knn_grid<1>(g, p.x, p.y, p.z, gate, best);
    const float d2 = knn_key_pos(best[0]) >= 0 && knn_key_d(best[0]) < gate ? knn_key_d(best[0]) : KNN_INF;
    k = (d2 > near2 && d2 < dmin2) || d2 > dmax2;
  }
  keep[i] = k ? 1 : 0;
}

// stand-alone exact 5-NN (tests / lisreg_knn5): the same flattened block scan + tracked shells as k_knn_search.
// safe (nullable, nq): the safe radius the search would record for the query.
__global__ void __launch_bounds__(LM_THREADS)
k_knn5(GridDev g, const float4* __restrict__ q, int nq, float gate, int* __restrict__ idx, float* __restrict__ sqd, float* __restrict__ safe) {
  __shared__ uint2 s_rng[KNN_WIDE_ROWS * LM_THREADS];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  float4 p = q[i];
  float bd[6]; unsigned bp[6]; float lb;
  if (knn6_block_flat(g, p.x, p.y, p.z, gate, bd, bp, s_rng + threadIdx.x, lb)) {
    const float d5_block = bd[4];
    // odd queries take the sequential ball walk so that the tests cover both deferred paths
    if ((i & 1) || !knn6_wide_flat(g, p.x, p.y, p.z, gate, d5_block, bd, bp, s_rng + threadIdx.x, lb)) {
      knn_key best[6];
      lb = knn_ball_walk<6, 4>(g, p.x, p.y, p.z, gate, KNN_PAD, best);
#pragma unroll
      for (int j = 0; j < 6; j++) { bd[j] = knn_key_d(best[j]); bp[j] = (unsigned)knn_key_pos(best[j]); }
    }
  }
  for (int j = 0; j < 5; j++) {
    const bool ok = bp[j] != 0xffffffffu && bd[j] < gate;
    idx[5 * i + j] = ok ? __float_as_int(g.pts[bp[j]].w) : -1;
    sqd[5 * i + j] = ok ? bd[j] : FLT_MAX;
  }
  if (safe) safe[i] = bd[4] < gate ? fminf(sqrtf(bd[5]), lb) * 0.99999f : 0.f;
}

}  // namespace lisreg
