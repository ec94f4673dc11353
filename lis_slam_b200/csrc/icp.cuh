// Batched point-to-point ICP (loop-closure verification) on the device.
//
// Reference: SubMapOdometryNode::detectLoopClosureForSubMap subMapOptmizationNode.cpp:2739-2916:
// pcl::IterativeClosestPoint with MaxCorrespondenceDistance 10, MaximumIterations 30,
// TransformationEpsilon 1e-4, EuclideanFitnessEpsilon 1e-4, RANSACIterations 0 (:2763-2769); the source is
// pre-transformed by the initial guess (:2822-2824); accept iff converged and getFitnessScore() <= 0.5.
// PCL's published algorithm is restated in oracle/orc_icp.cpp; the device follows the same flow:
//   k_icp_corr   : cur = T_pending * cur ; exact 1-NN in the target grid within the correspondence distance ;
//                  17 fp64 sums (sum p, sum q, sum q p^T, sum d^2, n) per block, fixed-order partials
//   k_icp_solve  : one warp per pair: means, cross-covariance, optimal rotation (Horn's quaternion form via a
//                  4x4 fp64 Jacobi == Umeyama/SVD with the determinant fix), final = T * final,
//                  DefaultConvergenceCriteria (iterations / transform epsilon / absolute + relative MSE)
//   k_icp_fitness: mean squared unbounded 1-NN distance of the ORIGINAL source under the final transformation
// Pairs are independent: grid.y = pair (candidate pairs shard across GPUs exactly like frames).
#pragma once
#include "grid.cuh"
#include "../../include/lisreg.h"

namespace lisreg {

struct IcpPair {
  const float4* src; int ns;        // source (already pre-transformed by the initial guess)
  float4* cur;                      // work copy (ns)
  int* nn;                          // nearest target point of every source point at the previous iteration (-1: none yet)
  int tgt_slot;                     // map slot whose SURF cloud is the target
  int pad;
};

struct IcpState {
  float Tpend[16];   // transform estimated in the previous iteration, applied to `cur` at the next k_icp_corr
  float Tfinal[16];
  double prev_mse;
  double fitness;
  int iters, done, converged, n_corr, pending;
};

struct IcpParamsDev { float max_d2; int max_iters; double rot_thr, trans_thr, fit_eps; };

constexpr int ICP_THREADS = 256;
constexpr int ICP_NSUM = 18;    // 3 sum p, 3 sum q, 9 sum q p^T, sum d^2, n, pad

__device__ __forceinline__ void icp_xform(const float* T, float4 p, float& x, float& y, float& z) {
  x = (T[0] * p.x + T[1] * p.y) + T[2] * p.z + T[3];     // Eigen: linear() * v (column order) + translation
  y = (T[4] * p.x + T[5] * p.y) + T[6] * p.z + T[7];
  z = (T[8] * p.x + T[9] * p.y) + T[10] * p.z + T[11];
}

__global__ void k_icp_init(const IcpPair* __restrict__ pairs, IcpState* __restrict__ st, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  IcpState s;
  for (int i = 0; i < 16; i++) { s.Tpend[i] = (i % 5 == 0) ? 1.f : 0.f; s.Tfinal[i] = s.Tpend[i]; }
  s.prev_mse = 1.7976931348623157e308; s.fitness = 1.7976931348623157e308;
  s.iters = 0; s.done = pairs[p].ns <= 0 ? 1 : 0; s.converged = 0; s.n_corr = 0; s.pending = 0;
  st[p] = s;
}

// copies the source into the work buffer.  grid = (blocks, P)
__global__ void k_icp_copy(const IcpPair* __restrict__ pairs) {
  const IcpPair pr = pairs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pr.ns; i += gridDim.x * blockDim.x) { pr.cur[i] = __ldg(&pr.src[i]); pr.nn[i] = -1; }
}

template <bool FITNESS>
__global__ void __launch_bounds__(ICP_THREADS)
k_icp_corr(const IcpPair* __restrict__ pairs, const IcpState* __restrict__ st, const MapDev* __restrict__ maps,
           IcpParamsDev prm, double* __restrict__ partials, int nblk) {
  const int p = blockIdx.y, tid = threadIdx.x;
  const IcpPair pr = pairs[p];
  __shared__ float sT[16];
  __shared__ int s_skip;
  __shared__ double swarp[ICP_THREADS / 32][ICP_NSUM];
  if (tid == 0) s_skip = FITNESS ? 0 : st[p].done;
  if (tid < 16) sT[tid] = FITNESS ? st[p].Tfinal[tid] : st[p].Tpend[tid];
  __syncthreads();
  if (s_skip) return;
  const bool apply = FITNESS ? true : (st[p].pending != 0);
  const GridDev& g = maps[pr.tgt_slot].surf;
  double acc[17];
#pragma unroll
  for (int i = 0; i < 17; i++) acc[i] = 0.0;
  for (int i = blockIdx.x * ICP_THREADS + tid; i < pr.ns; i += nblk * ICP_THREADS) {
    float4 c = FITNESS ? __ldg(&pr.src[i]) : pr.cur[i];
    float x = c.x, y = c.y, z = c.z;
    if (apply) { icp_xform(sT, c, x, y, z); if (!FITNESS) pr.cur[i] = make_float4(x, y, z, c.w); }
    knn_key best[1];
    float gate = FITNESS ? 3.0e38f : prm.max_d2 * 1.0000002f;                    // (gate is exclusive; PCL keeps d2 <= max2)
    // The neighbour of the previous iteration bounds the search: the nearest point is at most as far as that one is NOW, so
    // the exact search only has to look inside that radius (a point that moved a few centimetres no longer walks the
    // shells out to the 10 m correspondence distance in sparse regions).  Same result, by construction.
    const int prev = pr.nn[i];
    if (prev >= 0) {
      const float4 m = __ldg(&g.pts[prev]);
      const float ddx = x - m.x, ddy = y - m.y, ddz = z - m.z;
      float dp = ddx * ddx; dp = dp + ddy * ddy; dp = dp + ddz * ddz;        // the search's own expression (knn_scan_range)
      const float gp = dp * 1.0000002f + 1e-30f;
      if (gp < gate) gate = gp;
    }
    knn_grid<1>(g, x, y, z, gate, best);
    const int pos = knn_key_pos(best[0]);
    const float d2 = knn_key_d(best[0]);
    if (!FITNESS) pr.nn[i] = pos;
    if (pos < 0 || (!FITNESS && d2 > prm.max_d2)) continue;
    if (FITNESS) { acc[15] += (double)d2; acc[16] += 1.0; continue; }
    const float4 q = __ldg(&g.pts[pos]);
    acc[0] += x; acc[1] += y; acc[2] += z;
    acc[3] += q.x; acc[4] += q.y; acc[5] += q.z;
    acc[6] += (double)q.x * x; acc[7] += (double)q.x * y; acc[8] += (double)q.x * z;
    acc[9] += (double)q.y * x; acc[10] += (double)q.y * y; acc[11] += (double)q.y * z;
    acc[12] += (double)q.z * x; acc[13] += (double)q.z * y; acc[14] += (double)q.z * z;
    acc[15] += (double)d2; acc[16] += 1.0;
  }
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int i = 0; i < 17; i++) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) swarp[wid][i] = v;
  }
  __syncthreads();
  if (tid < 17) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < ICP_THREADS / 32; w++) v += swarp[w][tid];
    partials[((size_t)p * nblk + blockIdx.x) * ICP_NSUM + tid] = v;
  }
}

// cyclic Jacobi for a symmetric 4x4 (fp64, shared-memory operands): eigenvectors in the columns of V
__device__ inline void icp_jacobi4(double* A, double* W, double* V) {
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) V[i * 4 + j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) off += A[i * 4 + j] * A[i * 4 + j];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) {
      if (fabs(A[p * 4 + q]) < 1e-300) continue;
      const double theta = (A[q * 4 + q] - A[p * 4 + p]) / (2 * A[p * 4 + q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
      const double c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < 4; k++) { const double akp = A[k * 4 + p], akq = A[k * 4 + q]; A[k * 4 + p] = c * akp - s * akq; A[k * 4 + q] = s * akp + c * akq; }
      for (int k = 0; k < 4; k++) { const double apk = A[p * 4 + k], aqk = A[q * 4 + k]; A[p * 4 + k] = c * apk - s * aqk; A[q * 4 + k] = s * apk + c * aqk; }
      for (int k = 0; k < 4; k++) { const double vkp = V[k * 4 + p], vkq = V[k * 4 + q]; V[k * 4 + p] = c * vkp - s * vkq; V[k * 4 + q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < 4; i++) W[i] = A[i * 4 + i];
}

struct IcpScratch { double sums[ICP_NSUM]; double N[16], W[4], V[16]; };

// rigid fit of (source -> target) from the 17 correspondence sums (sum p, sum q, sum q p^T, sum d^2, n):
// means, cross-covariance, optimal rotation by Horn's quaternion form (4x4 fp64 Jacobi == Umeyama/SVD with the
// determinant fix), rounded to fp32 like PCL's Matrix4f.  T = row-major 4x4.
__device__ inline void icp_rigid_from_sums(IcpScratch& sc, double n, float* T) {
  double mp[3], mq[3], H[9];
  for (int a = 0; a < 3; a++) { mp[a] = sc.sums[a] / n; mq[a] = sc.sums[3 + a] / n; }
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) H[a * 3 + b] = sc.sums[6 + a * 3 + b] / n - mq[a] * mp[b];   // dst x src^T
  {
    const double Sxx = H[0], Sxy = H[3], Sxz = H[6], Syx = H[1], Syy = H[4], Syz = H[7], Szx = H[2], Szy = H[5], Szz = H[8];
    const double N[16] = {Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx,
                          Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz,
                          Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy,
                          Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz};
    for (int i = 0; i < 16; i++) sc.N[i] = N[i];
  }
  icp_jacobi4(sc.N, sc.W, sc.V);
  int b = 0; for (int i = 1; i < 4; i++) if (sc.W[i] > sc.W[b]) b = i;
  double q0 = sc.V[0 * 4 + b], q1 = sc.V[1 * 4 + b], q2 = sc.V[2 * 4 + b], q3 = sc.V[3 * 4 + b];
  const double nq = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  q0 /= nq; q1 /= nq; q2 /= nq; q3 /= nq;
  double R[9];
  R[0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
  R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; R[5] = 2 * (q2 * q3 - q0 * q1);
  R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  for (int i = 0; i < 16; i++) T[i] = 0.f;
  for (int a = 0; a < 3; a++) {
    for (int c = 0; c < 3; c++) T[a * 4 + c] = (float)R[a * 3 + c];
    T[a * 4 + 3] = (float)(mq[a] - (R[a * 3] * mp[0] + R[a * 3 + 1] * mp[1] + R[a * 3 + 2] * mp[2]));
  }
  T[15] = 1.f;
}

// one warp per pair.  FITNESS: only finalises the fitness score.
template <bool FITNESS>
__global__ void k_icp_solve(IcpState* __restrict__ states, IcpParamsDev prm, const double* __restrict__ partials, int nblk, int P) {
  __shared__ IcpScratch ssc[4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * 4 + wid;
  if (p >= P) return;
  if (!FITNESS && states[p].done) return;
  IcpScratch& sc = ssc[wid];
  if (lane < 17) {
    double v = 0.0;
    for (int b = 0; b < nblk; b++) v += partials[((size_t)p * nblk + b) * ICP_NSUM + lane];
    sc.sums[lane] = v;
  }
  __syncwarp();
  if (lane != 0) return;
  IcpState st = states[p];
  const double n = sc.sums[16];
  if (FITNESS) { st.fitness = n > 0 ? sc.sums[15] / n : 1.7976931348623157e308; states[p] = st; return; }
  st.n_corr = (int)n;
  if (n < 3) { st.converged = 0; st.done = 1; st.pending = 0; states[p] = st; return; }   // min_number_correspondences_ = 3
  float T[16];
  icp_rigid_from_sums(sc, n, T);
  float F[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0.f; for (int k = 0; k < 4; k++) s += T[i * 4 + k] * st.Tfinal[k * 4 + j]; F[i * 4 + j] = s; }
  for (int i = 0; i < 16; i++) { st.Tfinal[i] = F[i]; st.Tpend[i] = T[i]; }
  st.pending = 1;
  st.iters++;
  // DefaultConvergenceCriteria::hasConverged
  bool conv = false;
  if (st.iters >= prm.max_iters) conv = true;
  else {
    const double cos_angle = 0.5 * ((double)T[0] + (double)T[5] + (double)T[10] - 1);
    const double tr2 = (double)T[3] * T[3] + (double)T[7] * T[7] + (double)T[11] * T[11];
    if (cos_angle >= prm.rot_thr && tr2 <= prm.trans_thr) conv = true;
    else {
      const double mse = sc.sums[15] / n;
      if (fabs(mse - st.prev_mse) < 1e-12) conv = true;
      else if (fabs(mse - st.prev_mse) / st.prev_mse < prm.fit_eps) conv = true;
      else st.prev_mse = mse;
    }
  }
  if (conv) { st.converged = 1; st.done = 1; }
  states[p] = st;
}

__global__ void k_icp_finish(const IcpState* __restrict__ st, lisreg_icp_result* __restrict__ out, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  lisreg_icp_result r;
  for (int i = 0; i < 16; i++) r.T[i] = st[p].Tfinal[i];
  r.fitness = st[p].fitness; r.converged = st[p].converged; r.iters = st[p].iters; r.n_corr_last = st[p].n_corr; r.reserved = 0;
  out[p] = r;
}

}  // namespace lisreg
