// EPSC loop detector on the device (rows F17 / F18 of SURVEY.md 8a).
//
// Reference: EPSCGeneration::project epscGeneration.cpp:84-120, globalICP(ssc1, ssc2, yaw_diff) :258-401,
// loopDetection :663-992.  The reference re-bins the whole current cloud once per gated history candidate, one
// candidate after the other; here every candidate is one thread block and all candidates of a keyframe run
// concurrently, the transformed clouds are never materialised (transformPointCloud is fused into the binning):
//   k_loop_project  360-sector {count, last x, last y, last label} projection of the semantic cloud (one block)
//   k_loop_align    per candidate: sector-count shift search (60 shifts), rotation of the current sector points,
//                   PCL-default point-to-point ICP on <= 360 planar points (brute-force 1-NN in shared memory,
//                   fp64 correspondence sums in source-index order = the oracle's order, Horn rigid fit)
//                   -> T = T_icp * Rz(angle), bit-identical to the oracle
//   k_epsc_describe (epsc.cuh) with the per-candidate transform -> EPSC / SEPSC / FEPSC of the moved cloud
//   k_loop_score    per candidate: 20-shift byte SAD against the stored descriptors of that history keyframe
// The travel-distance gate, the best-candidate selection and the history bookkeeping are a few scalars per
// keyframe and stay on the host (csrc/lisreg.cu lisreg_loop_detect).  Semantics, quirks (Q8) and third-party
// resolutions are listed in oracle/orc_loop.cpp; both sides resolve them identically.
#pragma once
#include "icp.cuh"
#include "epsc.cuh"

namespace lisreg {

constexpr int LOOP_SECT = 360;
constexpr int LOOP_THREADS = 128;

__device__ __forceinline__ bool loop_label(unsigned l) { return l == 13u || l == 14u || l == 16u || l == 18u || l == 19u; }

// grid = 1, block = 256.  out: 360 x float4 {count, last x, last y, last label}
__global__ void k_loop_project(const float4* __restrict__ sem, const uint16_t* __restrict__ label, int n, float4* __restrict__ out) {
  __shared__ int s_cnt[LOOP_SECT], s_last[LOOP_SECT];
  for (int i = threadIdx.x; i < LOOP_SECT; i += blockDim.x) { s_cnt[i] = 0; s_last[i] = -1; }
  __syncthreads();
  const float step = (float)(2. * 3.14159265358979323846 / 360.f);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned l = label[i];
    if (!loop_label(l)) continue;
    const float4 p = __ldg(&sem[i]);
    const float distance = sqrtf(p.x * p.x + p.y * p.y);
    if ((double)distance < 1e-2) continue;
    const float angle = (float)(3.14159265358979323846 + (double)(float)atan2((double)p.y, (double)p.x));
    const int sector = (int)floorf(angle / step);
    if (sector >= LOOP_SECT || sector < 0) continue;
    atomicAdd(&s_cnt[sector], 1);
    atomicMax(&s_last[sector], i);          // the LAST point of a sector supplies x, y, label (:112-115)
  }
  __syncthreads();
  for (int s = threadIdx.x; s < LOOP_SECT; s += blockDim.x) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s_last[s] >= 0) { const float4 p = __ldg(&sem[s_last[s]]); o = make_float4((float)s_cnt[s], p.x, p.y, (float)label[s_last[s]]); }
    out[s] = o;
  }
}

struct LoopAlignOut { float T[16]; float diff_x, diff_y, yaw; int icp_iters; };

// grid = candidates, block = LOOP_THREADS.  hist: [frames][360] float4; cand_id / cand_yaw: history frame and
// yaw_t - yawArr[i] per candidate.
__global__ void __launch_bounds__(LOOP_THREADS)
k_loop_align(const float4* __restrict__ hist, const float4* __restrict__ cur, const int* __restrict__ cand_id,
             const float* __restrict__ cand_yaw, LoopAlignOut* __restrict__ out) {
  const int c = blockIdx.x, tid = threadIdx.x;
  const float4* h = hist + (size_t)cand_id[c] * LOOP_SECT;
  __shared__ float4 s1[LOOP_SECT], s2[LOOP_SECT];
  __shared__ float s_dis[64];
  __shared__ float s_tx[LOOP_SECT], s_ty[LOOP_SECT], s_sx[LOOP_SECT], s_sy[LOOP_SECT], s_sz[LOOP_SECT];   // target (cloud1), source (cloud2, moving)
  __shared__ int s_nt, s_ns, s_tmp_id, s_state;
  __shared__ float s_angle, s_T[16], s_F[16];
  __shared__ int s_bj[LOOP_SECT];
  __shared__ float s_bd[LOOP_SECT];
  __shared__ IcpScratch s_sc;
  __shared__ double s_prev_mse;
  for (int i = tid; i < LOOP_SECT; i += LOOP_THREADS) { s1[i] = __ldg(&h[i]); s2[i] = __ldg(&cur[i]); }
  const float step = (float)(2. * 3.14159265358979323846 / 360.f);
  if (tid == 0) {
    float angle = cand_yaw[c];
    if ((double)angle >= 2. * 3.14159265358979323846) angle = (float)((double)angle - 2. * 3.14159265358979323846);
    if (angle < 0.f) angle = (float)((double)angle + 2. * 3.14159265358979323846);
    s_angle = angle;
    s_tmp_id = (int)floorf(angle / step);
  }
  __syncthreads();
  // ---- sector-count shift search, i in [tmp_id - 30, tmp_id + 30) (:271-293); counts are small integers: exact in fp32
  if (tid < 60) {
    const int i = s_tmp_id - 30 + tid;
    float dis = 0.f;
    for (int j = 0; j < LOOP_SECT; j++) {
      const int col = ((j + i) % LOOP_SECT + LOOP_SECT) % LOOP_SECT;      // Q8: the reference wraps once only (UB beyond)
      dis += fabsf(s1[j].x - s2[col].x);
    }
    s_dis[tid] = dis;
  }
  __syncthreads();
  if (tid == 0) {
    double similarity = 100000;
    float angle = s_angle;
    for (int k = 0; k < 60; k++) if ((double)s_dis[k] < similarity) { similarity = s_dis[k]; angle = (float)(s_tmp_id - 30 + k); }
    angle = angle * step;
    s_angle = angle;
    // cloud1 = history sectors with a label, cloud2 = current sectors with a label rotated by angle (:305-318)
    const float cs = (float)cos((double)angle), sn = (float)sin((double)angle);
    int nt = 0, ns = 0;
    for (int i = 0; i < LOOP_SECT; i++) {
      if (s1[i].w > 0.f) { s_tx[nt] = s1[i].y; s_ty[nt] = s1[i].z; nt++; }
      if (s2[i].w > 0.f) { s_sx[ns] = s2[i].y * cs - s2[i].z * sn; s_sy[ns] = s2[i].y * sn + s2[i].z * cs; s_sz[ns] = 0.f; ns++; }
    }
    s_nt = nt; s_ns = ns;
    for (int i = 0; i < 16; i++) s_F[i] = (i % 5 == 0) ? 1.f : 0.f;
    s_prev_mse = 1.7976931348623157e308;
    s_state = 0;     // 0 running, 1 stopped
  }
  __syncthreads();
  // ---- pcl::IterativeClosestPoint, default parameters (:321-325): 1-NN over the whole target, <= 10 iterations,
  //      stop on |mse - prev| < 1e-12; transformation_epsilon 0 and euclidean_fitness_epsilon -DBL_MAX never fire
  const int nt = s_nt, ns = s_ns;
  int iters = 0;
  if (nt > 0 && ns > 0) {
    for (;;) {
      // correspondences: every thread its own source points (brute-force 1-NN, first minimum = smallest target index)
      for (int i = tid; i < ns; i += LOOP_THREADS) {
        const float x = s_sx[i], y = s_sy[i], z = s_sz[i];
        float bd = 3.0e38f; int bj = -1;
        for (int j = 0; j < nt; j++) {
          const float dx = x - s_tx[j], dy = y - s_ty[j], dz = z;      // target z = 0
          float d = dx * dx; d = d + dy * dy; d = d + dz * dz;
          if (d < bd) { bd = d; bj = j; }
        }
        s_bj[i] = bj; s_bd[i] = bd;
      }
      __syncthreads();
      // the 17 fp64 sums, each accumulated by ONE thread over the source points in index order: the summation order of
      // the CPU oracle (oracle/orc_icp.cpp), so the fitted transform - and every descriptor byte binned after it - is
      // bit-identical to it.  <= 360 terms per sum.
      if (tid < 17) {
        // sum tid: 0-2 source x y z, 3-5 target x y z, 6-14 target_a * source_b (row-major), 15 d^2, 16 count
        const int a = tid < 6 ? tid % 3 : (tid - 6) / 3, b = (tid - 6) % 3;
        const float* src_b = (tid < 3 ? tid : b) == 0 ? s_sx : (tid < 3 ? tid : b) == 1 ? s_sy : s_sz;
        double v = 0.0;
        for (int i = 0; i < ns; i++) {
          const int bj = s_bj[i];
          if (bj < 0) continue;
          const float qa = a == 0 ? s_tx[bj] : a == 1 ? s_ty[bj] : 0.f;
          double t;
          if (tid < 3) t = (double)src_b[i];
          else if (tid < 6) t = (double)qa;
          else if (tid < 15) t = (double)qa * (double)src_b[i];
          else if (tid == 15) t = (double)s_bd[i];
          else t = 1.0;
          v += t;
        }
        s_sc.sums[tid] = v;
      }
      __syncthreads();
      if (tid == 0) {
        const double n = s_sc.sums[16];
        if (n < 3) s_state = 1;                                   // min_number_correspondences_: not converged, stop
        else {
          icp_rigid_from_sums(s_sc, n, s_T);
          float F[16];
          for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0.f; for (int k = 0; k < 4; k++) s += s_T[i * 4 + k] * s_F[k * 4 + j]; F[i * 4 + j] = s; }
          for (int i = 0; i < 16; i++) s_F[i] = F[i];
          s_state = 2;                                            // apply s_T to the source, then test convergence
        }
      }
      __syncthreads();
      if (s_state == 1) break;
      for (int i = tid; i < ns; i += LOOP_THREADS) {
        const float x = s_sx[i], y = s_sy[i], z = s_sz[i];
        s_sx[i] = (s_T[0] * x + s_T[1] * y) + s_T[2] * z + s_T[3];
        s_sy[i] = (s_T[4] * x + s_T[5] * y) + s_T[6] * z + s_T[7];
        s_sz[i] = (s_T[8] * x + s_T[9] * y) + s_T[10] * z + s_T[11];
      }
      iters++;
      bool stop = false;
      if (iters >= 10) stop = true;
      else {
        const double cos_angle = 0.5 * ((double)s_T[0] + (double)s_T[5] + (double)s_T[10] - 1);
        const double tr2 = (double)s_T[3] * s_T[3] + (double)s_T[7] * s_T[7] + (double)s_T[11] * s_T[11];
        if (cos_angle >= 1.0 && tr2 <= 0.0) stop = true;          // rotation threshold 1 - 0, translation threshold 0
        else {
          const double mse = s_sc.sums[15] / s_sc.sums[16];
          if (fabs(mse - s_prev_mse) < 1e-12) stop = true;
          else if (fabs(mse - s_prev_mse) / s_prev_mse < -1.7976931348623157e308) stop = true;
        }
      }
      __syncthreads();                                            // everyone has read s_T / s_sc / s_prev_mse
      if (tid == 0 && !stop) s_prev_mse = s_sc.sums[15] / s_sc.sums[16];
      if (stop) break;
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0) {
    // trans * trans1 with trans1 = Rz(angle) (AngleAxisf about Z); then getTranslationAndEulerAngles (:327-336)
    const float angle = s_angle;
    const float cz = (float)cos((double)angle), sz = (float)sin((double)angle);
    float T1[16];
    for (int i = 0; i < 16; i++) T1[i] = (i % 5 == 0) ? 1.f : 0.f;
    T1[0] = cz; T1[1] = -sz; T1[4] = sz; T1[5] = cz; T1[10] = (1.f - cz) + cz;
    LoopAlignOut o;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float s = 0.f; for (int k = 0; k < 4; k++) s += s_F[i * 4 + k] * T1[k * 4 + j]; o.T[i * 4 + j] = s; }
    o.diff_x = o.T[3]; o.diff_y = o.T[7];
    o.yaw = (float)atan2((double)o.T[4], (double)o.T[0]);
    o.icp_iters = iters;
    out[c] = o;
  }
}

struct LoopScoreOut { int sad[3]; int shift[3]; };   // per kind (EPSC, SEPSC, FEPSC): minimal SAD and its shift i in [-10, 10); sad = -1: none

// grid = candidates, block = 64 (threads 0..59 = kind x shift).  cur_desc: [cand][3][1600] (moved current cloud),
// hist_desc: [frames][3][1600].  calculateDistance (:633-660): desc1 = history, desc2 = current, first minimum wins.
__global__ void k_loop_score(const uint8_t* __restrict__ cur_desc, const uint8_t* __restrict__ hist_desc,
                             const int* __restrict__ cand_id, LoopScoreOut* __restrict__ out) {
  const int c = blockIdx.x, tid = threadIdx.x;
  __shared__ uint8_t s_h[3 * EPSC_SIZE], s_c[3 * EPSC_SIZE];
  __shared__ int s_sad[60];
  const uint8_t* hd = hist_desc + (size_t)cand_id[c] * 3 * EPSC_SIZE;
  const uint8_t* cd = cur_desc + (size_t)c * 3 * EPSC_SIZE;
  for (int i = tid; i < 3 * EPSC_SIZE; i += blockDim.x) { s_h[i] = hd[i]; s_c[i] = cd[i]; }
  __syncthreads();
  if (tid < 60) {
    const int kind = tid / 20, i = tid % 20 - 10;
    const uint8_t* d1 = s_h + kind * EPSC_SIZE; const uint8_t* d2 = s_c + kind * EPSC_SIZE;
    int sad = 0;
    for (int p = 0; p < EPSC_SECTORS; p++) {
      int col = p + i;
      if (col >= EPSC_SECTORS) col -= EPSC_SECTORS;
      if (col < 0) col += EPSC_SECTORS;
      for (int q = 0; q < EPSC_RINGS; q++) sad += abs((int)d1[q * EPSC_SECTORS + p] - (int)d2[q * EPSC_SECTORS + col]);
    }
    s_sad[tid] = sad;
  }
  __syncthreads();
  if (tid < 3) {
    // difference starts at 1.0 and only a strictly smaller SAD / (80*20*255) replaces it
    int best = 80 * 20 * 255, bs = 0, found = -1;
    for (int k = 0; k < 20; k++) if (s_sad[tid * 20 + k] < best) { best = s_sad[tid * 20 + k]; bs = k - 10; found = best; }
    out[c].sad[tid] = found; out[c].shift[tid] = bs;
  }
}

}  // namespace lisreg
