// Sweep pre-treatment in front of the feature extractor (SURVEY.md 8f "next" #3).
//
// Reference: ring / time synthesis laserPretreatmentNode.cpp:60-230 (removeNaN, removeClosedPointCloud :244-272, scanID from
// the elevation angle :95-126 == feat_synth_ring, relTime from the azimuth :128-141); constant-velocity de-skew
// DistortionAdjust::AdjustCloud / UpdateMatrix distortionAdjust.cpp:419-479.  Semantics, third-party (Eigen AngleAxis /
// Quaternion) resolutions and the trigonometry convention are listed in oracle/orc_pretreat.cpp; both sides agree bit for bit.
//
// The only sequential piece upstream is the halfPassed flag of the azimuth unwrapping: it flips at the FIRST point (in
// cloud order, among the points that get a scanID) whose not-yet-passed azimuth is more than pi past startOri, and every
// point's own test does not depend on the flag.  So: every point evaluates its test, an atomicMin finds the first index
// that fires, and a second pass unwraps each point according to its position relative to that index.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "features.cuh"

namespace lisreg {

constexpr double PT_PI = 3.14159265358979323846;

// removeNaNFromPointCloud + removeClosedPointCloud: flags has n + 1 entries (the last one 0) for the scan
__global__ void k_pt_flags_range(const float4* __restrict__ pts, int n, float min_r, float max_r, uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t f = 0u;
  if (i < n) {
    const float4 p = __ldg(&pts[i]);
    const float r2 = p.x * p.x + p.y * p.y + p.z * p.z;
    f = (isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && !(r2 < min_r * min_r) && !(r2 > max_r * max_r)) ? 1u : 0u;
  }
  flags[i] = f;
}

struct PtOri { float startOri, endOri; int first_pass; int pad; };

// one thread: startOri / endOri from the first / last point of the range-filtered cloud (:79-83)
__global__ void k_pt_ori(const float4* __restrict__ cloud, int m, PtOri* __restrict__ o) {
  const float4 a = cloud[0], b = cloud[m - 1];
  const float startOri = -atan2f_cr(a.y, a.x);
  float endOri = (float)((double)-atan2f_cr(b.y, b.x) + 2 * PT_PI);
  if ((double)(endOri - startOri) > 3 * PT_PI) endOri = (float)((double)endOri - 2 * PT_PI);
  else if ((double)(endOri - startOri) < PT_PI) endOri = (float)((double)endOri + 2 * PT_PI);
  o->startOri = startOri; o->endOri = endOri; o->first_pass = 0x7fffffff;
}

// scanID validity (flags, n + 1 entries) + the index at which halfPassed flips
__global__ void k_pt_ring_cond(const float4* __restrict__ cloud, int m, int n_scan, PtOri* __restrict__ o, uint32_t* __restrict__ flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > m) return;
  uint32_t f = 0u;
  if (k < m) {
    const float4 p = __ldg(&cloud[k]);
    if (feat_synth_ring(p, n_scan) >= 0) {
      f = 1u;
      const float startOri = o->startOri;
      float ori = -atan2f_cr(p.y, p.x);
      if ((double)ori < (double)startOri - PT_PI / 2) ori = (float)((double)ori + 2 * PT_PI);
      else if ((double)ori > (double)startOri + PT_PI * 3 / 2) ori = (float)((double)ori - 2 * PT_PI);
      if ((double)(ori - startOri) > PT_PI) atomicMin(&o->first_pass, k);
    }
  }
  flags[k] = f;
}

// scanned = exclusive prefix of the scanID flags: emits point, ring, time of every surviving point in cloud order
__global__ void k_pt_emit(const float4* __restrict__ cloud, int m, int n_scan, double scan_period, const PtOri* __restrict__ o,
                          const uint32_t* __restrict__ scanned, float4* __restrict__ out, uint16_t* __restrict__ ring, float* __restrict__ time) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const uint32_t pos = scanned[k];
  if (scanned[k + 1] == pos) return;
  const float4 p = __ldg(&cloud[k]);
  const float startOri = o->startOri, endOri = o->endOri;
  float ori = -atan2f_cr(p.y, p.x);
  if (k <= o->first_pass) {                       // halfPassed still false when this point is processed (the flipping point included)
    if ((double)ori < (double)startOri - PT_PI / 2) ori = (float)((double)ori + 2 * PT_PI);
    else if ((double)ori > (double)startOri + PT_PI * 3 / 2) ori = (float)((double)ori - 2 * PT_PI);
  } else {
    ori = (float)((double)ori + 2 * PT_PI);
    if ((double)ori < (double)endOri - PT_PI * 3 / 2) ori = (float)((double)ori + 2 * PT_PI);
    else if ((double)ori > (double)endOri + PT_PI / 2) ori = (float)((double)ori - 2 * PT_PI);
  }
  const float relTime = (ori - startOri) / (endOri - startOri);
  out[pos] = p;
  ring[pos] = (uint16_t)feat_synth_ring(p, n_scan);
  time[pos] = (float)(scan_period * (double)relTime);
}

// ---- constant-velocity de-skew ----
struct PtQuat { float w, x, y, z; };
__device__ __forceinline__ PtQuat pt_q_axis(float angle, int axis) {
  const float ha = 0.5f * angle;
  PtQuat q{(float)cos((double)ha), 0.f, 0.f, 0.f};
  const float s = (float)sin((double)ha) * 1.f;
  if (axis == 0) q.x = s; else if (axis == 1) q.y = s; else q.z = s;
  return q;
}
__device__ __forceinline__ PtQuat pt_q_mul(const PtQuat& a, const PtQuat& b) {
  PtQuat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
struct PtMotion { float v[3], w[3]; float scan_period; };
// out[i - 1] = R(w * t) * p_i + v * t,  t = time_i - scan_period / 2,  i = 1 .. n - 1
__global__ void k_deskew_cv(const float4* __restrict__ pts, const float* __restrict__ time, int n, PtMotion mo, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i >= n) return;
  const float4 p = __ldg(&pts[i]);
  const float rt = (float)((double)time[i] - (double)mo.scan_period / 2.0);
  const float ax = mo.w[0] * rt, ay = mo.w[1] * rt, az = mo.w[2] * rt;
  const PtQuat q = pt_q_mul(pt_q_mul(pt_q_axis(az, 2), pt_q_axis(ay, 1)), pt_q_axis(ax, 0));
  float nn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
  float angle, ux, uy, uz;
  if (nn != 0.f) {
    angle = 2.f * (float)atan2((double)nn, (double)fabsf(q.w));
    if (q.w < 0.f) nn = -nn;
    ux = q.x / nn; uy = q.y / nn; uz = q.z / nn;
  } else { angle = 0.f; ux = 1.f; uy = 0.f; uz = 0.f; }
  const float s = (float)sin((double)angle), c = (float)cos((double)angle);
  const float sx = s * ux, sy = s * uy, sz = s * uz;
  const float cx = (1.f - c) * ux, cy = (1.f - c) * uy, cz = (1.f - c) * uz;
  float R[9], t;
  t = cx * uy; R[1] = t - sz; R[3] = t + sz;
  t = cx * uz; R[2] = t + sy; R[6] = t - sy;
  t = cy * uz; R[5] = t - sx; R[7] = t + sx;
  R[0] = cx * ux + c; R[4] = cy * uy + c; R[8] = cz * uz + c;
  float4 o;
  o.x = ((R[0] * p.x + R[1] * p.y) + R[2] * p.z) + mo.v[0] * rt;
  o.y = ((R[3] * p.x + R[4] * p.y) + R[5] * p.z) + mo.v[1] * rt;
  o.z = ((R[6] * p.x + R[7] * p.y) + R[8] * p.z) + mo.v[2] * rt;
  o.w = p.w;
  out[i - 1] = o;
}

}  // namespace lisreg
