// LOAM feature extraction on the device, batched over frames (blockIdx.y / frame index).
//
// Reference: LaserProcessing::projectPointCloud laserProcessing.cpp:467-510 (F1),
// cloudExtraction :515-539 (F2), calculateSmoothness :544-563 (F3), markOccludedPoints :568-605 (F4),
// extractFeatures :610-713 (F5); byte-identical duplicate in featureExtraction.cpp:125-366.
//
// Parallel decomposition: F1 point-parallel (first-hit-wins == atomicMin of the input index per
// range-image cell), F2 one block per ring (block prefix sum) after a 64-entry ring scan, F3/F4
// point-parallel (F4's writes are idempotent ORs), F5 one WARP per (frame, ring): its six segments are
// processed in order because the +-5 neighbour suppression of a pick crosses segment borders; inside a
// segment the sort + greedy walk of the reference is restated as a warp-wide selection loop over the
// picks (see k_feat_segments).
// Documented deviations from reference UB are listed in oracle/orc_features.cpp (Q3, Q5, sort ties,
// index -1 read): the device follows the same resolutions.
#pragma once
#include <cstdint>
#include <cfloat>
#include <cuda_runtime.h>
#include "../../include/lisreg.h"

namespace lisreg {

struct FeatParamsDev {
  int n_scan, horizon, downsample;
  float min_range, max_range, edge_thr, surf_thr;
  lisreg_cloud_layout lay;     // point_step == 0: packed float4 records + ring array
  int lean;                    // the per-point curvature / occlusion arrays are not used (k_feat_segments<true> computes both itself): do not initialise them
};

// Per-frame views into the batch work buffers (all device pointers).
struct FeatFrame {
  const float4* pts; const uint16_t* ring; int n;       // raw sweep
  int* owner;            // n_scan*horizon  : input index owning each range-image cell (INT_MAX = empty)
  float4* ext_pts;       // extracted cloud (capacity n_scan*horizon)
  int* ext_src;          // index of the input point in each extracted slot
  unsigned short* col; float* range; float* curv;   // column index (horizon <= 2048), range, curvature per extracted point
  unsigned char* picked; signed char* label;         // cloudNeighborPicked flag, cloudLabel (-1 / 0 / 1): bytes (the stage is DRAM-bound)
  int* ring_count;       // n_scan
  int* ring_start; int* ring_end;   // n_scan (startRingIndex / endRingIndex)
  int* M;                // number of extracted points
  int* seg_corner; int* seg_ncorner;   // [n_scan*6][20], [n_scan*6]
  int* seg_flat; int* seg_nflat;       // [n_scan*6][10], [n_scan*6]
  int* seg_valid;                      // [n_scan*6]  1 if sp < ep
  int* seg_sp; int* seg_ep;            // [n_scan*6]
  // compacted outputs, reference push order
  int* corner_idx; int* sharp_idx; int* flat_idx; int* surf_idx;
  int* counts;           // [4] n_corner, n_sharp, n_flat, n_surf
  // optional motion de-skew (deskewPoint, laserProcessing.cpp:427-462): n_imu = 0 disables it
  const float* time;     // per input point: PointXYZIRT::time
  const double* imu_time; const double* imu_rot;   // imuTime[n_imu], imuRotX/Y/Z[n_imu] interleaved
  int n_imu; double t_scan;                         // imuPointerCur + 1, timeScanCur
  float* start_inv;      // [9] transStartInverse (linear part; the translation is identically zero)
};

__device__ __forceinline__ float atan2f_cr(float y, float x) { return (float)atan2((double)y, (double)x); }

// raw point i of a sweep in the caller's layout (lisreg_cloud_layout): {x, y, z, intensity}
__device__ __forceinline__ float4 feat_load_point(const float4* __restrict__ pts, const lisreg_cloud_layout& lay, int i) {
  if (lay.point_step == 0) return __ldg(&pts[i]);
  const char* r = reinterpret_cast<const char*>(pts) + (size_t)i * (size_t)lay.point_step;
  float4 p;
  p.x = __ldg(reinterpret_cast<const float*>(r + lay.off_x));
  p.y = __ldg(reinterpret_cast<const float*>(r + lay.off_y));
  p.z = __ldg(reinterpret_cast<const float*>(r + lay.off_z));
  p.w = lay.off_intensity >= 0 ? __ldg(reinterpret_cast<const float*>(r + lay.off_intensity)) : 0.f;
  return p;
}
__device__ __forceinline__ float feat_load_time(const float4* __restrict__ pts, const float* __restrict__ time, const lisreg_cloud_layout& lay, int i) {
  if (lay.point_step != 0 && lay.off_time >= 0)
    return __ldg(reinterpret_cast<const float*>(reinterpret_cast<const char*>(pts) + (size_t)i * (size_t)lay.point_step + lay.off_time));
  return time[i];
}
// scanID from the elevation angle (laserPretreatmentNode.cpp:95-126); -1 = the reference drops the point
__device__ __forceinline__ int feat_synth_ring(float4 p, int n_scan) {
  const float angle = (float)((double)(float)atan((double)(p.z / sqrtf(p.x * p.x + p.y * p.y))) * 180 / 3.14159265358979323846);
  int id;
  if (n_scan == 16) id = (int)((double)((angle + 15) / 2) + 0.5);
  else if (n_scan == 32) id = (int)(((double)angle + 92.0 / 3.0) * 3.0 / 4.0);
  else if (n_scan == 64) {
    if ((double)angle >= -8.83) id = (int)((double)(2 - angle) * 3.0 + 0.5);
    else id = n_scan / 2 + (int)((-8.83 - (double)angle) * 2.0 + 0.5);
    if ((double)angle > 2 || (double)angle < -24.33 || id > 50 || id < 0) return -1;
    return id;
  } else return -1;
  if (id > n_scan - 1 || id < 0) return -1;
  return id;
}
__device__ __forceinline__ int feat_load_ring(const float4* __restrict__ pts, const uint16_t* __restrict__ ring, const lisreg_cloud_layout& lay,
                                              int n_scan, int i, float4 p) {
  if (lay.point_step == 0 || lay.off_ring == -1) return (int)ring[i];
  if (lay.off_ring >= 0) return (int)__ldg(reinterpret_cast<const unsigned short*>(reinterpret_cast<const char*>(pts) + (size_t)i * (size_t)lay.point_step + lay.off_ring));
  return feat_synth_ring(p, n_scan);
}

__global__ void k_feat_clear(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int cells = prm.n_scan * prm.horizon;
  // only the range-image owners need a value in EVERY cell; the per-point arrays are initialised for the M extracted
  // slots by k_feat_compact (nothing reads them beyond M)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) f.owner[i] = 0x7fffffff;
  if (blockIdx.x == 0 && threadIdx.x < 4) f.counts[threadIdx.x] = 0;
}

__device__ __forceinline__ bool feat_project(const FeatParamsDev& prm, float4 p, int ring, float& r, int& cell) {
  r = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
  if (r < prm.min_range || r > prm.max_range) return false;
  if (ring < 0 || ring >= prm.n_scan) return false;
  if (ring % prm.downsample != 0) return false;
  const float ang_res_x = (float)(360.0 / (double)(float)prm.horizon);
  // The column is the only consumer of the azimuth.  Fast path: fp32 atan2f (<= 2 ulp) and fp32 arithmetic, together
  // < 5e-4 columns off; only when the column coordinate lands within 4e-3 of a rounding boundary is the reference
  // expression evaluated (correctly rounded atan2 in fp64) - the result is identical, the fp64 path runs for ~0.8 %
  // of the points.
  const float t_fast = (atan2f(p.x, p.y) * 57.29577951f - 90.0f) / ang_res_x;    // fp32 end to end: < 5e-4 columns off
  const float fr = t_fast - floorf(t_fast);
  int col;
  if (fabsf(fr - 0.5f) > 4e-3f) {
    col = (int)(-roundf(t_fast)) + prm.horizon / 2;
  } else {
    const float horizonAngle = (float)((double)(atan2f_cr(p.x, p.y) * 180) / 3.14159265358979323846);
    col = (int)(-round(((double)horizonAngle - 90.0) / (double)ang_res_x) + (double)(prm.horizon / 2));
  }
  if (col >= prm.horizon) col -= prm.horizon;
  if (col < 0 || col >= prm.horizon) return false;
  cell = ring * prm.horizon + col;
  return true;
}

// F1: first point to hit a cell wins (:499) == smallest input index
__global__ void k_feat_project(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.n; i += gridDim.x * blockDim.x) {
    float r; int cell;
    const float4 p = feat_load_point(f.pts, prm.lay, i);
    if (feat_project(prm, p, feat_load_ring(f.pts, f.ring, prm.lay, prm.n_scan, i, p), r, cell)) atomicMin(&f.owner[cell], i);
  }
}

// ---- motion de-skew: findRotation :368-400, deskewPoint :427-462 ----
__device__ __forceinline__ void feat_rot_of(float roll, float pitch, float yaw, float* R) {   // pcl::getTransformation, linear part
  const float A = (float)cos((double)yaw), B = (float)sin((double)yaw), C = (float)cos((double)pitch), D = (float)sin((double)pitch),
              E = (float)cos((double)roll), F = (float)sin((double)roll);
  const float DE = D * E, DF = D * F;
  R[0] = A * C; R[1] = A * DF - B * E; R[2] = B * F + A * DE;
  R[3] = B * C; R[4] = A * E + B * DF; R[5] = B * DE - A * F;
  R[6] = -D;    R[7] = C * F;          R[8] = C * E;
}
__device__ __forceinline__ void feat_find_rotation(const FeatFrame& f, double pointTime, float* r) {
  const int cur = f.n_imu - 1;
  int front = 0;
  while (front < cur) { if (pointTime < f.imu_time[front]) break; ++front; }
  if (pointTime > f.imu_time[front] || front == 0) {
    for (int a = 0; a < 3; a++) r[a] = (float)f.imu_rot[3 * front + a];
  } else {
    const int back = front - 1;
    const double ratioFront = (pointTime - f.imu_time[back]) / (f.imu_time[front] - f.imu_time[back]);
    const double ratioBack = (f.imu_time[front] - pointTime) / (f.imu_time[front] - f.imu_time[back]);
    for (int a = 0; a < 3; a++) r[a] = (float)(f.imu_rot[3 * front + a] * ratioFront + f.imu_rot[3 * back + a] * ratioBack);
  }
}
__device__ __forceinline__ float feat_cof(const float* m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
// transStartInverse: the first point processed by projectPointCloud (smallest input index that owns a cell) fixes the
// reference orientation (:439-443).  grid = F, block = 256
__global__ void k_feat_deskew_start(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.x];
  if (f.n_imu <= 0) return;
  __shared__ int s_min[32];
  int m = 0x7fffffff;
  const int cells = prm.n_scan * prm.horizon;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) m = min(m, f.owner[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = min(m, s_min[w]);
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (m != 0x7fffffff) {
      float r[3];
      feat_find_rotation(f, f.t_scan + (double)feat_load_time(f.pts, f.time, prm.lay, m), r);
      feat_rot_of(r[0], r[1], r[2], R);
    }
    // Eigen::Affine3f::inverse(): cofactor inverse of the linear part
    const float c0 = feat_cof(R, 0, 0), c1 = feat_cof(R, 1, 0), c2 = feat_cof(R, 2, 0);
    const float det = (c0 * R[0] + c1 * R[3]) + c2 * R[6];
    const float invdet = 1.f / det;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) f.start_inv[i * 3 + j] = feat_cof(R, j, i) * invdet;
  }
}
__device__ __forceinline__ float4 feat_deskew_point(const FeatFrame& f, const float* Sinv, float4 p, float t_rel) {
  float r[3], Rc[9], Bt[9];
  feat_find_rotation(f, f.t_scan + (double)t_rel, r);
  feat_rot_of(r[0], r[1], r[2], Rc);
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) Bt[a * 3 + b] = (Sinv[a * 3] * Rc[b] + Sinv[a * 3 + 1] * Rc[3 + b]) + Sinv[a * 3 + 2] * Rc[6 + b];
  float4 o;
  o.x = Bt[0] * p.x + Bt[1] * p.y + Bt[2] * p.z + 0.f;
  o.y = Bt[3] * p.x + Bt[4] * p.y + Bt[5] * p.z + 0.f;
  o.z = Bt[6] * p.x + Bt[7] * p.y + Bt[8] * p.z + 0.f;
  o.w = p.w;
  return o;
}

// F2a: valid cells per ring.  grid = (n_scan, F), block = 256
__global__ void k_feat_ring_count(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int ring = blockIdx.x;
  int c = 0;
  for (int j = threadIdx.x; j < prm.horizon; j += blockDim.x) c += f.owner[ring * prm.horizon + j] != 0x7fffffff;
  __shared__ int s[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < (blockDim.x + 31) / 32; w++) t += s[w]; f.ring_count[ring] = t; }
}

// F2b: row-major compaction.  grid = (n_scan, F), block = 256 (8 warps); every warp scans chunks of 32 cells
template <bool DESKEW>   // the de-skew path (fp64 trigonometry) is compiled out of the common instantiation: it costs registers
__global__ void k_feat_compact(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int ring = blockIdx.x;
  __shared__ int s_base;
  __shared__ int s_chunk[64];   // horizon / 32 <= 64 chunks (horizon <= 2048)
  if (threadIdx.x == 0) {
    int base = 0;
    for (int r = 0; r < ring; r++) base += f.ring_count[r];
    s_base = base;
    f.ring_start[ring] = base - 1 + 5;                        // startRingIndex (:521)
    f.ring_end[ring] = base + f.ring_count[ring] - 1 - 5;     // endRingIndex   (:537)
    if (ring == prm.n_scan - 1) *f.M = base + f.ring_count[ring];
  }
  const int nchunk = (prm.horizon + 31) / 32;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // pass 1: per-chunk counts
  for (int ch = wid; ch < nchunk; ch += nw) {
    const int j = ch * 32 + lane;
    const bool v = j < prm.horizon && f.owner[ring * prm.horizon + j] != 0x7fffffff;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_chunk[ch] = __popc(m);
  }
  __syncthreads();
  if (threadIdx.x == 0) { int acc = 0; for (int ch = 0; ch < nchunk; ch++) { int t = s_chunk[ch]; s_chunk[ch] = acc; acc += t; } }
  __syncthreads();
  for (int ch = wid; ch < nchunk; ch += nw) {
    const int j = ch * 32 + lane;
    const int own = j < prm.horizon ? f.owner[ring * prm.horizon + j] : 0x7fffffff;
    const bool v = own != 0x7fffffff;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (v) {
      const int pos = s_base + s_chunk[ch] + __popc(m & ((1u << lane) - 1u));
      const float4 p = feat_load_point(f.pts, prm.lay, own);
      // range (and the column) come from the ORIGINAL point; only the stored coordinates are de-skewed (:489-507)
      f.ext_pts[pos] = (DESKEW && f.n_imu > 0) ? feat_deskew_point(f, f.start_inv, p, feat_load_time(f.pts, f.time, prm.lay, own)) : p;
      f.ext_src[pos] = own;
      f.col[pos] = (unsigned short)j;
      f.range[pos] = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
      f.label[pos] = 0;                                           // resetParameters (:61-75): cleared per extracted slot
      if (!prm.lean) { f.picked[pos] = 0; f.curv[pos] = 0.f; }
    }
  }
}

// ---- F1 + F2 fused: projection and row-major compaction without the range image ever leaving the SM ----
// One block per (frame, group of R consecutive rings).  The group's slice of the range image (the input index owning each
// cell) lives in SHARED memory: the block walks the ring ids of the whole sweep, projects the points of its own rings and
// resolves "first point wins" (:499) with shared-memory atomicMin; then it counts / scans its valid cells, learns how many
// cells the groups before it extracted (decoupled look-back over per-frame group totals: a block waits only for blocks
// that took an EARLIER ticket, which are resident or finished and publish their total before waiting themselves - no
// deadlock), and writes its rows of the extracted cloud at their final positions.  Replaces k_feat_clear, k_feat_project,
// k_feat_ring_count and k_feat_compact<false>: 118 MB of owner-image clears, 28 M global atomics and two re-reads of the
// owner image per 256-frame step stay on chip.
// ctl = [ticket counter, finished counter, group totals (F * G, 0 = not published, else total + 1)]: the last block to
// finish puts everything back to zero, so the kernel can be replayed (CUDA graph) without a host-side reset.
constexpr int FEAT_FRONT_THREADS = 384;
constexpr int FEAT_FRONT_SMEM_INTS = 15104;   // 59 KB of range-image slice + chunk offsets per block: three blocks per SM
__global__ void __launch_bounds__(FEAT_FRONT_THREADS, 3)
k_feat_front(FeatFrame* frames, FeatParamsDev prm, int R, int G, int F, int* ctl) {
  extern __shared__ int s_dyn[];
  __shared__ int s_ticket, s_base, s_last;
  __shared__ int s_queue[FEAT_FRONT_THREADS / 32][64];
  const int nchunk = (prm.horizon + 31) / 32;
  int* s_owner = s_dyn;                       // [R * horizon]
  int* s_chunk = s_dyn + R * prm.horizon;     // [R * nchunk + 1] valid cells per 32-cell chunk -> exclusive offsets
  if (threadIdx.x == 0) s_ticket = atomicAdd(&ctl[0], 1);
  __syncthreads();
  const int ticket = s_ticket, fi = ticket / G, g = ticket - fi * G;
  const FeatFrame f = frames[fi];
  const int r0 = g * R, r1 = min(r0 + R, prm.n_scan), nr = r1 - r0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < nr * prm.horizon; i += blockDim.x) s_owner[i] = 0x7fffffff;
  if (g == 0 && threadIdx.x < 4) f.counts[threadIdx.x] = 0;
  __syncthreads();
  // ---- F1 over the whole sweep, own rings only ----
  // A sweep in firing order has its rings interleaved: only nr of n_scan consecutive points belong to this block.  Each warp
  // therefore first collects the indices of its own points in a 64-entry queue and projects them 32 at a time with all
  // lanes busy.
  {
    const uint16_t* __restrict__ ring = f.ring;
    const int n = f.n;
    int* q = s_queue[wid];
    int qn = 0;                                                       // warp-uniform fill of the queue (< 32 between chunks)
    auto project32 = [&](int cnt) {
      if (lane < cnt) {
        const int i = q[lane];
        float r; int cell;
        if (feat_project(prm, feat_load_point(f.pts, prm.lay, i), (int)__ldg(&ring[i]), r, cell)) atomicMin(&s_owner[cell - r0 * prm.horizon], i);
      }
    };
    constexpr int U = 4;
    const int stride = blockDim.x;
    for (int i0 = threadIdx.x - lane; i0 < n; i0 += U * stride) {     // warp-uniform trip count
      int rg[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const int i = i0 + lane + u * stride; rg[u] = i < n ? (int)__ldg(&ring[i]) : -1; }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const bool own = rg[u] >= r0 && rg[u] < r1;
        const unsigned m = __ballot_sync(0xffffffffu, own);
        if (m == 0u) continue;
        if (own) q[qn + __popc(m & ((1u << lane) - 1u))] = i0 + lane + u * stride;
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) {
          project32(32);
          __syncwarp();
          const int rest = qn - 32;
          const int t = lane < rest ? q[32 + lane] : 0;
          __syncwarp();
          if (lane < rest) q[lane] = t;
          qn = rest;
          __syncwarp();
        }
      }
    }
    project32(qn);
  }
  __syncthreads();
  // ---- F2: valid cells per chunk, exclusive scan over the group (row-major = extraction order) ----
  const int nch = nr * nchunk;
  for (int ch = wid; ch < nch; ch += nw) {
    const int rr = ch / nchunk, j = (ch - rr * nchunk) * 32 + lane;
    const bool v = j < prm.horizon && s_owner[rr * prm.horizon + j] != 0x7fffffff;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_chunk[ch] = __popc(m);
  }
  __syncthreads();
  if (wid == 0) {
    int carry = 0;
    for (int b0 = 0; b0 < nch; b0 += 32) {
      const int v = b0 + lane < nch ? s_chunk[b0 + lane] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (b0 + lane < nch) s_chunk[b0 + lane] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) {
      s_chunk[nch] = carry;
      __threadfence();
      atomicExch(&ctl[2 + ticket], carry + 1);                       // publish before waiting on anybody
    }
    // cells extracted by the groups before this one (same frame: tickets fi * G .. ticket - 1)
    int before = 0;
    for (int l0 = 0; l0 < g; l0 += 32) {
      int v = 0;
      if (l0 + lane < g) {
        const volatile int* slot = &ctl[2 + fi * G + l0 + lane];
        while ((v = *slot) == 0) __nanosleep(64);
        v -= 1;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      before += v;
    }
    if (lane == 0) s_base = before;
  }
  __syncthreads();
  const int base = s_base;
  if (threadIdx.x < nr) {
    const int b = base + s_chunk[threadIdx.x * nchunk], c = s_chunk[(threadIdx.x + 1) * nchunk] - s_chunk[threadIdx.x * nchunk];
    f.ring_count[r0 + threadIdx.x] = c;
    f.ring_start[r0 + threadIdx.x] = b - 1 + 5;                      // startRingIndex (:521)
    f.ring_end[r0 + threadIdx.x] = b + c - 1 - 5;                    // endRingIndex   (:537)
  }
  if (g == G - 1 && threadIdx.x == 0) *f.M = base + s_chunk[nch];
  for (int ch = wid; ch < nch; ch += nw) {
    const int rr = ch / nchunk, j = (ch - rr * nchunk) * 32 + lane;
    const int own = j < prm.horizon ? s_owner[rr * prm.horizon + j] : 0x7fffffff;
    const bool v = own != 0x7fffffff;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (v) {
      const int pos = base + s_chunk[ch] + __popc(m & ((1u << lane) - 1u));
      const float4 p = feat_load_point(f.pts, prm.lay, own);
      f.ext_pts[pos] = p;
      f.ext_src[pos] = own;
      f.col[pos] = (unsigned short)j;
      f.range[pos] = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
      f.label[pos] = 0;                                             // resetParameters (:61-75): cleared per extracted slot
      if (!prm.lean) { f.picked[pos] = 0; f.curv[pos] = 0.f; }
    }
  }
  // ---- replay safety: the last block to finish clears the control words ----
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(&ctl[1], 1) == F * G - 1; }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < F * G; i += blockDim.x) ctl[2 + i] = 0;
    if (threadIdx.x == 0) { ctl[0] = 0; ctl[1] = 0; }
  }
}

// F3 + F4 in one pass over the extracted ranges.  grid = (blocks, F).  F4's writes are idempotent ORs of 1 into
// flags that k_feat_compact cleared, F3 writes only its own slot: no ordering between the two is needed.
__global__ void k_feat_curv_occl(FeatFrame* frames) {
  const FeatFrame f = frames[blockIdx.y];
  const int M = *f.M;
  const float* __restrict__ r = f.range;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    if (i >= 5 && i < M - 5) {
      const float d = r[i - 5] + r[i - 4] + r[i - 3] + r[i - 2] + r[i - 1] - r[i] * 10 +
                      r[i + 1] + r[i + 2] + r[i + 3] + r[i + 4] + r[i + 5];   // exact op order (:549-553)
      f.curv[i] = d * d;
    }
    if (i >= 5 && i < M - 6) {
      const float depth1 = r[i], depth2 = r[i + 1];
      const int columnDiff = abs((int)f.col[i + 1] - (int)f.col[i]);
      if (columnDiff < 10) {
        if ((double)(depth1 - depth2) > 0.3) { for (int k = -5; k <= 0; k++) f.picked[i + k] = 1; }
        else if ((double)(depth2 - depth1) > 0.3) { for (int k = 1; k <= 6; k++) f.picked[i + k] = 1; }
      }
      const float diff1 = fabsf(r[i - 1] - r[i]);
      const float diff2 = fabsf(r[i + 1] - r[i]);
      if ((double)diff1 > 0.02 * (double)r[i] && (double)diff2 > 0.02 * (double)r[i]) f.picked[i] = 1;
    }
  }
}

// ---- F5 ----
constexpr int FEAT_WARPS = 4;            // rings per block
constexpr int FEAT_RING_MAX = 2048 + 16; // staged ring window (horizon <= 2048, +-6 apron)
#define KNN_SEG_INF __int_as_float(0x7f800000)
constexpr int FEAT_CH = 12;              // segment elements per lane: a segment holds <= 32 * 12 points (horizon <= 2048 -> <= 341)

// one warp per (frame, ring).  grid = (ceil(n_scan / FEAT_WARPS), F), block = 32 * FEAT_WARPS.
//
// The reference sorts every segment by curvature and walks it twice, picking a point when it has not been
// suppressed yet and suppressing its +-5 neighbours (order-dependent greedy non-maximum suppression).  Only the
// PICKS change state, so the walk is restated as a selection loop: "take the best not-yet-suppressed candidate,
// pick it, suppress its neighbours", which visits exactly the same picks in exactly the same order without ever
// sorting.  The segment's curvatures sit in shared memory (element sp + 32 t + lane = slot t of a lane, an alive bit per slot);
// one step = per-lane best (lowest / highest alive rank of the lane's pre-ranked slots: one ffs / clz), two warp
// reductions (redux.sync) for the winning (curvature, index) key, the pre-computed suppression reach of the winner,
// and an O(1) alive-mask update per lane.  ~40 steps per
// segment instead of a 512-key bitonic sort plus a 300-element sequential walk by one lane.
// FUSED: smoothness (F3) and the occlusion / isolated-point marks (F4) are computed here from the extracted ranges and
// columns instead of being read back from k_feat_curv_occl's arrays (the batched pipelines, where neither is an output).
// The marks are "pulled": condition bits of every point i (A: marks i-5..i, B: marks i+1..i+6, :575-593) go into two
// bit masks per ring window by ballot; a point is marked when any A bit of [k, k+5] or B bit of [k-6, k-1] is set, or
// when it is isolated itself (:597-603).
__device__ __forceinline__ float feat_curv_at(const float* __restrict__ r, int i, int M) {
  if (i < 5 || i >= M - 5) return 0.f;                          // never computed upstream: the cleared value
  const float d = r[i - 5] + r[i - 4] + r[i - 3] + r[i - 2] + r[i - 1] - r[i] * 10 +
                  r[i + 1] + r[i + 2] + r[i + 3] + r[i + 4] + r[i + 5];   // exact op order (:549-553)
  return d * d;
}
template <bool FUSED>
__global__ void __launch_bounds__(32 * FEAT_WARPS, 7)
k_feat_segments(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ring = blockIdx.x * FEAT_WARPS + wid;
  __shared__ unsigned char s_pick[FEAT_WARPS][FEAT_RING_MAX];
  __shared__ unsigned char s_reach[FEAT_WARPS][FEAT_RING_MAX];   // suppression reach of every point: right | left << 4
  __shared__ float s_cv[FEAT_WARPS][32 * FEAT_CH];
  __shared__ unsigned s_mask[FUSED ? FEAT_WARPS : 1][2][FUSED ? (FEAT_RING_MAX + 12) / 32 + 3 : 1];   // FUSED: occlusion condition bits A / B of the ring window
  __shared__ unsigned s_gap[FEAT_WARPS][FEAT_RING_MAX / 32 + 3];  // column-gap bits of the ring window               // curvature of the current segment, slot-major
  if (ring >= prm.n_scan) return;
  const unsigned FULL = 0xffffffffu;
  const int M = *f.M;
  const int start = f.ring_start[ring], end = f.ring_end[ring];
  unsigned char* spick = s_pick[wid];
  // ring = extracted indices [first, last] with first = start - 4, last = end + 5; window adds a +-6 apron
  const int lo = max(start - 4 - 6, 0);
  const int hi = min(end + 5 + 6, M - 1);
  const int wlen = max(hi - lo + 1, 0);
  unsigned char* sreach = s_reach[wid];
  unsigned* sgap = s_gap[wid];
  if (!FUSED) {
    for (int t = lane; t < wlen; t += 32) spick[t] = f.picked[lo + t];
  } else {
    const float* __restrict__ rr = f.range;
    const unsigned short* __restrict__ cc = f.col;
    unsigned* sA = s_mask[wid][0]; unsigned* sB = s_mask[wid][1];
    const int mlo = lo - 6;                                     // bit u of the masks <-> extracted index mlo + u
    const int nbits = wlen + 12;
    for (int u0 = 0; u0 < nbits + 32; u0 += 32) {               // one spare word: the funnel shifts read word w + 1
      const int i = mlo + u0 + lane;
      bool a = false, b = false;
      if (u0 + lane < nbits && i >= 5 && i < M - 6) {
        const float depth1 = rr[i], depth2 = rr[i + 1];
        if (abs((int)cc[i + 1] - (int)cc[i]) < 10) {
          if ((double)(depth1 - depth2) > 0.3) a = true;
          else if ((double)(depth2 - depth1) > 0.3) b = true;
        }
      }
      const unsigned ma = __ballot_sync(FULL, a), mb = __ballot_sync(FULL, b);
      if (lane == 0) { sA[u0 >> 5] = ma; sB[u0 >> 5] = mb; }
    }
    __syncwarp();
    for (int t = lane; t < wlen; t += 32) {
      const int k = lo + t;
      const int ua = t + 6, ub = t;                             // A bits of [k, k+5] start at u = k - mlo; B bits of [k-6, k-1] at u - 6
      const unsigned abits = __funnelshift_r(sA[ua >> 5], sA[(ua >> 5) + 1], ua & 31) & 0x3fu;
      const unsigned bbits = __funnelshift_r(sB[ub >> 5], sB[(ub >> 5) + 1], ub & 31) & 0x3fu;
      bool pk = (abits | bbits) != 0u;
      if (!pk && k >= 5 && k < M - 6) {
        const float r0 = rr[k];
        const float diff1 = fabsf(rr[k - 1] - r0), diff2 = fabsf(rr[k + 1] - r0);
        pk = (double)diff1 > 0.02 * (double)r0 && (double)diff2 > 0.02 * (double)r0;
      }
      spick[t] = pk ? 1 : 0;
    }
  }
  __syncwarp();
  // How far a pick of each point suppresses to the right / left (:648-661) depends on the column indices only:
  // gap bit t = "the walk cannot step from t to t+1" (column jump > 10, or t+1 outside the window / the cloud); the
  // reach of a point is the run of clear gap bits next to it, capped at 5 - one ffs / clz per point.
  for (int base = 0; base < wlen + 32; base += 32) {
    const int t = base + lane;
    const bool gap = (t + 1 >= wlen) || abs((int)f.col[lo + t + 1] - (int)f.col[lo + t]) > 10;   // column indices straight from global memory
    const unsigned m = __ballot_sync(FULL, gap);
    if (lane == 0) sgap[base >> 5] = m;
  }
  __syncwarp();
  for (int t = lane; t < wlen; t += 32) {
    const int w = t >> 5, b = t & 31;
    const unsigned right = __funnelshift_r(sgap[w], sgap[w + 1], b);                 // bit k = gap[t + k]
    const unsigned below = w > 0 ? sgap[w - 1] : 0xffffffffu;                         // before the window: gaps
    const unsigned left = b ? __funnelshift_l(below, sgap[w], 32 - b) : below;        // bit 31 - k = gap[t - 1 - k]
    const int rr = min(5, __ffs(right | 0x20u) - 1);
    const int rl = min(5, __clz(left));
    sreach[t] = (unsigned char)(rr | (rl << 4));
  }
  __syncwarp();
  for (int j = 0; j < 6; j++) {
    const int seg = ring * 6 + j;
    const int sp = (start * (6 - j) + end * j) / 6;
    const int ep = (start * (5 - j) + end * (j + 1)) / 6 - 1;
    if (lane == 0) { f.seg_sp[seg] = sp; f.seg_ep[seg] = ep; f.seg_valid[seg] = (sp < ep) ? 1 : 0; f.seg_ncorner[seg] = 0; f.seg_nflat[seg] = 0; }
    if (sp >= ep) continue;
    if (ep - sp + 1 > 32 * FEAT_CH) continue;   // unreachable for horizon <= 2048 (checked by the host entry points)
    // elements sp .. ep - 1 are the reference's sorted range: slot t of this lane = sp + 32 t + lane.  Each lane ranks
    // its own <= 12 slots by (curvature, index) once per segment; its best candidate is then the lowest / highest
    // alive RANK (one ffs / clz), looked up through two 4-bit-per-entry permutations held in registers.
    // Element ep is not sorted (quirk Q3): it is visited first in pass 1 and last in pass 2, handled apart.
    float* scv = s_cv[wid];
    float cv[FEAT_CH];
#pragma unroll
    for (int t = 0; t < FEAT_CH; t++) {
      const int idx = sp + 32 * t + lane;
      cv[t] = idx < ep ? (FUSED ? feat_curv_at(f.range, idx, M) : f.curv[idx]) : KNN_SEG_INF;
      scv[32 * t + lane] = cv[t];
    }
    unsigned long long slot_of_rank = 0ull, rank_of_slot = 0ull;
#pragma unroll
    for (int t = 0; t < FEAT_CH; t++) {
      int rk = 0;
#pragma unroll
      for (int u = 0; u < FEAT_CH; u++) if (u != t) rk += (cv[u] < cv[t] || (cv[u] == cv[t] && u < t)) ? 1 : 0;
      slot_of_rank |= (unsigned long long)t << (4 * rk);
      rank_of_slot |= (unsigned long long)rk << (4 * t);
    }
    const float cv_ep = FUSED ? feat_curv_at(f.range, ep, M) : f.curv[ep];
    auto mark = [&](int ind, unsigned& alive_r) {
      const int rch = sreach[ind - lo];
      const int a0 = ind - (rch >> 4), b0 = ind + (rch & 15);
      if (lane <= b0 - a0) spick[a0 + lane - lo] = 1;
      const int o = a0 - sp - lane; const int t0 = o <= 0 ? 0 : (o + 31) >> 5;
      if (t0 < FEAT_CH && sp + 32 * t0 + lane <= b0) alive_r &= ~(1u << (int)((rank_of_slot >> (4 * t0)) & 15ull));
    };
    // ---------------- pass 1: edges, largest curvature first (:633-663) ----------------
    unsigned alive = 0u;
#pragma unroll
    for (int t = 0; t < FEAT_CH; t++) {
      const int idx = sp + 32 * t + lane;
      if (idx < ep && cv[t] > prm.edge_thr && spick[idx - lo] == 0) alive |= 1u << (int)((rank_of_slot >> (4 * t)) & 15ull);
    }
    int nc = 0;
    const bool ep_edge = cv_ep > prm.edge_thr && spick[ep - lo] == 0;
    __syncwarp();                                               // every lane has read the flags before any pick rewrites them
    if (ep_edge) {                                              // k = ep comes first
      if (lane == 0) { f.label[ep] = 1; f.seg_corner[seg * 20 + nc] = ep; }
      nc++;
      mark(ep, alive);
    }
    for (;;) {
      // per-lane best = highest alive rank; key = (curvature bits + 1, index); 0 = no candidate.  ">=" semantics of
      // the descending walk (larger index first on equal curvature) are in the ranking
      unsigned bh = 0u, bl = 0u;
      if (alive) {
        const int r = 31 - __clz(alive), slot = (int)((slot_of_rank >> (4 * r)) & 15ull);
        bh = __float_as_uint(scv[32 * slot + lane]) + 1u; bl = (unsigned)(sp + 32 * slot + lane);
      }
      const unsigned mh = __reduce_max_sync(FULL, bh);
      if (mh == 0u) break;
      const int ind = (int)__reduce_max_sync(FULL, bh == mh ? bl : 0u);
      if (nc == 20) break;                       // the 21st pick ends the pass without being marked (:640-645)
      if (lane == 0) { f.label[ind] = 1; f.seg_corner[seg * 20 + nc] = ind; }
      nc++;
      mark(ind, alive);
    }
    __syncwarp();
    // ---------------- pass 2: flat points, smallest curvature first (:665-694) ----------------
    alive = 0u;
#pragma unroll
    for (int t = 0; t < FEAT_CH; t++) {
      const int idx = sp + 32 * t + lane;
      if (idx < ep && cv[t] < prm.surf_thr && spick[idx - lo] == 0) alive |= 1u << (int)((rank_of_slot >> (4 * t)) & 15ull);
    }
    int nf = 0;
    __syncwarp();                                               // flags read, picks may rewrite them
    for (;;) {
      // per-lane best = lowest alive rank; key = (curvature bits, index) ascending; 0xffffffff = no candidate
      unsigned bh = 0xffffffffu, bl = 0xffffffffu;
      if (alive) {
        const int r = __ffs(alive) - 1, slot = (int)((slot_of_rank >> (4 * r)) & 15ull);
        bh = __float_as_uint(scv[32 * slot + lane]); bl = (unsigned)(sp + 32 * slot + lane);
      }
      const unsigned mh = __reduce_min_sync(FULL, bh);
      if (mh == 0xffffffffu) break;
      const int ind = (int)__reduce_min_sync(FULL, bh == mh ? bl : 0xffffffffu);
      if (lane == 0) { f.label[ind] = -1; if (nf < 10) f.seg_flat[seg * 10 + nf] = ind; }
      if (nf < 10) nf++;
      mark(ind, alive);
    }
    __syncwarp();
    const bool ep_flat = cv_ep < prm.surf_thr && spick[ep - lo] == 0;
    __syncwarp();
    if (ep_flat) {                                              // k = ep comes last
      if (lane == 0) { f.label[ep] = -1; if (nf < 10) f.seg_flat[seg * 10 + nf] = ep; }
      if (nf < 10) nf++;
      unsigned dummy = 0u;
      mark(ep, dummy);
    }
    if (lane == 0) { f.seg_ncorner[seg] = nc; f.seg_nflat[seg] = nf; }
    __syncwarp();
  }
}

// compaction of the per-segment lists into the reference push order.  grid = (FEAT_GATHER_SPLIT, F): every block
// rebuilds the (cheap) exclusive offsets over all segments of its frame and then copies its share of the segments,
// one WARP per segment (ballot compaction keeps the index order of the surface list).
constexpr int FEAT_GATHER_SPLIT = 8;
__global__ void __launch_bounds__(256)
k_feat_gather(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int nseg = prm.n_scan * 6;
  __shared__ int s_c[1024], s_s[1024], s_f[1024], s_u[1024];   // exclusive offsets per segment (nseg <= 1024)
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    const int valid = f.seg_valid[s];
    const int nc = valid ? f.seg_ncorner[s] : 0;
    s_c[s] = nc; s_s[s] = nc < 4 ? nc : 4;
    s_f[s] = valid ? f.seg_nflat[s] : 0;
    // surface points of a segment = label <= 0 over [sp, ep]; only this segment's corner picks carry label 1 there
    s_u[s] = valid ? (f.seg_ep[s] - f.seg_sp[s] + 1) - nc : 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (wid < 4) {                                               // four warps, one array each: exclusive scan, 32 segments per round
    int* a = wid == 0 ? s_c : wid == 1 ? s_s : wid == 2 ? s_f : s_u;
    int carry = 0;
    for (int b0 = 0; b0 < nseg; b0 += 32) {
      const int v = b0 + lane < nseg ? a[b0 + lane] : 0;
      int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      if (b0 + lane < nseg) a[b0 + lane] = carry + x - v;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (blockIdx.x == 0 && lane == 0) f.counts[wid] = carry;
  }
  __syncthreads();
  for (int s = blockIdx.x * nw + wid; s < nseg; s += gridDim.x * nw) {
    if (!f.seg_valid[s]) continue;
    const int nc = f.seg_ncorner[s], nf = f.seg_nflat[s];
    if (lane < nc) f.corner_idx[s_c[s] + lane] = f.seg_corner[s * 20 + lane];
    if (lane < (nc < 4 ? nc : 4)) f.sharp_idx[s_s[s] + lane] = f.seg_corner[s * 20 + lane];
    if (lane < nf) f.flat_idx[s_f[s] + lane] = f.seg_flat[s * 10 + lane];
    int o = s_u[s];
    const int sp = f.seg_sp[s], ep = f.seg_ep[s];
    for (int base = sp; base <= ep; base += 32) {
      const int k = base + lane;
      const bool keep = k <= ep && f.label[k] <= 0;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) f.surf_idx[o + __popc(m & ((1u << lane) - 1u))] = k;
      o += __popc(m);
    }
  }
}

}  // namespace lisreg
