// Streaming odometry front end on the device (SURVEY.md 8f "next" #1): the sliding-window local map of
// OdomEstimationNode (USING_MULTI_FRAME_TARGET) kept resident in HBM.
//
// Reference: laserCloudInfoHandler odomEstimationNode.cpp:163-239 (map = concatenation of the last <= 19 key-frame
// clouds, newest first, :190-193, then VoxelGrid 0.2 / 0.4 m, :196-201), saveKeyFrames :421-478
// (transformPointCloud of the FULL corner / surface clouds of the frame by the refined pose, common.cpp:134-160,
// window trimmed while size >= 20, :463-467).  Here a key frame is appended by ONE gather + transform kernel straight
// from the feature extractor's output (the clouds never exist in the sensor frame as separate buffers), the window
// is a ring of fixed-capacity slots, and the concatenation is one kernel over a by-value slot table.
#pragma once
#include <cuda_runtime.h>

namespace lisreg {

constexpr int ODOM_MAX_SLOTS = 32;

struct OdomConcat {
  const float4* src[ODOM_MAX_SLOTS];
  int n[ODOM_MAX_SLOTS];
  int off[ODOM_MAX_SLOTS];
  int count;
};

// grid = (blocks, count): slot blockIdx.y copied to dst + off (laserCloud*FromMap += *laserCloud*Vec[i])
__global__ void k_odom_concat(OdomConcat t, float4* __restrict__ dst) {
  const int s = blockIdx.y;
  const float4* __restrict__ src = t.src[s];
  float4* __restrict__ d = dst + t.off[s];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.n[s]; i += gridDim.x * blockDim.x) d[i] = __ldg(&src[i]);
}

struct OdomT12 { float m[12]; };

// transformPointCloud(cloudIn, PointTypePose*) (common.cpp:134-160) fused with the gather of the frame's feature
// list: dst[i] = T * ext[idx[i]], intensity kept.  Products and sums left to right in fp32 like upstream.
__global__ void k_odom_append(const float4* __restrict__ ext, const int* __restrict__ idx, const int* __restrict__ n_ptr,
                              OdomT12 T, float4* __restrict__ dst, int cap) {
  const int n = min(*n_ptr, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&ext[idx[i]]);
    float4 q;
    q.x = T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z + T.m[3];
    q.y = T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z + T.m[7];
    q.z = T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z + T.m[11];
    q.w = p.w;
    dst[i] = q;
  }
}

// ---- local map / submap assembly (SURVEY.md 8f "next" #2): transform, box filter, order-preserving compaction ----
// transformPointCloud(cloudIn, PointTypePose*) without a gather: dst[i] = T * src[i]
__global__ void k_sm_transform(const float4* __restrict__ src, int n, OdomT12 T, float4* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&src[i]);
    float4 q;
    q.x = T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z + T.m[3];
    q.y = T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z + T.m[7];
    q.z = T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z + T.m[11];
    q.w = p.w;
    dst[i] = q;
  }
}
struct SmBox { double lo[3], hi[3]; };
// bbx_filter (subMap.h:1125-1150): 1 = strictly inside the box; flags has n + 1 entries (the last one 0) for the scan
__global__ void k_sm_box_flags(const float4* __restrict__ pts, int n, SmBox b, uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t f = 0u;
  if (i < n) {
    const float4 p = __ldg(&pts[i]);
    f = ((double)p.x > b.lo[0] && (double)p.x < b.hi[0] && (double)p.y > b.lo[1] && (double)p.y < b.hi[1] &&
         (double)p.z > b.lo[2] && (double)p.z < b.hi[2]) ? 1u : 0u;
  }
  flags[i] = f;
}
__global__ void k_sm_widen_flags(const unsigned char* __restrict__ keep, int n, uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  flags[i] = i < n ? (keep[i] ? 1u : 0u) : 0u;
}
// scanned = exclusive prefix of the flags (n + 1 entries): survivors keep their order
__global__ void k_sm_scatter(const float4* __restrict__ pts, int n, const uint32_t* __restrict__ scanned, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t a = scanned[i];
  if (scanned[i + 1] != a) out[a] = __ldg(&pts[i]);
}

}  // namespace lisreg
