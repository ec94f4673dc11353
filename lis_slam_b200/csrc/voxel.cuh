// pcl::VoxelGrid centroid down-sampling on the device, batched over independent clouds ("segments").
//
// Reference call sites: downSizeFilterCorner/Surf on the scan features (odomEstimationNode.cpp:272-277)
// and on the assembled map (:196-201); leaf sizes 0.2 / 0.4 m (config/params.yaml:132-133).
// PCL semantics restated in oracle/orc_voxel.cpp: voxel index from floor(p * inv_leaf) - min_b, one output
// point per occupied voxel in ASCENDING voxel index, value = fp32 centroid of x, y, z, intensity with the
// points of a voxel accumulated in ascending input index.
//
// Implementation: (1) per-cloud bounding box -> plan, (2) key = voxel index per point, (2b) RUNS: consecutive input points
// with the same key (a LiDAR ring crosses a 0.4 m voxel with ~5 consecutive returns) are collapsed into one (key, run id)
// entry - the sort below then moves ~5x fewer elements and a voxel's points stay a handful of contiguous index ranges,
// (3) stable LSD radix sort of the runs (8-bit digits, only the passes the largest voxel index needs; per pass: tile
// histograms -> per-cloud scan -> stable scatter with warp match-any ranking), (4) head flags + per-cloud scan -> voxel
// starts, (5) one thread per voxel walks its runs in order (= ascending input index: the sort is stable and runs are
// index-ordered) and sums their points.  The sorted order is also spatially coherent (x fastest), which is what the kNN
// kernel wants from its query stream.
#pragma once
#include <cstdint>
#include <cfloat>
#include <cuda_runtime.h>

namespace lisreg {

struct VoxPlan { float inv; int minb[3]; int mul[3]; int overflow; int npass; int n_runs; };   // npass: 8-bit radix passes the largest key needs; n_runs: entries to sort

struct VoxSeg {
  const float4* src;       // source cloud
  const int* gather;       // optional index list into src (NULL = identity)
  const int* n_ptr;        // optional device count (overrides n when non-NULL)
  int n;                   // number of input points
  float leaf;
  uint32_t* key_a; uint32_t* val_a; uint32_t* key_b; uint32_t* val_b;   // capacity cap each
  uint32_t* hist;          // 256 * nblk_cap
  int* seg_start;          // cap + 1 : voxel starts (positions in the sorted run arrays)
  int* run_start;          // cap + 1 : first input index of every run (run r = input points [run_start[r], run_start[r + 1]))
  VoxPlan* plan;
  unsigned* bbox;          // 6 order-preserving encoded floats (min xyz, max xyz)
  float4* out;             // cap
  int* out_n;              // number of voxels
  int cap;
  int gather_increasing;   // the index list is strictly increasing (a surface list; NOT a corner list, which is in pick order): lets k_vox_block detect runs that are contiguous in src
  float bound;             // > 0: every coordinate is known to lie in [-bound, bound] (range-gated sweep points): lets k_vox_block skip its bounding-box pass
};

constexpr int RS_TILE = 2048;      // keys per block per pass
constexpr int RS_THREADS = 256;

__device__ __forceinline__ int vox_n(const VoxSeg& s) { return s.n_ptr ? *s.n_ptr : s.n; }
__device__ __forceinline__ float4 vox_point(const VoxSeg& s, int i) { return __ldg(&s.src[s.gather ? s.gather[i] : i]); }

// (1) bounding box (multi-block, order-preserving integer atomics) + plan
__device__ __forceinline__ unsigned vox_f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float vox_ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_vox_bbox_init(VoxSeg* segs, int nseg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseg * 6) return;
  segs[i / 6].bbox[i % 6] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

// grid = (blocks, nseg)
__global__ void k_vox_bbox(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.y];
  const int n = vox_n(s);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = vox_point(s, i);
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    any = true;
  }
  if (!__any_sync(0xffffffffu, any)) return;
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { atomicMin(&s.bbox[d], vox_f2ord(mn[d])); atomicMax(&s.bbox[3 + d], vox_f2ord(mx[d])); }
  }
}

// one thread per cloud
__global__ void k_vox_plan(VoxSeg* segs, int nseg) {
  const int si = blockIdx.x * blockDim.x + threadIdx.x;
  if (si >= nseg) return;
  const VoxSeg s = segs[si];
  const int n = vox_n(s);
  VoxPlan p;
  p.inv = 1.0f / s.leaf;
  p.overflow = 0;
  if (n > 0) {
    float mn[3], mx[3];
    for (int d = 0; d < 3; d++) { mn[d] = vox_ord2f(s.bbox[d]); mx[d] = vox_ord2f(s.bbox[3 + d]); }
    const long long dx = (long long)((mx[0] - mn[0]) * p.inv) + 1, dy = (long long)((mx[1] - mn[1]) * p.inv) + 1,
                    dz = (long long)((mx[2] - mn[2]) * p.inv) + 1;
    if (dx * dy * dz > 2147483647LL) p.overflow = 1;   // PCL: "leaf size is too small" -> output = input
    int divb[3];
    for (int d = 0; d < 3; d++) {
      p.minb[d] = (int)floorf(mn[d] * p.inv);
      divb[d] = (int)floorf(mx[d] * p.inv) - p.minb[d] + 1;
    }
    p.mul[0] = 1; p.mul[1] = divb[0]; p.mul[2] = divb[0] * divb[1];
    // keys are < divb0 * divb1 * divb2 (or < n in the overflow case): passes over all-zero high digits are skipped
    const unsigned long long kmax = p.overflow ? (unsigned long long)n : (unsigned long long)divb[0] * (unsigned long long)divb[1] * (unsigned long long)divb[2];
    p.npass = kmax > (1ull << 24) ? 4 : kmax > (1ull << 16) ? 3 : kmax > (1ull << 8) ? 2 : 1;
  } else {
    for (int d = 0; d < 3; d++) { p.minb[d] = 0; p.mul[d] = 0; }
    p.npass = 0;
  }
  p.n_runs = 0;
  *s.plan = p;
}

// (2) keys.  grid = (blocks, nseg)
__global__ void k_vox_keys(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.y];
  const int n = vox_n(s);
  const VoxPlan p = *s.plan;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t key;
    if (p.overflow) key = (uint32_t)i;
    else {
      const float4 q = vox_point(s, i);
      const int i0 = (int)(floorf(q.x * p.inv) - (float)p.minb[0]);
      const int i1 = (int)(floorf(q.y * p.inv) - (float)p.minb[1]);
      const int i2 = (int)(floorf(q.z * p.inv) - (float)p.minb[2]);
      key = (uint32_t)(i0 * p.mul[0] + i1 * p.mul[1] + i2 * p.mul[2]);
    }
    s.key_b[i] = key;          // per-point keys: input of the run detection, which fills key_a / val_a
  }
}

// (3a) tile histograms: hist[digit * nblk + blk].  grid = (nblk_max, nseg)
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(VoxSeg* segs, int shift, int flip) {
  const VoxSeg s = segs[blockIdx.y];
  const int n = s.plan->n_runs;
  const int nblk = (n + RS_TILE - 1) / RS_TILE;
  if ((int)blockIdx.x >= nblk || shift >= 8 * s.plan->npass) return;
  const uint32_t* key = flip ? s.key_b : s.key_a;
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
  for (int t = threadIdx.x; t < RS_TILE; t += RS_THREADS) {
    const int i = base + t;
    if (i < n) atomicAdd(&h[(key[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  s.hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// (3b) exclusive scan over the 256 * nblk counters of one cloud (digit-major: hist[digit * nblk + tile]), in two steps so
// that one 2 M-point cloud (the sliding-window map of the streaming odometry: 1000 tiles) is scanned by 256 warps instead
// of one block:  k_rs_scan_digit - one WARP per digit: exclusive scan over the tiles of its digit, digit total -> dbase;
// k_rs_scan_base - one warp per cloud: exclusive scan of the 256 digit totals.  k_rs_scatter adds dbase[digit].
// dbase lives behind the histogram matrix: hist[256 * nblk_cap .. + 256).
__device__ __forceinline__ int rs_nblk_cap(const VoxSeg& s) { return (s.cap + RS_TILE - 1) / RS_TILE + 1; }

// grid = (32, nseg), block = 256 (8 warps = 8 digits per block)
__global__ void __launch_bounds__(256)
k_rs_scan_digit(VoxSeg* segs, int shift) {
  const VoxSeg s = segs[blockIdx.y];
  if (shift >= 8 * s.plan->npass) return;
  const int n = s.plan->n_runs;
  const int nblk = (n + RS_TILE - 1) / RS_TILE;
  const int lane = threadIdx.x & 31, d = blockIdx.x * 8 + (threadIdx.x >> 5);
  uint32_t* __restrict__ h = s.hist + (size_t)d * nblk;
  uint32_t carry = 0u;
  for (int base = 0; base < nblk; base += 32) {
    const int i = base + lane;
    const uint32_t v = i < nblk ? h[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (i < nblk) h[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) s.hist[(size_t)256 * rs_nblk_cap(s) + d] = carry;
}
// grid = ceil(nseg / 8), block = 256: one warp per cloud
__global__ void __launch_bounds__(256)
k_rs_scan_base(VoxSeg* segs, int nseg, int shift) {
  const int si = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (si >= nseg) return;
  const VoxSeg s = segs[si];
  if (shift >= 8 * s.plan->npass) return;
  uint32_t* __restrict__ db = s.hist + (size_t)256 * rs_nblk_cap(s);
  uint32_t carry = 0u;
  for (int base = 0; base < 256; base += 32) {
    const uint32_t v = db[base + lane];
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    db[base + lane] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// (3b, many small clouds - the batched frame pipeline) exclusive scan over the 256 * nblk counters of one cloud by ONE block;
// zeroes the digit bases so that k_rs_scatter can add them unconditionally
__global__ void k_rs_scan(VoxSeg* segs, int shift) {
  const VoxSeg s = segs[blockIdx.x];
  if (shift >= 8 * s.plan->npass) return;
  const int n = s.plan->n_runs;
  const int nblk = (n + RS_TILE - 1) / RS_TILE;
  const int total = 256 * nblk;
  __shared__ uint32_t sm[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x < 256) s.hist[(size_t)256 * rs_nblk_cap(s) + threadIdx.x] = 0u;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < total ? s.hist[i] : 0u;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const uint32_t t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0u;
      __syncthreads();
      sm[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < total) s.hist[i] = sm[threadIdx.x] - v + carry;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sm[1023];
    __syncthreads();
  }
}

// (3c) stable scatter.  grid = (nblk_max, nseg).  Warp w owns keys [w*256, w*256+256) of the tile.
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(VoxSeg* segs, int shift, int flip) {
  const VoxSeg s = segs[blockIdx.y];
  const int n = s.plan->n_runs;
  const int nblk = (n + RS_TILE - 1) / RS_TILE;
  if ((int)blockIdx.x >= nblk || shift >= 8 * s.plan->npass) return;
  const uint32_t* key = flip ? s.key_b : s.key_a;
  const uint32_t* val = flip ? s.val_b : s.val_a;
  uint32_t* okey = flip ? s.key_a : s.key_b;
  uint32_t* oval = flip ? s.val_a : s.val_b;
  constexpr int NW = RS_THREADS / 32, PER_WARP = RS_TILE / NW, ROUNDS = PER_WARP / 32;
  __shared__ uint32_t wh[NW][256];     // per-warp digit counts, then running offsets
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) wh[wid][d] = 0;
  __syncwarp();
  const int base = blockIdx.x * RS_TILE + wid * PER_WARP;
  uint32_t k[ROUNDS], v[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int i = base + r * 32 + lane;
    const bool ok = i < n;
    k[r] = ok ? key[i] : 0xffffffffu; v[r] = ok ? val[i] : 0u;
    const uint32_t dgt = (k[r] >> shift) & 255u;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const unsigned m = __match_any_sync(act, dgt);
      if (lane == __ffs(m) - 1) wh[wid][dgt] += __popc(m);
    }
    __syncwarp();
  }
  __syncthreads();
  // exclusive prefix over warps per digit + global base of this (digit, tile)
  {
    const int d = threadIdx.x;   // 256 threads == 256 digits
    uint32_t acc = s.hist[d * nblk + blockIdx.x] + s.hist[(size_t)256 * rs_nblk_cap(s) + d];   // tile prefix inside the digit + digit base
#pragma unroll
    for (int w = 0; w < NW; w++) { const uint32_t t = wh[w][d]; wh[w][d] = acc; acc += t; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int i = base + r * 32 + lane;
    const bool ok = i < n;
    const uint32_t dgt = (k[r] >> shift) & 255u;
    const unsigned act = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const unsigned m = __match_any_sync(act, dgt);
      const uint32_t pos = wh[wid][dgt] + __popc(m & ((1u << lane) - 1u));
      okey[pos] = k[r]; oval[pos] = v[r];
    }
    __syncwarp();
    if (ok) {
      const unsigned m = __match_any_sync(act, dgt);
      if (lane == __ffs(m) - 1) wh[wid][dgt] += __popc(m);
    }
    __syncwarp();
  }
}

// (4) voxel starts (positions of the first entry of every voxel in the sorted arrays).  After npass passes the sorted
// data is in *_a (even) or *_b (odd).  Three fully parallel steps (a 2 M-point map has 1000 chunks):
//   k_vox_head_count  grid (chunks, nseg): heads per chunk of VH_CHUNK sorted keys -> hist[chunk] (the histogram matrix is free now)
//   k_vox_head_scan   one warp per cloud: exclusive scan of the chunk counts, out_n, seg_start[out_n] = n
//   k_vox_head_write  grid (chunks, nseg): rank of every head inside its chunk (ballot) + chunk offset -> seg_start
constexpr int VH_CHUNK = RS_TILE;      // same tiling as the sort: (cap + RS_TILE - 1) / RS_TILE + 1 counters fit the histogram area
__device__ __forceinline__ const uint32_t* vox_sorted_keys(const VoxSeg& s) { return (s.plan->npass & 1) ? s.key_b : s.key_a; }

// RUNS = true: the same three steps on the UNSORTED per-point keys (key_b) detect the runs of equal consecutive keys and
// fill the sort input: key_a[r] = key, val_a[r] = r, run_start[r] = first input index, plan->n_runs.
template <bool RUNS>
__global__ void __launch_bounds__(256)
k_vox_head_count(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.y];
  const int n = RUNS ? vox_n(s) : s.plan->n_runs;
  const int c0 = blockIdx.x * VH_CHUNK;
  if (c0 >= n) return;
  const uint32_t* __restrict__ skey = RUNS ? s.key_b : vox_sorted_keys(s);
  int cnt = 0;
  for (int i = c0 + threadIdx.x; i < min(c0 + VH_CHUNK, n); i += blockDim.x) cnt += (i == 0 || skey[i] != skey[i - 1]) ? 1 : 0;
  __shared__ int s_w[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_w[w]; s.hist[blockIdx.x] = (uint32_t)t; }
}
// grid = ceil(nseg / 8), block = 256: one warp per cloud
template <bool RUNS>
__global__ void __launch_bounds__(256)
k_vox_head_scan(VoxSeg* segs, int nseg) {
  const int si = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (si >= nseg) return;
  const VoxSeg s = segs[si];
  const int n = RUNS ? vox_n(s) : s.plan->n_runs;
  const int nchunk = (n + VH_CHUNK - 1) / VH_CHUNK;
  uint32_t carry = 0u;
  for (int base = 0; base < nchunk; base += 32) {
    const int i = base + lane;
    const uint32_t v = i < nchunk ? s.hist[i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (i < nchunk) s.hist[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    if (RUNS) { s.run_start[carry] = n; s.plan->n_runs = (int)carry; }
    else { s.seg_start[carry] = n; *s.out_n = (int)carry; }
  }
}
template <bool RUNS>
__global__ void __launch_bounds__(256)
k_vox_head_write(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.y];
  const int n = RUNS ? vox_n(s) : s.plan->n_runs;
  const int c0 = blockIdx.x * VH_CHUNK;
  if (c0 >= n) return;
  const uint32_t* __restrict__ skey = RUNS ? s.key_b : vox_sorted_keys(s);
  __shared__ int s_w[8];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = (int)s.hist[blockIdx.x];
  __syncthreads();
  for (int base = c0; base < min(c0 + VH_CHUNK, n); base += 256) {
    const int i = base + threadIdx.x;
    const int head = (i < n && i < c0 + VH_CHUNK && (i == 0 || skey[i] != skey[i - 1])) ? 1 : 0;
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (lane == 0) s_w[wid] = __popc(m);
    __syncthreads();
    int off = s_carry;
    for (int w = 0; w < wid; w++) off += s_w[w];
    if (head) {
      const int r = off + __popc(m & ((1u << lane) - 1u));
      if (RUNS) { s.run_start[r] = i; s.key_a[r] = skey[i]; s.val_a[r] = (uint32_t)r; }
      else s.seg_start[r] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; w++) t += s_w[w]; s_carry += t; }
    __syncthreads();
  }
}

// (4, many small clouds) voxel starts: one block (1024 threads) per cloud.  After npass passes the sorted data is in *_a (even) or *_b (odd).
template <bool RUNS>
__global__ void k_vox_heads(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.x];
  const int n = RUNS ? vox_n(s) : s.plan->n_runs;
  const uint32_t* skey = RUNS ? s.key_b : ((s.plan->npass & 1) ? s.key_b : s.key_a);
  __shared__ int s_w[32];
  __shared__ int carry, s_tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int head = (i < n && (i == 0 || skey[i] != skey[i - 1])) ? 1 : 0;
    // rank of a head = carry + heads before it: ballot inside the warp, the 32 warp totals scanned by warp 0
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (lane == 0) s_w[wid] = __popc(m);
    __syncthreads();
    if (wid == 0) {
      const int v = s_w[lane];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      s_w[lane] = incl - v;                      // exclusive warp offsets
      if (lane == 31) s_tot = incl;
    }
    __syncthreads();
    if (head) {
      const int r = carry + s_w[wid] + __popc(m & ((1u << lane) - 1u));
      if (RUNS) { s.run_start[r] = i; s.key_a[r] = skey[i]; s.val_a[r] = (uint32_t)r; }
      else s.seg_start[r] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += s_tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (RUNS) { s.run_start[carry] = n; s.plan->n_runs = carry; }
    else { s.seg_start[carry] = n; *s.out_n = carry; }
  }
}

// ---- steps (1)-(4) for ONE cloud in ONE block, with the sort in shared memory ----
// A frame's feature cloud (<= ~130 k points, ~17 k runs) is small enough that everything between "points" and "voxel
// starts" fits one SM: the block reads the points twice (bounding box, then keys + runs; the second read hits L2), keeps
// the run keys and two 16-bit index arrays in shared memory (8 B per run), runs the stable LSD radix passes there (per-warp
// digit counters, match-any ranking - the same scheme as k_rs_scatter without the histogram matrix in HBM) and finds
// the voxel heads on the sorted keys.  HBM sees the points, run_start, the sorted run ids and the voxel starts - no
// per-point key array, no ping-pong buffers, no histogram matrix, and 1 launch instead of 19.  A cloud with more than
// VB_CAP runs sorts through its global scratch arrays (same code, 32-bit ids): slow, correct, rare.
__device__ __forceinline__ VoxPlan vox_make_plan(const float* mn, const float* mx, int n, float leaf) {   // = k_vox_plan
  VoxPlan p;
  p.inv = 1.0f / leaf;
  p.overflow = 0;
  if (n > 0) {
    const long long dx = (long long)((mx[0] - mn[0]) * p.inv) + 1, dy = (long long)((mx[1] - mn[1]) * p.inv) + 1,
                    dz = (long long)((mx[2] - mn[2]) * p.inv) + 1;
    if (dx * dy * dz > 2147483647LL) p.overflow = 1;   // PCL: "leaf size is too small" -> output = input
    int divb[3];
    for (int d = 0; d < 3; d++) {
      p.minb[d] = (int)floorf(mn[d] * p.inv);
      divb[d] = (int)floorf(mx[d] * p.inv) - p.minb[d] + 1;
    }
    p.mul[0] = 1; p.mul[1] = divb[0]; p.mul[2] = divb[0] * divb[1];
    const unsigned long long kmax = p.overflow ? (unsigned long long)n : (unsigned long long)divb[0] * (unsigned long long)divb[1] * (unsigned long long)divb[2];
    p.npass = kmax > (1ull << 24) ? 4 : kmax > (1ull << 16) ? 3 : kmax > (1ull << 8) ? 2 : 1;
  } else {
    for (int d = 0; d < 3; d++) { p.minb[d] = 0; p.mul[d] = 0; }
    p.npass = 0;
  }
  p.n_runs = 0;
  return p;
}

constexpr int VB_THREADS = 1024;
constexpr int VB_LEN_SAT = 16383;           // run lengths are packed in 14 bits next to the 17-bit first index (clouds of <= 131072 points) and a flag
constexpr int VB_PER = 4;                 // points per thread and round of the two passes over the cloud (loads in flight)
constexpr int VB_CAP = 22528;              // runs sorted in shared memory: 8 B each + 32 KB of digit counters
constexpr size_t VB_SMEM = (size_t)VB_CAP * 8 + 32 * 256 * 4;
constexpr int VOX_BLOCK_MAX_N_FEW = 49152;  // ... and so do clouds above this size when there are only a few of them (no other block to fill the SMs)
constexpr int VOX_BLOCK_MAX_N = 131072;    // larger clouds (the 2 M-point window map) take the multi-block kernels

template <typename V>
__device__ __forceinline__ void vb_sort(const uint32_t* key, V* va, V* vb, int n, int npass, uint32_t* cnt, uint32_t* s_dtot) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int chunk = (((n + 31) / 32) + 31) & ~31;            // contiguous range of a warp, a multiple of 32
  const int e_begin = wid * chunk, e_end = min(n, e_begin + chunk);
  for (int p = 0; p < npass; p++) {
    const int shift = 8 * p;
    const V* in = (p & 1) ? va : vb;                         // pass 0 reads the identity
    V* out = (p & 1) ? vb : va;
    for (int d = lane; d < 256; d += 32) cnt[wid * 256 + d] = 0u;
    __syncwarp();
    for (int e0 = e_begin; e0 < e_end; e0 += 32) {
      const int e = e0 + lane;
      const bool ok = e < e_end;
      const uint32_t v = ok ? (p == 0 ? (uint32_t)e : (uint32_t)in[e]) : 0u;
      const uint32_t dgt = ok ? (key[v] >> shift) & 255u : 0u;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const unsigned m = __match_any_sync(act, dgt);
        if (lane == __ffs(m) - 1) cnt[wid * 256 + dgt] += __popc(m);
      }
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 256) {                                   // exclusive prefix over the warps of one digit
      uint32_t acc = 0u;
#pragma unroll 8
      for (int w = 0; w < VB_THREADS / 32; w++) { const uint32_t t = cnt[w * 256 + threadIdx.x]; cnt[w * 256 + threadIdx.x] = acc; acc += t; }
      s_dtot[threadIdx.x] = acc;
    }
    __syncthreads();
    if (wid == 0) {                                            // exclusive scan of the 256 digit totals: 8 digits per lane
      uint32_t loc[8], sum = 0u;
#pragma unroll
      for (int k = 0; k < 8; k++) { loc[k] = s_dtot[lane * 8 + k]; sum += loc[k]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      uint32_t acc = incl - sum;
#pragma unroll
      for (int k = 0; k < 8; k++) { s_dtot[lane * 8 + k] = acc; acc += loc[k]; }
    }
    __syncthreads();
    for (int e0 = e_begin; e0 < e_end; e0 += 32) {
      const int e = e0 + lane;
      const bool ok = e < e_end;
      const uint32_t v = ok ? (p == 0 ? (uint32_t)e : (uint32_t)in[e]) : 0u;
      const uint32_t dgt = ok ? (key[v] >> shift) & 255u : 0u;
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      unsigned m = 0u;
      if (ok) {
        m = __match_any_sync(act, dgt);
        out[s_dtot[dgt] + cnt[wid * 256 + dgt] + __popc(m & ((1u << lane) - 1u))] = (V)v;
      }
      __syncwarp();
      if (ok && lane == __ffs(m) - 1) cnt[wid * 256 + dgt] += __popc(m);
      __syncwarp();
    }
    __syncthreads();
  }
}

// exclusive block scan of one small count per thread (1024 threads): returns the offset of this thread and the block total
__device__ __forceinline__ int vb_rank(int count, int* s_w, int* s_tot, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_w[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int v = s_w[lane];
    int wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
    s_w[lane] = wi - v;
    if (lane == 31) *s_tot = wi;
  }
  __syncthreads();
  const int r = s_w[wid] + incl - count;
  total = *s_tot;
  __syncthreads();                                             // s_w / s_tot may be rewritten by the next call
  return r;
}

// CENTROID: the block also computes the centroids (batched frames: hundreds of clouds keep every SM busy); without it the
// voxel starts / sorted run ids go to global memory for k_vox_centroid_warp (a single frame: two clouds, the centroids
// are better spread over the whole GPU).
template <bool CENTROID>
__global__ void __launch_bounds__(VB_THREADS, 1)
k_vox_block(VoxSeg* segs) {
  extern __shared__ __align__(16) unsigned char vb_smem[];
  uint32_t* s_key = reinterpret_cast<uint32_t*>(vb_smem);
  uint16_t* s_va = reinterpret_cast<uint16_t*>(s_key + VB_CAP);
  uint16_t* s_vb = s_va + VB_CAP;
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_vb + VB_CAP);   // [32][256]; the key exchange of the run detection before the sort
  __shared__ float s_red[6][32];
  __shared__ int s_w[32];
  __shared__ int s_tot;
  __shared__ uint32_t s_lastkey;
  __shared__ uint32_t s_dtot[256];
  __shared__ VoxPlan s_plan;
  const VoxSeg s = segs[blockIdx.x];
  const int n = vox_n(s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // ---- (1) bounding box + plan (k_vox_bbox / k_vox_plan) ----
  // Only the ORDER of the voxel indices and the voxel a point falls into reach the output, and both are the same for any
  // box that contains the cloud (the index is lexicographic in (z, y, x) cells either way).  A caller that knows a bound
  // (sweep points passed the range gate) saves the pass over the points - unless the bound is so loose that the index
  // would overflow, where PCL's own check needs the real box.
  bool planned = false;
  if (s.bound > 0.f && n > 0) {
    const float bn[3] = {-s.bound, -s.bound, -s.bound}, bx[3] = {s.bound, s.bound, s.bound};
    const VoxPlan hp = vox_make_plan(bn, bx, n, s.leaf);
    if (!hp.overflow) { planned = true; if (threadIdx.x == 0) s_plan = hp; __syncthreads(); }
  }
  if (!planned) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i0 = threadIdx.x; i0 < n; i0 += VB_PER * VB_THREADS) {      // VB_PER points in flight per thread
      int idx[VB_PER]; float4 q[VB_PER];
#pragma unroll
      for (int k = 0; k < VB_PER; k++) { const int i = min(i0 + k * VB_THREADS, n - 1); idx[k] = s.gather ? __ldg(&s.gather[i]) : i; }
#pragma unroll
      for (int k = 0; k < VB_PER; k++) q[k] = __ldg(&s.src[idx[k]]);
#pragma unroll
      for (int k = 0; k < VB_PER; k++) {
        mn[0] = fminf(mn[0], q[k].x); mn[1] = fminf(mn[1], q[k].y); mn[2] = fminf(mn[2], q[k].z);
        mx[0] = fmaxf(mx[0], q[k].x); mx[1] = fmaxf(mx[1], q[k].y); mx[2] = fmaxf(mx[2], q[k].z);
      }
    }
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
        mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
      }
    if (lane == 0) { for (int d = 0; d < 3; d++) { s_red[d][wid] = mn[d]; s_red[3 + d][wid] = mx[d]; } }
    __syncthreads();
    if (threadIdx.x == 0) {
      float bn[3] = {0.f, 0.f, 0.f}, bx[3] = {0.f, 0.f, 0.f};
      if (n > 0)
        for (int d = 0; d < 3; d++) {
          bn[d] = s_red[d][0]; bx[d] = s_red[3 + d][0];
          for (int w = 1; w < VB_THREADS / 32; w++) { bn[d] = fminf(bn[d], s_red[d][w]); bx[d] = fmaxf(bx[d], s_red[3 + d][w]); }
        }
      s_plan = vox_make_plan(bn, bx, n, s.leaf);
    }
    __syncthreads();
  }
  const VoxPlan p = s_plan;
  // ---- (2) keys (k_vox_keys) and runs of equal consecutive keys: VB_PER consecutive points per thread ----
  int n_runs = 0;
  int idx_next[VB_PER];                                         // the index list runs one round ahead of the points
#pragma unroll
  for (int k = 0; k < VB_PER; k++) { const int i = max(0, min(VB_PER * (int)threadIdx.x + k, n - 1)); idx_next[k] = (s.gather && n > 0) ? __ldg(&s.gather[i]) : i; }
  for (int base = 0; base < n; base += VB_PER * VB_THREADS) {
    const int i0 = base + VB_PER * threadIdx.x;
    uint32_t key[VB_PER];
    if (p.overflow) {
#pragma unroll
      for (int k = 0; k < VB_PER; k++) key[k] = (uint32_t)(i0 + k);
    } else {
      int idx[VB_PER]; float4 q[VB_PER];
#pragma unroll
      for (int k = 0; k < VB_PER; k++) idx[k] = idx_next[k];
#pragma unroll
      for (int k = 0; k < VB_PER; k++) q[k] = __ldg(&s.src[idx[k]]);
#pragma unroll
      for (int k = 0; k < VB_PER; k++) { const int i = min(i0 + VB_PER * VB_THREADS + k, n - 1); idx_next[k] = s.gather ? __ldg(&s.gather[i]) : i; }
#pragma unroll
      for (int k = 0; k < VB_PER; k++) {
        const int c0 = (int)(floorf(q[k].x * p.inv) - (float)p.minb[0]);
        const int c1 = (int)(floorf(q[k].y * p.inv) - (float)p.minb[1]);
        const int c2 = (int)(floorf(q[k].z * p.inv) - (float)p.minb[2]);
        key[k] = (uint32_t)(c0 * p.mul[0] + c1 * p.mul[1] + c2 * p.mul[2]);
      }
    }
    s_cnt[threadIdx.x] = key[VB_PER - 1];                      // (a clamped index repeats the last valid point: same key)
    __syncthreads();
    uint32_t prev = threadIdx.x ? s_cnt[threadIdx.x - 1] : s_lastkey;
    int head[VB_PER], cnt = 0;
#pragma unroll
    for (int k = 0; k < VB_PER; k++) {
      const int i = i0 + k;
      head[k] = (i < n && (i == 0 || key[k] != prev)) ? 1 : 0;
      cnt += head[k];
      prev = key[k];
    }
    int total;
    int r = n_runs + vb_rank(cnt, s_w, &s_tot, total);          // (its barriers also order the exchange buffer)
#pragma unroll
    for (int k = 0; k < VB_PER; k++)
      if (head[k]) { s.run_start[r] = i0 + k; s.key_a[r] = key[k]; if (r < VB_CAP) s_key[r] = key[k]; r++; }
    if (threadIdx.x == VB_THREADS - 1) s_lastkey = key[VB_PER - 1];
    n_runs += total;
  }
  __syncthreads();
  if (threadIdx.x == 0) { s.run_start[n_runs] = n; VoxPlan q = p; q.n_runs = n_runs; *s.plan = q; }
  // ---- (3) stable LSD radix sort of the run ids by key; (4) voxel heads on the sorted order ----
  uint32_t* sval_out = (p.npass & 1) ? s.val_b : s.val_a;       // where k_vox_centroid* expect the sorted run ids
  const bool in_smem = n_runs <= VB_CAP;
  if (in_smem) vb_sort<uint16_t>(s_key, s_va, s_vb, n_runs, p.npass, s_cnt, s_dtot);
  else {
    __syncthreads();                                           // key_a of this block's own writes
    vb_sort<uint32_t>(s.key_a, s.val_b, s.val_a, n_runs, p.npass, s_cnt, s_dtot);   // pass 0 writes val_b: an odd pass count ends there
  }
  const uint16_t* fin16 = (p.npass & 1) ? s_va : s_vb;
  uint16_t* s_seg = (p.npass & 1) ? s_vb : s_va;                // CENTROID: voxel starts stay in shared memory (the idle id array)
  const bool meta_smem = CENTROID && in_smem;
  int n_vox = 0;
  for (int base = 0; base < n_runs; base += VB_THREADS) {
    const int i = base + threadIdx.x;
    int head = 0;
    uint32_t v = 0u;
    if (i < n_runs) {
      uint32_t k, kp = 0u;
      if (in_smem) { v = fin16[i]; k = s_key[v]; if (i) kp = s_key[fin16[i - 1]]; }
      else { v = sval_out[i]; k = s.key_a[v]; if (i) kp = s.key_a[sval_out[i - 1]]; }
      head = (i == 0 || k != kp) ? 1 : 0;
      if (in_smem && !meta_smem) sval_out[i] = v;
    }
    int total;
    const int r = n_vox + vb_rank(head, s_w, &s_tot, total);
    if (head) { if (meta_smem) s_seg[r] = (uint16_t)i; else s.seg_start[r] = i; }
    n_vox += total;
  }
  if (threadIdx.x == 0) { if (!meta_smem) s.seg_start[n_vox] = n_runs; *s.out_n = n_vox; }
  if (!CENTROID) return;
  __syncthreads();
  // ---- (5) centroids by this block: one thread per voxel, points added in ascending input index (see k_vox_centroid) ----
  // Shared-memory case: the keys are dead now, their array takes (first point, length) of every run in SORTED order, so a
  // voxel's thread reads its runs from shared memory and only the index list and the points themselves from L2 / HBM.
  // s_run[i] = first index (17 bits) | min(length, VB_LEN_SAT) << 17 | direct << 31.  direct: the run's points are CONSECUTIVE
  // in the source cloud (always without an index list; with one, whenever the list does not skip inside the run - a feature
  // list only skips the few corner picks), so the first index is stored as a SOURCE index and the points are read without
  // going through the list: one dependent load per batch instead of two.
  uint32_t* s_run = s_key;
  const float4* __restrict__ src = s.src;
  const int* __restrict__ gat = s.gather;
  if (meta_smem) {
    for (int i = threadIdx.x; i < n_runs; i += VB_THREADS) {
      const int v = fin16[i];
      const int a = s.run_start[v], len = s.run_start[v + 1] - a;
      uint32_t first = (uint32_t)a, direct = 1u;
      if (gat && !s.gather_increasing) direct = 0u;
      else if (gat) {
        const int g0 = __ldg(&gat[a]), g1 = __ldg(&gat[a + len - 1]);
        direct = (g1 - g0 == len - 1 && (unsigned)g0 < 0x20000u) ? 1u : 0u;   // the list is strictly increasing: equal span <=> no skip
        if (direct) first = (uint32_t)g0;
      }
      s_run[i] = first | ((uint32_t)min(len, VB_LEN_SAT) << 17) | (direct << 31);
    }
    __syncthreads();
  }
  for (int v = threadIdx.x; v < n_vox; v += VB_THREADS) {
    const int rb = meta_smem ? (int)s_seg[v] : s.seg_start[v];
    const int re = meta_smem ? (v + 1 < n_vox ? (int)s_seg[v + 1] : n_runs) : s.seg_start[v + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    for (int r = rb; r < re; r++) {
      int a, len;
      bool direct = false;
      if (meta_smem) {
        const uint32_t pk = s_run[r];
        a = (int)(pk & 0x1ffffu); len = (int)((pk >> 17) & (uint32_t)VB_LEN_SAT); direct = (pk >> 31) != 0u;
        if (len == VB_LEN_SAT) {                                 // a run longer than the packed field: the exact length from the table
          const int run = fin16[r];
          len = s.run_start[run + 1] - s.run_start[run];
        }
      } else {
        const int run = (int)sval_out[r];
        a = s.run_start[run]; len = s.run_start[run + 1] - a;
      }
      const bool use_list = gat != nullptr && !direct;
      const int e = a + len;
      int j = a;
      for (; j + 4 <= e; j += 4) {
        int i0 = j, i1 = j + 1, i2 = j + 2, i3 = j + 3;
        if (use_list) { i0 = __ldg(&gat[j]); i1 = __ldg(&gat[j + 1]); i2 = __ldg(&gat[j + 2]); i3 = __ldg(&gat[j + 3]); }
        const float4 p0 = __ldg(&src[i0]), p1 = __ldg(&src[i1]), p2 = __ldg(&src[i2]), p3 = __ldg(&src[i3]);
        sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
        sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w;
        sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w;
        sx += p3.x; sy += p3.y; sz += p3.z; si += p3.w;
      }
      if (j < e) {                                             // 1..3 points left: their loads go out together as well
        const int m = e - j;
        int i0 = j, i1 = j + 1, i2 = j + 2;
        if (use_list) { i0 = __ldg(&gat[j]); if (m > 1) i1 = __ldg(&gat[j + 1]); if (m > 2) i2 = __ldg(&gat[j + 2]); }
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 p0 = __ldg(&src[i0]);
        const float4 p1 = m > 1 ? __ldg(&src[i1]) : z4;
        const float4 p2 = m > 2 ? __ldg(&src[i2]) : z4;
        sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
        if (m > 1) { sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w; }
        if (m > 2) { sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w; }
      }
      cnt += len;
    }
    const float c = (float)cnt;
    s.out[v] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
}

// (5) centroids, fp32 accumulation in ascending input index: the runs of a voxel are adjacent in the sorted arrays and -
// the sort being stable - in ascending run id = ascending input index; inside a run the points are consecutive inputs.
// One thread per voxel (grid-stride).  A thread's loads are a chain of dependent round trips (index list -> point), so the
// points of a run are fetched four at a time (four index loads, then four point loads in flight) and added in order.
// Two other decompositions were measured on the 256-frame batch and dropped (profiles/r02_summary.md): staging the chunk's
// points in shared memory first (the barriers and the per-chunk binary search cost more than the loads they save) and the
// round-1 kernel on a re-expanded point list.  grid = (blocks, nseg)
constexpr int VC_CHUNK = 2048;     // host: blocks = ceil(max_n / VC_CHUNK) per cloud (voxels are dealt out grid-stride)
__global__ void __launch_bounds__(256)
k_vox_centroid(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.y];
  const int m = *s.out_n;
  const uint32_t* __restrict__ sval = (s.plan->npass & 1) ? s.val_b : s.val_a;
  const float4* __restrict__ src = s.src;
  const int* __restrict__ gat = s.gather;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < m; v += gridDim.x * blockDim.x) {
    const int rb = s.seg_start[v], re = s.seg_start[v + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    for (int r = rb; r < re; r++) {
      const int run = (int)sval[r];
      const int a = s.run_start[run], e = s.run_start[run + 1];
      int j = a;
      for (; j + 4 <= e; j += 4) {
        int i0 = j, i1 = j + 1, i2 = j + 2, i3 = j + 3;
        if (gat) { i0 = gat[j]; i1 = gat[j + 1]; i2 = gat[j + 2]; i3 = gat[j + 3]; }
        const float4 p0 = __ldg(&src[i0]), p1 = __ldg(&src[i1]), p2 = __ldg(&src[i2]), p3 = __ldg(&src[i3]);
        sx += p0.x; sy += p0.y; sz += p0.z; si += p0.w;
        sx += p1.x; sy += p1.y; sz += p1.z; si += p1.w;
        sx += p2.x; sy += p2.y; sz += p2.z; si += p2.w;
        sx += p3.x; sy += p3.y; sz += p3.z; si += p3.w;
      }
      for (; j < e; j++) { const float4 p = vox_point(s, j); sx += p.x; sy += p.y; sz += p.z; si += p.w; }
      cnt += e - a;
    }
    const float c = (float)cnt;
    s.out[v] = make_float4(sx / c, sy / c, sz / c, si / c);
  }
}

// (5, a few large clouds: the sliding-window map, submap classes) the same sums with one WARP per voxel.  A map voxel holds
// ~100 points in ~20 runs (one or two per key frame); a single thread would walk them through ~50 dependent round trips.
// Here the lanes fetch the points of up to 32 runs at a time in parallel into shared memory (flattened point list, the run
// of a point found by comparing against the shuffled run offsets) and lane 0 adds them in order.  grid = (blocks, nseg)
constexpr int VCW_PTS = 256;       // staged points per warp and pass
__global__ void __launch_bounds__(256)
k_vox_centroid_warp(VoxSeg* segs) {
  const VoxSeg s = segs[blockIdx.y];
  const int m = *s.out_n;
  const uint32_t* __restrict__ sval = (s.plan->npass & 1) ? s.val_b : s.val_a;
  __shared__ float4 s_pts[8][VCW_PTS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  for (int v = blockIdx.x * 8 + wid; v < m; v += gridDim.x * 8) {
    const int rb = s.seg_start[v], re = s.seg_start[v + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    for (int r0 = rb; r0 < re; r0 += 32) {                       // up to 32 runs per pass, in sorted (= input) order
      const int nr = min(32, re - r0);
      int a = 0, len = 0;
      if (lane < nr) { const int run = (int)sval[r0 + lane]; a = s.run_start[run]; len = s.run_start[run + 1] - a; }
      int incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
      const int off = incl - len;                                // exclusive offset of this lane's run
      const int T = __shfl_sync(FULL, incl, 31);
      for (int q0 = 0; q0 < T; q0 += VCW_PTS) {                  // batches of the flattened point list
        const int qn = min(VCW_PTS, T - q0);
        for (int q = q0 + lane; q < q0 + VCW_PTS; q += 32) {     // uniform trip count: the shuffles below need every lane
          int rr = 0;                                            // run of point q = number of runs whose offset is <= q, minus 1
#pragma unroll 1
          for (int r = 1; r < nr; r++) rr += (__shfl_sync(FULL, off, r) <= q) ? 1 : 0;
          const int ra = __shfl_sync(FULL, a, rr), ro = __shfl_sync(FULL, off, rr);
          if (q < q0 + qn) s_pts[wid][q - q0] = vox_point(s, ra + (q - ro));
        }
        __syncwarp();
        if (lane == 0)
          for (int k = 0; k < qn; k++) { const float4 p = s_pts[wid][k]; sx += p.x; sy += p.y; sz += p.z; si += p.w; }
        __syncwarp();
      }
      cnt += T;
    }
    if (lane == 0) { const float c = (float)cnt; s.out[v] = make_float4(sx / c, sy / c, sz / c, si / c); }
  }
}

}  // namespace lisreg
