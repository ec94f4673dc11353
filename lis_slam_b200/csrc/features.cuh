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
// segment the warp bitonic-sorts (curvature, index) keys in shared memory and lane 0 runs the two
// greedy passes (the non-maximum suppression is order-dependent, hence sequential).
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
};

// Per-frame views into the batch work buffers (all device pointers).
struct FeatFrame {
  const float4* pts; const uint16_t* ring; int n;       // raw sweep
  int* owner;            // n_scan*horizon  : input index owning each range-image cell (INT_MAX = empty)
  float4* ext_pts;       // extracted cloud (capacity n_scan*horizon)
  int* ext_src;          // index of the input point in each extracted slot
  int* col; float* range; float* curv; int* picked; int* label;
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
};

__device__ __forceinline__ float atan2f_cr(float y, float x) { return (float)atan2((double)y, (double)x); }

__global__ void k_feat_clear(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int cells = prm.n_scan * prm.horizon;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
    f.owner[i] = 0x7fffffff;
    f.picked[i] = 0; f.label[i] = 0; f.curv[i] = 0.f; f.col[i] = 0; f.range[i] = 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x < 4) f.counts[threadIdx.x] = 0;
}

__device__ __forceinline__ bool feat_project(const FeatParamsDev& prm, float4 p, int ring, float& r, int& cell) {
  r = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
  if (r < prm.min_range || r > prm.max_range) return false;
  if (ring < 0 || ring >= prm.n_scan) return false;
  if (ring % prm.downsample != 0) return false;
  const float ang_res_x = (float)(360.0 / (double)(float)prm.horizon);
  const float horizonAngle = (float)((double)(atan2f_cr(p.x, p.y) * 180) / 3.14159265358979323846);
  int col = (int)(-round(((double)horizonAngle - 90.0) / (double)ang_res_x) + (double)(prm.horizon / 2));
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
    if (feat_project(prm, __ldg(&f.pts[i]), (int)f.ring[i], r, cell)) atomicMin(&f.owner[cell], i);
  }
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
      const float4 p = __ldg(&f.pts[own]);
      f.ext_pts[pos] = p;
      f.ext_src[pos] = own;
      f.col[pos] = j;
      f.range[pos] = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
    }
  }
}

// F3 + F4.  grid = (blocks, F)
__global__ void k_feat_curvature(FeatFrame* frames) {
  const FeatFrame f = frames[blockIdx.y];
  const int M = *f.M;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    if (i >= 5 && i < M - 5) {
      const float* r = f.range;
      const float d = r[i - 5] + r[i - 4] + r[i - 3] + r[i - 2] + r[i - 1] - r[i] * 10 +
                      r[i + 1] + r[i + 2] + r[i + 3] + r[i + 4] + r[i + 5];   // exact op order (:549-553)
      f.curv[i] = d * d;
    }
  }
}
__global__ void k_feat_occlusion(FeatFrame* frames) {
  const FeatFrame f = frames[blockIdx.y];
  const int M = *f.M;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    if (i >= 5 && i < M - 6) {
      const float depth1 = f.range[i], depth2 = f.range[i + 1];
      const int columnDiff = abs(f.col[i + 1] - f.col[i]);
      if (columnDiff < 10) {
        if ((double)(depth1 - depth2) > 0.3) { for (int k = -5; k <= 0; k++) f.picked[i + k] = 1; }
        else if ((double)(depth2 - depth1) > 0.3) { for (int k = 1; k <= 6; k++) f.picked[i + k] = 1; }
      }
      const float diff1 = fabsf(f.range[i - 1] - f.range[i]);
      const float diff2 = fabsf(f.range[i + 1] - f.range[i]);
      if ((double)diff1 > 0.02 * (double)f.range[i] && (double)diff2 > 0.02 * (double)f.range[i]) f.picked[i] = 1;
    }
  }
}

// ---- F5 ----
constexpr int FEAT_HI_MAX = 128;        // edge candidates (curvature > edgeThreshold) of one segment (power of two)
constexpr int FEAT_LO_MAX = 512;        // flat candidates (curvature < surfThreshold) of one segment (power of two)
constexpr int FEAT_SEG_MAX = FEAT_HI_MAX + FEAT_LO_MAX;   // key slots per warp in shared memory
constexpr int FEAT_WARPS = 4;           // rings per block
constexpr int FEAT_RING_MAX = 2048 + 16; // staged ring window (horizon <= 2048, +-6 apron)

// Neighbour suppression of a pick (:648-661) on the ring window staged in shared memory.
// lo = global index of window slot 0; indices outside [0, M) or outside the window end the walk
// (the window covers [first-6, last+6] of the ring, the only indices a pick of this ring can reach).
__device__ __forceinline__ void feat_mark_neighbours(unsigned char* spick, const unsigned short* scol, int lo, int wlen, int ind, int M) {
  for (int l = 1; l <= 5; l++) {
    const int a = ind + l, b = ind + l - 1;
    if (a < 0 || a >= M || b < 0 || b >= M || a - lo >= wlen || b - lo < 0) break;
    if (abs((int)scol[a - lo] - (int)scol[b - lo]) > 10) break;
    spick[a - lo] = 1;
  }
  for (int l = -1; l >= -5; l--) {
    const int a = ind + l, b = ind + l + 1;
    if (a < 0 || a >= M || b < 0 || b >= M || a - lo < 0 || b - lo >= wlen) break;
    if (abs((int)scol[a - lo] - (int)scol[b - lo]) > 10) break;
    spick[a - lo] = 1;
  }
}

// one warp per (frame, ring).  grid = (ceil(n_scan / FEAT_WARPS), F), block = 32 * FEAT_WARPS.
// The ring's picked flags and column indices are staged in shared memory so that the sequential greedy
// passes of lane 0 never wait on global memory; curvature comes from the sorted key itself.
__global__ void __launch_bounds__(32 * FEAT_WARPS)
k_feat_segments(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.y];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ring = blockIdx.x * FEAT_WARPS + wid;
  __shared__ unsigned long long s_key[FEAT_WARPS][FEAT_SEG_MAX];
  __shared__ unsigned char s_pick[FEAT_WARPS][FEAT_RING_MAX];
  __shared__ unsigned short s_col[FEAT_WARPS][FEAT_RING_MAX];
  if (ring >= prm.n_scan) return;
  const int M = *f.M;
  const int start = f.ring_start[ring], end = f.ring_end[ring];
  unsigned long long* key = s_key[wid];
  unsigned char* spick = s_pick[wid];
  unsigned short* scol = s_col[wid];
  // ring = extracted indices [first, last] with first = start - 4, last = end + 5; window adds a +-6 apron
  const int lo = max(start - 4 - 6, 0);
  const int hi = min(end + 5 + 6, M - 1);
  const int wlen = max(hi - lo + 1, 0);
  for (int t = lane; t < wlen; t += 32) { spick[t] = (unsigned char)f.picked[lo + t]; scol[t] = (unsigned short)f.col[lo + t]; }
  __syncwarp();
  for (int j = 0; j < 6; j++) {
    const int seg = ring * 6 + j;
    const int sp = (start * (6 - j) + end * j) / 6;
    const int ep = (start * (5 - j) + end * (j + 1)) / 6 - 1;
    if (lane == 0) { f.seg_sp[seg] = sp; f.seg_ep[seg] = ep; f.seg_valid[seg] = (sp < ep) ? 1 : 0; f.seg_ncorner[seg] = 0; f.seg_nflat[seg] = 0; }
    if (sp >= ep) continue;
    const int len = ep - sp;            // sorted range [sp, ep); element ep stays in place (Q3)
    // Only elements that can ever be picked need ordering: curvature > edge_thr (pass 1, descending) and
    // curvature < surf_thr (pass 2, ascending); the band in between is never visited with effect.  Both
    // candidate sets are compacted (warp ballot) into the two halves of the key buffer and bitonic-sorted
    // separately: keys = curvature bits << 32 | index (curvature >= 0 => integer order == (value, index)).
    unsigned long long* khi = key;                   // [0, FEAT_HI_MAX)
    unsigned long long* klo = key + FEAT_HI_MAX;     // [FEAT_HI_MAX, FEAT_SEG_MAX)
    int nhi = 0, nlo = 0;
    for (int base = 0; base < len; base += 32) {
      const int t = base + lane;
      float cv = 0.f; bool ishi = false, islo = false;
      if (t < len) { cv = f.curv[sp + t]; ishi = cv > prm.edge_thr; islo = cv < prm.surf_thr; }
      const unsigned mh = __ballot_sync(0xffffffffu, ishi), ml = __ballot_sync(0xffffffffu, islo);
      const unsigned long long kk = ((unsigned long long)__float_as_uint(cv) << 32) | (unsigned)(sp + t);
      if (ishi) { const int o = nhi + __popc(mh & ((1u << lane) - 1u)); if (o < FEAT_HI_MAX) khi[o] = kk; }
      if (islo) { const int o = nlo + __popc(ml & ((1u << lane) - 1u)); if (o < FEAT_LO_MAX) klo[o] = kk; }
      nhi += __popc(mh); nlo += __popc(ml);
    }
    // mode 0: split candidate lists (common); mode 1: too many edge candidates -> sort the whole segment in
    // shared memory (both passes then walk the same array and stop at their threshold); mode 2: oversize
    // segment (never for <= 2048 columns / 6) -> slow insertion sort by lane 0 in global scratch.
    int npad_full = 1; while (npad_full < len) npad_full <<= 1;
    const int mode = (nhi <= FEAT_HI_MAX && nlo <= FEAT_LO_MAX) ? 0 : (npad_full <= FEAT_SEG_MAX ? 1 : 2);
    if (mode == 1) {
      __syncwarp();
      for (int t = lane; t < npad_full; t += 32) {
        const int i = sp + t;
        key[t] = t < len ? (((unsigned long long)__float_as_uint(f.curv[i]) << 32) | (unsigned)i) : ~0ull;
      }
    }
    if (mode <= 1) {
      for (int pass = 0; pass < (mode == 0 ? 2 : 1); pass++) {
        unsigned long long* kb = mode == 1 ? key : (pass == 0 ? khi : klo);
        const int cnt = mode == 1 ? len : (pass == 0 ? nhi : nlo);
        int npad = 1; while (npad < cnt) npad <<= 1;
        if (mode == 0) for (int t = cnt + lane; t < npad; t += 32) kb[t] = ~0ull;
        __syncwarp();
        for (int k = 2; k <= npad; k <<= 1)
          for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int t = lane; t < npad; t += 32) {
              const int ixj = t ^ jj;
              if (ixj > t) {
                const unsigned long long a = kb[t], bq = kb[ixj];
                const bool up = (t & k) == 0;
                if ((a > bq) == up) { kb[t] = bq; kb[ixj] = a; }
              }
            }
            __syncwarp();
          }
      }
    } else {
      if (lane == 0) {
        int* o = f.owner + ring * prm.horizon;      // owner[] is free after compaction
        for (int t = 0; t < len; t++) o[t] = sp + t;
        for (int a = 1; a < len; a++) {
          const int v = o[a]; const float cv = f.curv[v]; int bb = a;
          while (bb > 0 && (f.curv[o[bb - 1]] > cv || (f.curv[o[bb - 1]] == cv && o[bb - 1] > v))) { o[bb] = o[bb - 1]; bb--; }
          o[bb] = v;
        }
      }
      __syncwarp();
    }
    // ---- greedy passes by lane 0 (order-dependent non-maximum suppression) ----
    if (lane == 0) {
      const int* o = f.owner + ring * prm.horizon;
      const unsigned long long* l1 = mode == 0 ? khi : key;   // pass-1 list (ascending; walked from the top)
      const unsigned long long* l2 = mode == 0 ? klo : key;   // pass-2 list (ascending; walked from the bottom)
      const int n1 = mode == 0 ? nhi : len, n2 = mode == 0 ? nlo : len;
      const float cv_ep = f.curv[ep];
      int largestPickedNum = 0, nc = 0;
      // pass 1: k = ep first, then the sorted candidates from the largest curvature down
      for (int k = n1; k >= 0; k--) {
        int ind; float cv;
        if (k == n1) { ind = ep; cv = cv_ep; }
        else {
          if (mode <= 1) { const unsigned long long kk = l1[k]; ind = (int)(unsigned)(kk & 0xffffffffull); cv = __uint_as_float((unsigned)(kk >> 32)); }
          else { ind = o[k]; cv = f.curv[ind]; }
          if (!(cv > prm.edge_thr)) break;   // ascending order: nothing below can qualify
        }
        if (spick[ind - lo] == 0 && cv > prm.edge_thr) {
          largestPickedNum++;
          if (largestPickedNum <= 20) {
            f.label[ind] = 1;
            f.seg_corner[seg * 20 + nc] = ind; nc++;
          } else break;
          spick[ind - lo] = 1;
          feat_mark_neighbours(spick, scol, lo, wlen, ind, M);
        }
      }
      f.seg_ncorner[seg] = nc;
      largestPickedNum = 0; int nf = 0;
      // pass 2: the sorted candidates from the smallest curvature up, then k = ep
      for (int k = 0; k <= n2; k++) {
        int ind; float cv;
        if (k == n2) { ind = ep; cv = cv_ep; }
        else {
          if (mode <= 1) { const unsigned long long kk = l2[k]; ind = (int)(unsigned)(kk & 0xffffffffull); cv = __uint_as_float((unsigned)(kk >> 32)); }
          else { ind = o[k]; cv = f.curv[ind]; }
          if (!(cv < prm.surf_thr)) { k = n2 - 1; continue; }   // nothing above can qualify: jump to the unsorted element ep
        }
        if (spick[ind - lo] == 0 && cv < prm.surf_thr) {
          largestPickedNum++;
          f.label[ind] = -1;
          spick[ind - lo] = 1;
          if (largestPickedNum <= 10) { f.seg_flat[seg * 10 + nf] = ind; nf++; }
          feat_mark_neighbours(spick, scol, lo, wlen, ind, M);
        }
      }
      f.seg_nflat[seg] = nf;
    }
    __syncwarp();
  }
}

// compaction of the per-segment lists into the reference push order; one block per frame
__global__ void k_feat_gather(FeatFrame* frames, FeatParamsDev prm) {
  const FeatFrame f = frames[blockIdx.x];
  const int nseg = prm.n_scan * 6;
  __shared__ int s_c[1024], s_s[1024], s_f[1024], s_u[1024];   // exclusive offsets per segment (nseg <= 1024)
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    const int nc = f.seg_valid[s] ? f.seg_ncorner[s] : 0;
    s_c[s] = nc; s_s[s] = nc < 4 ? nc : 4;
    s_f[s] = f.seg_valid[s] ? f.seg_nflat[s] : 0;
  }
  __syncthreads();
  // surf count per segment: label <= 0 over [sp, ep]
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    int c = 0;
    if (f.seg_valid[s]) for (int k = f.seg_sp[s]; k <= f.seg_ep[s]; k++) c += f.label[k] <= 0;
    s_u[s] = c;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    int* a = threadIdx.x == 0 ? s_c : threadIdx.x == 1 ? s_s : threadIdx.x == 2 ? s_f : s_u;
    int acc = 0;
    for (int s = 0; s < nseg; s++) { int t = a[s]; a[s] = acc; acc += t; }
    f.counts[threadIdx.x] = acc;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    if (!f.seg_valid[s]) continue;
    const int nc = f.seg_ncorner[s], nf = f.seg_nflat[s];
    for (int i = 0; i < nc; i++) f.corner_idx[s_c[s] + i] = f.seg_corner[s * 20 + i];
    for (int i = 0; i < (nc < 4 ? nc : 4); i++) f.sharp_idx[s_s[s] + i] = f.seg_corner[s * 20 + i];
    for (int i = 0; i < nf; i++) f.flat_idx[s_f[s] + i] = f.seg_flat[s * 10 + i];
    int o = s_u[s];
    for (int k = f.seg_sp[s]; k <= f.seg_ep[s]; k++) if (f.label[k] <= 0) f.surf_idx[o++] = k;
  }
}

}  // namespace lisreg
