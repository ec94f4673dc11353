// Uniform-grid spatial index + exact gated 5-NN.
//
// Replaces the per-frame FLANN kd-tree of the reference
// (kdtree*FromMap->setInputCloud / nearestKSearch, odomEstimationNode.cpp:602-603, :650, :766).
// The reference only ever USES a neighbour set when its 5th squared distance is below a
// gate (1.0 for variant A :657/:776, 2.0 for B/C subMapOptmizationNode.cpp:1610/:1760), so an
// exact search restricted to the gate radius returns identical results wherever they
// matter; points whose 5th neighbour lies beyond the gate are rejected on both paths.
//
// Layout in HBM: map points are counting-sorted by linear cell id (x fastest) into a packed
// float4 array {x, y, z, bits(original index)}; cell_start[] (ncells + 1, uint32) is the CSR
// offset table.  With x fastest, the cells (cx-s .. cx+s, cy, cz) of one row are one
// contiguous point range, so a shell of the search is a handful of coalesced streaks.
// The search visits Chebyshev shells s = 0, 1, ... and stops as soon as the 5th best
// distance is below the lower bound of everything not yet visited (or the gate).
#pragma once
#include <cstdint>
#include <cfloat>
#include <cstring>
#include <cuda_runtime.h>
#include <cmath>
#ifndef LISREG_HD
#define LISREG_HD __host__ __device__
#endif
#ifdef __CUDA_ARCH__
#define LISREG_LDG(p) __ldg(p)
#else
#define LISREG_LDG(p) (*(p))
#endif

namespace lisreg {

struct GridDev {
  float ox, oy, oz;      // origin (min corner)
  float h, inv_h;        // cell size
  int nx, ny, nz;
  int n;                 // number of points
  int ncells;
  const uint32_t* cell_start;  // ncells + 1
  const float4* pts;           // sorted, w = original index bits
};

LISREG_HD __forceinline__ int cell_coord(float v, float o, float inv_h) {
  return (int)floorf((v - o) * inv_h);
}

LISREG_HD __forceinline__ int f2i(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int i; memcpy(&i, &f, 4); return i;
#endif
}

// lexicographic (distance, original index) order makes the result independent of the
// (atomic, hence unordered) placement of points inside a cell.
LISREG_HD __forceinline__ bool knn_less(float d, int i, float d2, int i2) {
  return d < d2 || (d == d2 && i < i2);
}

// Sorted insert into the running best-5 (ascending): replace the worst entry, then bubble it
// down with select-based compare-exchanges (static indices only, so the arrays stay in registers).
LISREG_HD __forceinline__ void knn5_insert(float (&bd)[5], int (&bi)[5], int (&bp)[5], float d, int idx, int pos) {
  if (!knn_less(d, idx, bd[4], bi[4])) return;
  bd[4] = d; bi[4] = idx; bp[4] = pos;
#pragma unroll
  for (int j = 4; j > 0; j--) {
    const bool sw = knn_less(bd[j], bi[j], bd[j - 1], bi[j - 1]);
    const float d0 = bd[j - 1], d1 = bd[j];
    const int i0 = bi[j - 1], i1 = bi[j], p0 = bp[j - 1], p1 = bp[j];
    bd[j - 1] = sw ? d1 : d0; bd[j] = sw ? d0 : d1;
    bi[j - 1] = sw ? i1 : i0; bi[j] = sw ? i0 : i1;
    bp[j - 1] = sw ? p1 : p0; bp[j] = sw ? p0 : p1;
  }
}

LISREG_HD __forceinline__ void knn5_scan_range(const GridDev& g, uint32_t b, uint32_t e, float qx, float qy, float qz,
                                                float (&bd)[5], int (&bi)[5], int (&bp)[5]) {
  for (uint32_t p = b; p < e; p++) {
    float4 m = LISREG_LDG(&g.pts[p]);
    float dx = qx - m.x, dy = qy - m.y, dz = qz - m.z;
    float d = dx * dx; d = d + dy * dy; d = d + dz * dz;   // FLANN L2 functor op order, no FMA
    if (d <= bd[4]) knn5_insert(bd, bi, bp, d, f2i(m.w), (int)p);
  }
}

// Exact 5-NN restricted to squared distance < gate.  On return bd[] ascending; slots that
// found no neighbour inside the gate keep bd = gate, bi = INT_MAX, bp = -1.
LISREG_HD __forceinline__ void knn5_grid(const GridDev& g, float qx, float qy, float qz, float gate,
                                          float (&bd)[5], int (&bi)[5], int (&bp)[5]) {
#pragma unroll
  for (int j = 0; j < 5; j++) { bd[j] = gate; bi[j] = 0x7fffffff; bp[j] = -1; }
  if (g.n <= 0) return;
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
  // distance from the query to the nearest face of its own cell (in cells), conservative
  float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
  float minf = fminf(fminf(fminf(rx, 1.f - rx), fminf(ry, 1.f - ry)), fminf(rz, 1.f - rz));
  const int max_shell = (int)ceilf(sqrtf(gate) * g.inv_h) + 1;
  for (int s = 0; s <= max_shell; s++) {
    const int z0 = cz - s > 0 ? cz - s : 0, z1 = cz + s < g.nz - 1 ? cz + s : g.nz - 1;
    const int y0 = cy - s > 0 ? cy - s : 0, y1 = cy + s < g.ny - 1 ? cy + s : g.ny - 1;
    const int xa = cx - s, xb = cx + s;
    if (!(xb < 0 || xa >= g.nx)) {
      for (int z = z0; z <= z1; z++) {
        const bool zface = (z == cz - s) || (z == cz + s);
        for (int y = y0; y <= y1; y++) {
          const bool face = zface || (y == cy - s) || (y == cy + s);
          const int rowbase = (z * g.ny + y) * g.nx;
          if (face) {
            const int x0 = xa > 0 ? xa : 0, x1 = xb < g.nx - 1 ? xb : g.nx - 1;
            uint32_t b = LISREG_LDG(&g.cell_start[rowbase + x0]);
            uint32_t e = LISREG_LDG(&g.cell_start[rowbase + x1 + 1]);
            knn5_scan_range(g, b, e, qx, qy, qz, bd, bi, bp);
          } else {
            if (xa >= 0 && xa < g.nx) {
              uint32_t b = LISREG_LDG(&g.cell_start[rowbase + xa]);
              uint32_t e = LISREG_LDG(&g.cell_start[rowbase + xa + 1]);
              knn5_scan_range(g, b, e, qx, qy, qz, bd, bi, bp);
            }
            if (xb >= 0 && xb < g.nx && xb != xa) {
              uint32_t b = LISREG_LDG(&g.cell_start[rowbase + xb]);
              uint32_t e = LISREG_LDG(&g.cell_start[rowbase + xb + 1]);
              knn5_scan_range(g, b, e, qx, qy, qz, bd, bi, bp);
            }
          }
        }
      }
    }
    // everything not yet visited is farther than lb (with a safety margin for the float
    // rounding of cell assignment)
    float lb = ((float)s + minf - 1e-3f) * g.h;
    if (lb > 0.f) {
      float lb2 = lb * lb;
      if (bd[4] < lb2 || lb2 >= gate) break;
    }
  }
}

}  // namespace lisreg
