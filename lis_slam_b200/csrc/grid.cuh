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
// float4 array {x, y, z, bits(original index)}, points of one cell ordered by original index;
// cell_start[] (ncells + 1, uint32) is the CSR offset table.  With x fastest, the cells
// (cx-s .. cx+s, cy, cz) of one row are ONE contiguous point range, so the 3x3x3 block around a
// query is 9 coalesced streaks.  The search visits that block first, then Chebyshev shells
// s = 2, 3, ... and stops as soon as the 5th best distance is below the lower bound of
// everything not yet visited (or the bound reaches the gate).
//
// The running best-5 is kept as five 64-bit keys (float bits of d^2 << 32 | position in the
// sorted array): d^2 >= 0 so the integer order of the key is the order of d^2, and a sorted insert is
// five min/max stages held entirely in registers.  Bit-equal distances are ordered by the ORIGINAL point
// index (knn_key_less), like the CPU oracle.
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

typedef unsigned long long knn_key;

LISREG_HD __forceinline__ int f2i(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int i; memcpy(&i, &f, 4); return i;
#endif
}
LISREG_HD __forceinline__ float i2f(int i) {
#ifdef __CUDA_ARCH__
  return __int_as_float(i);
#else
  float f; memcpy(&f, &i, 4); return f;
#endif
}
LISREG_HD __forceinline__ float knn_key_d(knn_key k) { return i2f((int)(unsigned)(k >> 32)); }
LISREG_HD __forceinline__ int knn_key_pos(knn_key k) { return (int)(unsigned)(k & 0xffffffffull); }

LISREG_HD __forceinline__ int cell_coord(float v, float o, float inv_h) {
  return (int)floorf((v - o) * inv_h);
}

// original (caller-side) index of the point stored at `pos` of the cell-sorted array
LISREG_HD __forceinline__ int knn_orig(const float4* __restrict__ pts, unsigned pos) { return f2i(LISREG_LDG(&pts[pos].w)); }

// Total order of the candidates of one query: (d^2, ORIGINAL index) - the order of the CPU oracle's kd-tree
// (oracle/orc_lm.cpp KdTree::insert), so that maps with exactly equidistant points (lattices, duplicated voxel
// centroids) give the same neighbour set and the same neighbour order on both sides.  Keys carry the POSITION in
// the sorted array (what the gathers need); the original index is only fetched when two squared distances are
// bit-equal, which real clouds almost never produce.
LISREG_HD __forceinline__ bool knn_key_less(const float4* __restrict__ pts, knn_key a, knn_key b) {
  const unsigned da = (unsigned)(a >> 32), db = (unsigned)(b >> 32);
  if (da != db) return da < db;
  const unsigned pa = (unsigned)a, pb = (unsigned)b;
  if (pa == 0xffffffffu || pb == 0xffffffffu) return pa < pb;      // sentinel: after every real candidate
  return knn_orig(pts, pa) < knn_orig(pts, pb);
}

// keeps the K smallest keys, ascending
template <int K>
LISREG_HD __forceinline__ void knn_insert(const float4* __restrict__ pts, knn_key (&b)[K], knn_key k) {
#pragma unroll
  for (int j = 0; j < K; j++) {
    bool lt = b[j] < k;
    if ((unsigned)(b[j] >> 32) == (unsigned)(k >> 32)) lt = knn_key_less(pts, b[j], k);
    const knn_key lo = lt ? b[j] : k;
    k = lt ? k : b[j];
    b[j] = lo;
  }
}

template <int K>
LISREG_HD __forceinline__ void knn_scan_range(const float4* __restrict__ pts, uint32_t b, uint32_t e,
                                               float qx, float qy, float qz, knn_key (&best)[K]) {
  for (uint32_t p = b; p < e; p++) {
    const float4 m = LISREG_LDG(&pts[p]);
    const float dx = qx - m.x, dy = qy - m.y, dz = qz - m.z;
    float d = dx * dx; d = d + dy * dy; d = d + dz * dz;   // FLANN L2 functor op order, no FMA
    const knn_key k = ((knn_key)(unsigned)f2i(d) << 32) | (knn_key)p;
    if ((unsigned)f2i(d) <= (unsigned)(best[K - 1] >> 32)) knn_insert<K>(pts, best, k);   // "<=": a tie is decided inside
  }
}

LISREG_HD __forceinline__ void knn_row_range(const GridDev& g, int x0, int x1, int y, int z, uint32_t& b, uint32_t& e) {
  b = 0u; e = 0u;
  if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) return;
  x0 = x0 > 0 ? x0 : 0; x1 = x1 < g.nx - 1 ? x1 : g.nx - 1;
  if (x0 > x1) return;
  const int rowbase = (z * g.ny + y) * g.nx;
  b = LISREG_LDG(&g.cell_start[rowbase + x0]);
  e = LISREG_LDG(&g.cell_start[rowbase + x1 + 1]);
}

// Outer shells t = 2, 3, ... (rare: sparse neighbourhoods).  Kept out of line so that the hot
// 3x3x3 loop stays small in the instruction cache.  best[KD] is the key that drives the stop test (K > KD + 1
// tracks extra runners-up without changing what is visited).  Returns a lower bound on the distance from the
// query to every map point NOT visited (FLT_MAX when the whole grid has been visited).
template <int K, int KD = K - 1>
LISREG_HD __noinline__ float knn_outer_shells(const GridDev& g, float qx, float qy, float qz, float gate,
                                             int cx, int cy, int cz, float minf, knn_key (&best)[K]) {
  // shells beyond the grid extent contain nothing
  int reach = cx > g.nx - 1 - cx ? cx : g.nx - 1 - cx;
  reach = reach > cy ? reach : cy; reach = reach > g.ny - 1 - cy ? reach : g.ny - 1 - cy;
  reach = reach > cz ? reach : cz; reach = reach > g.nz - 1 - cz ? reach : g.nz - 1 - cz;
  const float fs = ceilf(sqrtf(gate) * g.inv_h) + 1.f;
  int max_shell = fs < 1.0e6f ? (int)fs : 1000000;
  const bool whole_grid = max_shell >= reach;
  if (max_shell > reach) max_shell = reach;
  for (int s = 1; s <= max_shell; s++) {
    // everything not yet visited is farther than lb (margin covers the float rounding of cell assignment)
    const float lb = ((float)s + minf - 1e-3f) * g.h;
    const float lb2 = lb * lb;
    if (lb > 0.f && (knn_key_d(best[KD]) < lb2 || lb2 >= gate)) return lb;
    const int t = s + 1;   // visit shell t
    const int za = cz - t > 0 ? cz - t : 0, zb = cz + t < g.nz - 1 ? cz + t : g.nz - 1;
    const int ya = cy - t > 0 ? cy - t : 0, yb = cy + t < g.ny - 1 ? cy + t : g.ny - 1;
    for (int z = za; z <= zb; z++) {
      const bool zface = (z == cz - t) || (z == cz + t);
      for (int y = ya; y <= yb; y++) {
        const bool face = zface || (y == cy - t) || (y == cy + t);
        // a face row is one streak; an inner row contributes only its two end cells
        for (int part = 0; part < (face ? 1 : 2); part++) {
          uint32_t b, e;
          if (face) knn_row_range(g, cx - t, cx + t, y, z, b, e);
          else knn_row_range(g, part == 0 ? cx - t : cx + t, part == 0 ? cx - t : cx + t, y, z, b, e);
          knn_scan_range<K>(g.pts, b, e, qx, qy, qz, best);
        }
      }
    }
  }
  // shells 2 .. max_shell + 1 visited: either that was the whole grid, or the next shell starts beyond the gate radius
  if (whole_grid) return FLT_MAX;
  const float lbn = ((float)(max_shell + 1) + minf - 1e-3f) * g.h;
  return lbn > 0.f ? lbn : 0.f;
}

// The 3x3x3 block of the search.  best[] ascending; unused slots keep the sentinel (d^2 = gate - or +inf when
// INF_SENTINEL, which lets runners-up beyond the gate be tracked - position 0xffffffff).  Returns true when
// something outside the block could still beat best[KD], i.e. the outer shells must be visited
// (knn_outer_shells) for the result to be exact.
LISREG_HD __forceinline__ float knn_block_lb(const GridDev& g, float minf) { return (1.f + minf - 1e-3f) * g.h; }

template <int K, int KD = K - 1, bool INF_SENTINEL = false>
LISREG_HD __forceinline__ bool knn_grid_block(const GridDev& g, float qx, float qy, float qz, float gate, knn_key (&best)[K],
                                               int& cx, int& cy, int& cz, float& minf) {
  const knn_key sentinel = ((knn_key)(unsigned)(INF_SENTINEL ? 0x7f800000 : f2i(gate)) << 32) | 0xffffffffull;
#pragma unroll
  for (int j = 0; j < K; j++) best[j] = sentinel;
  cx = cy = cz = 0; minf = 0.f;
  if (g.n <= 0) return false;
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  cx = (int)floorf(fx); cy = (int)floorf(fy); cz = (int)floorf(fz);
  // distance from the query to the nearest face of its own cell (in cells)
  const float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
  minf = fminf(fminf(fminf(rx, 1.f - rx), fminf(ry, 1.f - ry)), fminf(rz, 1.f - rz));
  // 9 row streaks; ONE copy of the scan loop (instruction-cache friendly), the offsets of row r+1 are
  // fetched while row r is scanned
  uint32_t b, e, nb = 0u, ne = 0u;
  knn_row_range(g, cx - 1, cx + 1, cy - 1, cz - 1, b, e);
#pragma unroll 1
  for (int r = 0; r < 9; r++) {
    if (r < 8) knn_row_range(g, cx - 1, cx + 1, cy + ((r + 1) % 3) - 1, cz + ((r + 1) / 3) - 1, nb, ne);
    knn_scan_range<K>(g.pts, b, e, qx, qy, qz, best);
    b = nb; e = ne;
  }
  const float lb = knn_block_lb(g, minf);
  const float lb2 = lb * lb;
  return !(knn_key_d(best[KD]) < lb2 || lb2 >= gate);
}

// Exact K-NN restricted to squared distance < gate ("K neighbours inside the gate" <=> key_d(best[K-1]) < gate).
template <int K>
LISREG_HD __forceinline__ void knn_grid(const GridDev& g, float qx, float qy, float qz, float gate, knn_key (&best)[K]) {
  int cx, cy, cz; float minf;
  if (knn_grid_block<K>(g, qx, qy, qz, gate, best, cx, cy, cz, minf)) knn_outer_shells<K>(g, qx, qy, qz, gate, cx, cy, cz, minf, best);
}

// Exact KD+1 nearest neighbours plus the K - KD - 1 runners-up among everything visited (+inf sentinel, so the
// runners-up may lie beyond the gate).  Returns the lower bound on the distance to every point NOT visited.
template <int K, int KD>
LISREG_HD __forceinline__ float knn_grid_tracked(const GridDev& g, float qx, float qy, float qz, float gate, knn_key (&best)[K]) {
  int cx, cy, cz; float minf;
  if (knn_grid_block<K, KD, true>(g, qx, qy, qz, gate, best, cx, cy, cz, minf))
    return knn_outer_shells<K, KD>(g, qx, qy, qz, gate, cx, cy, cz, minf, best);
  return g.n > 0 ? knn_block_lb(g, minf) : FLT_MAX;
}

// Ball walk: exact KD+1 nearest neighbours (plus K-KD-1 runners-up among the visited points, +inf sentinel) for
// queries in sparse neighbourhoods.  Visits the Chebyshev shells t = 0, 1, 2, ... around the query's cell, but
// inside a shell only the rows - and inside a row only the cells - whose distance to the query is within the
// pruning radius Rp = sqrt(min(d2(best[KD]), gate)) + pad, which shrinks as the result improves.  Every point
// NOT visited is therefore farther than the returned radius (>= sqrt(d2 of the KD-th result) + pad when found
// inside the gate): the caller gets an exact result and a proof margin of `pad` metres around it.
// eps (in cells) absorbs the fp32 rounding of the cell assignment.
template <int K, int KD>
LISREG_HD __noinline__ float knn_ball_walk(const GridDev& g, float qx, float qy, float qz, float gate, float pad, knn_key (&best)[K]) {
  const knn_key sentinel = ((knn_key)0x7f800000u << 32) | 0xffffffffull;
#pragma unroll
  for (int j = 0; j < K; j++) best[j] = sentinel;
  if (g.n <= 0) return FLT_MAX;
  const float eps = 1e-3f;
  const float fx = (qx - g.ox) * g.inv_h, fy = (qy - g.oy) * g.inv_h, fz = (qz - g.oz) * g.inv_h;
  const int cx = (int)floorf(fx), cy = (int)floorf(fy), cz = (int)floorf(fz);
  const float rx = fx - (float)cx, ry = fy - (float)cy, rz = fz - (float)cz;
  const float minf = fminf(fminf(fminf(rx, 1.f - rx), fminf(ry, 1.f - ry)), fminf(rz, 1.f - rz));
  int reach = cx > g.nx - 1 - cx ? cx : g.nx - 1 - cx;
  reach = reach > cy ? reach : cy; reach = reach > g.ny - 1 - cy ? reach : g.ny - 1 - cy;
  reach = reach > cz ? reach : cz; reach = reach > g.nz - 1 - cz ? reach : g.nz - 1 - cz;
  const float sg = sqrtf(gate);
  float Rp = sg + pad;
  for (int t = 0; t <= reach; t++) {
    if (t >= 1 && ((float)(t - 1) + minf - eps) * g.h > Rp) break;   // the whole shell is out of reach
    const int za = cz - t > 0 ? cz - t : 0, zb = cz + t < g.nz - 1 ? cz + t : g.nz - 1;
    const int ya = cy - t > 0 ? cy - t : 0, yb = cy + t < g.ny - 1 ? cy + t : g.ny - 1;
    for (int z = za; z <= zb; z++) {
      const bool zface = (z == cz - t) || (z == cz + t);
      float dz = z < cz ? fz - (float)(z + 1) : (z > cz ? (float)z - fz : 0.f);
      dz = dz - eps > 0.f ? dz - eps : 0.f;
      for (int y = ya; y <= yb; y++) {
        const bool face = zface || (y == cy - t) || (y == cy + t);
        float dy = y < cy ? fy - (float)(y + 1) : (y > cy ? (float)y - fy : 0.f);
        dy = dy - eps > 0.f ? dy - eps : 0.f;
        const float dyz2 = (dy * dy + dz * dz) * g.h * g.h;
        const float rem = Rp * Rp - dyz2;
        if (rem < 0.f) continue;                                      // the whole row is out of reach
        const float rad = sqrtf(rem) * g.inv_h + eps;                 // reach along x, in cells
        const int xlo = (int)floorf(fx - rad), xhi = (int)floorf(fx + rad);
        bool scanned = false;
        if (face) {                                                   // a face row is one streak
          const int x0 = cx - t > xlo ? cx - t : xlo, x1 = cx + t < xhi ? cx + t : xhi;
          uint32_t b, e;
          knn_row_range(g, x0, x1, y, z, b, e);
          if (e > b) { knn_scan_range<K>(g.pts, b, e, qx, qy, qz, best); scanned = true; }
        } else {                                                      // an inner row contributes its two end cells
          for (int part = 0; part < 2; part++) {
            const int x = part == 0 ? cx - t : cx + t;
            if (x < xlo || x > xhi) continue;
            uint32_t b, e;
            knn_row_range(g, x, x, y, z, b, e);
            if (e > b) { knn_scan_range<K>(g.pts, b, e, qx, qy, qz, best); scanned = true; }
          }
        }
        if (scanned) {
          const float dk = knn_key_d(best[KD]);
          if (dk < gate) { const float r2 = sqrtf(dk) + pad; Rp = r2 < Rp ? r2 : Rp; }
        }
      }
    }
  }
  return Rp;
}

LISREG_HD __forceinline__ void knn5_grid(const GridDev& g, float qx, float qy, float qz, float gate, knn_key (&best)[5]) {
  knn_grid<5>(g, qx, qy, qz, gate, best);
}

}  // namespace lisreg
