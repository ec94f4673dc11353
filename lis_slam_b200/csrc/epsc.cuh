// EPSC ring-sector loop-closure descriptors and their shift-invariant scoring on the device.
//
// Reference: EPSCGeneration::calculateEPSC epscGeneration.cpp:478-520, calculateSEPSC :522-562,
// calculateFEPSC :591-607 (F15); calculateDistance :633-660 (F16); constants epscGeneration.h:9-43.
// Quirk Q4 reproduced: unsigned-char counters wrap at 256 and the quotient is narrowed mod 256.
//
// k_epsc_describe: one block per submap, 32-bit shared-memory histograms (count mod 256 == the
// reference's wrapping u8 counter), then the integer quotient and the FEPSC blend.
// k_epsc_score: byte SAD of 20 x 80 descriptors for the 20 column shifts i in [-10, 10).  One thread per
// (history j, query q) pair; for each of the 20 rings d1 = history row (20 words) and d2 = query row
// extended by the wrap-around (25 words) both live in REGISTERS; a shift s = i + 10 = 4a + b is one funnel shift of two adjacent d2 words and one
// __vsadu4 (4 byte-SADs per instruction), so the inner loop has no memory traffic at all: the kernel is
// bound by the integer ALU (SHF + VABSDIFF4 issue), not by HBM (all descriptors = 8 MB, L2 resident).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "tma.cuh"

namespace lisreg {

constexpr int EPSC_RINGS = 20, EPSC_SECTORS = 80, EPSC_SIZE = 1600;

struct EpscCloud {
  const float4* corner; const float4* surf; const float4* sem; const uint16_t* sem_label;
  int nc, ns, nsem;
};

__device__ __forceinline__ bool epsc_bin(float x, float y, int& bin) {
  const double distance = (double)sqrtf(x * x + y * y);
  if (distance >= 60.0 || distance < 3.0) return false;
  const double ring_step = (60.0 - 3.0) / 20;
  const double sector_step = 2 * 3.14159265358979323846 / 80;
  const int ring_id = (int)floor((distance - 3.0) / ring_step);
  const double angle = 3.14159265358979323846 + (double)(float)atan2((double)y, (double)x);
  const int sector_id = (int)floor(angle / sector_step);
  if (ring_id >= EPSC_RINGS || ring_id < 0) return false;
  if (sector_id >= EPSC_SECTORS || sector_id < 0) return false;
  bin = ring_id * EPSC_SECTORS + sector_id;
  return true;
}

// grid = nsubmaps, block = 256.  out: [n][3][1600] = epsc, sepsc, fepsc.
// xf (nullable): one row-major 4x4 transform per BLOCK (stride xf_stride floats) applied to every point before
// binning - the fused pcl::transformPointCloud of loopDetection (epscGeneration.cpp:764-770), fp32, left to right;
// block b then describes cloud b * cloud_stride (cloud_stride 0: every block re-describes cloud 0 under its own
// transform, one block per loop candidate).
__global__ void k_epsc_describe(const EpscCloud* __restrict__ clouds, const uint8_t* __restrict__ using_map, uint8_t* __restrict__ out,
                                const float* __restrict__ xf, int xf_stride, int cloud_stride) {
  const EpscCloud c = clouds[(size_t)blockIdx.x * cloud_stride];
  __shared__ unsigned esc[EPSC_SIZE], psc[EPSC_SIZE];
  __shared__ uint8_t s_epsc[EPSC_SIZE];
  __shared__ uint8_t s_lut[256];
  __shared__ float sT[8];
  if (threadIdx.x < 8) sT[threadIdx.x] = xf ? xf[(size_t)blockIdx.x * xf_stride + threadIdx.x] : (threadIdx.x == 0 || threadIdx.x == 5 ? 1.f : 0.f);
  s_lut[threadIdx.x & 255] = using_map[threadIdx.x & 255];
  for (int i = threadIdx.x; i < EPSC_SIZE; i += blockDim.x) { esc[i] = 0u; psc[i] = 0u; }
  __syncthreads();
  const bool moved = xf != nullptr;
  auto bin_of = [&](float4 p, int& bin) -> bool {
    float x = p.x, y = p.y;
    if (moved) { x = ((sT[0] * p.x + sT[1] * p.y) + sT[2] * p.z) + sT[3]; y = ((sT[4] * p.x + sT[5] * p.y) + sT[6] * p.z) + sT[7]; }
    return epsc_bin(x, y, bin);
  };
  int bin;
  for (int i = threadIdx.x; i < c.nc; i += blockDim.x) { if (bin_of(__ldg(&c.corner[i]), bin)) atomicAdd(&esc[bin], 1u); }
  for (int i = threadIdx.x; i < c.ns; i += blockDim.x) { if (bin_of(__ldg(&c.surf[i]), bin)) atomicAdd(&psc[bin], 1u); }
  __syncthreads();
  uint8_t* o = out + (size_t)blockIdx.x * 3 * EPSC_SIZE;
  for (int i = threadIdx.x; i < EPSC_SIZE; i += blockDim.x) {
    const int p8 = psc[i] & 255u, e8 = esc[i] & 255u;          // unsigned char counters wrap (Q4)
    const uint8_t v = (uint8_t)((100 * p8 / (1 + e8)) & 255);   // int quotient narrowed mod 256
    s_epsc[i] = v; o[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < EPSC_SIZE; i += blockDim.x) { esc[i] = 0u; psc[i] = 0u; }
  __syncthreads();
  for (int i = threadIdx.x; i < c.nsem; i += blockDim.x) {
    if (!bin_of(__ldg(&c.sem[i]), bin)) continue;
    const unsigned l = c.sem_label[i];
    const int cls = l < 256u ? s_lut[l] : 0;
    if (cls == 40 || cls == 50) atomicAdd(&psc[bin], 1u);
    else if (cls == 81) atomicAdd(&esc[bin], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < EPSC_SIZE; i += blockDim.x) {
    const int p8 = psc[i] & 255u, e8 = esc[i] & 255u;
    const uint8_t sv = (uint8_t)((100 * p8 / (1 + e8)) & 255);
    o[EPSC_SIZE + i] = sv;
    o[2 * EPSC_SIZE + i] = (uint8_t)((double)sv * 0.4 + (double)s_epsc[i] * 0.6);   // truncation (:603)
  }
}

// Pair scoring.  Work item = (query q, history j < q).  Pairs are enumerated over a [Q_TILE x J_TILE] tile
// per block; thread t handles query (t / J_TILE), history (t % J_TILE) of the tile.
constexpr int EPSC_QT = 8, EPSC_JT = 16, EPSC_THREADS = EPSC_QT * EPSC_JT;
// words per descriptor row in shared memory: 400 payload words + 4 pad = 1616 B, a multiple of 16 B so that every row is
// a legal bulk-copy destination (16 threads with different rows and the same word then touch 8 banks: 2-way conflicts on
// 40 LDS per ring against 600 ALU instructions - negligible)
constexpr int EPSC_STRIDE_W = 404;
constexpr int EPSC_CLUSTER = 2;      // CTAs per cluster along the history axis: they share the tile's query rows (multicast)

// Query rows are q = q_begin + r * q_stride for local rows r in [0, n_rows): the whole matrix is (0, 1, N); a rank
// of a multi-GPU run owns the cyclic rows (rank, world, ...) (triangular load => cyclic assignment).
//
// Staging: the 8 query rows + 16 history rows of a tile (24 x 1600 B) are copied by the bulk-copy engine
// (cp.async.bulk, one 1600-byte transaction per descriptor row, completion counted on an mbarrier) - no registers, no
// LSU instructions.  The EPSC_CLUSTER CTAs of a cluster sit next to each other on the history axis and need the SAME
// query rows: the cluster leader fetches them once and multicasts them into every CTA's shared memory.
//
constexpr int EPSC_TOPK_SLOTS = 8;
constexpr unsigned EPSC_SAD_GATE = 102000u;     // 1 - SAD / (80 * 20 * 255) > DISTANCE_THRESHOLD 0.75
__global__ void __launch_bounds__(EPSC_THREADS)
k_epsc_score(const uint8_t* __restrict__ desc, int N, int q_begin, int q_stride, int n_rows, unsigned long long* __restrict__ row_top) {
  const int r0 = blockIdx.y * EPSC_QT, j0 = blockIdx.x * EPSC_JT;
  const int r_last = min(r0 + EPSC_QT, n_rows) - 1;
  // the cluster's first history tile decides for the whole cluster (j0 ascends inside it): on / above the diagonal
  // (needs j < q) nobody has work and everybody leaves before any barrier exists
  const unsigned crank = tma::cluster_ctarank(), csize = tma::cluster_nctarank();
  const int j0_first = ((int)blockIdx.x - (int)crank) * EPSC_JT;
  if (r_last < r0 || j0_first >= q_begin + r_last * q_stride) return;
  __shared__ __align__(16) unsigned s_q[EPSC_QT * EPSC_STRIDE_W];
  __shared__ __align__(16) unsigned s_j[EPSC_JT * EPSC_STRIDE_W];
  __shared__ __align__(8) unsigned long long s_bar;
  const int nq = r_last - r0 + 1;                                   // valid query rows of the tile (same for the cluster)
  const int nj = max(0, min(EPSC_JT, N - j0));                      // valid history rows of this CTA
  if (threadIdx.x == 0) { tma::mbar_init(&s_bar, 1); tma::fence_barrier_init(); }
  if (csize > 1) tma::cluster_sync(); else __syncthreads();        // every CTA's barrier exists before the leader multicasts into it
  if (threadIdx.x == 0) {
    tma::mbar_arrive_expect_tx(&s_bar, (unsigned)(nq + nj) * EPSC_SIZE);
    for (int d = 0; d < nj; d++) tma::bulk_g2s(&s_j[d * EPSC_STRIDE_W], desc + (size_t)(j0 + d) * EPSC_SIZE, EPSC_SIZE, &s_bar);
    if (csize == 1)
      for (int d = 0; d < nq; d++) tma::bulk_g2s(&s_q[d * EPSC_STRIDE_W], desc + (size_t)(q_begin + (r0 + d) * q_stride) * EPSC_SIZE, EPSC_SIZE, &s_bar);
    else if (crank == 0)
      for (int d = 0; d < nq; d++)
        tma::bulk_g2s_multicast(&s_q[d * EPSC_STRIDE_W], desc + (size_t)(q_begin + (r0 + d) * q_stride) * EPSC_SIZE, EPSC_SIZE, &s_bar,
                                (unsigned short)((1u << csize) - 1u));
  }
  tma::mbar_wait(&s_bar, 0);
  if (csize > 1) tma::cluster_sync();                               // nobody exits while copies into its shared memory may be in flight
  const int ql = threadIdx.x / EPSC_JT, jl = threadIdx.x % EPSC_JT;
  const int r = r0 + ql, q = q_begin + r * q_stride, j = j0 + jl;
  if (r >= n_rows || j >= q) return;
  const unsigned* dq = s_q + ql * EPSC_STRIDE_W;   // desc2 = current/query  (shifted columns)
  const unsigned* dj = s_j + jl * EPSC_STRIDE_W;   // desc1 = history
  unsigned sad[20];
#pragma unroll
  for (int s = 0; s < 20; s++) sad[s] = 0u;
#pragma unroll 1
  for (int ring = 0; ring < EPSC_RINGS; ring++) {
    unsigned a[20], r[25];
#pragma unroll
    for (int w = 0; w < 20; w++) a[w] = dj[ring * 20 + w];
    // extended query row: byte t of r = d2[ring][(t - 10) mod 80], t in [0, 100)
    //   bytes 0..9   <- columns 70..79 ; bytes 10..89 <- columns 0..79 ; bytes 90..99 <- columns 0..9
    unsigned c[20];
#pragma unroll
    for (int w = 0; w < 20; w++) c[w] = dq[ring * 20 + w];
    // r is c rotated right by 10 bytes with wrap: r_word[k] = bytes (4k-10 .. 4k-7) mod 80 of c
#pragma unroll
    for (int k = 0; k < 25; k++) {
      // source byte offset (4k - 10) mod 80 = 4 * ((k - 3 + 20) % 20) + 2  => funnel of words m, m+1 at 16 bits
      const int m = (k + 17) % 20, m1 = (k + 18) % 20;
      r[k] = __funnelshift_r(c[m], c[m1], 16);
    }
#pragma unroll
    for (int s = 0; s < 20; s++) {
      const int aa = s >> 2, bb = s & 3;
      unsigned acc = sad[s];
#pragma unroll
      for (int w = 0; w < 20; w++) {
        const unsigned x = bb == 0 ? r[w + aa] : __funnelshift_r(r[w + aa], r[w + aa + 1], 8 * bb);
        acc = __vsadu4(a[w], x) + acc;
      }
      sad[s] = acc;
    }
  }
  // first minimal shift wins (strict <), i = s - 10
  unsigned best = sad[0]; int bs = 0;
#pragma unroll
  for (int s = 1; s < 20; s++) if (sad[s] < best) { best = sad[s]; bs = s; }
  if (best < EPSC_SAD_GATE) {
    unsigned long long key = ((unsigned long long)best << 40) | ((unsigned long long)(unsigned)j << 8) | (unsigned long long)(unsigned)bs;
    unsigned long long* top = row_top + (size_t)r * EPSC_TOPK_SLOTS;
#pragma unroll 1
    for (int k = 0; k < EPSC_TOPK_SLOTS; k++) {
      const unsigned long long old = atomicMin(&top[k], key);
      key = old > key ? old : key;                 // the larger one moves on
      if (key == ~0ull) break;                      // displaced an empty slot: done
    }
  }
}

// row_top -> the interface arrays: per query the topk (<= 8) best candidates, best first (idx -1 when fewer qualify)
__global__ void k_epsc_topk(const unsigned long long* __restrict__ row_top, int n_rows, int topk,
                            int* __restrict__ idx, float* __restrict__ score, int8_t* __restrict__ shift) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * topk) return;
  const int r = t / topk, k = t % topk;
  const unsigned long long key = row_top[(size_t)r * EPSC_TOPK_SLOTS + k];
  if (key != ~0ull) {
    const int s = (int)(key >> 40), j = (int)((key >> 8) & 0xffffffull), bs = (int)(key & 0xffull);
    idx[t] = j; score[t] = (float)(1.0 - (double)s / (80 * 20 * 255)); shift[t] = (int8_t)(bs - 10);
  } else { idx[t] = -1; score[t] = 0.f; shift[t] = 0; }
}

// integer-ALU roofline of the pair scoring: a register-only chain of the kernel's inner-loop mix
// (one funnel shift + one VABSDIFF4.ACC per 4 byte-pairs) with no memory traffic; out[blockIdx] keeps the result live.
// grid = SMs * 8, block = 256; every thread issues iters * 64 SAD instructions on 8 independent accumulators.
__global__ void __launch_bounds__(256)
k_epsc_alu_peak(unsigned* __restrict__ out, int iters, unsigned seed) {
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, x0 = a0 ^ 0x9e3779b9u, x1 = a1 ^ 0x7f4a7c15u;
  unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const unsigned y0 = __funnelshift_r(x0, x1, 8), y1 = __funnelshift_r(x1, x0, 16);
      s0 = __vsadu4(a0, y0) + s0; s1 = __vsadu4(a1, y1) + s1; s2 = __vsadu4(a2, y0) + s2; s3 = __vsadu4(a3, y1) + s3;
      s4 = __vsadu4(a0, y1) + s4; s5 = __vsadu4(a1, y0) + s5; s6 = __vsadu4(a2, y1) + s6; s7 = __vsadu4(a3, y0) + s7;
      x0 += s0; x1 ^= s4;
    }
  }
  if (((s0 ^ s1) + (s2 ^ s3) + (s4 ^ s5) + (s6 ^ s7)) == 0x12345u) out[blockIdx.x] = x0;   // practically never: keeps the chain alive
}

}  // namespace lisreg
