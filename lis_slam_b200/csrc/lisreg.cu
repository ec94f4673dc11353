// lisreg C-ABI implementation (include/lisreg.h): context, map index build, LM driver.
// sm_100a only; there is no CPU fallback — every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/lisreg.h"
#include "lm.cuh"
#include "features.cuh"
#include "voxel.cuh"
#include "epsc.cuh"
#include "icp.cuh"
#include "loop.cuh"
#include "odom.cuh"
#include "pretreat.cuh"

using namespace lisreg;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct CloudIndex {   // one cloud of a map
  float4* sorted = nullptr;
  uint32_t* cell_start = nullptr;
  size_t cap_pts = 0, cap_cells = 0;   // allocated capacities: an index rebuilt in place (streaming odometry) only grows
  GridDev g{};
};

struct MapSlot {
  bool used = false;
  CloudIndex corner, surf;
};

// Everything one in-flight batch (or chunk of a batch) of the frame pipeline / LM loop needs.  A context owns three:
// ws[0] runs on the context stream (every synchronous entry point); ws[1..4] have private streams (created on first use); ws[1] and ws[2]
// alternate between the chunks of the pipelined e2e path, so that two chunks are in flight while a third uploads.
struct WorkSet {
  cudaStream_t stream = nullptr;
  DevBuf d_descs, d_states, d_partials, d_nbr, d_kstate, d_klist, d_geom;
  DevBuf d_feat, d_feat_frames, d_vox, d_vox_segs, d_imu;
  DevBuf d_feat_ctl;                       // k_feat_front control words (ticket, finished, group totals): all zero between launches
  int feat_cap_frames = 0, feat_cells = 0, feat_nscan = 0;
  void release() {
    for (DevBuf* b : {&d_descs, &d_states, &d_partials, &d_nbr, &d_kstate, &d_klist, &d_geom, &d_feat, &d_feat_frames, &d_vox, &d_vox_segs, &d_imu, &d_feat_ctl}) b->release();
  }
};

struct lisreg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int max_cells = 4 << 20;
  std::string err;
  int64_t launches = 0;
  std::vector<MapSlot> maps;
  MapDev* d_maps = nullptr; int d_maps_cap = 0; bool maps_dirty = true;
  // scratch
  DevBuf d_stage, d_logs, d_pose, d_res, d_tmp, d_bbox;
  PinBuf h_stage, h_out;
  WorkSet ws[5];   // 0: context stream; 1, 2: private streams (e2e chunks, submit / wait slots); 1..4: sub-batches of lisreg_frames_batch_dev
  WorkSet* cur = &ws[0];   // work set (and stream) the run_* drivers use; only the pipelined e2e path switches it
  DevBuf d_epsc, d_epsc2, d_icp;
  // e2e pipeline: H2D of chunk c+1 on copy_stream overlaps the compute of chunk c on `stream`
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  PinBuf h_desc;       // pinned descriptor staging of the arena entry points (one slice per chunk)
  int e2e_chunk = 128; // frames per chunk (LISREG_E2E_CHUNK; 0 = one copy, no overlap)
  int dev_split = 4;   // sub-batches lisreg_frames_batch_dev runs concurrently on private streams (LISREG_DEV_SPLIT; <= 1 = off)
  // asynchronous submit / wait pipeline: two slots, each with its own staging arena, result buffers and work set
  // (ws[1 + slot]); the upload of one batch overlaps the compute of the other
  struct AsyncSlot { DevBuf d_stage, d_res; PinBuf h_out, h_desc; cudaEvent_t done = nullptr, fence = nullptr; bool busy = false; int F = 0; };
  AsyncSlot slot[2];
  int next_slot = 0;
  int n_sm = 148;
  struct LoopDet {
    bool used = false;
    lisreg_loop_params prm{};
    int n = 0, cap = 0;                    // keyframes stored / capacity of the device history
    float4* d_proj = nullptr;              // [cap][360] sector projections
    uint8_t* d_desc = nullptr;             // [cap][3][1600] EPSC / SEPSC / FEPSC
    uint8_t* d_lut = nullptr;              // 256-entry using_label LUT
    std::vector<double> travel, px, py;    // travelDistanceArr, posArr (host bookkeeping, a few scalars per keyframe)
    std::vector<float> yaw;                // yawArr
  };
  std::vector<LoopDet> loops;
  DevBuf d_loop;                           // per-call scratch of lisreg_loop_detect
  struct Odom {                            // streaming odometry state (lisreg_odom_*)
    bool used = false;
    lisreg_odom_params prm{};
    float pose[6] = {0, 0, 0, 0, 0, 0}, last_pose[6] = {0, 0, 0, 0, 0, 0}, key_pose[6] = {0, 0, 0, 0, 0, 0};   // transformTobeMapped, lastTransformTobeMapped, transformPriFrame
    bool first_trans = false, have_last = false, first_flag = true;
    float last_imu[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};      // lastImuTransformation
    float last_imu_pre[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // lastImuPreTransformation
    bool have_imu_pre = false;             // lastImuPreTransAvailable
    float rot_tol = 0.f, z_tol = 0.f;      // transformUpdate clamps: applied on the host after the IMU slerp (the device loop runs without)
    float deltaR = 100.f, deltaT = 100.f;
    int keyframe_id = 0, frame_id = 0;
    int slots = 0, ccap = 0, scap = 0;     // ring of key-frame slots: capacities per slot (corner / surface points)
    float4* d_win_c = nullptr; float4* d_win_s = nullptr;
    std::vector<int> order, nc, ns;        // slot ids oldest .. newest; points per slot
    int map_id = -1; bool map_dirty = true; int n_map_c = 0, n_map_s = 0;
    DevBuf d_cat, d_mapvox, d_mapseg, d_in, d_io;
    PinBuf h_io, h_desc;
    int64_t graph_kernels = 0;
    cudaGraphExec_t gexec = nullptr; std::vector<const void*> gkey;   // captured per-frame graph + the buffer addresses it was captured with
  };
  std::vector<Odom> odoms;
  struct Submap {                          // localMap_t / submap_t class clouds in HBM (lisreg_submap_*)
    bool used = false;
    DevBuf cls[5]; int n[5] = {0, 0, 0, 0, 0};
    double bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
    CloudIndex dyn_index;                  // scratch index of the dynamic cloud (map-based dynamic removal)
    int icp_slot = -1; bool icp_dirty = true;   // cached ICP target (map slot) of lisreg_loop_verify
  };
  std::vector<Submap> submaps;
  DevBuf d_smvox, d_smcat;
  // multi-GPU exchange: NCCL communicator (own or adopted), private stream, fence / done events
  void* comm = nullptr; bool comm_owned = false; int comm_world = 1, comm_rank = 0;
  cudaStream_t comm_stream = nullptr; cudaEvent_t comm_fence = nullptr, comm_done[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t comm_seq = 0;                    // gathers issued so far; gather k signals comm_done[k % 4]
  int knn_coop_max = 16384;   // scan lists shorter than this are searched warp-per-query (LISREG_KNN_COOP_MAX; 0 = never)
  bool attr_front_set = false, attr_vox_set = false;   // opt-in to > 48 KB of dynamic shared memory done on this context's device
  int feat_fused = 0;     // LISREG_FEAT_FUSED=1: projection + compaction in one kernel with the range-image slice in shared memory (k_feat_front);
                          // measured slower than the global range image on firing-order sweeps (8x redundant ring-id scans), so off by default
  int feat_seg_unfused = 0; // LISREG_FEAT_SEG_UNFUSED=1: the batched pipelines run k_feat_curv_occl + the plain selection kernel too (parity check)
  int vox_batch_form = 0; // LISREG_VOX_BATCH_FORM=1: a handful of clouds also take the kernels meant for hundreds (tests reach them through lisreg_voxel_grid)
  int vox_unfused = 0;    // LISREG_VOX_UNFUSED=1: frame-sized clouds take the multi-kernel voxel path too (parity check of k_vox_block)
  int knn_noskip = 0;   // LISREG_KNN_NOSKIP=1: search every query from scratch at every iteration (parity check of the CHECK path)
  // profiling
  bool prof_on = false;
  struct EvPair { cudaEvent_t a, b; int kind; double bytes; int64_t launches; };
  std::vector<EvPair> ev_pending;
  std::vector<cudaEvent_t> ev_free;
  lisreg_profile prof{};
};

enum { PROF_LM = 0, PROF_FEAT = 1, PROF_VOXEL = 2, PROF_INDEX = 3 };

static cudaEvent_t ev_get(lisreg_ctx* ctx) {
  if (!ctx->ev_free.empty()) { cudaEvent_t e = ctx->ev_free.back(); ctx->ev_free.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {   // records an event pair around a group of launches when profiling is on
  lisreg_ctx* ctx; lisreg_ctx::EvPair p; bool on;
  ProfScope(lisreg_ctx* c, int kind, double bytes, int64_t launches) : ctx(c), on(c->prof_on) {
    if (!on) return;
    p.a = ev_get(c); p.b = ev_get(c); p.kind = kind; p.bytes = bytes; p.launches = launches;
    cudaEventRecord(p.a, c->cur->stream);
  }
  ~ProfScope() { if (on) { cudaEventRecord(p.b, ctx->cur->stream); ctx->ev_pending.push_back(p); } }
};

// [off, off + bytes) inside an arena of `total` bytes, without wrapping for offsets near 2^64
static inline bool in_arena(uint64_t off, uint64_t bytes, uint64_t total) { return off <= total && bytes <= total - off; }

static int fail(lisreg_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return fail(ctx, LISREG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define LAUNCH_CK() do { ctx->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
  return fail(ctx, LISREG_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ------------------------------------------------------------------------------------------------
// grid build kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_bbox_init(unsigned* bb) { if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu; else if (threadIdx.x < 6) bb[threadIdx.x] = 0u; }

__global__ void k_bbox(const float4* __restrict__ pts, int n, const int* __restrict__ n_ptr, unsigned* __restrict__ bb) {
  if (n_ptr) n = *n_ptr;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[i]);
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { atomicMin(&bb[d], f2ord(mn[d])); atomicMax(&bb[3 + d], f2ord(mx[d])); }
  }
}

// one thread: choose the cell size (>= h_req, grown until the grid fits max_cells) and dims
__global__ void k_grid_plan(const unsigned* __restrict__ bb, float h_req, int max_cells, int n, const int* __restrict__ n_ptr, GridDev* __restrict__ out) {
  GridDev g;
  if (n_ptr) n = *n_ptr;
  float mn[3], mx[3];
  for (int d = 0; d < 3; d++) { mn[d] = ord2f(bb[d]); mx[d] = ord2f(bb[3 + d]); }
  if (n <= 0) { for (int d = 0; d < 3; d++) { mn[d] = 0.f; mx[d] = 0.f; } }
  float h = h_req;
  int nx = 0, ny = 0, nz = 0;
  for (int grow = 0; grow < 400; grow++) {   // bounded: non-finite boxes never satisfy the test (the host rejects the plan)
    nx = (int)floorf((mx[0] - mn[0]) / h) + 2; ny = (int)floorf((mx[1] - mn[1]) / h) + 2; nz = (int)floorf((mx[2] - mn[2]) / h) + 2;
    if ((double)nx * (double)ny * (double)nz <= (double)max_cells) break;
    h *= 1.25f;
  }
  g.ox = mn[0] - 0.5f * h; g.oy = mn[1] - 0.5f * h; g.oz = mn[2] - 0.5f * h;   // half-cell apron keeps boundary points interior
  g.h = h; g.inv_h = 1.0f / h; g.nx = nx; g.ny = ny; g.nz = nz; g.n = n; g.ncells = nx * ny * nz;
  g.cell_start = nullptr; g.pts = nullptr;
  *out = g;
}

__device__ __forceinline__ int cell_of(const GridDev& g, float4 p) {
  int cx = min(max(cell_coord(p.x, g.ox, g.inv_h), 0), g.nx - 1);
  int cy = min(max(cell_coord(p.y, g.oy, g.inv_h), 0), g.ny - 1);
  int cz = min(max(cell_coord(p.z, g.oz, g.inv_h), 0), g.nz - 1);
  return (cz * g.ny + cy) * g.nx + cx;
}

__global__ void k_cell_count(const float4* __restrict__ pts, int n, GridDev g, int* __restrict__ cell_id, uint32_t* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_of(g, __ldg(&pts[i]));
  cell_id[i] = c;
  atomicAdd(&counts[c], 1u);
}

// exclusive scan, three phases, 1024 elements per block
constexpr int SCAN_BLOCK = 1024;
__global__ void k_scan_local(uint32_t* __restrict__ data, int n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s[SCAN_BLOCK];
  int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  uint32_t v = i < n ? data[i] : 0u;
  s[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
    uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
    __syncthreads();
    s[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < n) data[i] = s[threadIdx.x] - v;   // exclusive
  if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void k_scan_sums(uint32_t* __restrict__ block_sums, int nb) {   // single block, sequential chunks
  __shared__ uint32_t s[SCAN_BLOCK];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_BLOCK) {
    int i = base + threadIdx.x;
    uint32_t v = i < nb ? block_sums[i] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) block_sums[i] = s[threadIdx.x] - v + carry;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) carry += s[threadIdx.x];
    __syncthreads();
  }
}
__global__ void k_scan_add(uint32_t* __restrict__ data, int n, const uint32_t* __restrict__ block_sums) {
  int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (i < n) data[i] += block_sums[blockIdx.x];
}

__global__ void k_cell_scatter(const float4* __restrict__ pts, int n, const int* __restrict__ cell_id,
                               const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ fill, float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_id[i];
  uint32_t pos = cell_start[c] + atomicAdd(&fill[c], 1u);
  float4 p = __ldg(&pts[i]);
  p.w = __int_as_float(i);
  sorted[pos] = p;
}

// orders the points of each cell by original index (the atomic scatter leaves them unordered), so
// the index - and every result that depends on tie-breaking by position - is reproducible
__global__ void k_cell_order(float4* __restrict__ sorted, const uint32_t* __restrict__ cell_start, int ncells) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint32_t b = cell_start[c], e = cell_start[c + 1];
  for (uint32_t i = b + 1; i < e; i++) {
    float4 v = sorted[i];
    const int key = __float_as_int(v.w);
    uint32_t j = i;
    while (j > b && __float_as_int(sorted[j - 1].w) > key) { sorted[j] = sorted[j - 1]; j--; }
    sorted[j] = v;
  }
}

// builds one cloud index from device-resident points.  d_n (nullable): the point count lives on the device (output of a
// voxel grid); n is then only the capacity bound used to size launches and buffers.  The buffers of *ci are re-used
// when large enough (in-place rebuild).
static int build_cloud_index(lisreg_ctx* ctx, const float4* d_pts, int n, float h_req, CloudIndex* ci, const int* d_n = nullptr) {
  cudaStream_t st = ctx->cur->stream;
  ProfScope ps(ctx, PROF_INDEX, 16.0 * n, 1);
  CK(ctx->d_bbox.reserve(6 * sizeof(unsigned) + sizeof(GridDev)));
  unsigned* bb = (unsigned*)ctx->d_bbox.p;
  GridDev* d_g = (GridDev*)((char*)ctx->d_bbox.p + 32);
  k_bbox_init<<<1, 32, 0, st>>>(bb); LAUNCH_CK();
  if (n > 0) { k_bbox<<<std::min(1184, (n + 255) / 256), 256, 0, st>>>(d_pts, n, d_n, bb); LAUNCH_CK(); }
  k_grid_plan<<<1, 1, 0, st>>>(bb, h_req, ctx->max_cells, n, d_n, d_g); LAUNCH_CK();
  GridDev g;
  CK(cudaMemcpyAsync(&g, d_g, sizeof(g), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  // a NaN / Inf coordinate makes the planned dimensions meaningless (overflowed or negative): reject the cloud
  if (!(g.nx > 0 && g.ny > 0 && g.nz > 0) || !std::isfinite(g.ox) || !std::isfinite(g.oy) || !std::isfinite(g.oz) || !std::isfinite(g.h) ||
      (double)g.nx * (double)g.ny * (double)g.nz > (double)ctx->max_cells || g.ncells != g.nx * g.ny * g.nz || g.n < 0 || g.n > n)
    return fail(ctx, LISREG_ERR_ARG, "map cloud has non-finite coordinates (grid %d x %d x %d)", g.nx, g.ny, g.nz);
  n = g.n;
  const int ncells = g.ncells;
  if ((size_t)std::max(n, 1) > ci->cap_pts) {
    if (ci->sorted) cudaFree(ci->sorted);
    ci->sorted = nullptr; ci->cap_pts = 0;
    const size_t want = (size_t)std::max(n, 1) + (ci->cap_cells ? (size_t)n / 4 : 0);   // in-place rebuilds grow with slack
    CK(cudaMalloc(&ci->sorted, sizeof(float4) * want));
    ci->cap_pts = want;
  }
  if ((size_t)ncells + 1 > ci->cap_cells) {
    const bool regrow = ci->cap_cells != 0;
    if (ci->cell_start) cudaFree(ci->cell_start);
    ci->cell_start = nullptr; ci->cap_cells = 0;
    const size_t want = (size_t)ncells + 1 + (regrow ? (size_t)ncells / 4 : 0);
    cudaError_t e2 = cudaMalloc(&ci->cell_start, sizeof(uint32_t) * want);
    if (e2 != cudaSuccess) return fail(ctx, LISREG_ERR_CUDA, "cudaMalloc(cell_start) failed: %s", cudaGetErrorString(e2));
    ci->cap_cells = want;
  }
  const int nblk = (ncells + 1 + SCAN_BLOCK - 1) / SCAN_BLOCK;
  CK(ctx->d_tmp.reserve(sizeof(int) * (size_t)std::max(n, 1) + sizeof(uint32_t) * ((size_t)ncells + 1) + sizeof(uint32_t) * (size_t)(nblk + 1)));
  int* cell_id = (int*)ctx->d_tmp.p;
  uint32_t* fill = (uint32_t*)(cell_id + std::max(n, 1));
  uint32_t* bsums = fill + (ncells + 1);
  CK(cudaMemsetAsync(ci->cell_start, 0, sizeof(uint32_t) * ((size_t)ncells + 1), st));
  CK(cudaMemsetAsync(fill, 0, sizeof(uint32_t) * ((size_t)ncells + 1), st));
  if (n > 0) { k_cell_count<<<(n + 255) / 256, 256, 0, st>>>(d_pts, n, g, cell_id, ci->cell_start); LAUNCH_CK(); }
  k_scan_local<<<nblk, SCAN_BLOCK, 0, st>>>(ci->cell_start, ncells + 1, bsums); LAUNCH_CK();
  k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(bsums, nblk); LAUNCH_CK();
  k_scan_add<<<nblk, SCAN_BLOCK, 0, st>>>(ci->cell_start, ncells + 1, bsums); LAUNCH_CK();
  if (n > 0) {
    k_cell_scatter<<<(n + 255) / 256, 256, 0, st>>>(d_pts, n, cell_id, ci->cell_start, fill, ci->sorted); LAUNCH_CK();
    k_cell_order<<<(ncells + 127) / 128, 128, 0, st>>>(ci->sorted, ci->cell_start, ncells); LAUNCH_CK();
  }
  g.cell_start = ci->cell_start; g.pts = ci->sorted;
  ci->g = g;
  return LISREG_OK;
}

// Cell size: the search is exact for ANY cell size (shell expansion); ~0.6 m keeps the 3x3x3 block
// at ~20-30 candidates for 0.4 m voxel-grid surf maps while the 5th neighbour usually lies inside it.
static float cell_size_for_gate(float gate) {
  const char* e = getenv("LISREG_CELL");
  float h = e ? (float)atof(e) : 0.6f;
  const float r = sqrtf(gate > 0.f ? gate : 1.f);
  return fminf(h, 1.002f * r + 1e-3f);
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* lisreg_version(void) { return "lisreg 0.1 (sm_100a)"; }

int32_t lisreg_create(const lisreg_config* cfg, lisreg_ctx** out) {
  if (!out) return LISREG_ERR_ARG;
  *out = nullptr;
  lisreg_ctx* ctx = new lisreg_ctx();
  ctx->device = cfg ? cfg->device : 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || ctx->device >= ndev) {
    // no CPU fallback: fail loudly
    fprintf(stderr, "lisreg_create: no usable CUDA device (%s); this engine has no CPU path\n",
            e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
    delete ctx;
    return LISREG_ERR_CUDA;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) { delete ctx; return LISREG_ERR_CUDA; }
  if (cfg && cfg->own_stream) {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LISREG_ERR_CUDA; }
    ctx->own_stream = true;
  } else { ctx->stream = cfg ? (cudaStream_t)cfg->stream : nullptr; ctx->own_stream = false; }   // NULL = legacy default stream
  ctx->ws[0].stream = ctx->stream;
  if (cfg && cfg->max_grid_cells > 0) ctx->max_cells = cfg->max_grid_cells;
  { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, ctx->device) == cudaSuccess && v > 0) ctx->n_sm = v; }
  if (const char* e2 = getenv("LISREG_KNN_NOSKIP")) ctx->knn_noskip = atoi(e2) ? 1 : 0;
  if (const char* e3 = getenv("LISREG_FEAT_FUSED")) ctx->feat_fused = atoi(e3) ? 1 : 0;
  if (const char* e5 = getenv("LISREG_VOX_UNFUSED")) ctx->vox_unfused = atoi(e5) ? 1 : 0;
  if (const char* e6 = getenv("LISREG_VOX_BATCH_FORM")) ctx->vox_batch_form = atoi(e6) ? 1 : 0;
  if (const char* e7 = getenv("LISREG_FEAT_SEG_UNFUSED")) ctx->feat_seg_unfused = atoi(e7) ? 1 : 0;
  if (const char* e4 = getenv("LISREG_KNN_COOP_MAX")) ctx->knn_coop_max = std::max(0, atoi(e4));
  if (const char* e3 = getenv("LISREG_E2E_CHUNK")) ctx->e2e_chunk = std::max(0, atoi(e3));
  if (const char* e4 = getenv("LISREG_DEV_SPLIT")) ctx->dev_split = std::min(4, std::max(0, atoi(e4)));
  *out = ctx;
  return LISREG_OK;
}

void lisreg_destroy(lisreg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& m : ctx->maps) if (m.used) {
    cudaFree(m.corner.sorted); cudaFree(m.corner.cell_start); cudaFree(m.surf.sorted); cudaFree(m.surf.cell_start);
  }
  if (ctx->d_maps) cudaFree(ctx->d_maps);
  for (DevBuf* b : {&ctx->d_stage, &ctx->d_logs, &ctx->d_pose, &ctx->d_res, &ctx->d_tmp, &ctx->d_bbox, &ctx->d_epsc, &ctx->d_epsc2, &ctx->d_icp}) b->release();
  for (int i = 0; i < 5; i++) { if (i > 0 && ctx->ws[i].stream) { cudaStreamSynchronize(ctx->ws[i].stream); cudaStreamDestroy(ctx->ws[i].stream); } ctx->ws[i].release(); }
  ctx->h_stage.release(); ctx->h_out.release(); ctx->h_desc.release();
  for (auto& L : ctx->loops) if (L.used) { cudaFree(L.d_proj); cudaFree(L.d_desc); cudaFree(L.d_lut); }
  lisreg_comm_destroy(ctx);
  for (auto& S : ctx->submaps) if (S.used) { for (auto& b : S.cls) b.release(); cudaFree(S.dyn_index.sorted); cudaFree(S.dyn_index.cell_start); }
  ctx->d_smvox.release(); ctx->d_smcat.release();
  for (auto& O : ctx->odoms) if (O.used) {
    if (O.gexec) cudaGraphExecDestroy(O.gexec);
    cudaFree(O.d_win_c); cudaFree(O.d_win_s);
    for (DevBuf* b : {&O.d_cat, &O.d_mapvox, &O.d_mapseg, &O.d_in, &O.d_io}) b->release();
    O.h_io.release(); O.h_desc.release();
  }
  ctx->d_loop.release();
  for (auto e : ctx->chunk_ev) cudaEventDestroy(e);
  for (auto& sl : ctx->slot) { sl.d_stage.release(); sl.d_res.release(); sl.h_out.release(); sl.h_desc.release(); if (sl.done) cudaEventDestroy(sl.done); if (sl.fence) cudaEventDestroy(sl.fence); }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& p : ctx->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : ctx->ev_free) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* lisreg_last_error(const lisreg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t lisreg_launch_count(const lisreg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t lisreg_sync(lisreg_ctx* ctx) {
  if (!ctx) return LISREG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return LISREG_OK;
}

void lisreg_lm_params_preset(lisreg_lm_params* p, char variant) {
  static const float score[20] = {1.0f, 1.0f, 0.6f, 0.5f, 0.8f, 0.5f, 0.5f, 0.5f, 0.5f, 1.2f,
                                  1.2f, 1.2f, 0.5f, 1.0f, 0.8f, 0.5f, 1.3f, 0.5f, 1.5f, 1.5f};   // config/label.yaml:214-234
  memset(p, 0, sizeof(*p));
  p->early_exit = 1; p->edge_min_valid = -1; p->surf_min_valid = 100; p->min_sel = 50; p->degenerate_eig = 100.f;
  p->rot_tolerance = 1000.f; p->z_tolerance = 1000.f;   // config/params.yaml:123-124
  for (int i = 0; i < 20; i++) p->label_score[i] = score[i];
  switch (variant) {
    case 'B': p->max_iters = 20; p->sqdist_gate = 2.0f; p->conv_rot_deg = 0.003f; p->conv_trans_cm = 0.03f; p->use_label_weight = 1; break;
    case 'C': p->max_iters = 30; p->sqdist_gate = 2.0f; p->conv_rot_deg = 0.002f; p->conv_trans_cm = 0.02f; p->use_label_weight = 1; break;
    default:  p->max_iters = 15; p->sqdist_gate = 1.0f; p->conv_rot_deg = 0.005f; p->conv_trans_cm = 0.05f; p->use_label_weight = 0; break;
  }
}

static int map_alloc_slot(lisreg_ctx* ctx) {
  for (size_t i = 0; i < ctx->maps.size(); i++) if (!ctx->maps[i].used) return (int)i;
  ctx->maps.emplace_back();
  return (int)ctx->maps.size() - 1;
}

int32_t lisreg_map_create_dev(lisreg_ctx* ctx, const float* d_corner, int32_t mc, const float* d_surf, int32_t ms,
                              float gate_hint, int32_t* map_id) {
  if (!ctx || !map_id || mc < 0 || ms < 0 || (mc > 0 && !d_corner) || (ms > 0 && !d_surf)) return fail(ctx, LISREG_ERR_ARG, "lisreg_map_create: bad argument");
  CK(cudaSetDevice(ctx->device));
  const int slot = map_alloc_slot(ctx);
  MapSlot& m = ctx->maps[slot];
  const float h = cell_size_for_gate(gate_hint);
  auto drop = [](CloudIndex& ci) { cudaFree(ci.sorted); cudaFree(ci.cell_start); ci = CloudIndex(); };
  int rc = build_cloud_index(ctx, (const float4*)d_corner, mc, h, &m.corner);
  if (rc) { cudaStreamSynchronize(ctx->stream); drop(m.corner); return rc; }
  rc = build_cloud_index(ctx, (const float4*)d_surf, ms, h, &m.surf);
  if (rc) { cudaStreamSynchronize(ctx->stream); drop(m.corner); drop(m.surf); return rc; }
  m.used = true;
  ctx->maps_dirty = true;
  *map_id = slot;
  return LISREG_OK;
}

int32_t lisreg_map_create(lisreg_ctx* ctx, const float* corner, int32_t mc, const float* surf, int32_t ms,
                          float gate_hint, int32_t* map_id) {
  if (!ctx || !map_id || mc < 0 || ms < 0 || (mc > 0 && !corner) || (ms > 0 && !surf)) return fail(ctx, LISREG_ERR_ARG, "lisreg_map_create: bad argument");
  CK(cudaSetDevice(ctx->device));
  const size_t bc = sizeof(float4) * (size_t)mc, bs = sizeof(float4) * (size_t)ms;
  CK(ctx->d_stage.reserve(bc + bs + 32));
  char* d = (char*)ctx->d_stage.p;
  if (mc) CK(cudaMemcpyAsync(d, corner, bc, cudaMemcpyHostToDevice, ctx->stream));
  if (ms) CK(cudaMemcpyAsync(d + bc, surf, bs, cudaMemcpyHostToDevice, ctx->stream));
  return lisreg_map_create_dev(ctx, (const float*)d, mc, (const float*)(d + bc), ms, gate_hint, map_id);
}

int32_t lisreg_map_destroy(lisreg_ctx* ctx, int32_t map_id) {
  if (!ctx || map_id < 0 || map_id >= (int)ctx->maps.size() || !ctx->maps[map_id].used) return fail(ctx, LISREG_ERR_ARG, "lisreg_map_destroy: bad map id");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 1; i < 5; i++) if (ctx->ws[i].stream) CK(cudaStreamSynchronize(ctx->ws[i].stream));   // batches still in flight may read the map
  MapSlot& m = ctx->maps[map_id];
  cudaFree(m.corner.sorted); cudaFree(m.corner.cell_start); cudaFree(m.surf.sorted); cudaFree(m.surf.cell_start);
  m = MapSlot();
  ctx->maps_dirty = true;
  return LISREG_OK;
}

static int sync_maps(lisreg_ctx* ctx) {
  if (!ctx->maps_dirty) return LISREG_OK;
  const int n = (int)ctx->maps.size();
  if (n > ctx->d_maps_cap) {
    if (ctx->d_maps) cudaFree(ctx->d_maps);
    ctx->d_maps_cap = std::max(16, 2 * n);
    CK(cudaMalloc(&ctx->d_maps, sizeof(MapDev) * ctx->d_maps_cap));
  }
  std::vector<MapDev> h(n);
  for (int i = 0; i < n; i++) { h[i].corner = ctx->maps[i].corner.g; h[i].surf = ctx->maps[i].surf.g; }
  if (n) {
    CK(cudaMemcpyAsync(ctx->d_maps, h.data(), sizeof(MapDev) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // h is a stack-lifetime pageable buffer
  }
  ctx->maps_dirty = false;
  return LISREG_OK;
}

int32_t lisreg_knn5(lisreg_ctx* ctx, int32_t map_id, int32_t which, const float* queries, int32_t nq, float sqdist_gate,
                    int32_t* idx, float* sqd) {
  if (!ctx || map_id < 0 || map_id >= (int)ctx->maps.size() || !ctx->maps[map_id].used || nq < 0) return fail(ctx, LISREG_ERR_ARG, "lisreg_knn5: bad argument");
  if (nq == 0) return LISREG_OK;
  CK(cudaSetDevice(ctx->device));
  const GridDev g = which == 0 ? ctx->maps[map_id].corner.g : ctx->maps[map_id].surf.g;
  const size_t bq = sizeof(float4) * (size_t)nq, bi = sizeof(int) * 5 * (size_t)nq, bd = sizeof(float) * 5 * (size_t)nq;
  CK(ctx->d_stage.reserve(bq + bi + bd));
  char* d = (char*)ctx->d_stage.p;
  CK(cudaMemcpyAsync(d, queries, bq, cudaMemcpyHostToDevice, ctx->stream));
  k_knn5<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(g, (const float4*)d, nq, sqdist_gate, (int*)(d + bq), (float*)(d + bq + bi), nullptr); LAUNCH_CK();
  CK(cudaMemcpyAsync(idx, d + bq, bi, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(sqd, d + bq + bi, bd, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LISREG_OK;
}

int32_t lisreg_map_distance_filter(lisreg_ctx* ctx, int32_t map_id, int32_t which, const float* feat, int32_t n, float center_radius,
                                   float dyn_min, float dyn_max, float near_thre, uint8_t* keep, int32_t* n_kept) {
  if (!ctx || map_id < 0 || map_id >= (int)ctx->maps.size() || !ctx->maps[map_id].used || n < 0 || (n > 0 && (!feat || !keep)))
    return fail(ctx, LISREG_ERR_ARG, "lisreg_map_distance_filter: bad argument");
  if (n_kept) *n_kept = n;
  if (n <= 10) { for (int i = 0; i < n; i++) keep[i] = 1; return LISREG_OK; }     // subMap.h:1069-1070: small clouds are left alone
  CK(cudaSetDevice(ctx->device));
  const GridDev g = which == 0 ? ctx->maps[map_id].corner.g : ctx->maps[map_id].surf.g;
  const size_t bq = sizeof(float4) * (size_t)n;
  CK(ctx->d_stage.reserve(bq + (size_t)n + 64));
  char* d = (char*)ctx->d_stage.p;
  CK(cudaMemcpyAsync(d, feat, bq, cudaMemcpyHostToDevice, ctx->stream));
  const float near2 = near_thre * near_thre, dmin2 = dyn_min * dyn_min, dmax2 = dyn_max * dyn_max;   // fp32 squares, like upstream
  float top = dmin2;
  if (std::isfinite(dmax2) && dmax2 > top) top = dmax2;
  if (near2 > top) top = near2;
  const float gate = top * 1.000001f + 1e-12f;
  k_map_distance_filter<<<(n + 127) / 128, 128, 0, ctx->stream>>>(g, (const float4*)d, n, center_radius * center_radius, near2, dmin2, dmax2,
                                                                     gate, (unsigned char*)(d + bq)); LAUNCH_CK();
  CK(cudaMemcpyAsync(keep, d + bq, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (n_kept) { int c = 0; for (int i = 0; i < n; i++) c += keep[i] != 0; *n_kept = c; }
  return LISREG_OK;
}

static void to_dev_params(const lisreg_lm_params* p, LmParamsDev* d) {
  d->max_iters = std::min(p->max_iters, (int)LISREG_MAX_ITERS); d->early_exit = p->early_exit;
  d->gate = p->sqdist_gate; d->conv_rot = p->conv_rot_deg; d->conv_trans = p->conv_trans_cm;
  d->edge_min = p->edge_min_valid; d->surf_min = p->surf_min_valid; d->min_sel = p->min_sel;
  d->degenerate_eig = p->degenerate_eig; d->use_w = p->use_label_weight;
  d->rot_tol = p->rot_tolerance; d->z_tol = p->z_tolerance; d->degenerate_in = p->degenerate_in;
  memcpy(d->label_score, p->label_score, sizeof(d->label_score));
}

// core driver: descs already on the device. max_n = largest nc+ns of the batch.
// tiling_B: batch size that decides the tile size (a chunk of a larger batch must tile like the whole batch so
// that the fixed summation order - hence every bit of the result - does not depend on the chunking); 0 = B
// it0 / it1: run only the iterations [it0, it1) of the loop (it1 < 0 = all): the streaming odometry launches the loop in
// slices and looks at the convergence flag in between instead of launching 15 x 6 kernels that mostly find `done` set;
// it0 == 0 includes k_lm_init and the counter reset, every slice ends with k_lm_finish (the result is valid after each)
static int run_lm(lisreg_ctx* ctx, int B, const RegDesc* d_descs, int max_n, double alg_bytes_per_iter, float* d_pose,
                  const lisreg_lm_params* prm, lisreg_lm_result* d_res, lisreg_lm_iter* d_logs, int tiling_B = 0, int it0 = 0, int it1 = -1) {
  cudaStream_t st = ctx->cur->stream;
  int rc = sync_maps(ctx);
  if (rc) return rc;
  LmParamsDev dp; to_dev_params(prm, &dp);
  // tile size: big batches amortise the 27-term reduction over 4 queries per thread; a lone
  // registration is spread over as many SMs as possible
  if (tiling_B <= 0) tiling_B = B;
  const int tile_pts = (tiling_B >= 32) ? LM_MAX_TILE : LM_THREADS;
  const int max_tiles = std::max(1, (max_n + tile_pts - 1) / tile_pts);
  CK(ctx->cur->d_states.reserve(sizeof(RegState) * (size_t)B));
  CK(ctx->cur->d_partials.reserve(sizeof(double) * LM_NSUM * (size_t)B * max_tiles));
  RegState* states = (RegState*)ctx->cur->d_states.p;
  double* partials = (double*)ctx->cur->d_partials.p;
  if (it1 < 0 || it1 > dp.max_iters) it1 = dp.max_iters;
  if (it0 == 0) { k_lm_init<<<(B + 127) / 128, 128, 0, st>>>(d_descs, states, d_pose, dp, B); LAUNCH_CK(); }
  // a block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...: batches get 32 blocks per registration
  // (the real tile count is only known on the device in the frame pipeline), a lone registration gets
  // one block per tile so that it spreads over the whole GPU
  dim3 grid(std::min(max_tiles, tiling_B >= 32 ? 32 : 1024), B);
  {
    ProfScope ps(ctx, PROF_LM, alg_bytes_per_iter * (it1 - it0), it1 - it0);
    const size_t slots = (size_t)tile_pts * max_tiles * B;
    if (slots >= (size_t)0xffffffffu) return fail(ctx, LISREG_ERR_CAPACITY, "batch too large: %zu query slots (split the batch)", slots);
    CK(ctx->cur->d_nbr.reserve(sizeof(int) * 5 * slots));
    CK(ctx->cur->d_kstate.reserve(sizeof(KnnState) * slots));
    CK(ctx->cur->d_geom.reserve(sizeof(GeomCache) * slots));
    CK(ctx->cur->d_klist.reserve(sizeof(unsigned) * 2 * slots + sizeof(int) * 2 * LISREG_MAX_ITERS));
    int* nbr = (int*)ctx->cur->d_nbr.p;
    KnnState* kstate = (KnnState*)ctx->cur->d_kstate.p;
    GeomCache* geom = (GeomCache*)ctx->cur->d_geom.p;
    unsigned* scan_list = (unsigned*)ctx->cur->d_klist.p;
    unsigned* shell_list = scan_list + slots;
    int* counters = (int*)(shell_list + slots);          // [it][2]: scan / shell list lengths of iteration it
    if (it0 == 0) CK(cudaMemsetAsync(counters, 0, sizeof(int) * 2 * LISREG_MAX_ITERS, st));
    const int tile_shift = tile_pts == LM_MAX_TILE ? 9 : 7;
    static_assert(LM_MAX_TILE == 512 && LM_THREADS == 128, "tile_shift assumes 512 / 128 query tiles");
    for (int it = it0; it < it1; it++) {
      // iteration 0 searches every query; later iterations first try to PROVE that the neighbours did not change
      k_knn_check<<<grid, LM_THREADS, 0, st>>>(d_descs, states, ctx->d_maps, dp.gate, nbr, kstate, geom, scan_list, counters + 2 * it,
                                               max_tiles, tile_shift, (it > 0 && !ctx->knn_noskip) ? 1 : 0); LAUNCH_CK();
      // short scan lists (late iterations) go to the warp-per-query kernel, long ones to the thread-per-query scan
      k_knn_coop<<<ctx->n_sm * 8, LM_THREADS, 0, st>>>(d_descs, states, ctx->d_maps, dp.gate, nbr, kstate, geom, scan_list, counters + 2 * it,
                                               ctx->knn_coop_max, max_tiles, tile_shift); LAUNCH_CK();
      k_knn_search<false><<<ctx->n_sm * LM_KNN_MIN_BLOCKS, LM_THREADS, 0, st>>>(d_descs, states, ctx->d_maps, dp.gate, nbr, kstate, geom, scan_list,
                                               counters + 2 * it, shell_list, counters + 2 * it + 1, ctx->knn_coop_max, max_tiles, tile_shift); LAUNCH_CK();
      k_knn_search<true><<<ctx->n_sm * 4, LM_THREADS, 0, st>>>(d_descs, states, ctx->d_maps, dp.gate, nbr, kstate, geom, shell_list,
                                               counters + 2 * it + 1, nullptr, nullptr, 0, max_tiles, tile_shift); LAUNCH_CK();
      k_lm_resid<<<grid, LM_THREADS, 0, st>>>(d_descs, states, ctx->d_maps, dp, nbr, geom, partials, max_tiles, tile_pts); LAUNCH_CK();
      k_lm_solve<<<B, LM_SOLVE_THREADS, 0, st>>>(
          d_descs, states, dp, partials, d_logs, max_tiles, tile_pts, B); LAUNCH_CK();
    }
  }
  k_lm_finish<<<(B + 127) / 128, 128, 0, st>>>(states, d_pose, d_res, B); LAUNCH_CK();
  return LISREG_OK;
}

int32_t lisreg_scan2map_batch_dev(lisreg_ctx* ctx, int32_t B, const lisreg_batch_item* items, float* d_pose6xB,
                                  const lisreg_lm_params* prm, lisreg_lm_result* d_resxB) {
  if (!ctx || B <= 0 || !items || !d_pose6xB || !prm || !d_resxB) return fail(ctx, LISREG_ERR_ARG, "lisreg_scan2map_batch_dev: bad argument");
  if (prm->max_iters <= 0 || prm->max_iters > LISREG_MAX_ITERS) return fail(ctx, LISREG_ERR_ARG, "max_iters must be in 1..%d", LISREG_MAX_ITERS);
  CK(cudaSetDevice(ctx->device));
  CK(ctx->cur->d_descs.reserve(sizeof(RegDesc) * (size_t)B));
  // pageable staging on purpose: cudaMemcpyAsync returns once a pageable source has been
  // consumed, so back-to-back asynchronous calls cannot race on the descriptor staging
  std::vector<RegDesc> hvec((size_t)B);
  RegDesc* h = hvec.data();
  int max_n = 0; double alg = 0;
  for (int b = 0; b < B; b++) {
    const lisreg_batch_item& it = items[b];
    if (it.nc < 0 || it.ns < 0 || it.map_id < 0 || it.map_id >= (int)ctx->maps.size() || !ctx->maps[it.map_id].used)
      return fail(ctx, LISREG_ERR_ARG, "batch item %d: bad sizes or map id", b);
    alg += 96.0 * (it.nc + it.ns);
    h[b].corner = (const float4*)it.corner; h[b].surf = (const float4*)it.surf;
    h[b].clabel = it.clabel; h[b].slabel = it.slabel; h[b].nc = it.nc; h[b].ns = it.ns; h[b].map_slot = it.map_id; h[b].pad = 0; h[b].nc_ptr = nullptr; h[b].ns_ptr = nullptr;
    max_n = std::max(max_n, it.nc + it.ns);
  }
  CK(cudaMemcpyAsync(ctx->cur->d_descs.p, h, sizeof(RegDesc) * (size_t)B, cudaMemcpyHostToDevice, ctx->stream));
  lisreg_lm_iter* d_logs = nullptr;
  return run_lm(ctx, B, (const RegDesc*)ctx->cur->d_descs.p, max_n, alg, d_pose6xB, prm, d_resxB, d_logs);
}

int32_t lisreg_scan2map_batch(lisreg_ctx* ctx, int32_t B, const lisreg_batch_item* items, float* pose6xB,
                              const lisreg_lm_params* prm, lisreg_lm_result* resxB, lisreg_lm_iter* iter_log) {
  if (!ctx || B <= 0 || !items || !pose6xB || !prm || !resxB) return fail(ctx, LISREG_ERR_ARG, "lisreg_scan2map_batch: bad argument");
  if (prm->max_iters <= 0 || prm->max_iters > LISREG_MAX_ITERS) return fail(ctx, LISREG_ERR_ARG, "max_iters must be in 1..%d", LISREG_MAX_ITERS);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  // pack: [descs][poses][clouds...][labels...] into one pinned staging buffer -> one H2D copy
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
  const size_t o_desc = take(sizeof(RegDesc) * (size_t)B);
  const size_t o_pose = take(sizeof(float) * 6 * (size_t)B);
  std::vector<size_t> oc(B), os(B), ocl(B), osl(B);
  int max_n = 0; double alg = 0;
  for (int b = 0; b < B; b++) {
    const lisreg_batch_item& it = items[b];
    if (it.nc < 0 || it.ns < 0 || (it.nc > 0 && !it.corner) || (it.ns > 0 && !it.surf) ||
        it.map_id < 0 || it.map_id >= (int)ctx->maps.size() || !ctx->maps[it.map_id].used)
      return fail(ctx, LISREG_ERR_ARG, "batch item %d: bad sizes, pointers or map id", b);
    oc[b] = take(sizeof(float4) * (size_t)it.nc);
    os[b] = take(sizeof(float4) * (size_t)it.ns);
    ocl[b] = it.clabel ? take(sizeof(uint16_t) * (size_t)it.nc) : (size_t)-1;
    osl[b] = it.slabel ? take(sizeof(uint16_t) * (size_t)it.ns) : (size_t)-1;
    max_n = std::max(max_n, it.nc + it.ns); alg += 96.0 * (it.nc + it.ns);
  }
  const size_t total = off;
  CK(ctx->h_stage.reserve(total));
  CK(ctx->d_stage.reserve(total));
  char* h = (char*)ctx->h_stage.p; char* d = (char*)ctx->d_stage.p;
  RegDesc* hd = (RegDesc*)(h + o_desc);
  memcpy(h + o_pose, pose6xB, sizeof(float) * 6 * (size_t)B);
  for (int b = 0; b < B; b++) {
    const lisreg_batch_item& it = items[b];
    if (it.nc) memcpy(h + oc[b], it.corner, sizeof(float4) * (size_t)it.nc);
    if (it.ns) memcpy(h + os[b], it.surf, sizeof(float4) * (size_t)it.ns);
    if (it.clabel && it.nc) memcpy(h + ocl[b], it.clabel, sizeof(uint16_t) * (size_t)it.nc);
    if (it.slabel && it.ns) memcpy(h + osl[b], it.slabel, sizeof(uint16_t) * (size_t)it.ns);
    hd[b].corner = (const float4*)(d + oc[b]); hd[b].surf = (const float4*)(d + os[b]);
    hd[b].clabel = it.clabel ? (const uint16_t*)(d + ocl[b]) : nullptr;
    hd[b].slabel = it.slabel ? (const uint16_t*)(d + osl[b]) : nullptr;
    hd[b].nc = it.nc; hd[b].ns = it.ns; hd[b].map_slot = it.map_id; hd[b].pad = 0; hd[b].nc_ptr = nullptr; hd[b].ns_ptr = nullptr;
  }
  CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, st));
  CK(ctx->d_res.reserve(sizeof(lisreg_lm_result) * (size_t)B));
  lisreg_lm_iter* d_logs = nullptr;
  const bool want_log = iter_log && prm->want_iter_log;
  if (want_log) {
    CK(ctx->d_logs.reserve(sizeof(lisreg_lm_iter) * (size_t)B * LISREG_MAX_ITERS));
    d_logs = (lisreg_lm_iter*)ctx->d_logs.p;
    CK(cudaMemsetAsync(d_logs, 0, sizeof(lisreg_lm_iter) * (size_t)B * LISREG_MAX_ITERS, st));
  }
  int rc = run_lm(ctx, B, (const RegDesc*)(d + o_desc), max_n, alg, (float*)(d + o_pose), prm, (lisreg_lm_result*)ctx->d_res.p, d_logs);
  if (rc) return rc;
  CK(ctx->h_out.reserve(sizeof(lisreg_lm_result) * (size_t)B + (want_log ? sizeof(lisreg_lm_iter) * (size_t)B * LISREG_MAX_ITERS : 0)));
  CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_res.p, sizeof(lisreg_lm_result) * (size_t)B, cudaMemcpyDeviceToHost, st));
  if (want_log)
    CK(cudaMemcpyAsync((char*)ctx->h_out.p + sizeof(lisreg_lm_result) * (size_t)B, d_logs,
                       sizeof(lisreg_lm_iter) * (size_t)B * LISREG_MAX_ITERS, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const lisreg_lm_result* hr = (const lisreg_lm_result*)ctx->h_out.p;
  int worst = LISREG_OK;
  for (int b = 0; b < B; b++) {
    resxB[b] = hr[b];
    if (hr[b].status != LISREG_NOT_ENOUGH_FEATURES) memcpy(pose6xB + 6 * (size_t)b, hr[b].pose, sizeof(float) * 6);
    worst = std::max(worst, hr[b].status);
  }
  if (want_log) {
    const lisreg_lm_iter* hl = (const lisreg_lm_iter*)((const char*)ctx->h_out.p + sizeof(lisreg_lm_result) * (size_t)B);
    for (int b = 0; b < B; b++)
      memcpy(iter_log + (size_t)b * prm->max_iters, hl + (size_t)b * LISREG_MAX_ITERS, sizeof(lisreg_lm_iter) * (size_t)prm->max_iters);
  }
  return worst;
}

int32_t lisreg_scan2map_batch_arena(lisreg_ctx* ctx, int32_t B, const lisreg_batch_item* items, const void* host_arena,
                                    uint64_t arena_bytes, float* pose6xB, const lisreg_lm_params* prm, lisreg_lm_result* resxB) {
  if (!ctx || B <= 0 || !items || !host_arena || !pose6xB || !prm || !resxB) return fail(ctx, LISREG_ERR_ARG, "lisreg_scan2map_batch_arena: bad argument");
  if (prm->max_iters <= 0 || prm->max_iters > LISREG_MAX_ITERS) return fail(ctx, LISREG_ERR_ARG, "max_iters must be in 1..%d", LISREG_MAX_ITERS);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t head = ((sizeof(RegDesc) * (size_t)B + 15) & ~size_t(15)) + ((sizeof(float) * 6 * (size_t)B + 15) & ~size_t(15));
  CK(ctx->h_stage.reserve(head));
  CK(ctx->d_stage.reserve(head + (size_t)arena_bytes + 16));
  char* h = (char*)ctx->h_stage.p; char* d = (char*)ctx->d_stage.p;
  char* d_arena = d + head;
  RegDesc* hd = (RegDesc*)h;
  float* hp = (float*)(h + ((sizeof(RegDesc) * (size_t)B + 15) & ~size_t(15)));
  memcpy(hp, pose6xB, sizeof(float) * 6 * (size_t)B);
  int max_n = 0; double alg = 0;
  for (int b = 0; b < B; b++) {
    const lisreg_batch_item& it = items[b];
    const size_t oc = (size_t)it.corner, os = (size_t)it.surf, ocl = (size_t)it.clabel, osl = (size_t)it.slabel;
    if (it.nc < 0 || it.ns < 0 || it.map_id < 0 || it.map_id >= (int)ctx->maps.size() || !ctx->maps[it.map_id].used ||
        (oc & 15) || (os & 15) || !in_arena(oc, 16ull * it.nc, arena_bytes) || !in_arena(os, 16ull * it.ns, arena_bytes) ||
        (ocl != (size_t)-1 && !in_arena(ocl, 2ull * it.nc, arena_bytes)) || (osl != (size_t)-1 && !in_arena(osl, 2ull * it.ns, arena_bytes)))
      return fail(ctx, LISREG_ERR_ARG, "arena item %d: bad sizes, offsets or map id", b);
    hd[b].corner = (const float4*)(d_arena + oc); hd[b].surf = (const float4*)(d_arena + os);
    hd[b].clabel = ocl != (size_t)-1 ? (const uint16_t*)(d_arena + ocl) : nullptr;
    hd[b].slabel = osl != (size_t)-1 ? (const uint16_t*)(d_arena + osl) : nullptr;
    hd[b].nc = it.nc; hd[b].ns = it.ns; hd[b].map_slot = it.map_id; hd[b].pad = 0; hd[b].nc_ptr = nullptr; hd[b].ns_ptr = nullptr;
    max_n = std::max(max_n, it.nc + it.ns); alg += 96.0 * (it.nc + it.ns);
  }
  CK(cudaMemcpyAsync(d, h, head, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_arena, host_arena, (size_t)arena_bytes, cudaMemcpyHostToDevice, st));
  CK(ctx->d_res.reserve(sizeof(lisreg_lm_result) * (size_t)B));
  float* d_pose = (float*)(d + ((sizeof(RegDesc) * (size_t)B + 15) & ~size_t(15)));
  int rc = run_lm(ctx, B, (const RegDesc*)d, max_n, alg, d_pose, prm, (lisreg_lm_result*)ctx->d_res.p, nullptr);
  if (rc) return rc;
  CK(ctx->h_out.reserve(sizeof(lisreg_lm_result) * (size_t)B));
  CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_res.p, sizeof(lisreg_lm_result) * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const lisreg_lm_result* hr = (const lisreg_lm_result*)ctx->h_out.p;
  int worst = LISREG_OK;
  for (int b = 0; b < B; b++) {
    resxB[b] = hr[b];
    memcpy(pose6xB + 6 * (size_t)b, hr[b].pose, sizeof(float) * 6);
    worst = std::max(worst, hr[b].status);
  }
  return worst;
}

// ------------------------------------------------------------------------------------------------
// feature extraction
// ------------------------------------------------------------------------------------------------
void lisreg_cloud_layout_preset(lisreg_cloud_layout* l, int32_t which) {
  memset(l, 0, sizeof(*l));
  l->off_intensity = -1; l->off_ring = -1; l->off_time = -1;
  switch (which) {
    case 1: l->point_step = 12; l->off_x = 0; l->off_y = 4; l->off_z = 8; break;                                   // xyz + ring array
    case 2: l->point_step = 32; l->off_x = 0; l->off_y = 4; l->off_z = 8; l->off_intensity = 16; l->off_ring = 20; l->off_time = 24; break;   // PointXYZIRT
    case 3: l->point_step = 12; l->off_x = 0; l->off_y = 4; l->off_z = 8; l->off_ring = -2; break;                  // xyz only, ring synthesised
    default: memset(l, 0, sizeof(*l)); break;                                                                       // packed float4 + ring array
  }
}

// bytes per record and validity of a layout
static inline int layout_step(const lisreg_cloud_layout& l) { return l.point_step == 0 ? 16 : l.point_step; }
static bool layout_ok(const lisreg_cloud_layout& l, int n_scan) {
  if (l.point_step == 0) return true;
  if (l.point_step < 12 || (l.point_step & 3)) return false;
  for (int o : {l.off_x, l.off_y, l.off_z}) if (o < 0 || (o & 3) || o + 4 > l.point_step) return false;
  if (l.off_intensity >= 0 && ((l.off_intensity & 3) || l.off_intensity + 4 > l.point_step)) return false;
  if (l.off_ring >= 0 && ((l.off_ring & 1) || l.off_ring + 2 > l.point_step)) return false;
  if (l.off_ring < -2) return false;
  if (l.off_ring == -2 && !(n_scan == 16 || n_scan == 32 || n_scan == 64)) return false;   // the formulas upstream knows
  if (l.off_time >= 0 && ((l.off_time & 3) || l.off_time + 4 > l.point_step)) return false;
  return true;
}
static inline bool layout_needs_ring_array(const lisreg_cloud_layout& l) { return l.point_step == 0 || l.off_ring == -1; }

void lisreg_feat_params_default(lisreg_feat_params* p) {
  memset(p, 0, sizeof(*p));
  p->n_scan = 64; p->horizon = 1800; p->downsample_rate = 1;   // benchmark pins downsampleRate = 1
  p->min_range = 0.0f; p->max_range = 70.0f;                    // config/params.yaml:73-74
  p->edge_thr = 1.0f; p->surf_thr = 0.1f;                       // config/params.yaml:117-118
}

static size_t feat_frame_bytes(int cells, int nscan) {
  // owner, ext_src, col, picked, label, surf_idx (int), range, curv (float) : 8 x 4 B per cell; ext_pts 16 B per cell
  size_t b = (size_t)cells * (8 * 4 + 16);
  b += sizeof(int) * (size_t)nscan * 3;                 // ring_count, ring_start, ring_end
  b += sizeof(int) * (size_t)nscan * 6 * (20 + 1 + 10 + 1 + 1 + 1 + 1);   // seg_corner, ncorner, seg_flat, nflat, valid, sp, ep
  b += sizeof(int) * (size_t)nscan * (120 + 24 + 60);   // corner_idx, sharp_idx, flat_idx
  b += sizeof(int) * 8 + 64;                            // M, counts[4], start_inv[9]
  return (b + 255) & ~size_t(255);
}

// carves frame i's work area; raw input pointers are filled by the caller
static void feat_carve(char* base, int cells, int nscan, FeatFrame* f) {
  char* p = base;
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 15) & ~size_t(15); return r; };
  f->ext_pts = (float4*)take(sizeof(float4) * (size_t)cells);
  f->owner = (int*)take(4 * (size_t)cells); f->ext_src = (int*)take(4 * (size_t)cells); f->col = (unsigned short*)take(4 * (size_t)cells);
  f->range = (float*)take(4 * (size_t)cells); f->curv = (float*)take(4 * (size_t)cells);
  f->picked = (unsigned char*)take(4 * (size_t)cells); f->label = (signed char*)take(4 * (size_t)cells); f->surf_idx = (int*)take(4 * (size_t)cells);
  f->ring_count = (int*)take(4 * (size_t)nscan); f->ring_start = (int*)take(4 * (size_t)nscan); f->ring_end = (int*)take(4 * (size_t)nscan);
  const size_t ns = (size_t)nscan * 6;
  f->seg_corner = (int*)take(4 * ns * 20); f->seg_ncorner = (int*)take(4 * ns);
  f->seg_flat = (int*)take(4 * ns * 10); f->seg_nflat = (int*)take(4 * ns);
  f->seg_valid = (int*)take(4 * ns); f->seg_sp = (int*)take(4 * ns); f->seg_ep = (int*)take(4 * ns);
  f->corner_idx = (int*)take(4 * (size_t)nscan * 120); f->sharp_idx = (int*)take(4 * (size_t)nscan * 24); f->flat_idx = (int*)take(4 * (size_t)nscan * 60);
  f->M = (int*)take(4); f->counts = (int*)take(16); f->start_inv = (float*)take(36);
  f->time = nullptr; f->imu_time = nullptr; f->imu_rot = nullptr; f->n_imu = 0; f->t_scan = 0.0;
}

static int feat_reserve(lisreg_ctx* ctx, int F, int cells, int nscan) {
  const size_t per = feat_frame_bytes(cells, nscan) + 4096;
  CK(ctx->cur->d_feat.reserve(per * (size_t)F));
  CK(ctx->cur->d_feat_frames.reserve(sizeof(FeatFrame) * (size_t)F));
  {
    const void* before = ctx->cur->d_feat_ctl.p;
    CK(ctx->cur->d_feat_ctl.reserve(sizeof(int) * (2 + (size_t)F * (size_t)nscan)));
    if (ctx->cur->d_feat_ctl.p != before) CK(cudaMemsetAsync(ctx->cur->d_feat_ctl.p, 0, ctx->cur->d_feat_ctl.cap, ctx->cur->stream));
  }
  ctx->cur->feat_cap_frames = F; ctx->cur->feat_cells = cells; ctx->cur->feat_nscan = nscan;
  return LISREG_OK;
}

// runs F1-F5 for F frames whose FeatFrame descriptors (device) are ready
// lean: the caller consumes the feature lists only (not the per-point curvature / label arrays of lisreg_extract_features)
static int run_features(lisreg_ctx* ctx, FeatFrame* d_frames, int F, const lisreg_feat_params* prm, int max_n, double alg_bytes,
                        bool with_deskew = false, bool lean = false) {
  cudaStream_t st = ctx->cur->stream;
  // (a single frame keeps the separate 4-us k_feat_curv_occl launch: inside the selection kernel the same work sits on the
  // critical path of 64 lone warps)
  const bool seg_fused = lean && !ctx->feat_seg_unfused && F >= 16;
  FeatParamsDev dp{prm->n_scan, prm->horizon, prm->downsample_rate, prm->min_range, prm->max_range, prm->edge_thr, prm->surf_thr, prm->layout, seg_fused ? 1 : 0};
  const int cells = prm->n_scan * prm->horizon;
  ProfScope ps(ctx, PROF_FEAT, alg_bytes, 7);
  // F1 + F2 on chip (k_feat_front, opt-in) when the ring ids come as their own array and no de-skew reference has to be found
  // across the whole sweep first; otherwise (default) the range image goes through global memory
  const bool fused = !with_deskew && ctx->feat_fused && layout_needs_ring_array(prm->layout) &&
                     ctx->cur->d_feat_ctl.cap >= sizeof(int) * (2 + (size_t)F * (size_t)prm->n_scan);
  if (fused) {
    const int nchunk = (prm->horizon + 31) / 32;
    const int rmax = std::max(1, std::min(prm->n_scan, FEAT_FRONT_SMEM_INTS / (prm->horizon + nchunk)));
    int R = rmax;                                                 // rings per block: as many as fit, fewer when the batch is small
    while (R > 1 && (long long)F * ((prm->n_scan + R - 1) / R) < 2 * 148) R--;
    const int G = (prm->n_scan + R - 1) / R;
    const size_t smem = sizeof(int) * ((size_t)R * (size_t)(prm->horizon + nchunk) + 1);
    if (!ctx->attr_front_set) { CK(cudaFuncSetAttribute(k_feat_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * (FEAT_FRONT_SMEM_INTS + 1)))); ctx->attr_front_set = true; }
    k_feat_front<<<F * G, FEAT_FRONT_THREADS, smem, st>>>(d_frames, dp, R, G, F, (int*)ctx->cur->d_feat_ctl.p); LAUNCH_CK();
  } else {
    k_feat_clear<<<dim3((cells + 255) / 256, F), 256, 0, st>>>(d_frames, dp); LAUNCH_CK();
    k_feat_project<<<dim3(std::max(1, (max_n + 255) / 256), F), 256, 0, st>>>(d_frames, dp); LAUNCH_CK();
    if (with_deskew) { k_feat_deskew_start<<<F, 256, 0, st>>>(d_frames, dp); LAUNCH_CK(); }
    k_feat_ring_count<<<dim3(prm->n_scan, F), 256, 0, st>>>(d_frames, dp); LAUNCH_CK();
    if (with_deskew) { k_feat_compact<true><<<dim3(prm->n_scan, F), 256, 0, st>>>(d_frames, dp); LAUNCH_CK(); }
    else { k_feat_compact<false><<<dim3(prm->n_scan, F), 256, 0, st>>>(d_frames, dp); LAUNCH_CK(); }
  }
  if (seg_fused) {
    // pipelines that only consume the feature lists: smoothness and occlusion marks are computed inside the selection kernel
    k_feat_segments<true><<<dim3((prm->n_scan + FEAT_WARPS - 1) / FEAT_WARPS, F), 32 * FEAT_WARPS, 0, st>>>(d_frames, dp); LAUNCH_CK();
  } else {
    k_feat_curv_occl<<<dim3((cells + 255) / 256, F), 256, 0, st>>>(d_frames); LAUNCH_CK();
    k_feat_segments<false><<<dim3((prm->n_scan + FEAT_WARPS - 1) / FEAT_WARPS, F), 32 * FEAT_WARPS, 0, st>>>(d_frames, dp); LAUNCH_CK();
  }
  k_feat_gather<<<dim3(FEAT_GATHER_SPLIT, F), 256, 0, st>>>(d_frames, dp); LAUNCH_CK();
  return LISREG_OK;
}

static int extract_features_impl(lisreg_ctx* ctx, const float* pts, const uint16_t* ring, const float* time, int32_t n,
                                 const lisreg_feat_params* prm, const lisreg_deskew* dsk, lisreg_feat_out* out, float* ext_xyzi) {
  if (!ctx || n < 0 || (n > 0 && !pts) || !prm || !out) return fail(ctx, LISREG_ERR_ARG, "lisreg_extract_features: bad argument");
  if (prm->n_scan <= 0 || prm->horizon <= 0 || prm->horizon > 2048 || prm->n_scan * 6 > 1024 || prm->downsample_rate <= 0)
    return fail(ctx, LISREG_ERR_ARG, "lisreg_extract_features: unsupported n_scan/horizon/downsample_rate");
  if (!layout_ok(prm->layout, prm->n_scan)) return fail(ctx, LISREG_ERR_ARG, "lisreg_extract_features: bad cloud layout");
  if (n > 0 && layout_needs_ring_array(prm->layout) && !ring) return fail(ctx, LISREG_ERR_ARG, "lisreg_extract_features: ring array missing");
  const bool deskew = dsk && dsk->n_imu > 0;
  const bool time_in_rec = prm->layout.point_step != 0 && prm->layout.off_time >= 0;
  if (deskew && ((!time && !time_in_rec) || !dsk->imu_time || !dsk->imu_rot)) return fail(ctx, LISREG_ERR_ARG, "lisreg_extract_features_deskew: time / IMU table missing");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int cells = prm->n_scan * prm->horizon;
  int rc = feat_reserve(ctx, 1, cells, prm->n_scan);
  if (rc) return rc;
  const bool ring_arr = layout_needs_ring_array(prm->layout);
  const size_t bp = ((size_t)layout_step(prm->layout) * (size_t)n + 15) & ~size_t(15), br = (sizeof(uint16_t) * (size_t)n + 15) & ~size_t(15);
  const size_t bt = deskew ? sizeof(float) * (size_t)n : 0, bi = deskew ? sizeof(double) * 4 * (size_t)dsk->n_imu : 0;
  CK(ctx->d_stage.reserve(bp + br + ((bt + 15) & ~size_t(15)) + bi + 64));
  char* d = (char*)ctx->d_stage.p;
  char* d_time = d + bp + br; char* d_imu = d_time + ((bt + 15) & ~size_t(15));
  if (n) {
    CK(cudaMemcpyAsync(d, pts, (size_t)layout_step(prm->layout) * (size_t)n, cudaMemcpyHostToDevice, st));
    if (ring_arr) CK(cudaMemcpyAsync(d + bp, ring, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, st));
  }
  if (deskew) {
    if (n && time) CK(cudaMemcpyAsync(d_time, time, bt, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_imu, dsk->imu_time, sizeof(double) * (size_t)dsk->n_imu, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_imu + sizeof(double) * (size_t)dsk->n_imu, dsk->imu_rot, sizeof(double) * 3 * (size_t)dsk->n_imu, cudaMemcpyHostToDevice, st));
  }
  FeatFrame f{};
  feat_carve((char*)ctx->cur->d_feat.p, cells, prm->n_scan, &f);
  f.pts = (const float4*)d; f.ring = (const uint16_t*)(d + bp); f.n = n;
  if (deskew) {
    f.time = (const float*)d_time; f.imu_time = (const double*)d_imu; f.imu_rot = (const double*)(d_imu + sizeof(double) * (size_t)dsk->n_imu);
    f.n_imu = dsk->n_imu; f.t_scan = dsk->time_scan_cur;
  }
  CK(cudaMemcpyAsync(ctx->cur->d_feat_frames.p, &f, sizeof(f), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));   // f is a stack object
  rc = run_features(ctx, (FeatFrame*)ctx->cur->d_feat_frames.p, 1, prm, n, 17.0 * n, deskew);
  if (rc) return rc;
  int h[5];
  CK(cudaMemcpyAsync(&h[0], f.M, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&h[1], f.counts, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  out->n_extracted = h[0]; out->n_corner = h[1]; out->n_sharp = h[2]; out->n_flat = h[3]; out->n_surf = h[4];
  auto dl = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    return (dst && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
  const size_t M = (size_t)h[0];
  // column indices and labels live as 16-bit / 8-bit arrays on the device: widened to the int32 of the interface here
  std::vector<unsigned short> h_col(out->col_ind ? M : 0); std::vector<signed char> h_label(out->label ? M : 0);
  CK(dl(out->src_index, f.ext_src, 4 * M)); CK(dl(h_col.data(), f.col, 2 * h_col.size())); CK(dl(out->range, f.range, 4 * M));
  CK(dl(out->start_ring, f.ring_start, 4 * (size_t)prm->n_scan)); CK(dl(out->end_ring, f.ring_end, 4 * (size_t)prm->n_scan));
  CK(dl(out->corner_idx, f.corner_idx, 4 * (size_t)h[1])); CK(dl(out->sharp_idx, f.sharp_idx, 4 * (size_t)h[2]));
  CK(dl(out->flat_idx, f.flat_idx, 4 * (size_t)h[3])); CK(dl(out->surf_idx, f.surf_idx, 4 * (size_t)h[4]));
  CK(dl(out->curvature, f.curv, 4 * M)); CK(dl(h_label.data(), f.label, h_label.size()));
  CK(dl(ext_xyzi, f.ext_pts, 16 * M));
  CK(cudaStreamSynchronize(st));
  for (size_t i = 0; i < h_col.size(); i++) out->col_ind[i] = h_col[i];
  for (size_t i = 0; i < h_label.size(); i++) out->label[i] = h_label[i];
  return LISREG_OK;
}

int32_t lisreg_extract_features(lisreg_ctx* ctx, const float* pts, const uint16_t* ring, int32_t n,
                                const lisreg_feat_params* prm, lisreg_feat_out* out) {
  return extract_features_impl(ctx, pts, ring, nullptr, n, prm, nullptr, out, nullptr);
}

int32_t lisreg_extract_features_deskew(lisreg_ctx* ctx, const float* pts, const uint16_t* ring, const float* time, int32_t n,
                                       const lisreg_feat_params* prm, const lisreg_deskew* dsk, lisreg_feat_out* out, float* ext_xyzi) {
  return extract_features_impl(ctx, pts, ring, time, n, prm, dsk, out, ext_xyzi);
}

// ------------------------------------------------------------------------------------------------
// voxel grid
// ------------------------------------------------------------------------------------------------
static size_t vox_seg_bytes(int cap) {
  const int nblk = (cap + RS_TILE - 1) / RS_TILE + 1;
  size_t b = 4 * (size_t)cap * 4;                 // key_a, val_a, key_b, val_b
  b += 4 * 256 * (size_t)(nblk + 1);              // hist + the 256 digit bases behind it
  b += 2 * 4 * ((size_t)cap + 1) + 32;            // seg_start, run_start
  b += sizeof(VoxPlan) + 16 + 32;                 // plan, out_n, bbox
  b += sizeof(float4) * (size_t)cap;              // out
  return (b + 1024) & ~size_t(255);
}
static void vox_carve(char* base, int cap, VoxSeg* s) {
  char* p = base;
  auto take = [&](size_t bytes) { char* r = p; p += (bytes + 15) & ~size_t(15); return r; };
  const int nblk = (cap + RS_TILE - 1) / RS_TILE + 1;
  s->out = (float4*)take(sizeof(float4) * (size_t)cap);
  s->key_a = (uint32_t*)take(4 * (size_t)cap); s->val_a = (uint32_t*)take(4 * (size_t)cap);
  s->key_b = (uint32_t*)take(4 * (size_t)cap); s->val_b = (uint32_t*)take(4 * (size_t)cap);
  s->hist = (uint32_t*)take(4 * 256 * (size_t)(nblk + 1));
  s->seg_start = (int*)take(4 * ((size_t)cap + 1));
  s->run_start = (int*)take(4 * ((size_t)cap + 1));
  s->plan = (VoxPlan*)take(sizeof(VoxPlan));
  s->bbox = (unsigned*)take(24);
  s->out_n = (int*)take(4);
  s->cap = cap;
  s->bound = 0.f; s->gather_increasing = 0;
}

// runs the voxel grid for nseg clouds whose VoxSeg descriptors (device) are ready; max_n bounds every n
static int run_voxel(lisreg_ctx* ctx, VoxSeg* d_segs, int nseg, int max_n, double alg_bytes) {
  cudaStream_t st = ctx->cur->stream;
  if (nseg <= 0) return LISREG_OK;
  const int nblk = std::max(1, (max_n + RS_TILE - 1) / RS_TILE);
  const int pblk = std::max(1, std::min(64, (max_n + 1023) / 1024));
  ProfScope ps(ctx, PROF_VOXEL, alg_bytes, 16);
  // single-block scans are the cheaper choice when there are many clouds to keep the GPU busy; with a handful of clouds
  // (streaming odometry: one frame's two clouds, the 2 M-point window map) they serialise, so those take the parallel forms
  const bool big = nseg < 64 && !ctx->vox_batch_form;
  // a handful of clouds of full HDL-64 size (the streaming odometry's single frame) finish sooner spread over many blocks
  // than on two SMs (measured: 0.57 vs 0.63 ms per frame), smaller ones (VLP-16) the other way round
  if (max_n <= (big ? VOX_BLOCK_MAX_N_FEW : VOX_BLOCK_MAX_N) && !ctx->vox_unfused) {
    // clouds of frame size: bounding box -> keys -> runs -> sort -> voxel heads by ONE block per cloud, the sort in shared memory
    if (!ctx->attr_vox_set) {                       // per context: the attribute belongs to the (function, device) pair
      CK(cudaFuncSetAttribute(k_vox_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VB_SMEM));
      CK(cudaFuncSetAttribute(k_vox_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VB_SMEM));
      ctx->attr_vox_set = true;
    }
    if (big) {
      k_vox_block<false><<<nseg, VB_THREADS, VB_SMEM, st>>>(d_segs); LAUNCH_CK();
      k_vox_centroid_warp<<<dim3(std::max(1, std::min(2 * ctx->n_sm, (max_n + 255) / 256)), nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
    } else {
      k_vox_block<true><<<nseg, VB_THREADS, VB_SMEM, st>>>(d_segs); LAUNCH_CK();
    }
    return LISREG_OK;
  }
  k_vox_bbox_init<<<(nseg * 6 + 255) / 256, 256, 0, st>>>(d_segs, nseg); LAUNCH_CK();
  k_vox_bbox<<<dim3(pblk, nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
  k_vox_plan<<<(nseg + 127) / 128, 128, 0, st>>>(d_segs, nseg); LAUNCH_CK();
  k_vox_keys<<<dim3(pblk, nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
  // runs of equal consecutive keys -> the entries the sort moves
  if (big) {
    k_vox_head_count<true><<<dim3(nblk, nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
    k_vox_head_scan<true><<<(nseg + 7) / 8, 256, 0, st>>>(d_segs, nseg); LAUNCH_CK();
    k_vox_head_write<true><<<dim3(nblk, nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
  } else {
    k_vox_heads<true><<<nseg, 1024, 0, st>>>(d_segs); LAUNCH_CK();
  }
  for (int pass = 0; pass < 4; pass++) {                  // passes beyond a cloud's plan->npass return at once
    k_rs_hist<<<dim3(nblk, nseg), RS_THREADS, 0, st>>>(d_segs, 8 * pass, pass & 1); LAUNCH_CK();
    if (big) {      // a few large clouds (sliding-window map, submap classes): scan with 256 warps per cloud
      k_rs_scan_digit<<<dim3(32, nseg), 256, 0, st>>>(d_segs, 8 * pass); LAUNCH_CK();
      k_rs_scan_base<<<(nseg + 7) / 8, 256, 0, st>>>(d_segs, nseg, 8 * pass); LAUNCH_CK();
    } else {        // many small clouds (batched frames): one block per cloud is the cheaper launch
      k_rs_scan<<<nseg, 1024, 0, st>>>(d_segs, 8 * pass); LAUNCH_CK();
    }
    k_rs_scatter<<<dim3(nblk, nseg), RS_THREADS, 0, st>>>(d_segs, 8 * pass, pass & 1); LAUNCH_CK();
  }
  if (big) {
    k_vox_head_count<false><<<dim3(nblk, nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
    k_vox_head_scan<false><<<(nseg + 7) / 8, 256, 0, st>>>(d_segs, nseg); LAUNCH_CK();
    k_vox_head_write<false><<<dim3(nblk, nseg), 256, 0, st>>>(d_segs); LAUNCH_CK();
  } else {
    k_vox_heads<false><<<nseg, 1024, 0, st>>>(d_segs); LAUNCH_CK();
  }
  if (big) { k_vox_centroid_warp<<<dim3(std::max(1, std::min(2 * ctx->n_sm, (max_n + 255) / 256)), nseg), 256, 0, st>>>(d_segs); LAUNCH_CK(); }
  else { k_vox_centroid<<<dim3(std::max(1, (max_n + VC_CHUNK - 1) / VC_CHUNK), nseg), 256, 0, st>>>(d_segs); LAUNCH_CK(); }
  return LISREG_OK;
}

int32_t lisreg_voxel_grid(lisreg_ctx* ctx, const float* pts, int32_t n, float leaf, float* out, int32_t* m) {
  if (!ctx || n < 0 || (n > 0 && (!pts || !out)) || !m || !(leaf > 0.f)) return fail(ctx, LISREG_ERR_ARG, "lisreg_voxel_grid: bad argument");
  *m = 0;
  if (n == 0) return LISREG_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CK(ctx->d_stage.reserve(sizeof(float4) * (size_t)n));
  CK(ctx->cur->d_vox.reserve(vox_seg_bytes(n)));
  CK(ctx->cur->d_vox_segs.reserve(sizeof(VoxSeg)));
  CK(cudaMemcpyAsync(ctx->d_stage.p, pts, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  VoxSeg s{};
  vox_carve((char*)ctx->cur->d_vox.p, n, &s);
  s.src = (const float4*)ctx->d_stage.p; s.gather = nullptr; s.n_ptr = nullptr; s.n = n; s.leaf = leaf;
  CK(cudaMemcpyAsync(ctx->cur->d_vox_segs.p, &s, sizeof(s), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  int rc = run_voxel(ctx, (VoxSeg*)ctx->cur->d_vox_segs.p, 1, n, 32.0 * n);
  if (rc) return rc;
  int cnt = 0;
  CK(cudaMemcpyAsync(&cnt, s.out_n, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaMemcpyAsync(out, s.out, sizeof(float4) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *m = cnt;
  return LISREG_OK;
}

// ------------------------------------------------------------------------------------------------
// whole-frame pipeline
// ------------------------------------------------------------------------------------------------
void lisreg_frame_params_default(lisreg_frame_params* p) {
  lisreg_feat_params_default(&p->feat);
  p->corner_leaf = 0.2f; p->surf_leaf = 0.4f;       // config/params.yaml:132-133
  lisreg_lm_params_preset(&p->lm, 'A');
  p->deskew = nullptr;
}

// d_pts_base/d_ring_base: if arena != nullptr the item pointers are byte offsets into it
// h_pinned: optional pinned staging (frame_desc_bytes(F) bytes, private to this call slice) for the descriptor
// upload; NULL = pageable vectors (cudaMemcpyAsync returns once a small pageable source has been consumed).
static size_t frame_desc_bytes(int F) {
  return ((sizeof(FeatFrame) + 2 * sizeof(VoxSeg) + sizeof(RegDesc)) * (size_t)F + 255) & ~size_t(255);
}
// launch_n > 0: size every launch / buffer for sweeps of up to launch_n points whatever the items say (the per-frame CUDA
// graph of the streaming odometry is captured once and must fit every later frame; all kernels bound themselves by the
// device-resident counts, so the results do not depend on it)
static int run_frames(lisreg_ctx* ctx, int F, const lisreg_frame_item* items, const char* d_arena, uint64_t arena_bytes,
                      float* d_pose, const lisreg_frame_params* prm, lisreg_lm_result* d_res, char* h_pinned = nullptr,
                      int tiling_B = 0, int launch_n = 0, int frame_offset = 0, int lm_it1 = -1) {
  cudaStream_t st = ctx->cur->stream;
  const lisreg_feat_params* fp = &prm->feat;
  if (fp->n_scan <= 0 || fp->horizon <= 0 || fp->horizon > 2048 || fp->n_scan * 6 > 1024 || fp->downsample_rate <= 0)
    return fail(ctx, LISREG_ERR_ARG, "frame pipeline: unsupported n_scan/horizon/downsample_rate");
  if (!(prm->corner_leaf > 0.f) || !(prm->surf_leaf > 0.f)) return fail(ctx, LISREG_ERR_ARG, "frame pipeline: leaf sizes must be > 0");
  if (!layout_ok(fp->layout, fp->n_scan)) return fail(ctx, LISREG_ERR_ARG, "frame pipeline: bad cloud layout");
  // optional motion de-skew (projectPointCloud's deskewPoint, laserProcessing.cpp:427-462, :501): one IMU rotation table per
  // frame, the per-point time read from the records
  const lisreg_deskew* dsk = prm->deskew ? prm->deskew + frame_offset : nullptr;
  size_t imu_doubles = 0;
  if (dsk) {
    if (!(fp->layout.point_step != 0 && fp->layout.off_time >= 0))
      return fail(ctx, LISREG_ERR_ARG, "frame pipeline: de-skew needs the per-point time inside the records (layout.off_time)");
    for (int i = 0; i < F; i++) {
      if (dsk[i].n_imu < 0 || (dsk[i].n_imu > 0 && (!dsk[i].imu_time || !dsk[i].imu_rot))) return fail(ctx, LISREG_ERR_ARG, "frame %d: bad IMU table", i);
      imu_doubles += 4 * (size_t)dsk[i].n_imu;
    }
    CK(ctx->cur->d_imu.reserve(sizeof(double) * std::max<size_t>(imu_doubles, 1)));
  }
  std::vector<double> h_imu(imu_doubles);
  size_t imu_off = 0;
  const uint64_t rec = (uint64_t)layout_step(fp->layout);
  const bool ring_arr = layout_needs_ring_array(fp->layout);
  const size_t pts_align = fp->layout.point_step == 0 ? 15 : 3;
  const int cells = fp->n_scan * fp->horizon;
  const int ccap = fp->n_scan * 120;
  int rc = feat_reserve(ctx, F, cells, fp->n_scan);
  if (rc) return rc;
  const size_t feat_per = feat_frame_bytes(cells, fp->n_scan) + 4096;
  const size_t vc_per = vox_seg_bytes(ccap), vs_per = vox_seg_bytes(cells);
  CK(ctx->cur->d_vox.reserve((vc_per + vs_per) * (size_t)F));
  CK(ctx->cur->d_vox_segs.reserve(sizeof(VoxSeg) * 2 * (size_t)F));
  CK(ctx->cur->d_descs.reserve(sizeof(RegDesc) * (size_t)F));
  std::vector<FeatFrame> vf; std::vector<VoxSeg> vv; std::vector<RegDesc> vd;
  FeatFrame* hf; VoxSeg* hv; RegDesc* hd;
  if (h_pinned) {
    hf = (FeatFrame*)h_pinned; hv = (VoxSeg*)(hf + F); hd = (RegDesc*)(hv + 2 * (size_t)F);
  } else {
    vf.resize((size_t)F); vv.resize(2 * (size_t)F); vd.resize((size_t)F);
    hf = vf.data(); hv = vv.data(); hd = vd.data();
  }
  int max_n = 0; double feat_bytes = 0;
  for (int i = 0; i < F; i++) {
    const lisreg_frame_item& it = items[i];
    if (it.n < 0 || it.map_id < 0 || it.map_id >= (int)ctx->maps.size() || !ctx->maps[it.map_id].used)
      return fail(ctx, LISREG_ERR_ARG, "frame item %d: bad size or map id", i);
    const float4* pts; const uint16_t* ring;
    if (d_arena) {
      const size_t op = (size_t)it.pts, orr = (size_t)it.ring;
      if ((op & pts_align) || !in_arena(op, rec * it.n, arena_bytes) || (ring_arr && ((orr & 1) || !in_arena(orr, 2ull * it.n, arena_bytes))))
        return fail(ctx, LISREG_ERR_ARG, "frame item %d: bad arena offsets", i);
      pts = (const float4*)(d_arena + op); ring = ring_arr ? (const uint16_t*)(d_arena + orr) : nullptr;
    } else { pts = (const float4*)it.pts; ring = it.ring; }
    FeatFrame& f = hf[i];
    feat_carve((char*)ctx->cur->d_feat.p + feat_per * (size_t)i, cells, fp->n_scan, &f);
    f.pts = pts; f.ring = ring; f.n = it.n;
    if (dsk && dsk[i].n_imu > 0) {
      double* hd_t = h_imu.data() + imu_off; double* dd_t = (double*)ctx->cur->d_imu.p + imu_off;
      memcpy(hd_t, dsk[i].imu_time, sizeof(double) * (size_t)dsk[i].n_imu);
      memcpy(hd_t + dsk[i].n_imu, dsk[i].imu_rot, sizeof(double) * 3 * (size_t)dsk[i].n_imu);
      f.imu_time = dd_t; f.imu_rot = dd_t + dsk[i].n_imu; f.n_imu = dsk[i].n_imu; f.t_scan = dsk[i].time_scan_cur;
      imu_off += 4 * (size_t)dsk[i].n_imu;
    }
    VoxSeg& vc = hv[2 * i]; VoxSeg& vs = hv[2 * i + 1];
    char* vb = (char*)ctx->cur->d_vox.p + (vc_per + vs_per) * (size_t)i;
    vox_carve(vb, ccap, &vc); vox_carve(vb + vc_per, cells, &vs);
    vc.src = f.ext_pts; vc.gather = f.corner_idx; vc.n_ptr = f.counts + 0; vc.n = 0; vc.leaf = prm->corner_leaf;
    vs.src = f.ext_pts; vs.gather = f.surf_idx;   vs.n_ptr = f.counts + 3; vs.n = 0; vs.leaf = prm->surf_leaf;
    vs.gather_increasing = 1;                              // surf_idx ascends (k_feat_gather); corner_idx is in pick order
    vc.bound = vs.bound = fp->max_range + 1.0f;            // extracted points passed the range gate (the de-skew is a pure rotation)
    RegDesc& d = hd[i];
    d.corner = vc.out; d.surf = vs.out; d.clabel = nullptr; d.slabel = nullptr; d.nc = 0; d.ns = 0;
    d.map_slot = it.map_id; d.pad = 0; d.nc_ptr = vc.out_n; d.ns_ptr = vs.out_n;
    max_n = std::max(max_n, it.n); feat_bytes += 17.0 * it.n;
  }
  if (launch_n > 0) max_n = std::max(max_n, launch_n);
  // pageable sources: cudaMemcpyAsync returns once they are consumed
  CK(cudaMemcpyAsync(ctx->cur->d_feat_frames.p, hf, sizeof(FeatFrame) * (size_t)F, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->cur->d_vox_segs.p, hv, sizeof(VoxSeg) * 2 * (size_t)F, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->cur->d_descs.p, hd, sizeof(RegDesc) * (size_t)F, cudaMemcpyHostToDevice, st));
  if (imu_doubles) CK(cudaMemcpyAsync(ctx->cur->d_imu.p, h_imu.data(), sizeof(double) * imu_doubles, cudaMemcpyHostToDevice, st));   // pageable: consumed on return
  rc = run_features(ctx, (FeatFrame*)ctx->cur->d_feat_frames.p, F, fp, max_n, feat_bytes, dsk != nullptr, true);
  if (rc) return rc;
  rc = run_voxel(ctx, (VoxSeg*)ctx->cur->d_vox_segs.p, 2 * F, std::min(max_n, cells), 2.0 * feat_bytes);
  if (rc) return rc;
  // the per-frame query counts live on the device; size the LM grid for the largest possible tile count
  // of this batch (blocks beyond a frame's real tile count exit immediately).  Voxel output never exceeds
  // its input, and the input never exceeds the sweep size.
  const int lm_max_n = std::min(max_n, cells);
  return run_lm(ctx, F, (const RegDesc*)ctx->cur->d_descs.p, lm_max_n, 0.0, d_pose, &prm->lm, d_res, nullptr, tiling_B, 0, lm_it1);
}

int32_t lisreg_frames_batch_dev(lisreg_ctx* ctx, int32_t F, const lisreg_frame_item* items, float* d_pose6xF,
                                const lisreg_frame_params* prm, lisreg_lm_result* d_resxF) {
  if (!ctx || F <= 0 || !items || !d_pose6xF || !prm || !d_resxF) return fail(ctx, LISREG_ERR_ARG, "lisreg_frames_batch_dev: bad argument");
  if (prm->lm.max_iters <= 0 || prm->lm.max_iters > LISREG_MAX_ITERS) return fail(ctx, LISREG_ERR_ARG, "max_iters must be in 1..%d", LISREG_MAX_ITERS);
  CK(cudaSetDevice(ctx->device));
  // A large batch runs as up to four sub-batches on private streams: the stages of one frame pipeline are bound by different
  // things (selection loop: issue; ordered centroids and the 6x6 solves: latency; launch gaps), so kernels of different
  // sub-batches fill each other's idle SM time (measured: 6.65 -> 5.8 ms per 256 frames).  Every registration is reduced
  // in the same fixed order, so the results do not depend on the split.  Profiling keeps the batch whole (stage times
  // would overlap), and so does a submit / wait ticket in flight (it owns work sets 1 / 2).
  const int S = (ctx->dev_split > 1 && !ctx->prof_on && !ctx->slot[0].busy && !ctx->slot[1].busy) ? std::min(ctx->dev_split, F / 32) : 1;
  if (S <= 1) return run_frames(ctx, F, items, nullptr, 0, d_pose6xF, prm, d_resxF);
  cudaStream_t st = ctx->stream;
  for (int i = 1; i <= S; i++) if (!ctx->ws[i].stream) CK(cudaStreamCreateWithFlags(&ctx->ws[i].stream, cudaStreamNonBlocking));
  while ((int)ctx->chunk_ev.size() < S + 1) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
  const int C = (F + S - 1) / S;
  int rc = sync_maps(ctx);
  if (rc) return rc;
  CK(cudaEventRecord(ctx->chunk_ev[S], st));                       // fork: everything the caller queued before this call
  rc = LISREG_OK;
  for (int c = 0; c < S && rc == LISREG_OK; c++) {
    const int f0 = c * C, fc = std::min(C, F - f0);
    if (fc <= 0) break;
    ctx->cur = &ctx->ws[1 + c];
    cudaError_t e = cudaStreamWaitEvent(ctx->cur->stream, ctx->chunk_ev[S], 0);
    if (e != cudaSuccess) rc = fail(ctx, LISREG_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
    // descriptors go through pageable staging (consumed before cudaMemcpyAsync returns): this call does not synchronise,
    // so a pinned block could still be in use by the previous call's copies
    else rc = run_frames(ctx, fc, items + f0, nullptr, 0, d_pose6xF + 6 * (size_t)f0, prm, d_resxF + f0, nullptr, F, 0, f0);
  }
  ctx->cur = &ctx->ws[0];
  for (int c = 0; c < S; c++) {                                    // join
    cudaEventRecord(ctx->chunk_ev[c], ctx->ws[1 + c].stream);
    cudaStreamWaitEvent(st, ctx->chunk_ev[c], 0);
  }
  return rc;
}

int32_t lisreg_frames_batch_arena(lisreg_ctx* ctx, int32_t F, const lisreg_frame_item* items, const void* host_arena,
                                  uint64_t arena_bytes, float* pose6xF, const lisreg_frame_params* prm, lisreg_lm_result* resxF) {
  if (!ctx || F <= 0 || !items || !host_arena || !pose6xF || !prm || !resxF) return fail(ctx, LISREG_ERR_ARG, "lisreg_frames_batch_arena: bad argument");
  if (prm->lm.max_iters <= 0 || prm->lm.max_iters > LISREG_MAX_ITERS) return fail(ctx, LISREG_ERR_ARG, "max_iters must be in 1..%d", LISREG_MAX_ITERS);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t head = (sizeof(float) * 6 * (size_t)F + 255) & ~size_t(255);
  CK(ctx->d_stage.reserve(head + (size_t)arena_bytes + 16));
  CK(ctx->d_res.reserve(sizeof(lisreg_lm_result) * (size_t)F));
  char* d = (char*)ctx->d_stage.p;
  char* d_arena = d + head;
  float* d_pose = (float*)d;
  lisreg_lm_result* d_res = (lisreg_lm_result*)ctx->d_res.p;
  CK(cudaMemcpyAsync(d, pose6xF, sizeof(float) * 6 * (size_t)F, cudaMemcpyHostToDevice, st));

  // Chunk plan: frames [c*C, (c+1)*C) form chunk c; its arena extent is the byte range covering its sweeps.  When
  // the extents are disjoint and ascending (frames packed in order, the normal case) every chunk is uploaded by its
  // own copy on a second stream and the pipeline of chunk c (features -> voxel grid -> LM) starts as soon as its
  // bytes have landed, overlapping the PCIe transfer of the chunks behind it.
  const int C = ctx->e2e_chunk;
  const int nchunk = (C > 0 && F > C) ? (F + C - 1) / C : 1;
  std::vector<uint64_t> lo((size_t)nchunk, ~0ull), hi((size_t)nchunk, 0ull);
  // work sets 1 / 2 are shared with the submit / wait pipeline: while a ticket is in flight this call stays on set 0
  bool pipelined = nchunk > 1 && !ctx->slot[0].busy && !ctx->slot[1].busy;
  if (pipelined) {
    for (int i = 0; i < F; i++) {
      const lisreg_frame_item& it = items[i];
      if (it.n <= 0) continue;
      const int c = i / C;
      const uint64_t recb = (uint64_t)layout_step(prm->feat.layout);
      const bool rarr = layout_needs_ring_array(prm->feat.layout);
      const uint64_t op = (uint64_t)(size_t)it.pts, orr = rarr ? (uint64_t)(size_t)it.ring : op;
      if (!in_arena(op, recb * it.n, arena_bytes) || (rarr && !in_arena(orr, 2ull * it.n, arena_bytes))) { pipelined = false; break; }   // run_frames reports it
      lo[c] = std::min(lo[c], std::min(op, orr));
      hi[c] = std::max(hi[c], std::max(op + recb * (uint64_t)it.n, rarr ? orr + (uint64_t)2 * (uint64_t)it.n : op));
    }
    uint64_t prev = 0;
    for (int c = 0; c < nchunk && pipelined; c++) {
      if (hi[c] == 0) { lo[c] = hi[c] = prev; continue; }           // chunk without points
      lo[c] &= ~uint64_t(15);
      if (lo[c] < prev || hi[c] > arena_bytes) pipelined = false;   // overlapping / out of range: one copy (run_frames reports bad offsets)
      prev = hi[c];
    }
  }
  if (!pipelined) {
    CK(cudaMemcpyAsync(d_arena, host_arena, (size_t)arena_bytes, cudaMemcpyHostToDevice, st));
    int rc = run_frames(ctx, F, items, d_arena, arena_bytes, d_pose, prm, d_res);
    if (rc) return rc;
  } else {
    if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 1; i < 3; i++) if (!ctx->ws[i].stream) CK(cudaStreamCreateWithFlags(&ctx->ws[i].stream, cudaStreamNonBlocking));
    while ((int)ctx->chunk_ev.size() < nchunk + 3) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
    const size_t desc_per = frame_desc_bytes(C);
    CK(ctx->h_desc.reserve(desc_per * (size_t)nchunk));
    int rc = sync_maps(ctx);
    if (rc) return rc;
    // neither the copy stream nor the two chunk streams may touch the staging buffers while earlier work of this
    // context (incl. the pose upload above) is still in flight on the context stream
    cudaEvent_t ev_start = ctx->chunk_ev[nchunk];
    CK(cudaEventRecord(ev_start, st));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->ws[1].stream, ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->ws[2].stream, ev_start, 0));
    for (int c = 0; c < nchunk; c++) {
      if (hi[c] > lo[c])
        CK(cudaMemcpyAsync(d_arena + lo[c], (const char*)host_arena + lo[c], (size_t)(hi[c] - lo[c]), cudaMemcpyHostToDevice, ctx->copy_stream));
      CK(cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
    }
    // chunks alternate between two work sets with private streams: while one chunk sits in a latency-bound phase
    // (segment selection, the 6x6 solves, launch gaps) the other one keeps the SMs busy
    rc = LISREG_OK;
    for (int c = 0; c < nchunk && rc == LISREG_OK; c++) {
      const int f0 = c * C, fc = std::min(C, F - f0);
      ctx->cur = &ctx->ws[1 + (c & 1)];
      cudaError_t e = cudaStreamWaitEvent(ctx->cur->stream, ctx->chunk_ev[c], 0);
      if (e != cudaSuccess) rc = fail(ctx, LISREG_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
      else rc = run_frames(ctx, fc, items + f0, d_arena, arena_bytes, d_pose + 6 * (size_t)f0, prm, d_res + f0,
                           (char*)ctx->h_desc.p + desc_per * (size_t)c, F, 0, f0);
    }
    ctx->cur = &ctx->ws[0];
    // join: the context stream continues after both chunk streams
    for (int i = 1; i < 3; i++) {
      cudaEventRecord(ctx->chunk_ev[nchunk + i], ctx->ws[i].stream);
      cudaStreamWaitEvent(st, ctx->chunk_ev[nchunk + i], 0);
    }
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(st); return rc; }
  }
  CK(ctx->h_out.reserve(sizeof(lisreg_lm_result) * (size_t)F));
  CK(cudaMemcpyAsync(ctx->h_out.p, d_res, sizeof(lisreg_lm_result) * (size_t)F, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const lisreg_lm_result* hr = (const lisreg_lm_result*)ctx->h_out.p;
  int worst = LISREG_OK;
  for (int b = 0; b < F; b++) {
    resxF[b] = hr[b];
    memcpy(pose6xF + 6 * (size_t)b, hr[b].pose, sizeof(float) * 6);
    worst = std::max(worst, hr[b].status);
  }
  return worst;
}

int32_t lisreg_frames_batch_submit(lisreg_ctx* ctx, int32_t F, const lisreg_frame_item* items, const void* host_arena,
                                   uint64_t arena_bytes, const float* pose6xF, const lisreg_frame_params* prm, int32_t* ticket) {
  if (!ctx || F <= 0 || !items || !host_arena || !pose6xF || !prm || !ticket) return fail(ctx, LISREG_ERR_ARG, "lisreg_frames_batch_submit: bad argument");
  if (prm->lm.max_iters <= 0 || prm->lm.max_iters > LISREG_MAX_ITERS) return fail(ctx, LISREG_ERR_ARG, "max_iters must be in 1..%d", LISREG_MAX_ITERS);
  CK(cudaSetDevice(ctx->device));
  const int si = ctx->next_slot;
  lisreg_ctx::AsyncSlot& sl = ctx->slot[si];
  if (sl.busy) return fail(ctx, LISREG_ERR_ARG, "lisreg_frames_batch_submit: both pipeline slots are in flight, wait for a ticket first");
  WorkSet& w = ctx->ws[1 + si];
  if (!w.stream) CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
  if (!sl.done) CK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  int rc = sync_maps(ctx);
  if (rc) return rc;
  // the slot's stream must see everything already enqueued on the context stream (e.g. a map index still being built)
  if (!sl.fence) CK(cudaEventCreateWithFlags(&sl.fence, cudaEventDisableTiming));
  CK(cudaEventRecord(sl.fence, ctx->stream));
  CK(cudaStreamWaitEvent(w.stream, sl.fence, 0));
  const size_t head = (sizeof(float) * 6 * (size_t)F + 255) & ~size_t(255);
  CK(sl.d_stage.reserve(head + (size_t)arena_bytes + 16));
  CK(sl.d_res.reserve(sizeof(lisreg_lm_result) * (size_t)F));
  CK(sl.h_out.reserve(sizeof(lisreg_lm_result) * (size_t)F + head));
  CK(sl.h_desc.reserve(frame_desc_bytes(F)));
  char* d = (char*)sl.d_stage.p;
  // the guesses go through the slot's pinned buffer: the caller may reuse its array as soon as this call returns
  float* h_pose = (float*)((char*)sl.h_out.p + sizeof(lisreg_lm_result) * (size_t)F);
  memcpy(h_pose, pose6xF, sizeof(float) * 6 * (size_t)F);
  CK(cudaMemcpyAsync(d, h_pose, sizeof(float) * 6 * (size_t)F, cudaMemcpyHostToDevice, w.stream));
  CK(cudaMemcpyAsync(d + head, host_arena, (size_t)arena_bytes, cudaMemcpyHostToDevice, w.stream));
  ctx->cur = &w;
  rc = run_frames(ctx, F, items, d + head, arena_bytes, (float*)d, prm, (lisreg_lm_result*)sl.d_res.p, (char*)sl.h_desc.p);
  ctx->cur = &ctx->ws[0];
  if (rc) { cudaStreamSynchronize(w.stream); return rc; }
  CK(cudaMemcpyAsync(sl.h_out.p, sl.d_res.p, sizeof(lisreg_lm_result) * (size_t)F, cudaMemcpyDeviceToHost, w.stream));
  CK(cudaEventRecord(sl.done, w.stream));
  sl.busy = true; sl.F = F;
  *ticket = si;
  ctx->next_slot = si ^ 1;
  return LISREG_OK;
}

int32_t lisreg_frames_batch_wait(lisreg_ctx* ctx, int32_t ticket, float* pose6xF, lisreg_lm_result* resxF) {
  if (!ctx || ticket < 0 || ticket > 1 || !pose6xF || !resxF) return fail(ctx, LISREG_ERR_ARG, "lisreg_frames_batch_wait: bad argument");
  lisreg_ctx::AsyncSlot& sl = ctx->slot[ticket];
  if (!sl.busy) return fail(ctx, LISREG_ERR_ARG, "lisreg_frames_batch_wait: ticket %d is not in flight", ticket);
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(sl.done));
  sl.busy = false;
  const lisreg_lm_result* hr = (const lisreg_lm_result*)sl.h_out.p;
  int worst = LISREG_OK;
  for (int b = 0; b < sl.F; b++) {
    resxF[b] = hr[b];
    memcpy(pose6xF + 6 * (size_t)b, hr[b].pose, sizeof(float) * 6);
    worst = std::max(worst, hr[b].status);
  }
  return worst;
}

// ------------------------------------------------------------------------------------------------
// EPSC
// ------------------------------------------------------------------------------------------------
int32_t lisreg_epsc_describe(lisreg_ctx* ctx, int32_t n, const lisreg_epsc_cloud* clouds, const uint8_t using_map[256],
                             uint8_t* epsc, uint8_t* sepsc, uint8_t* fepsc) {
  if (!ctx || n <= 0 || !clouds || !using_map) return fail(ctx, LISREG_ERR_ARG, "lisreg_epsc_describe: bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  size_t total = 256;
  auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
  for (int i = 0; i < n; i++) {
    const lisreg_epsc_cloud& c = clouds[i];
    if (c.nc < 0 || c.ns < 0 || c.nsem < 0 || (c.nc && !c.corner) || (c.ns && !c.surf) || (c.nsem && (!c.sem || !c.sem_label)))
      return fail(ctx, LISREG_ERR_ARG, "lisreg_epsc_describe: cloud %d bad", i);
    total += al(16 * (size_t)c.nc) + al(16 * (size_t)c.ns) + al(16 * (size_t)c.nsem) + al(2 * (size_t)c.nsem);
  }
  total += al(sizeof(EpscCloud) * (size_t)n);
  CK(ctx->h_stage.reserve(total));
  CK(ctx->d_stage.reserve(total));
  CK(ctx->d_epsc.reserve(3 * (size_t)EPSC_SIZE * n));
  char* h = (char*)ctx->h_stage.p; char* d = (char*)ctx->d_stage.p;
  size_t off = 0;
  memcpy(h, using_map, 256); off = 256;
  EpscCloud* hc = (EpscCloud*)(h + off); const size_t o_desc = off; off += al(sizeof(EpscCloud) * (size_t)n);
  for (int i = 0; i < n; i++) {
    const lisreg_epsc_cloud& c = clouds[i];
    EpscCloud e{};
    e.nc = c.nc; e.ns = c.ns; e.nsem = c.nsem;
    if (c.nc) memcpy(h + off, c.corner, 16 * (size_t)c.nc); e.corner = (const float4*)(d + off); off += al(16 * (size_t)c.nc);
    if (c.ns) memcpy(h + off, c.surf, 16 * (size_t)c.ns); e.surf = (const float4*)(d + off); off += al(16 * (size_t)c.ns);
    if (c.nsem) memcpy(h + off, c.sem, 16 * (size_t)c.nsem); e.sem = (const float4*)(d + off); off += al(16 * (size_t)c.nsem);
    if (c.nsem) memcpy(h + off, c.sem_label, 2 * (size_t)c.nsem); e.sem_label = (const uint16_t*)(d + off); off += al(2 * (size_t)c.nsem);
    hc[i] = e;
  }
  CK(cudaMemcpyAsync(d, h, off, cudaMemcpyHostToDevice, st));
  k_epsc_describe<<<n, 256, 0, st>>>((const EpscCloud*)(d + o_desc), (const uint8_t*)d, (uint8_t*)ctx->d_epsc.p, nullptr, 0, 1); LAUNCH_CK();
  CK(ctx->h_out.reserve(3 * (size_t)EPSC_SIZE * n));
  CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_epsc.p, 3 * (size_t)EPSC_SIZE * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint8_t* o = (const uint8_t*)ctx->h_out.p;
  for (int i = 0; i < n; i++) {
    if (epsc) memcpy(epsc + (size_t)i * EPSC_SIZE, o + (size_t)i * 3 * EPSC_SIZE, EPSC_SIZE);
    if (sepsc) memcpy(sepsc + (size_t)i * EPSC_SIZE, o + (size_t)i * 3 * EPSC_SIZE + EPSC_SIZE, EPSC_SIZE);
    if (fepsc) memcpy(fepsc + (size_t)i * EPSC_SIZE, o + (size_t)i * 3 * EPSC_SIZE + 2 * EPSC_SIZE, EPSC_SIZE);
  }
  return LISREG_OK;
}

static int epsc_score_rows_dev(lisreg_ctx* ctx, const uint8_t* d_desc, int32_t N, int32_t row_begin, int32_t row_stride, int32_t topk,
                               int32_t* d_idx, float* d_score, int8_t* d_shift) {
  cudaStream_t st = ctx->stream;
  const int n_rows = row_begin < N ? (N - row_begin + row_stride - 1) / row_stride : 0;
  if (n_rows == 0) return LISREG_OK;
  CK(ctx->d_epsc2.reserve(sizeof(unsigned long long) * EPSC_TOPK_SLOTS * (size_t)n_rows));
  unsigned long long* row_top = (unsigned long long*)ctx->d_epsc2.p;
  CK(cudaMemsetAsync(row_top, 0xff, sizeof(unsigned long long) * EPSC_TOPK_SLOTS * (size_t)n_rows, st));
  {
    // clusters of EPSC_CLUSTER CTAs along the history axis share the tile's query rows (bulk-copy multicast)
    const int gx = (((N + EPSC_JT - 1) / EPSC_JT) + EPSC_CLUSTER - 1) / EPSC_CLUSTER * EPSC_CLUSTER;
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(gx, (n_rows + EPSC_QT - 1) / EPSC_QT); lc.blockDim = dim3(EPSC_THREADS); lc.dynamicSmemBytes = 0; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = EPSC_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CK(cudaLaunchKernelEx(&lc, k_epsc_score, d_desc, (int)N, (int)row_begin, (int)row_stride, (int)n_rows, row_top));
    ctx->launches++;
  }
  k_epsc_topk<<<(n_rows * topk + 127) / 128, 128, 0, st>>>(row_top, n_rows, topk, d_idx, d_score, d_shift); LAUNCH_CK();
  return LISREG_OK;
}

int32_t lisreg_epsc_score_all_dev(lisreg_ctx* ctx, const uint8_t* d_desc, int32_t N, int32_t topk,
                                  int32_t* d_idx, float* d_score, int8_t* d_shift) {
  if (!ctx || N <= 0 || N >= (1 << 24) || !d_desc || topk <= 0 || topk > 8 || !d_idx || !d_score || !d_shift) return fail(ctx, LISREG_ERR_ARG, "lisreg_epsc_score_all: bad argument");
  if (((size_t)d_desc & 15) != 0) return fail(ctx, LISREG_ERR_ARG, "lisreg_epsc_score_all_dev: descriptors must be 16-byte aligned (bulk-copy source)");
  CK(cudaSetDevice(ctx->device));
  return epsc_score_rows_dev(ctx, d_desc, N, 0, 1, topk, d_idx, d_score, d_shift);
}

// host buffers; rows q = row_begin + r * row_stride (the whole matrix: 0, 1); outputs n_rows x topk
int32_t lisreg_epsc_score_rows(lisreg_ctx* ctx, const uint8_t* desc, int32_t N, int32_t row_begin, int32_t row_stride, int32_t topk,
                               int32_t* idx, float* score, int8_t* shift) {
  if (!ctx || N <= 0 || N >= (1 << 24) || !desc || row_begin < 0 || row_stride <= 0 || topk <= 0 || topk > 8 || !idx || !score || !shift)
    return fail(ctx, LISREG_ERR_ARG, "lisreg_epsc_score_rows: bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n_rows = row_begin < N ? (N - row_begin + row_stride - 1) / row_stride : 0;
  if (n_rows == 0) return LISREG_OK;
  const size_t bd = (size_t)N * EPSC_SIZE, bi = 4 * (size_t)n_rows * topk, bs = 4 * (size_t)n_rows * topk, bh = (size_t)n_rows * topk;
  CK(ctx->d_epsc.reserve(bd + bi + bs + bh + 64));
  char* d = (char*)ctx->d_epsc.p;
  CK(cudaMemcpyAsync(d, desc, bd, cudaMemcpyHostToDevice, st));
  int rc = epsc_score_rows_dev(ctx, (const uint8_t*)d, N, row_begin, row_stride, topk, (int32_t*)(d + bd), (float*)(d + bd + bi), (int8_t*)(d + bd + bi + bs));
  if (rc) return rc;
  CK(cudaMemcpyAsync(idx, d + bd, bi, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(score, d + bd + bi, bs, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(shift, d + bd + bi + bs, bh, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return LISREG_OK;
}

int32_t lisreg_epsc_score_all(lisreg_ctx* ctx, const uint8_t* desc, int32_t N, int32_t topk, int32_t* idx, float* score, int8_t* shift) {
  return lisreg_epsc_score_rows(ctx, desc, N, 0, 1, topk, idx, score, shift);
}

// ------------------------------------------------------------------------------------------------
// EPSC loop detector (B3)
// ------------------------------------------------------------------------------------------------
void lisreg_loop_params_default(lisreg_loop_params* p) {
  memset(p, 0, sizeof(*p));
  p->use_fepsc = 1;                          // config/params.yaml:22-28: only UsingFEPSCFlag is true
  p->skip_neighbour_distance = 20.0f;        // SKIP_NEIBOUR_DISTANCE  (epscGeneration.h:9)
  p->inflation_covariance = 0.01f;           // INFLATION_COVARIANCE   (:11)
  p->distance_threshold = 0.75f;             // DISTANCE_THRESHOLD     (:16)
}

int32_t lisreg_loop_create(lisreg_ctx* ctx, const lisreg_loop_params* prm, const uint8_t using_map[256], int32_t* det_id) {
  if (!ctx || !prm || !using_map || !det_id) return fail(ctx, LISREG_ERR_ARG, "lisreg_loop_create: bad argument");
  CK(cudaSetDevice(ctx->device));
  int slot = -1;
  for (size_t i = 0; i < ctx->loops.size(); i++) if (!ctx->loops[i].used) { slot = (int)i; break; }
  if (slot < 0) { ctx->loops.emplace_back(); slot = (int)ctx->loops.size() - 1; }
  lisreg_ctx::LoopDet& L = ctx->loops[slot];
  L = lisreg_ctx::LoopDet();
  L.prm = *prm;
  CK(cudaMalloc(&L.d_lut, 256));
  CK(cudaMemcpyAsync(L.d_lut, using_map, 256, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  L.used = true;
  *det_id = slot;
  return LISREG_OK;
}

int32_t lisreg_loop_destroy(lisreg_ctx* ctx, int32_t det_id) {
  if (!ctx || det_id < 0 || det_id >= (int)ctx->loops.size() || !ctx->loops[det_id].used) return fail(ctx, LISREG_ERR_ARG, "lisreg_loop_destroy: bad detector id");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  lisreg_ctx::LoopDet& L = ctx->loops[det_id];
  cudaFree(L.d_proj); cudaFree(L.d_desc); cudaFree(L.d_lut);
  L = lisreg_ctx::LoopDet();
  return LISREG_OK;
}

static inline float f_sin64(float a) { return (float)sin((double)a); }
static inline float f_cos64(float a) { return (float)cos((double)a); }

// Identity; translation << dx, dy, 0; rotate(AngleAxisf(angle, UnitZ))  (epscGeneration.cpp:822-826)
static void loop_planar_transform(float dx, float dy, float angle, float* T) {
  const float c = f_cos64(angle), s = f_sin64(angle);
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.f : 0.f;
  T[0] = c; T[1] = -s; T[4] = s; T[5] = c; T[10] = (1.f - c) + c;
  T[3] = dx; T[7] = dy; T[11] = 0.f;
}

int32_t lisreg_loop_detect(lisreg_ctx* ctx, int32_t det_id, const float* corner, int32_t nc, const float* surf, int32_t ns,
                           const float* sem, const uint16_t* sem_label, int32_t nsem, const float odom[16], lisreg_loop_result* res) {
  if (!ctx || det_id < 0 || det_id >= (int)ctx->loops.size() || !ctx->loops[det_id].used || !odom || !res ||
      nc < 0 || ns < 0 || nsem < 0 || (nc && !corner) || (ns && !surf) || (nsem && (!sem || !sem_label)))
    return fail(ctx, LISREG_ERR_ARG, "lisreg_loop_detect: bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  lisreg_ctx::LoopDet& L = ctx->loops[det_id];
  memset(res, 0, sizeof(*res));
  // ---- pose bookkeeping (:687-701): pcl::getTranslationAndEulerAngles, travelled distance ----
  const float x_t = odom[3], y_t = odom[7];
  const float yaw_t = (float)atan2((double)odom[4], (double)odom[0]);
  const double cx = x_t, cy = y_t;
  // travelDistanceArr entry of THIS keyframe; committed together with posArr / yawArr / the device history only after
  // every fallible step below has succeeded, so an error return leaves the detector exactly as it was
  double travel_cur = 0;
  if (!L.travel.empty()) { const double ex = L.px.back() - cx, ey = L.py.back() - cy; travel_cur = L.travel.back() + sqrt(ex * ex + ey * ey + 0.0); }
  const int cur = L.n;
  res->current_frame_id = cur;
  // ---- travel gate (:736-741); posArr.back() is still the PREVIOUS keyframe here ----
  std::vector<int> cand; std::vector<float> cand_yaw; std::vector<double> cand_pos;
  for (int i = 0; i < cur; i++) {
    const double delta_travel = travel_cur - L.travel[i];
    const double qx = L.px[i] - L.px.back(), qy = L.py[i] - L.py.back();
    const double pos_distance = sqrt(qx * qx + qy * qy + 0.0);
    if (delta_travel > (double)L.prm.skip_neighbour_distance && pos_distance < delta_travel * (double)L.prm.inflation_covariance) {
      cand.push_back(i); cand_yaw.push_back(yaw_t - L.yaw[i]); cand_pos.push_back(pos_distance);
    }
  }
  const int P = (int)cand.size();
  res->n_candidates = P;
  // ---- stage the three clouds (one pinned buffer -> one H2D copy) ----
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t o_c = 0, o_s = o_c + al(16 * (size_t)nc), o_m = o_s + al(16 * (size_t)ns), o_l = o_m + al(16 * (size_t)nsem),
               o_desc = o_l + al(2 * (size_t)nsem), o_cid = o_desc + al(sizeof(EpscCloud)), o_cyaw = o_cid + al(4 * (size_t)std::max(P, 1)),
               in_total = o_cyaw + al(4 * (size_t)std::max(P, 1));
  CK(ctx->h_stage.reserve(in_total));
  CK(ctx->d_stage.reserve(in_total));
  char* h = (char*)ctx->h_stage.p; char* d = (char*)ctx->d_stage.p;
  if (nc) memcpy(h + o_c, corner, 16 * (size_t)nc);
  if (ns) memcpy(h + o_s, surf, 16 * (size_t)ns);
  if (nsem) { memcpy(h + o_m, sem, 16 * (size_t)nsem); memcpy(h + o_l, sem_label, 2 * (size_t)nsem); }
  EpscCloud ec{};
  ec.corner = (const float4*)(d + o_c); ec.surf = (const float4*)(d + o_s); ec.sem = (const float4*)(d + o_m);
  ec.sem_label = (const uint16_t*)(d + o_l); ec.nc = nc; ec.ns = ns; ec.nsem = nsem;
  memcpy(h + o_desc, &ec, sizeof(ec));
  if (P) { memcpy(h + o_cid, cand.data(), 4 * (size_t)P); memcpy(h + o_cyaw, cand_yaw.data(), 4 * (size_t)P); }
  CK(cudaMemcpyAsync(d, h, in_total, cudaMemcpyHostToDevice, st));
  // scratch: current projection, per-candidate align / score results, per-candidate descriptors of the moved cloud
  const size_t s_proj = 0, s_align = s_proj + al(sizeof(float4) * LOOP_SECT), s_score = s_align + al(sizeof(LoopAlignOut) * (size_t)std::max(P, 1)),
               s_cdesc = s_score + al(sizeof(LoopScoreOut) * (size_t)std::max(P, 1)), s_total = s_cdesc + al(3 * (size_t)EPSC_SIZE * std::max(P, 1));
  CK(ctx->d_loop.reserve(s_total));
  char* sc = (char*)ctx->d_loop.p;
  float4* d_curproj = (float4*)(sc + s_proj);
  k_loop_project<<<1, 256, 0, st>>>((const float4*)(d + o_m), (const uint16_t*)(d + o_l), nsem, d_curproj); LAUNCH_CK();
  std::vector<LoopAlignOut> ha((size_t)P); std::vector<LoopScoreOut> hs((size_t)P);
  if (P) {
    k_loop_align<<<P, LOOP_THREADS, 0, st>>>(L.d_proj, d_curproj, (const int*)(d + o_cid), (const float*)(d + o_cyaw), (LoopAlignOut*)(sc + s_align)); LAUNCH_CK();
    k_epsc_describe<<<P, 256, 0, st>>>((const EpscCloud*)(d + o_desc), L.d_lut, (uint8_t*)(sc + s_cdesc),
                                       (const float*)(sc + s_align), (int)(sizeof(LoopAlignOut) / sizeof(float)), 0); LAUNCH_CK();
    k_loop_score<<<P, 64, 0, st>>>((const uint8_t*)(sc + s_cdesc), L.d_desc, (const int*)(d + o_cid), (LoopScoreOut*)(sc + s_score)); LAUNCH_CK();
    CK(cudaMemcpyAsync(ha.data(), sc + s_align, sizeof(LoopAlignOut) * (size_t)P, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hs.data(), sc + s_score, sizeof(LoopScoreOut) * (size_t)P, cudaMemcpyDeviceToHost, st));
  }
  // ---- append the current keyframe to the device history (:899-967): projection + descriptors of the UNmoved clouds ----
  if (L.n == L.cap) {
    const int ncap = std::max(256, 2 * L.cap);
    float4* np = nullptr; uint8_t* nd = nullptr;
    cudaError_t ge = cudaMalloc(&np, sizeof(float4) * LOOP_SECT * (size_t)ncap);
    if (ge == cudaSuccess) ge = cudaMalloc(&nd, 3 * (size_t)EPSC_SIZE * ncap);
    if (ge == cudaSuccess && L.n) ge = cudaMemcpyAsync(np, L.d_proj, sizeof(float4) * LOOP_SECT * (size_t)L.n, cudaMemcpyDeviceToDevice, st);
    if (ge == cudaSuccess && L.n) ge = cudaMemcpyAsync(nd, L.d_desc, 3 * (size_t)EPSC_SIZE * L.n, cudaMemcpyDeviceToDevice, st);
    if (ge == cudaSuccess) ge = cudaStreamSynchronize(st);
    if (ge != cudaSuccess) {                     // the old history stays valid and owned; nothing leaks
      cudaFree(np); cudaFree(nd);
      return fail(ctx, LISREG_ERR_CUDA, "lisreg_loop_detect: growing the history failed: %s", cudaGetErrorString(ge));
    }
    cudaFree(L.d_proj); cudaFree(L.d_desc);
    L.d_proj = np; L.d_desc = nd; L.cap = ncap;
  }
  CK(cudaMemcpyAsync(L.d_proj + (size_t)LOOP_SECT * L.n, d_curproj, sizeof(float4) * LOOP_SECT, cudaMemcpyDeviceToDevice, st));
  k_epsc_describe<<<1, 256, 0, st>>>((const EpscCloud*)(d + o_desc), L.d_lut, L.d_desc + 3 * (size_t)EPSC_SIZE * L.n, nullptr, 0, 1); LAUNCH_CK();
  CK(cudaStreamSynchronize(st));
  L.travel.push_back(travel_cur); L.px.push_back(cx); L.py.push_back(cy); L.yaw.push_back(yaw_t); L.n++;
  // ---- best candidate per descriptor kind (:812-897): first maximum above the threshold wins ----
  const int use[3] = {L.prm.use_epsc, L.prm.use_sepsc, L.prm.use_fepsc};
  int best_id[3] = {-1, -1, -1}; double best_score[3] = {0, 0, 0}; float best_T[3][16];
  double min_distance = 1000000; int best_pose = -1;
  const double sector_step = 2 * 3.14159265358979323846 / 80;
  for (int k = 0; k < P; k++) {
    for (int kind = 0; kind < 3; kind++) {
      if (!use[kind]) continue;
      const int sad = hs[k].sad[kind];
      const double score = sad >= 0 ? 1 - (double)sad / (80 * 20 * 255) : 1 - 1.0;
      if (score > (double)L.prm.distance_threshold && score > best_score[kind]) {
        best_score[kind] = score; best_id[kind] = cand[k];
        // EPSC rotates by yaw_diff + shift (:816-829), SEPSC by the ICP yaw + shift (:837-850), FEPSC by the ICP yaw (:857-869)
        double a = kind == 0 ? (double)cand_yaw[k] : (double)ha[k].yaw;
        if (kind != 2 && sad >= 0) a = a + hs[k].shift[kind] * sector_step;
        loop_planar_transform(ha[k].diff_x, ha[k].diff_y, (float)a, best_T[kind]);
      }
    }
    if (L.prm.use_pose && cand_pos[k] < min_distance) { min_distance = cand_pos[k]; best_pose = k; }
  }
  int m = 0;
  for (int kind = 0; kind < 3; kind++) {
    if (!use[kind] || best_id[kind] < 0) continue;
    res->match[m].kind = kind; res->match[m].frame_id = best_id[kind]; res->match[m].score = best_score[kind];
    memcpy(res->match[m].T, best_T[kind], sizeof(float) * 16); m++;
  }
  if (L.prm.use_pose && best_pose >= 0) {
    res->match[m].kind = 3; res->match[m].frame_id = cand[best_pose]; res->match[m].score = min_distance;
    memcpy(res->match[m].T, ha[best_pose].T, sizeof(float) * 16); m++;
  }
  res->n_matched = m;
  return LISREG_OK;
}

// ------------------------------------------------------------------------------------------------
// ICP verify
// ------------------------------------------------------------------------------------------------
void lisreg_icp_params_default(lisreg_icp_params* p) { p->max_corr_dist = 10.f; p->max_iters = 30; p->trans_eps = 1e-4; p->fitness_eps = 1e-4; }

// ICP of P pairs whose (pre-transformed) sources already sit in ctx->d_stage at byte offsets off[i] (256-byte aligned slots of
// a region of src_bytes bytes): work copies, per-iteration correspondence + solve launches, results to the host
static int icp_run_dev(lisreg_ctx* ctx, int P, const std::vector<size_t>& off, const std::vector<int>& ns, const std::vector<int>& tgt,
                       size_t src_bytes, const lisreg_icp_params* prm, lisreg_icp_result* out) {
  cudaStream_t st = ctx->stream;
  int rc = sync_maps(ctx);
  if (rc) return rc;
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  int max_ns = 0;
  for (int i = 0; i < P; i++) max_ns = std::max(max_ns, ns[i]);
  const int nblk = std::max(1, std::min(64, (max_ns + ICP_THREADS - 1) / ICP_THREADS));
  const size_t o_pairs = 0, o_states = al(sizeof(IcpPair) * (size_t)P), o_part = o_states + al(sizeof(IcpState) * (size_t)P),
               o_res = o_part + al(sizeof(double) * ICP_NSUM * (size_t)P * nblk), o_cur = o_res + al(sizeof(lisreg_icp_result) * (size_t)P),
               o_nn = o_cur + al(src_bytes);
  CK(ctx->d_icp.reserve(o_nn + src_bytes / 4 + 256));
  char* dsrc = (char*)ctx->d_stage.p; char* d = (char*)ctx->d_icp.p;
  std::vector<IcpPair> hp((size_t)P);
  for (int i = 0; i < P; i++) { hp[i].src = (const float4*)(dsrc + off[i]); hp[i].ns = ns[i]; hp[i].cur = (float4*)(d + o_cur + off[i]); hp[i].nn = (int*)(d + o_nn + off[i] / 4); hp[i].tgt_slot = tgt[i]; hp[i].pad = 0; }
  CK(cudaMemcpyAsync(d + o_pairs, hp.data(), sizeof(IcpPair) * (size_t)P, cudaMemcpyHostToDevice, st));   // pageable: consumed on return
  IcpPair* dp = (IcpPair*)(d + o_pairs); IcpState* ds = (IcpState*)(d + o_states); double* part = (double*)(d + o_part);
  lisreg_icp_result* dres = (lisreg_icp_result*)(d + o_res);
  IcpParamsDev kp{prm->max_corr_dist * prm->max_corr_dist, prm->max_iters, 1.0 - prm->trans_eps, prm->trans_eps, prm->fitness_eps};
  k_icp_init<<<(P + 127) / 128, 128, 0, st>>>(dp, ds, P); LAUNCH_CK();
  k_icp_copy<<<dim3(nblk, P), ICP_THREADS, 0, st>>>(dp); LAUNCH_CK();
  for (int it = 0; it < prm->max_iters; it++) {
    k_icp_corr<false><<<dim3(nblk, P), ICP_THREADS, 0, st>>>(dp, ds, ctx->d_maps, kp, part, nblk); LAUNCH_CK();
    k_icp_solve<false><<<(P + 3) / 4, 128, 0, st>>>(ds, kp, part, nblk, P); LAUNCH_CK();
  }
  k_icp_corr<true><<<dim3(nblk, P), ICP_THREADS, 0, st>>>(dp, ds, ctx->d_maps, kp, part, nblk); LAUNCH_CK();
  k_icp_solve<true><<<(P + 3) / 4, 128, 0, st>>>(ds, kp, part, nblk, P); LAUNCH_CK();
  k_icp_finish<<<(P + 127) / 128, 128, 0, st>>>(ds, dres, P); LAUNCH_CK();
  CK(cudaMemcpyAsync(out, dres, sizeof(lisreg_icp_result) * (size_t)P, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return LISREG_OK;
}

int32_t lisreg_icp_verify_batch(lisreg_ctx* ctx, int32_t P, const lisreg_icp_pair* pairs, const lisreg_icp_params* prm,
                                lisreg_icp_result* out) {
  if (!ctx || P <= 0 || !pairs || !prm || !out || prm->max_iters <= 0) return fail(ctx, LISREG_ERR_ARG, "lisreg_icp_verify_batch: bad argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  size_t src_bytes = 0;
  std::vector<size_t> off((size_t)P); std::vector<int> ns((size_t)P), tgt((size_t)P);
  for (int i = 0; i < P; i++) {
    const lisreg_icp_pair& pr = pairs[i];
    if (pr.ns < 0 || (pr.ns > 0 && !pr.src) || pr.target_id < 0 || pr.target_id >= (int)ctx->maps.size() || !ctx->maps[pr.target_id].used)
      return fail(ctx, LISREG_ERR_ARG, "icp pair %d: bad size, pointer or target id", i);
    off[i] = src_bytes; ns[i] = pr.ns; tgt[i] = pr.target_id;
    src_bytes += al(16 * (size_t)pr.ns);
  }
  CK(ctx->d_stage.reserve(src_bytes + 256));
  // sources in page-locked memory go to the device straight from the caller's buffer; pageable ones through the pinned
  // staging block (one host copy more)
  std::vector<char> pinned((size_t)P, 0);
  size_t staged = 0;
  for (int i = 0; i < P; i++) {
    if (!pairs[i].ns) continue;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, pairs[i].src) == cudaSuccess && at.type == cudaMemoryTypeHost) pinned[i] = 1;
    else { cudaGetLastError(); staged += 16 * (size_t)pairs[i].ns; }
  }
  if (staged) CK(ctx->h_stage.reserve(src_bytes + 256));
  char* h = (char*)ctx->h_stage.p;
  for (int i = 0; i < P; i++) {
    if (!pairs[i].ns) continue;
    const void* from = pairs[i].src;
    if (!pinned[i]) { memcpy(h + off[i], pairs[i].src, 16 * (size_t)pairs[i].ns); from = h + off[i]; }
    CK(cudaMemcpyAsync((char*)ctx->d_stage.p + off[i], from, 16 * (size_t)pairs[i].ns, cudaMemcpyHostToDevice, st));
  }
  return icp_run_dev(ctx, P, off, ns, tgt, src_bytes, prm, out);
}

int32_t lisreg_selftest_smallmat(lisreg_ctx* ctx, const float* A36, const float* b6, float* out98) {
  if (!ctx || !A36 || !b6 || !out98) return LISREG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_stage.reserve(sizeof(float) * (36 + 6 + 98)));
  float* d = (float*)ctx->d_stage.p;
  CK(cudaMemcpyAsync(d, A36, sizeof(float) * 36, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d + 36, b6, sizeof(float) * 6, cudaMemcpyHostToDevice, ctx->stream));
  k_selftest_smallmat<<<1, 32, 0, ctx->stream>>>(d, d + 36, d + 42); LAUNCH_CK();
  CK(cudaMemcpyAsync(out98, d + 42, sizeof(float) * 98, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LISREG_OK;
}

int32_t lisreg_selftest_alu_peak(lisreg_ctx* ctx, double* gsad_per_s) {
  if (!ctx || !gsad_per_s) return LISREG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->d_stage.reserve(4 * 4096));
  cudaStream_t st = ctx->stream;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int blocks = ctx->n_sm * 8, iters = 4096;
  k_epsc_alu_peak<<<blocks, 256, 0, st>>>((unsigned*)ctx->d_stage.p, 64, 1u); LAUNCH_CK();     // warm-up
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    CK(cudaEventRecord(a, st));
    k_epsc_alu_peak<<<blocks, 256, 0, st>>>((unsigned*)ctx->d_stage.p, iters, 2u + rep); LAUNCH_CK();
    CK(cudaEventRecord(b, st));
    CK(cudaEventSynchronize(b));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, a, b));
    const double ops = (double)blocks * 256.0 * iters * 64.0;            // VABSDIFF4 instructions (4 byte-SADs each)
    best = std::max(best, ops / (ms * 1e-3) / 1e9);
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  *gsad_per_s = best;
  return LISREG_OK;
}

int32_t lisreg_profile_enable(lisreg_ctx* ctx, int32_t on) {
  if (!ctx) return LISREG_ERR_ARG;
  ctx->prof_on = on != 0;
  return LISREG_OK;
}

int32_t lisreg_profile_get(lisreg_ctx* ctx, lisreg_profile* out, int32_t reset) {
  if (!ctx || !out) return LISREG_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (auto& p : ctx->ev_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      switch (p.kind) {
        case PROF_LM:    ctx->prof.lm_iter_ms += ms; ctx->prof.lm_iter_launches += p.launches; ctx->prof.lm_alg_bytes += p.bytes; break;
        case PROF_FEAT:  ctx->prof.feat_ms += ms; ctx->prof.feat_launches += p.launches; ctx->prof.feat_alg_bytes += p.bytes; break;
        case PROF_VOXEL: ctx->prof.voxel_ms += ms; ctx->prof.voxel_launches += p.launches; ctx->prof.voxel_alg_bytes += p.bytes; break;
        default:         ctx->prof.index_ms += ms; ctx->prof.index_launches += p.launches; ctx->prof.index_alg_bytes += p.bytes; break;
      }
    }
    ctx->ev_free.push_back(p.a); ctx->ev_free.push_back(p.b);
  }
  ctx->ev_pending.clear();
  *out = ctx->prof;
  if (reset) ctx->prof = lisreg_profile{};
  return LISREG_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU exchange: NCCL bound at run time
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclId { char b[LISREG_NCCL_ID_BYTES]; };                          // ncclUniqueId: 128 opaque bytes, passed by value
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;                                   // ncclResult_t ncclGetUniqueId(ncclUniqueId*)
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
typedef decltype(NcclApi::CommInitRank) nccl_init_fn;
NcclApi g_nccl;
bool nccl_load(std::string* why) {
  if (g_nccl.h) return true;
  void* h = nullptr;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
  if (!h) { if (why) *why = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?"); return false; }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (nccl_init_fn)dlsym(h, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) { if (why) *why = "libnccl lacks a required symbol"; return false; }
  g_nccl.h = h;
  return true;
}
const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"; }
}  // namespace

int32_t lisreg_comm_unique_id(uint8_t id[LISREG_NCCL_ID_BYTES]) {
  if (!id || !nccl_load(nullptr)) return LISREG_ERR_CUDA;
  return g_nccl.GetUniqueId(id) == 0 ? LISREG_OK : LISREG_ERR_CUDA;
}

static int comm_streams(lisreg_ctx* ctx) {
  if (!ctx->comm_stream) CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  if (!ctx->comm_fence) CK(cudaEventCreateWithFlags(&ctx->comm_fence, cudaEventDisableTiming));
  for (auto& e : ctx->comm_done) if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return LISREG_OK;
}

int32_t lisreg_comm_init(lisreg_ctx* ctx, int32_t world, int32_t rank, const uint8_t id[LISREG_NCCL_ID_BYTES]) {
  if (!ctx || world < 1 || rank < 0 || rank >= world || !id) return fail(ctx, LISREG_ERR_ARG, "lisreg_comm_init: bad argument");
  std::string why;
  if (!nccl_load(&why)) return fail(ctx, LISREG_ERR_CUDA, "lisreg_comm_init: %s", why.c_str());
  CK(cudaSetDevice(ctx->device));
  if (ctx->comm) lisreg_comm_destroy(ctx);
  NcclId uid; memcpy(uid.b, id, LISREG_NCCL_ID_BYTES);
  void* comm = nullptr;
  const int rc = g_nccl.CommInitRank(&comm, world, uid, rank);
  if (rc != 0) return fail(ctx, LISREG_ERR_CUDA, "ncclCommInitRank failed: %s", nccl_err(rc));
  ctx->comm = comm; ctx->comm_owned = true; ctx->comm_world = world; ctx->comm_rank = rank;
  return comm_streams(ctx);
}

int32_t lisreg_comm_adopt(lisreg_ctx* ctx, void* nccl_comm, int32_t world, int32_t rank) {
  if (!ctx || !nccl_comm || world < 1 || rank < 0 || rank >= world) return fail(ctx, LISREG_ERR_ARG, "lisreg_comm_adopt: bad argument");
  std::string why;
  if (!nccl_load(&why)) return fail(ctx, LISREG_ERR_CUDA, "lisreg_comm_adopt: %s", why.c_str());
  CK(cudaSetDevice(ctx->device));
  if (ctx->comm) lisreg_comm_destroy(ctx);
  ctx->comm = nccl_comm; ctx->comm_owned = false; ctx->comm_world = world; ctx->comm_rank = rank;
  return comm_streams(ctx);
}

int32_t lisreg_comm_destroy(lisreg_ctx* ctx) {
  if (!ctx) return LISREG_ERR_ARG;
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  if (ctx->comm && ctx->comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  ctx->comm = nullptr; ctx->comm_owned = false; ctx->comm_world = 1; ctx->comm_rank = 0; ctx->comm_seq = 0;
  if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
  if (ctx->comm_fence) { cudaEventDestroy(ctx->comm_fence); ctx->comm_fence = nullptr; }
  for (auto& e : ctx->comm_done) if (e) { cudaEventDestroy(e); e = nullptr; }
  return LISREG_OK;
}

int32_t lisreg_allgather_results(lisreg_ctx* ctx, const void* d_send, void* d_recv, uint64_t bytes_per_rank) {
  if (!ctx || !d_send || !d_recv || bytes_per_rank == 0) return fail(ctx, LISREG_ERR_ARG, "lisreg_allgather_results: bad argument");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->comm) {                                  // a single-process run: the gather is a copy
    if (d_send != d_recv) CK(cudaMemcpyAsync(d_recv, d_send, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return LISREG_OK;
  }
  CK(cudaEventRecord(ctx->comm_fence, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_fence, 0));
  const int rc = g_nccl.AllGather(d_send, d_recv, (size_t)bytes_per_rank, /* ncclInt8 */ 0, ctx->comm, ctx->comm_stream);
  if (rc != 0) return fail(ctx, LISREG_ERR_CUDA, "ncclAllGather failed: %s", nccl_err(rc));
  CK(cudaEventRecord(ctx->comm_done[ctx->comm_seq % 4], ctx->comm_stream));
  ctx->comm_seq++;
  return LISREG_OK;
}

int32_t lisreg_allgather_wait(lisreg_ctx* ctx, int32_t back) {
  if (!ctx || back < 0 || back > 2) return fail(ctx, LISREG_ERR_ARG, "lisreg_allgather_wait: back must be 0..2");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->comm) { if (back == 0) CK(cudaStreamSynchronize(ctx->stream)); return LISREG_OK; }
  const int64_t k = ctx->comm_seq - 1 - back;          // the gather to wait for (and with it every earlier one)
  if (k >= 0) CK(cudaEventSynchronize(ctx->comm_done[k % 4]));
  return LISREG_OK;
}

// ------------------------------------------------------------------------------------------------
// streaming odometry (device-resident sliding-window map)
// ------------------------------------------------------------------------------------------------
// Host-side pose arithmetic of updateInitialGuess / calculateTranslation (odomEstimationNode.cpp:284-391):
// Eigen::Affine3f products and inverses in fp32.  Eigen's vectorised evaluation order is not pinnable; the natural
// order is used (row times column, k ascending; cofactor inverse of the linear part) - mirrored by
// lis_slam_b200/stream.py, which drives the CPU oracle through the same flow for the parity tests.
static void odom_T16(const float pose[6], float T[16]) {       // pcl::getTransformation(x, y, z, roll, pitch, yaw)
  RegState s; for (int i = 0; i < 6; i++) s.pose[i] = pose[i];
  state_refresh(s);
  for (int i = 0; i < 12; i++) T[i] = s.T[i];
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}
static inline float odom_cof(const float* T, int i, int j) {    // cofactor (i, j) of the 3x3 linear part of a row-major 4x4
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return T[i1 * 4 + j1] * T[i2 * 4 + j2] - T[i1 * 4 + j2] * T[i2 * 4 + j1];
}
static void odom_inv(const float T[16], float R[16]) {          // Eigen::Affine3f::inverse(): linear^-1, -linear^-1 * t
  const float c0 = odom_cof(T, 0, 0), c1 = odom_cof(T, 1, 0), c2 = odom_cof(T, 2, 0);
  const float det = (c0 * T[0] + c1 * T[4]) + c2 * T[8];
  const float invdet = 1.f / det;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i * 4 + j] = odom_cof(T, j, i) * invdet;
  for (int i = 0; i < 3; i++) R[i * 4 + 3] = ((-R[i * 4 + 0]) * T[3] + (-R[i * 4 + 1]) * T[7]) + (-R[i * 4 + 2]) * T[11];
  R[12] = 0.f; R[13] = 0.f; R[14] = 0.f; R[15] = 1.f;
}
static void odom_mul(const float A[16], const float B[16], float Cm[16]) {
  float r[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float a = 0.f; for (int k = 0; k < 4; k++) a += A[i * 4 + k] * B[k * 4 + j]; r[i * 4 + j] = a; }
  memcpy(Cm, r, sizeof(r));
}
static void odom_euler(const float T[16], float pose[6]) {      // pcl::getTranslationAndEulerAngles
  pose[3] = T[3]; pose[4] = T[7]; pose[5] = T[11];
  pose[0] = (float)atan2((double)T[9], (double)T[10]);
  pose[1] = (float)asin((double)-T[8]);
  pose[2] = (float)atan2((double)T[4], (double)T[0]);
}

void lisreg_odom_params_default(lisreg_odom_params* p) {
  memset(p, 0, sizeof(*p));
  lisreg_frame_params_default(&p->frame);
  p->keyframe_min_distance = 1.4f; p->keyframe_min_yaw = 0.5f;   // config/params.yaml:140-141
  p->window = 19;                                                 // odomEstimationNode.cpp:463
  p->use_graph = 1;
  p->use_imu_heading_initialization = 0;                          // config/params.yaml:77
  p->imu_rpy_weight = 0.01f;                                      // utility.h:405 (the shipped yaml sets 0.1)
}

int32_t lisreg_odom_create(lisreg_ctx* ctx, const lisreg_odom_params* prm, int32_t* odom_id) {
  if (!ctx || !prm || !odom_id) return fail(ctx, LISREG_ERR_ARG, "lisreg_odom_create: bad argument");
  const lisreg_feat_params& fp = prm->frame.feat;
  if (fp.n_scan <= 0 || fp.horizon <= 0 || fp.horizon > 2048 || fp.n_scan * 6 > 1024 || fp.downsample_rate <= 0 ||
      prm->window < 1 || prm->window + 1 > ODOM_MAX_SLOTS || !(prm->frame.corner_leaf > 0.f) || !(prm->frame.surf_leaf > 0.f) ||
      !layout_ok(fp.layout, fp.n_scan) ||
      prm->frame.lm.max_iters <= 0 || prm->frame.lm.max_iters > LISREG_MAX_ITERS)
    return fail(ctx, LISREG_ERR_ARG, "lisreg_odom_create: unsupported parameters");
  CK(cudaSetDevice(ctx->device));
  int slot = -1;
  for (size_t i = 0; i < ctx->odoms.size(); i++) if (!ctx->odoms[i].used) { slot = (int)i; break; }
  if (slot < 0) { ctx->odoms.emplace_back(); slot = (int)ctx->odoms.size() - 1; }
  lisreg_ctx::Odom& O = ctx->odoms[slot];
  O = lisreg_ctx::Odom();
  O.prm = *prm;
  O.rot_tol = prm->frame.lm.rot_tolerance; O.z_tol = prm->frame.lm.z_tolerance;
  O.prm.frame.lm.rot_tolerance = 0.f; O.prm.frame.lm.z_tolerance = 0.f;   // see odom_transform_update
  O.slots = prm->window + 1;
  O.ccap = fp.n_scan * 120; O.scap = fp.n_scan * fp.horizon;
  CK(cudaMalloc(&O.d_win_c, sizeof(float4) * (size_t)O.ccap * O.slots));
  { cudaError_t e = cudaMalloc(&O.d_win_s, sizeof(float4) * (size_t)O.scap * O.slots);
    if (e != cudaSuccess) { cudaFree(O.d_win_c); O.d_win_c = nullptr; return fail(ctx, LISREG_ERR_CUDA, "lisreg_odom_create: window allocation failed: %s", cudaGetErrorString(e)); } }
  O.nc.assign(O.slots, 0); O.ns.assign(O.slots, 0);
  O.used = true;
  // every buffer the stream will need, sized for a full window now: no cudaMalloc / cudaFree (a device-wide synchronisation)
  // while frames are flowing
  {
    const size_t tot_c = (size_t)O.ccap * prm->window, tot_s = (size_t)O.scap * prm->window;
    const size_t cat_c = (sizeof(float4) * tot_c + 255) & ~size_t(255);
    cudaError_t e = O.d_cat.reserve(cat_c + sizeof(float4) * tot_s);
    if (e == cudaSuccess) e = O.d_mapvox.reserve(vox_seg_bytes((int)tot_c) + vox_seg_bytes((int)tot_s));
    if (e == cudaSuccess) e = O.d_mapseg.reserve(sizeof(VoxSeg) * 2);
    if (e == cudaSuccess) e = O.d_in.reserve(((size_t)layout_step(fp.layout) * O.scap + 255) + sizeof(uint16_t) * (size_t)O.scap + 64);
    if (e != cudaSuccess) { lisreg_odom_destroy(ctx, slot); return fail(ctx, LISREG_ERR_CUDA, "lisreg_odom_create: buffer allocation failed: %s", cudaGetErrorString(e)); }
    O.map_id = map_alloc_slot(ctx);
    ctx->maps[O.map_id] = MapSlot();
    MapSlot& m = ctx->maps[O.map_id];
    m.used = true;                                  // reserved (empty) until the first rebuild
    const size_t cells = (size_t)ctx->max_cells + 1;
    for (CloudIndex* ci : {&m.corner, &m.surf}) {
      const size_t npts = ci == &m.corner ? tot_c : tot_s;
      if (cudaMalloc(&ci->sorted, sizeof(float4) * npts) == cudaSuccess) ci->cap_pts = npts; else ci->sorted = nullptr;
      if (cudaMalloc(&ci->cell_start, sizeof(uint32_t) * cells) == cudaSuccess) ci->cap_cells = cells; else ci->cell_start = nullptr;
    }
    ctx->maps_dirty = true;
  }
  *odom_id = slot;
  return LISREG_OK;
}

int32_t lisreg_odom_destroy(lisreg_ctx* ctx, int32_t odom_id) {
  if (!ctx || odom_id < 0 || odom_id >= (int)ctx->odoms.size() || !ctx->odoms[odom_id].used) return fail(ctx, LISREG_ERR_ARG, "lisreg_odom_destroy: bad id");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  lisreg_ctx::Odom& O = ctx->odoms[odom_id];
  if (O.gexec) cudaGraphExecDestroy(O.gexec);
  if (O.map_id >= 0) lisreg_map_destroy(ctx, O.map_id);
  cudaFree(O.d_win_c); cudaFree(O.d_win_s);
  for (DevBuf* b : {&O.d_cat, &O.d_mapvox, &O.d_mapseg, &O.d_in, &O.d_io}) b->release();
  O.h_io.release(); O.h_desc.release();
  O = lisreg_ctx::Odom();
  return LISREG_OK;
}

// updateInitialGuess, cloudInfo.odomAvailable == false branch (:298-313, :343-383)
static void odom_imu_T16(const lisreg_cloud_info* info, float T[16]) {
  const float p[6] = {info->imu_roll_init, info->imu_pitch_init, info->imu_yaw_init, 0.f, 0.f, 0.f};
  odom_T16(p, T);
}
// updateInitialGuess (:297-419).  info == NULL: no hints (odomAvailable = imuAvailable = false), first pose = init_pose6 or 0
static void odom_update_initial_guess(lisreg_ctx::Odom& O, const float* init_pose6, const lisreg_cloud_info* info) {
  if (!O.first_trans) {
    if (info) {
      O.pose[0] = info->imu_roll_init; O.pose[1] = info->imu_pitch_init;
      O.pose[2] = O.prm.use_imu_heading_initialization ? info->imu_yaw_init : 0.f;
      O.pose[3] = O.pose[4] = O.pose[5] = 0.f;
      odom_imu_T16(info, O.last_imu);
    } else {
      for (int i = 0; i < 6; i++) O.pose[i] = init_pose6 ? init_pose6[i] : 0.f;
    }
    O.first_trans = true;
    return;
  }
  float Tb[16], Tl[16], Tli[16], Ti[16], Tt[16], Tf[16];
  if (info && info->odom_available) {
    const float g[6] = {info->initial_guess[3], info->initial_guess[4], info->initial_guess[5],
                        info->initial_guess[0], info->initial_guess[1], info->initial_guess[2]};
    odom_T16(g, Tb);
    if (!O.have_imu_pre) {
      memcpy(O.last_imu_pre, Tb, sizeof(Tb)); O.have_imu_pre = true;     // falls through to the imuAvailable block (:327-330)
    } else {
      odom_inv(O.last_imu_pre, Tli); odom_mul(Tli, Tb, Ti);              // transIncre = lastImuPreTransformation.inverse() * transBack
      odom_T16(O.pose, Tt); odom_mul(Tt, Ti, Tf);                        // transFinal = transTobe * transIncre
      odom_euler(Tf, O.pose);
      memcpy(O.last_imu_pre, Tb, sizeof(Tb));
      odom_imu_T16(info, O.last_imu);
      return;
    }
  }
  if (!info || !info->odom_available) {
    if (!O.have_last) { memcpy(O.last_pose, O.pose, sizeof(O.pose)); O.have_last = true; return; }
    odom_T16(O.pose, Tb); odom_T16(O.last_pose, Tl);
    memcpy(O.last_pose, O.pose, sizeof(O.pose));
    odom_inv(Tl, Tli); odom_mul(Tli, Tb, Ti);       // transIncre = transLast.inverse() * transBack
    odom_mul(Tb, Ti, Tf);                           // transFinal = transTobe * transIncre
    odom_euler(Tf, O.pose);
    return;
  }
  if (info->imu_available) {                        // rotation increment of the IMU attitude (:394-417)
    odom_imu_T16(info, Tb);
    odom_inv(O.last_imu, Tli); odom_mul(Tli, Tb, Ti);
    odom_T16(O.pose, Tt); odom_mul(Tt, Ti, Tf);
    odom_euler(Tf, O.pose);
    memcpy(O.last_imu, Tb, sizeof(Tb));
  }
}

// transformUpdate (:976-1006).  tf (ROS geometry LinearMath) is not part of the reference tree; its Quaternion::setRPY /
// slerp / Matrix3x3::getRPY are written out here in double (tfScalar) for the two single-axis cases the function uses:
// for a pure roll (or pure pitch) pair the quaternions are (sin h, 0, 0, cos h), and getRPY of the slerp result reads the
// angle back through the rotation matrix exactly as tf does.
struct OdomQuat { double x, y, z, w; };
static OdomQuat odom_quat_rpy(double roll, double pitch, double yaw) {
  const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
  const double cy = cos(hy), sy = sin(hy), cp = cos(hp), sp = sin(hp), cr = cos(hr), sr = sin(hr);
  return {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy};
}
static double odom_quat_dot(const OdomQuat& a, const OdomQuat& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static OdomQuat odom_quat_slerp(const OdomQuat& a, const OdomQuat& b, double t) {
  const double s = sqrt(odom_quat_dot(a, a) * odom_quat_dot(b, b)), d = odom_quat_dot(a, b);
  const double theta = (d < 0 ? acos(-d / s) * 2.0 : acos(d / s) * 2.0) / 2.0;     // angleShortestPath(q) / 2
  if (theta == 0.0) return a;
  const double inv = 1.0 / sin(theta), s0 = sin((1.0 - t) * theta), s1 = sin(t * theta);
  const double sg = d < 0 ? -1.0 : 1.0;
  return {(a.x * s0 + sg * b.x * s1) * inv, (a.y * s0 + sg * b.y * s1) * inv, (a.z * s0 + sg * b.z * s1) * inv, (a.w * s0 + sg * b.w * s1) * inv};
}
static void odom_quat_get_rpy(const OdomQuat& q, double* roll, double* pitch, double* yaw) {
  const double s = 2.0 / odom_quat_dot(q, q);
  const double xs = q.x * s, ys = q.y * s, zs = q.z * s;
  const double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs, xx = q.x * xs, xy = q.x * ys, xz = q.x * zs, yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
  const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  if (fabs(m20) >= 1) {                                                // gimbal lock branch of getEulerYPR
    *yaw = 0; *roll = atan2(m21, m22); *pitch = m20 < 0 ? M_PI / 2.0 : -M_PI / 2.0;
  } else {
    *pitch = -asin(m20);
    const double c = cos(*pitch);
    *roll = atan2(m21 / c, m22 / c); *yaw = atan2(m10 / c, m00 / c);
  }
}
static float odom_clamp(float v, float lim) {      // constraintTransformation (src/core/common.cpp:286-292); lim <= 0 disables
  if (!(lim > 0.f)) return v;
  if (v < -lim) v = -lim;
  if (v > lim) v = lim;
  return v;
}
void lisreg_transform_update(const lisreg_cloud_info* info, float imu_rpy_weight, float rot_tolerance, float z_tolerance, float pose6[6]) {
  if (!pose6) return;
  if (info && info->imu_available && std::abs(info->imu_pitch_init) < 1.4) {
    const double w = imu_rpy_weight;
    double r, p, y;
    odom_quat_get_rpy(odom_quat_slerp(odom_quat_rpy(pose6[0], 0, 0), odom_quat_rpy(info->imu_roll_init, 0, 0), w), &r, &p, &y);
    pose6[0] = (float)r;
    odom_quat_get_rpy(odom_quat_slerp(odom_quat_rpy(0, pose6[1], 0), odom_quat_rpy(0, info->imu_pitch_init, 0), w), &r, &p, &y);
    pose6[1] = (float)p;
  }
  pose6[0] = odom_clamp(pose6[0], rot_tolerance);
  pose6[1] = odom_clamp(pose6[1], rot_tolerance);
  pose6[5] = odom_clamp(pose6[5], z_tolerance);
}
static void odom_transform_update(lisreg_ctx::Odom& O, const lisreg_cloud_info* info) {
  lisreg_transform_update(info, O.prm.imu_rpy_weight, O.rot_tol, O.z_tol, O.pose);
}

// saveKeyFrames (:421-478): the frame's full corner / surface clouds, moved by the refined pose, enter the window
static int odom_save_keyframe(lisreg_ctx* ctx, lisreg_ctx::Odom& O, const FeatFrame& f, int n_corner, int n_surf) {
  cudaStream_t st = ctx->cur->stream;
  // a free slot: the ring holds window + 1 slots and at most `window` are live after the trim below
  std::vector<char> live((size_t)O.slots, 0);
  for (int s : O.order) live[s] = 1;
  int slot = 0; while (slot < O.slots && live[slot]) slot++;
  if (slot >= O.slots) return fail(ctx, LISREG_ERR_CAPACITY, "odom window ring is full");
  float T[16]; odom_T16(O.pose, T);
  OdomT12 t12; for (int i = 0; i < 12; i++) t12.m[i] = T[i];
  const int nc = std::min(n_corner, O.ccap), ns = std::min(n_surf, O.scap);
  if (nc > 0) { k_odom_append<<<std::min(296, (nc + 255) / 256), 256, 0, st>>>(f.ext_pts, f.corner_idx, f.counts + 0, t12, O.d_win_c + (size_t)slot * O.ccap, O.ccap); LAUNCH_CK(); }
  if (ns > 0) { k_odom_append<<<std::min(592, (ns + 255) / 256), 256, 0, st>>>(f.ext_pts, f.surf_idx, f.counts + 3, t12, O.d_win_s + (size_t)slot * O.scap, O.scap); LAUNCH_CK(); }
  O.nc[slot] = nc; O.ns[slot] = ns;
  O.order.push_back(slot);
  while ((int)O.order.size() >= O.prm.window + 1) O.order.erase(O.order.begin());
  memcpy(O.key_pose, O.pose, sizeof(O.pose));
  O.keyframe_id++;
  O.map_dirty = true;
  return LISREG_OK;
}

// map = window concatenated newest first (:190-193) -> VoxelGrid (:196-201) -> spatial index (:602-603)
static int odom_rebuild_map(lisreg_ctx* ctx, lisreg_ctx::Odom& O) {
  cudaStream_t st = ctx->cur->stream;
  OdomConcat tc{}, ts{};
  int total_c = 0, total_s = 0, k = 0, max_c = 0, max_s = 0;
  for (int i = (int)O.order.size() - 1; i >= 0; i--, k++) {
    const int s = O.order[i];
    tc.src[k] = O.d_win_c + (size_t)s * O.ccap; tc.n[k] = O.nc[s]; tc.off[k] = total_c; total_c += O.nc[s]; max_c = std::max(max_c, O.nc[s]);
    ts.src[k] = O.d_win_s + (size_t)s * O.scap; ts.n[k] = O.ns[s]; ts.off[k] = total_s; total_s += O.ns[s]; max_s = std::max(max_s, O.ns[s]);
  }
  tc.count = ts.count = k;
  const size_t cat_c = (sizeof(float4) * (size_t)std::max(total_c, 1) + 255) & ~size_t(255);
  CK(O.d_cat.reserve(cat_c + sizeof(float4) * (size_t)std::max(total_s, 1)));
  float4* d_cc = (float4*)O.d_cat.p; float4* d_cs = (float4*)((char*)O.d_cat.p + cat_c);
  if (total_c > 0) { k_odom_concat<<<dim3(std::max(1, std::min(64, (max_c + 255) / 256)), k), 256, 0, st>>>(tc, d_cc); LAUNCH_CK(); }
  if (total_s > 0) { k_odom_concat<<<dim3(std::max(1, std::min(148, (max_s + 255) / 256)), k), 256, 0, st>>>(ts, d_cs); LAUNCH_CK(); }
  const size_t vc_per = vox_seg_bytes(std::max(total_c, 1)), vs_per = vox_seg_bytes(std::max(total_s, 1));
  CK(O.d_mapvox.reserve(vc_per + vs_per));
  CK(O.d_mapseg.reserve(sizeof(VoxSeg) * 2));
  VoxSeg seg[2];
  memset(seg, 0, sizeof(seg));
  vox_carve((char*)O.d_mapvox.p, std::max(total_c, 1), &seg[0]); vox_carve((char*)O.d_mapvox.p + vc_per, std::max(total_s, 1), &seg[1]);
  seg[0].src = d_cc; seg[0].gather = nullptr; seg[0].n_ptr = nullptr; seg[0].n = total_c; seg[0].leaf = O.prm.frame.corner_leaf;
  seg[1].src = d_cs; seg[1].gather = nullptr; seg[1].n_ptr = nullptr; seg[1].n = total_s; seg[1].leaf = O.prm.frame.surf_leaf;
  CK(cudaMemcpyAsync(O.d_mapseg.p, seg, sizeof(seg), cudaMemcpyHostToDevice, st));   // pageable source: consumed on return
  CK(cudaMemsetAsync(seg[0].out_n, 0, 4, st)); CK(cudaMemsetAsync(seg[1].out_n, 0, 4, st));
  int rc = run_voxel(ctx, (VoxSeg*)O.d_mapseg.p, 2, std::max(std::max(total_c, total_s), 1), 32.0 * (total_c + total_s));
  if (rc) return rc;
  if (O.map_id < 0) { O.map_id = map_alloc_slot(ctx); ctx->maps[O.map_id] = MapSlot(); }
  MapSlot& m = ctx->maps[O.map_id];
  const float h = cell_size_for_gate(O.prm.frame.lm.sqdist_gate);
  rc = build_cloud_index(ctx, seg[0].out, total_c, h, &m.corner, seg[0].out_n);
  if (rc) return rc;
  rc = build_cloud_index(ctx, seg[1].out, total_s, h, &m.surf, seg[1].out_n);
  if (rc) return rc;
  m.used = true;
  ctx->maps_dirty = true;
  O.n_map_c = m.corner.g.n; O.n_map_s = m.surf.g.n;
  O.map_dirty = false;
  return sync_maps(ctx);
}

constexpr int ODOM_LM_SLICE = 4;   // Gauss-Newton iterations per launch slice of the streaming odometry

// addresses a captured frame graph depends on: when one of them moves the graph is dropped and captured again
static std::vector<const void*> odom_graph_key(lisreg_ctx* ctx, lisreg_ctx::Odom& O) {
  WorkSet& w = *ctx->cur;
  return {w.d_descs.p, w.d_states.p, w.d_partials.p, w.d_nbr.p, w.d_kstate.p, w.d_klist.p, w.d_geom.p, w.d_feat.p, w.d_feat_frames.p,
          w.d_vox.p, w.d_vox_segs.p, w.d_feat_ctl.p, ctx->d_maps, O.d_io.p, O.h_desc.p};
}

static int odom_push_impl(lisreg_ctx* ctx, int32_t odom_id, const float* pts, const uint16_t* ring, int32_t n, bool on_device,
                          const float* init_pose6, const lisreg_cloud_info* info, float pose6[6], lisreg_odom_result* res) {
  if (!ctx || odom_id < 0 || odom_id >= (int)ctx->odoms.size() || !ctx->odoms[odom_id].used || n < 0 || (n > 0 && !pts) || !pose6 || !res)
    return fail(ctx, LISREG_ERR_ARG, "lisreg_odom_push: bad argument");
  CK(cudaSetDevice(ctx->device));
  lisreg_ctx::Odom& O = ctx->odoms[odom_id];
  const bool ring_arr = layout_needs_ring_array(O.prm.frame.feat.layout);
  if (n > 0 && ring_arr && !ring) return fail(ctx, LISREG_ERR_ARG, "lisreg_odom_push: ring array missing");
  cudaStream_t st = ctx->cur->stream;
  const lisreg_feat_params& fp = O.prm.frame.feat;
  const int cells = fp.n_scan * fp.horizon;
  memset(res, 0, sizeof(*res));
  // ---- the sweep ----
  const float4* d_pts = (const float4*)pts; const uint16_t* d_ring = ring;
  if (!on_device) {
    const size_t rec = (size_t)layout_step(O.prm.frame.feat.layout);
    const size_t bp = (rec * (size_t)n + 255) & ~size_t(255);
    CK(O.d_in.reserve(bp + sizeof(uint16_t) * (size_t)n + 64));
    if (n) {
      CK(cudaMemcpyAsync(O.d_in.p, pts, rec * (size_t)n, cudaMemcpyHostToDevice, st));
      if (ring_arr) CK(cudaMemcpyAsync((char*)O.d_in.p + bp, ring, sizeof(uint16_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    }
    d_pts = (const float4*)O.d_in.p; d_ring = ring_arr ? (const uint16_t*)((char*)O.d_in.p + bp) : nullptr;
  }
  odom_update_initial_guess(O, init_pose6, info);
  O.frame_id++;
  memcpy(res->guess, O.pose, sizeof(O.pose));
  // io block: [pose6 in/out][lm result][4 feature counts], device + pinned mirror
  const size_t o_res = 32, o_cnt = o_res + ((sizeof(lisreg_lm_result) + 15) & ~size_t(15)), io_bytes = o_cnt + 16;
  CK(O.d_io.reserve(io_bytes)); CK(O.h_io.reserve(io_bytes));
  CK(O.h_desc.reserve(frame_desc_bytes(1)));
  char* dio = (char*)O.d_io.p; char* hio = (char*)O.h_io.p;
  int rc = feat_reserve(ctx, 1, cells, fp.n_scan);
  if (rc) return rc;
  FeatFrame f{};
  feat_carve((char*)ctx->cur->d_feat.p, cells, fp.n_scan, &f);        // frame 0 of the work set: where run_frames puts it
  if (O.first_flag) {
    // FirstFlag (:175-183): features and the first key frame only, no registration
    f.pts = d_pts; f.ring = d_ring; f.n = n;
    CK(cudaMemcpyAsync(ctx->cur->d_feat_frames.p, &f, sizeof(f), cudaMemcpyHostToDevice, st));   // pageable: consumed on return
    rc = run_features(ctx, (FeatFrame*)ctx->cur->d_feat_frames.p, 1, &fp, n, 17.0 * n, false, true);
    if (rc) return rc;
    CK(cudaMemcpyAsync(hio + o_cnt, f.counts, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int* cnt = (const int*)(hio + o_cnt);
    rc = odom_save_keyframe(ctx, O, f, cnt[0], cnt[3]);
    if (rc) return rc;
    O.first_flag = false;
    res->keyframe_saved = 1;
  } else {
    if (O.map_dirty) { rc = odom_rebuild_map(ctx, O); if (rc) return rc; res->map_rebuilt = 1; }
    res->n_map_corner = O.n_map_c; res->n_map_surf = O.n_map_s;
    memcpy(hio, O.pose, sizeof(O.pose));
    CK(cudaMemcpyAsync(dio, hio, sizeof(O.pose), cudaMemcpyHostToDevice, st));
    lisreg_frame_item it; it.pts = (const float*)d_pts; it.ring = d_ring; it.n = n; it.map_id = O.map_id;
    const bool graph_ok = O.prm.use_graph && st != nullptr && !ctx->prof_on && n <= cells;
    bool launched = false;
    if (graph_ok && O.gexec && O.gkey == odom_graph_key(ctx, O)) {
      // replay: the sweep (address, size) reaches the kernels through the frame descriptor, which the graph uploads from
      // the pinned block - only that block's CONTENT changes; every address the graph itself holds is the captured one
      FeatFrame* hf = (FeatFrame*)O.h_desc.p;
      hf->pts = d_pts; hf->ring = d_ring; hf->n = n;
      cudaError_t e = cudaGraphLaunch(O.gexec, st);
      if (e != cudaSuccess) return fail(ctx, LISREG_ERR_CUDA, "cudaGraphLaunch failed: %s", cudaGetErrorString(e));
      ctx->launches += O.graph_kernels;
      launched = true;
    }
    if (!launched) {
      if (O.gexec) { cudaGraphExecDestroy(O.gexec); O.gexec = nullptr; }
      // capture needs every buffer at its final size: the first registered frame runs eagerly (sized for a full sweep),
      // the graph is captured on a later frame once nothing would have to be allocated inside the capture
      const bool sized = graph_ok && ctx->cur->feat_cap_frames >= 1 && O.gkey.size() && O.gkey == odom_graph_key(ctx, O);
      if (sized) {
        cudaGraph_t g = nullptr;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int64_t l0 = ctx->launches;
        rc = run_frames(ctx, 1, &it, nullptr, 0, (float*)dio, &O.prm.frame, (lisreg_lm_result*)(dio + o_res), (char*)O.h_desc.p, 0, cells, 0, ODOM_LM_SLICE);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        O.graph_kernels = ctx->launches - l0;                              // kernels one replay launches
        ctx->launches = l0;
        if (rc != LISREG_OK || e != cudaSuccess || !g) {
          if (g) cudaGraphDestroy(g);
          cudaGetLastError();
          O.prm.use_graph = 0;                                             // fall back to eager launches for good
          if (rc == LISREG_OK) rc = run_frames(ctx, 1, &it, nullptr, 0, (float*)dio, &O.prm.frame, (lisreg_lm_result*)(dio + o_res), (char*)O.h_desc.p, 0, cells, 0, ODOM_LM_SLICE);
          if (rc) return rc;
        } else {
          e = cudaGraphInstantiate(&O.gexec, g, 0);
          cudaGraphDestroy(g);
          if (e != cudaSuccess) { O.gexec = nullptr; return fail(ctx, LISREG_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
          e = cudaGraphLaunch(O.gexec, st);
          if (e != cudaSuccess) return fail(ctx, LISREG_ERR_CUDA, "cudaGraphLaunch failed: %s", cudaGetErrorString(e));
          ctx->launches += O.graph_kernels;
        }
      } else {
        rc = run_frames(ctx, 1, &it, nullptr, 0, (float*)dio, &O.prm.frame, (lisreg_lm_result*)(dio + o_res), (char*)O.h_desc.p, 0, cells, 0, ODOM_LM_SLICE);
        if (rc) return rc;
        O.gkey = odom_graph_key(ctx, O);
      }
    }
    CK(cudaMemcpyAsync(hio + o_res, dio + o_res, sizeof(lisreg_lm_result), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hio + o_cnt, f.counts, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const lisreg_lm_result* lr = (const lisreg_lm_result*)(hio + o_res);
    const int* cnt = (const int*)(hio + o_cnt);
    // The frame graph holds the first ODOM_LM_SLICE iterations (a 10 Hz / 100 Hz stream converges in 2-3); the rare frame that
    // needs more continues slice by slice with eager launches - same kernels, same order, so the result does not depend on it.
    const lisreg_lm_params& lp = O.prm.frame.lm;
    const int max_it = std::min(lp.max_iters, (int)LISREG_MAX_ITERS);
    while (lr->status != LISREG_NOT_ENOUGH_FEATURES && lr->iters < max_it && !(lr->converged && lp.early_exit)) {
      const int it0 = lr->iters;
      rc = run_lm(ctx, 1, (const RegDesc*)ctx->cur->d_descs.p, cells, 0.0, (float*)dio, &lp, (lisreg_lm_result*)(dio + o_res), nullptr, 0, it0,
                  it0 + ODOM_LM_SLICE);
      if (rc) return rc;
      CK(cudaMemcpyAsync(hio + o_res, dio + o_res, sizeof(lisreg_lm_result), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (lr->iters <= it0) break;                                   // no progress (cannot happen): never spin
    }
    res->lm = *lr;
    if (lr->status != LISREG_NOT_ENOUGH_FEATURES) {
      memcpy(O.pose, lr->pose, sizeof(O.pose));
      odom_transform_update(O, info);
      memcpy(res->lm.pose, O.pose, sizeof(O.pose));
      if (!(lr->deltaR == 100.f && lr->deltaT == 100.f)) { O.deltaR = lr->deltaR; O.deltaT = lr->deltaT; }   // members (:70-71) keep the last solved step
    }
    // key-frame rule (:216-229)
    if ((double)O.deltaR < 0.005 || (double)O.deltaT < 0.05) {
      float Tk[16], Tki[16], Tc[16], Ti[16], inc[6];
      odom_T16(O.key_pose, Tk); odom_T16(O.pose, Tc);
      odom_inv(Tk, Tki); odom_mul(Tki, Tc, Ti); odom_euler(Ti, inc);       // calculateTranslation (:284-295)
      if (O.keyframe_id <= 5 || fabsf(inc[2]) >= O.prm.keyframe_min_yaw || fabsf(inc[3]) >= O.prm.keyframe_min_distance ||
          fabsf(inc[4]) >= O.prm.keyframe_min_distance) {
        rc = odom_save_keyframe(ctx, O, f, cnt[0], cnt[3]);
        if (rc) return rc;
        res->keyframe_saved = 1;
      }
    }
  }
  memcpy(pose6, O.pose, sizeof(O.pose));
  res->frame_id = O.frame_id; res->keyframe_id = O.keyframe_id;
  return res->lm.status > 0 ? res->lm.status : LISREG_OK;
}

int32_t lisreg_odom_push(lisreg_ctx* ctx, int32_t odom_id, const float* pts, const uint16_t* ring, int32_t n,
                         const float* init_pose6, float pose6[6], lisreg_odom_result* res) {
  return odom_push_impl(ctx, odom_id, pts, ring, n, false, init_pose6, nullptr, pose6, res);
}
int32_t lisreg_odom_push_dev(lisreg_ctx* ctx, int32_t odom_id, const float* d_pts, const uint16_t* d_ring, int32_t n,
                             const float* init_pose6, float pose6[6], lisreg_odom_result* res) {
  return odom_push_impl(ctx, odom_id, d_pts, d_ring, n, true, init_pose6, nullptr, pose6, res);
}
int32_t lisreg_odom_push_info(lisreg_ctx* ctx, int32_t odom_id, const float* pts, const uint16_t* ring, int32_t n, int32_t on_device,
                              const lisreg_cloud_info* info, float pose6[6], lisreg_odom_result* res) {
  return odom_push_impl(ctx, odom_id, pts, ring, n, on_device != 0, nullptr, info, pose6, res);
}

// ------------------------------------------------------------------------------------------------
// device-resident local map / submap (T4, SURVEY.md 8f "next" #2)
// ------------------------------------------------------------------------------------------------
// exclusive scan in place over n_entries uint32 (three-phase block scan of the grid build)
static int scan_u32(lisreg_ctx* ctx, uint32_t* d_data, int n_entries, uint32_t* d_bsums) {
  cudaStream_t st = ctx->cur->stream;
  const int nblk = (n_entries + SCAN_BLOCK - 1) / SCAN_BLOCK;
  k_scan_local<<<nblk, SCAN_BLOCK, 0, st>>>(d_data, n_entries, d_bsums); LAUNCH_CK();
  k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(d_bsums, nblk); LAUNCH_CK();
  k_scan_add<<<nblk, SCAN_BLOCK, 0, st>>>(d_data, n_entries, d_bsums); LAUNCH_CK();
  return LISREG_OK;
}
// order-preserving compaction of d_in (n points) by d_flags (n + 1 entries, last 0; scanned in place) into d_out; *kept on the host
static int compact_points(lisreg_ctx* ctx, const float4* d_in, int n, uint32_t* d_flags, uint32_t* d_bsums, float4* d_out, int* kept) {
  cudaStream_t st = ctx->cur->stream;
  *kept = 0;
  if (n <= 0) return LISREG_OK;
  int rc = scan_u32(ctx, d_flags, n + 1, d_bsums);
  if (rc) return rc;
  k_sm_scatter<<<(n + 255) / 256, 256, 0, st>>>(d_in, n, d_flags, d_out); LAUNCH_CK();
  uint32_t total = 0;
  CK(cudaMemcpyAsync(&total, d_flags + n, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *kept = (int)total;
  return LISREG_OK;
}
static size_t sm_flag_bytes(int n) { return ((sizeof(uint32_t) * ((size_t)n + 1) + 255) & ~size_t(255)) + sizeof(uint32_t) * ((size_t)(n + 1) / SCAN_BLOCK + 2); }

int32_t lisreg_submap_create(lisreg_ctx* ctx, int32_t* submap_id) {
  if (!ctx || !submap_id) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_create: bad argument");
  int slot = -1;
  for (size_t i = 0; i < ctx->submaps.size(); i++) if (!ctx->submaps[i].used) { slot = (int)i; break; }
  if (slot < 0) { ctx->submaps.emplace_back(); slot = (int)ctx->submaps.size() - 1; }
  ctx->submaps[slot] = lisreg_ctx::Submap();
  ctx->submaps[slot].used = true;
  *submap_id = slot;
  return LISREG_OK;
}
#define SUBMAP_CK(name) if (!ctx || submap_id < 0 || submap_id >= (int)ctx->submaps.size() || !ctx->submaps[submap_id].used) \
  return fail(ctx, LISREG_ERR_ARG, name ": bad submap id")
int32_t lisreg_submap_destroy(lisreg_ctx* ctx, int32_t submap_id) {
  SUBMAP_CK("lisreg_submap_destroy");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (auto& b : ctx->submaps[submap_id].cls) b.release();
  if (ctx->submaps[submap_id].icp_slot >= 0) lisreg_map_destroy(ctx, ctx->submaps[submap_id].icp_slot);
  cudaFree(ctx->submaps[submap_id].dyn_index.sorted); cudaFree(ctx->submaps[submap_id].dyn_index.cell_start);
  ctx->submaps[submap_id] = lisreg_ctx::Submap();
  return LISREG_OK;
}
int32_t lisreg_submap_clear(lisreg_ctx* ctx, int32_t submap_id) {
  SUBMAP_CK("lisreg_submap_clear");
  lisreg_ctx::Submap& S = ctx->submaps[submap_id];
  for (int c = 0; c < 5; c++) S.n[c] = 0;
  for (int d = 0; d < 3; d++) { S.bmin[d] = 0; S.bmax[d] = 0; }
  S.icp_dirty = true;
  return LISREG_OK;
}
static void submap_fill_info(const lisreg_ctx::Submap& S, lisreg_submap_info* info) {
  if (!info) return;
  memset(info, 0, sizeof(*info));
  for (int c = 0; c < 5; c++) { info->n[c] = S.n[c]; info->feature_point_num += S.n[c]; }
  for (int d = 0; d < 3; d++) { info->bound_min[d] = S.bmin[d]; info->bound_max[d] = S.bmax[d]; }
}
// grows a class buffer to hold `want` points, keeping its content
static int submap_reserve(lisreg_ctx* ctx, lisreg_ctx::Submap& S, int c, size_t want) {
  if (sizeof(float4) * want <= S.cls[c].cap) return LISREG_OK;
  DevBuf nb;
  CK(nb.reserve(sizeof(float4) * (want + want / 2 + 1024)));
  if (S.n[c] > 0) {
    CK(cudaMemcpyAsync(nb.p, S.cls[c].p, sizeof(float4) * (size_t)S.n[c], cudaMemcpyDeviceToDevice, ctx->cur->stream));
    CK(cudaStreamSynchronize(ctx->cur->stream));
  }
  S.cls[c].release();
  S.cls[c] = nb;
  return LISREG_OK;
}

int32_t lisreg_submap_insert(lisreg_ctx* ctx, int32_t submap_id, const float* const pts[LISREG_SUBMAP_CLASSES], const int32_t n[LISREG_SUBMAP_CLASSES],
                             const float pose6[6], const lisreg_submap_insert_params* prm, lisreg_submap_info* info) {
  SUBMAP_CK("lisreg_submap_insert");
  if (!pts || !n || !pose6) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_insert: bad argument");
  for (int c = 0; c < 5; c++) if (n[c] < 0 || (n[c] > 0 && !pts[c])) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_insert: class %d bad", c);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->cur->stream;
  lisreg_ctx::Submap& S = ctx->submaps[submap_id];
  S.icp_dirty = true;
  float T[16]; odom_T16(pose6, T);
  OdomT12 t12; for (int i = 0; i < 12; i++) t12.m[i] = T[i];
  int feature_point_num = 0;
  for (int c = 0; c < 5; c++) feature_point_num += S.n[c];
  for (int c = 0; c < 5; c++) {
    if (n[c] == 0) continue;
    const size_t bytes = sizeof(float4) * (size_t)n[c];
    const size_t o_moved = (bytes + 255) & ~size_t(255), o_keep = 2 * o_moved, o_flags = o_keep + (((size_t)n[c] + 255) & ~size_t(255));
    CK(ctx->d_stage.reserve(o_flags + sm_flag_bytes(n[c])));
    char* d = (char*)ctx->d_stage.p;
    CK(cudaMemcpyAsync(d, pts[c], bytes, cudaMemcpyHostToDevice, st));
    int rc = submap_reserve(ctx, S, c, (size_t)S.n[c] + n[c]);
    if (rc) return rc;
    float4* dst = (float4*)S.cls[c].p + S.n[c];
    const bool filter = c == 0 && prm && prm->dynamic_removal_on && feature_point_num > prm->max_num_pts / 5 && n[c] > 10;   // subMap.h:980, :1069
    if (!filter) {
      k_sm_transform<<<std::min(592, (n[c] + 255) / 256), 256, 0, st>>>((const float4*)d, n[c], t12, dst); LAUNCH_CK();
      S.n[c] += n[c];
      continue;
    }
    // dynamic class: transform -> 1-NN distance test against the map's current dynamic cloud -> append the survivors
    float4* moved = (float4*)(d + o_moved);
    k_sm_transform<<<std::min(592, (n[c] + 255) / 256), 256, 0, st>>>((const float4*)d, n[c], t12, moved); LAUNCH_CK();
    CloudIndex& ci = S.dyn_index;
    rc = build_cloud_index(ctx, (const float4*)S.cls[0].p, S.n[0], cell_size_for_gate(1.0f), &ci);
    if (rc) return rc;
    const float dist_max = std::max(prm->dist_max, (float)((double)prm->dist_min + 0.1));             // :979
    const float near2 = prm->near_dist * prm->near_dist, dmin2 = prm->dist_min * prm->dist_min, dmax2 = dist_max * dist_max;
    float top = dmin2;
    if (std::isfinite(dmax2) && dmax2 > top) top = dmax2;
    if (near2 > top) top = near2;
    const float gate = top * 1.000001f + 1e-12f;
    unsigned char* keep = (unsigned char*)(d + o_keep);
    uint32_t* flags = (uint32_t*)(d + o_flags);
    uint32_t* bsums = (uint32_t*)(d + o_flags + ((sizeof(uint32_t) * ((size_t)n[c] + 1) + 255) & ~size_t(255)));
    k_map_distance_filter<<<(n[c] + 127) / 128, 128, 0, st>>>(ci.g, moved, n[c], prm->center_radius * prm->center_radius, near2, dmin2, dmax2, gate, keep); LAUNCH_CK();
    k_sm_widen_flags<<<(n[c] + 256) / 256, 256, 0, st>>>(keep, n[c], flags); LAUNCH_CK();
    int kept = 0;
    rc = compact_points(ctx, moved, n[c], flags, bsums, dst, &kept);
    if (rc) return rc;
    S.n[c] += kept;
  }
  // get_cloud_bbx over the five clouds (coordinates are floats: the double bounds upstream hold exactly these values)
  CK(ctx->d_bbox.reserve(6 * sizeof(unsigned) + sizeof(GridDev)));
  unsigned* bb = (unsigned*)ctx->d_bbox.p;
  k_bbox_init<<<1, 32, 0, st>>>(bb); LAUNCH_CK();
  bool any = false;
  for (int c = 0; c < 5; c++) if (S.n[c] > 0) { k_bbox<<<std::min(1184, (S.n[c] + 255) / 256), 256, 0, st>>>((const float4*)S.cls[c].p, S.n[c], nullptr, bb); LAUNCH_CK(); any = true; }
  unsigned hb[6];
  CK(cudaMemcpyAsync(hb, bb, sizeof(hb), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int dd = 0; dd < 3; dd++) {
    auto ord2f_h = [](unsigned u) { unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
    S.bmin[dd] = any ? (double)ord2f_h(hb[dd]) : 1.7976931348623157e308;       // DBL_MAX / -DBL_MAX for an empty map, as upstream
    S.bmax[dd] = any ? (double)ord2f_h(hb[3 + dd]) : -1.7976931348623157e308;
  }
  submap_fill_info(S, info);
  return LISREG_OK;
}

int32_t lisreg_submap_extract(lisreg_ctx* ctx, int32_t submap_id, const float cur_pose6[6], const float leaf[LISREG_SUBMAP_CLASSES], float gate_hint,
                              int32_t* map_id, lisreg_submap_info* info) {
  SUBMAP_CK("lisreg_submap_extract");
  if (!cur_pose6 || !map_id) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_extract: bad argument");
  if (*map_id >= 0 && (*map_id >= (int)ctx->maps.size() || !ctx->maps[*map_id].used)) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_extract: bad map id");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->cur->stream;
  lisreg_ctx::Submap& S = ctx->submaps[submap_id];
  S.icp_dirty = true;
  static const float kLeaf[5] = {0.1f, 0.05f, 0.4f, 0.2f, 0.6f};                      // subMapOptmizationNode.cpp:1393-1397
  const float* lf = leaf ? leaf : kLeaf;
  for (int c = 0; c < 5; c++) if (!(lf[c] > 0.f)) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_extract: leaf sizes must be > 0");
  // sensor box moved by the current pose (transform_bbx: float matrix entries, double coordinates), intersected with the map box
  float T[16]; odom_T16(cur_pose6, T);
  const double bmin[3] = {-70.0, -70.0, -10.0}, bmax[3] = {70.0, 70.0, 20.0};
  double cp[3], cpo[3];
  for (int d = 0; d < 3; d++) cp[d] = 0.5 * (bmin[d] + bmax[d]);
  for (int d = 0; d < 3; d++) cpo[d] = (double)T[4 * d] * cp[0] + (double)T[4 * d + 1] * cp[1] + (double)T[4 * d + 2] * cp[2] + (double)T[4 * d + 3];
  SmBox box;
  const float pad = 2.0f;
  for (int d = 0; d < 3; d++) {
    const double hi = bmax[d] - cp[d] + cpo[d], lo = bmin[d] - cp[d] + cpo[d];
    box.lo[d] = std::max(lo, S.bmin[d]) - pad; box.hi[d] = std::min(hi, S.bmax[d]) + pad;
  }
  // ---- voxel filter of every class (one batched call), then the box filter back into the class buffers ----
  int max_n = 0; size_t vox_total = 0; size_t vox_off[5];
  for (int c = 0; c < 5; c++) { vox_off[c] = vox_total; vox_total += vox_seg_bytes(std::max(S.n[c], 1)); max_n = std::max(max_n, S.n[c]); }
  CK(ctx->d_smvox.reserve(vox_total + sizeof(VoxSeg) * 5 + 256));
  VoxSeg seg[5];
  memset(seg, 0, sizeof(seg));
  for (int c = 0; c < 5; c++) {
    vox_carve((char*)ctx->d_smvox.p + vox_off[c], std::max(S.n[c], 1), &seg[c]);
    seg[c].src = (const float4*)S.cls[c].p; seg[c].gather = nullptr; seg[c].n_ptr = nullptr; seg[c].n = S.n[c]; seg[c].leaf = lf[c];
  }
  VoxSeg* d_seg = (VoxSeg*)((char*)ctx->d_smvox.p + ((vox_total + 255) & ~size_t(255)));
  CK(cudaMemcpyAsync(d_seg, seg, sizeof(seg), cudaMemcpyHostToDevice, st));
  for (int c = 0; c < 5; c++) CK(cudaMemsetAsync(seg[c].out_n, 0, 4, st));
  int rc = run_voxel(ctx, d_seg, 5, std::max(max_n, 1), 32.0 * (S.n[0] + S.n[1] + S.n[2] + S.n[3] + S.n[4]));
  if (rc) return rc;
  int vn[5];
  for (int c = 0; c < 5; c++) CK(cudaMemcpyAsync(&vn[c], seg[c].out_n, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int c = 0; c < 5; c++) {
    if (S.n[c] == 0) continue;                                                        // voxel_downsample_pcl leaves empty clouds alone
    const int m = vn[c];
    CK(ctx->d_stage.reserve(sm_flag_bytes(m) + 256));
    uint32_t* flags = (uint32_t*)ctx->d_stage.p;
    uint32_t* bsums = (uint32_t*)((char*)ctx->d_stage.p + ((sizeof(uint32_t) * ((size_t)m + 1) + 255) & ~size_t(255)));
    k_sm_box_flags<<<(m + 256) / 256, 256, 0, st>>>(seg[c].out, m, box, flags); LAUNCH_CK();
    int kept = 0;
    rc = compact_points(ctx, seg[c].out, m, flags, bsums, (float4*)S.cls[c].p, &kept);
    if (rc) return rc;
    S.n[c] = kept;
  }
  // ---- registration map: corner = pole, surf = ground + building + dynamic ----
  const int ns = S.n[2] + S.n[3] + S.n[0];
  CK(ctx->d_smcat.reserve(sizeof(float4) * (size_t)std::max(ns, 1)));
  size_t o = 0;
  for (int c : {2, 3, 0}) {
    if (S.n[c] > 0) CK(cudaMemcpyAsync((float4*)ctx->d_smcat.p + o, S.cls[c].p, sizeof(float4) * (size_t)S.n[c], cudaMemcpyDeviceToDevice, st));
    o += (size_t)S.n[c];
  }
  int slot = *map_id;
  if (slot < 0) { slot = map_alloc_slot(ctx); ctx->maps[slot] = MapSlot(); }
  MapSlot& m = ctx->maps[slot];
  const float h = cell_size_for_gate(gate_hint);
  rc = build_cloud_index(ctx, (const float4*)S.cls[1].p, S.n[1], h, &m.corner);
  if (rc) return rc;
  rc = build_cloud_index(ctx, (const float4*)ctx->d_smcat.p, ns, h, &m.surf);
  if (rc) return rc;
  m.used = true;
  ctx->maps_dirty = true;
  *map_id = slot;
  submap_fill_info(S, info);
  if (info) { info->n_map_corner = S.n[1]; info->n_map_surf = ns; }
  return sync_maps(ctx);
}

int32_t lisreg_submap_download(lisreg_ctx* ctx, int32_t submap_id, int32_t cls, float* out, int32_t cap, int32_t* n) {
  SUBMAP_CK("lisreg_submap_download");
  if (cls < 0 || cls >= 5 || !n) return fail(ctx, LISREG_ERR_ARG, "lisreg_submap_download: bad argument");
  lisreg_ctx::Submap& S = ctx->submaps[submap_id];
  *n = S.n[cls];
  if (!out || cap < S.n[cls] || S.n[cls] == 0) return LISREG_OK;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, S.cls[cls].p, sizeof(float4) * (size_t)S.n[cls], cudaMemcpyDeviceToHost, ctx->cur->stream));
  CK(cudaStreamSynchronize(ctx->cur->stream));
  return LISREG_OK;
}

// ------------------------------------------------------------------------------------------------
// sweep pre-treatment (SURVEY.md 8f "next" #3)
// ------------------------------------------------------------------------------------------------
int32_t lisreg_pretreat(lisreg_ctx* ctx, const float* pts, int32_t n, int32_t n_scan, double scan_period, float min_range, float max_range,
                        float* pts_out, uint16_t* ring_out, float* time_out, int32_t* n_out) {
  if (!ctx || n < 0 || (n > 0 && (!pts || !pts_out || !ring_out || !time_out)) || !n_out) return fail(ctx, LISREG_ERR_ARG, "lisreg_pretreat: bad argument");
  if (!(n_scan == 16 || n_scan == 32 || n_scan == 64)) return fail(ctx, LISREG_ERR_ARG, "lisreg_pretreat: N_SCAN must be 16, 32 or 64 (laserPretreatmentNode.cpp:98-126)");
  *n_out = 0;
  if (n == 0) return LISREG_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->cur->stream;
  const size_t bp = (sizeof(float4) * (size_t)n + 255) & ~size_t(255), bf = (sizeof(uint32_t) * ((size_t)n + 1) + 255) & ~size_t(255);
  const size_t o_cloud = bp, o_out = 2 * bp, o_flags = 3 * bp, o_bs = o_flags + bf, o_ring = o_bs + sm_flag_bytes(n), o_time = o_ring + ((2 * (size_t)n + 255) & ~size_t(255)),
               o_ori = o_time + ((4 * (size_t)n + 255) & ~size_t(255)), total = o_ori + 256;
  CK(ctx->d_stage.reserve(total));
  char* d = (char*)ctx->d_stage.p;
  CK(cudaMemcpyAsync(d, pts, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  uint32_t* flags = (uint32_t*)(d + o_flags); uint32_t* bsums = (uint32_t*)(d + o_bs);
  float4* cloud = (float4*)(d + o_cloud); float4* out = (float4*)(d + o_out);
  k_pt_flags_range<<<(n + 256) / 256, 256, 0, st>>>((const float4*)d, n, min_range, max_range, flags); LAUNCH_CK();
  int m = 0;
  int rc = compact_points(ctx, (const float4*)d, n, flags, bsums, cloud, &m);
  if (rc) return rc;
  if (m == 0) return LISREG_OK;
  PtOri* ori = (PtOri*)(d + o_ori);
  k_pt_ori<<<1, 1, 0, st>>>(cloud, m, ori); LAUNCH_CK();
  k_pt_ring_cond<<<(m + 256) / 256, 256, 0, st>>>(cloud, m, n_scan, ori, flags); LAUNCH_CK();
  rc = scan_u32(ctx, flags, m + 1, bsums);
  if (rc) return rc;
  k_pt_emit<<<(m + 255) / 256, 256, 0, st>>>(cloud, m, n_scan, scan_period, ori, flags, out, (uint16_t*)(d + o_ring), (float*)(d + o_time)); LAUNCH_CK();
  uint32_t cnt = 0;
  CK(cudaMemcpyAsync(&cnt, flags + m, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (cnt) {
    CK(cudaMemcpyAsync(pts_out, out, sizeof(float4) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(ring_out, d + o_ring, sizeof(uint16_t) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(time_out, d + o_time, sizeof(float) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  *n_out = (int32_t)cnt;
  return LISREG_OK;
}

int32_t lisreg_deskew_constant_velocity(lisreg_ctx* ctx, const float* pts, const float* time, int32_t n, float scan_period,
                                        const float lin_vel[3], const float ang_vel[3], float* out) {
  if (!ctx || n < 0 || (n > 1 && (!pts || !time || !out)) || !lin_vel || !ang_vel) return fail(ctx, LISREG_ERR_ARG, "lisreg_deskew_constant_velocity: bad argument");
  if (n <= 1) return LISREG_OK;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->cur->stream;
  const size_t bp = (sizeof(float4) * (size_t)n + 255) & ~size_t(255), bt = (sizeof(float) * (size_t)n + 255) & ~size_t(255);
  CK(ctx->d_stage.reserve(2 * bp + bt));
  char* d = (char*)ctx->d_stage.p;
  CK(cudaMemcpyAsync(d, pts, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d + bp, time, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, st));
  PtMotion mo; for (int k = 0; k < 3; k++) { mo.v[k] = lin_vel[k]; mo.w[k] = ang_vel[k]; } mo.scan_period = scan_period;
  k_deskew_cv<<<(n - 1 + 255) / 256, 256, 0, st>>>((const float4*)d, (const float*)(d + bp), n, mo, (float4*)(d + bp + bt)); LAUNCH_CK();
  CK(cudaMemcpyAsync(out, d + bp + bt, sizeof(float4) * (size_t)(n - 1), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return LISREG_OK;
}

// ------------------------------------------------------------------------------------------------
// loop-closure verification against candidate submaps (B4)
// ------------------------------------------------------------------------------------------------
// ICP target of a submap = dynamic + pole + ground + building (subMapOptmizationNode.cpp:2785-2790), indexed once per change
static int submap_icp_target(lisreg_ctx* ctx, lisreg_ctx::Submap& S, int* slot_out) {
  cudaStream_t st = ctx->cur->stream;
  if (S.icp_slot >= 0 && !S.icp_dirty) { *slot_out = S.icp_slot; return LISREG_OK; }
  const int nt = S.n[0] + S.n[1] + S.n[2] + S.n[3];
  CK(ctx->d_smcat.reserve(sizeof(float4) * (size_t)std::max(nt, 1)));
  size_t o = 0;
  for (int c = 0; c < 4; c++) {
    if (S.n[c] > 0) CK(cudaMemcpyAsync((float4*)ctx->d_smcat.p + o, S.cls[c].p, sizeof(float4) * (size_t)S.n[c], cudaMemcpyDeviceToDevice, st));
    o += (size_t)S.n[c];
  }
  if (S.icp_slot < 0) { S.icp_slot = map_alloc_slot(ctx); ctx->maps[S.icp_slot] = MapSlot(); ctx->maps[S.icp_slot].used = true; }
  MapSlot& m = ctx->maps[S.icp_slot];
  int rc = build_cloud_index(ctx, nullptr, 0, 1.0f, &m.corner);
  if (rc) return rc;
  rc = build_cloud_index(ctx, (const float4*)ctx->d_smcat.p, nt, cell_size_for_gate(4.0f), &m.surf);
  if (rc) return rc;
  m.used = true; ctx->maps_dirty = true; S.icp_dirty = false;
  *slot_out = S.icp_slot;
  return LISREG_OK;
}

int32_t lisreg_loop_verify(lisreg_ctx* ctx, const float* key_cloud, int32_t n, const float key_pose6[6], const float key_rel_pose6[6],
                           int32_t P, const lisreg_loop_candidate* cand, float fitness_threshold, const lisreg_icp_params* prm,
                           lisreg_loop_verify_result* out, lisreg_icp_result* per_candidate) {
  if (!ctx || n < 0 || (n > 0 && !key_cloud) || !key_pose6 || !key_rel_pose6 || P < 0 || (P > 0 && !cand) || !prm || !out || prm->max_iters <= 0)
    return fail(ctx, LISREG_ERR_ARG, "lisreg_loop_verify: bad argument");
  memset(out, 0, sizeof(*out));
  out->best = -1; out->best_score = 1.7976931348623157e308;
  if (P == 0) return LISREG_OK;
  for (int i = 0; i < P; i++)
    if (cand[i].submap_id < 0 || cand[i].submap_id >= (int)ctx->submaps.size() || !ctx->submaps[cand[i].submap_id].used)
      return fail(ctx, LISREG_ERR_ARG, "lisreg_loop_verify: candidate %d: bad submap id", i);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  // targets first (they use the staging buffers themselves)
  std::vector<int> tgt((size_t)P), ns((size_t)P, n);
  for (int i = 0; i < P; i++) { int rc = submap_icp_target(ctx, ctx->submaps[cand[i].submap_id], &tgt[i]); if (rc) return rc; }
  // sources: the key-frame cloud moved by each candidate's initial alignment, written straight into the ICP staging slots
  const size_t slot = al(sizeof(float4) * (size_t)std::max(n, 1));
  CK(ctx->d_stage.reserve(slot * ((size_t)P + 1) + 256));
  char* d = (char*)ctx->d_stage.p;
  float4* d_key = (float4*)(d + slot * (size_t)P);
  if (n) CK(cudaMemcpyAsync(d_key, key_cloud, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, st));
  std::vector<size_t> off((size_t)P);
  std::vector<float> K((size_t)P * 16);
  for (int i = 0; i < P; i++) {
    float* T = &K[16 * (size_t)i];
    if (cand[i].use_epsc_init) { float A[16]; odom_T16(cand[i].prekey_pose6, A); odom_mul(A, cand[i].epsc_T, T); }            // :2801-2802
    else { float A[16], Ai[16], B[16]; odom_T16(cand[i].submap_pose6, A); odom_T16(key_pose6, B); odom_inv(A, Ai); odom_mul(Ai, B, T); }   // :2807-2809
    OdomT12 t12; for (int k = 0; k < 12; k++) t12.m[k] = T[k];
    off[i] = slot * (size_t)i;
    if (n) { k_sm_transform<<<std::min(592, (n + 255) / 256), 256, 0, st>>>(d_key, n, t12, (float4*)(d + off[i])); LAUNCH_CK(); }
  }
  std::vector<lisreg_icp_result> res((size_t)P);
  int rc = icp_run_dev(ctx, P, off, ns, tgt, slot * (size_t)P, prm, res.data());
  if (rc) return rc;
  if (per_candidate) memcpy(per_candidate, res.data(), sizeof(lisreg_icp_result) * (size_t)P);
  for (int i = 0; i < P; i++) {                                        // :2835-2842
    if (!res[i].converged || res[i].fitness > out->best_score) continue;
    out->best_score = res[i].fitness; out->best = i;
  }
  if (out->best < 0) return LISREG_OK;
  memcpy(out->correction, res[out->best].T, sizeof(float) * 16);
  memcpy(out->key2pre, &K[16 * (size_t)out->best], sizeof(float) * 16);
  if (out->best_score > (double)fitness_threshold) return LISREG_OK;  // "loop not found" (:2855)
  float R[16], Ri[16], M[16];
  odom_T16(key_rel_pose6, R); odom_inv(R, Ri);                          // curSubMap2KeyTrans
  odom_mul(out->correction, out->key2pre, M); odom_mul(M, Ri, out->t_correct);   // tCorrect (:2876)
  float e[6]; odom_euler(out->t_correct, e);
  out->constraint6[0] = e[3]; out->constraint6[1] = e[4]; out->constraint6[2] = e[5];
  out->constraint6[3] = e[0]; out->constraint6[4] = e[1]; out->constraint6[5] = e[2];
  out->found = 1;
  return LISREG_OK;
}

int32_t lisreg_scan2map(lisreg_ctx* ctx, int32_t map_id, const float* corner, const uint16_t* clabel, int32_t nc,
                        const float* surf, const uint16_t* slabel, int32_t ns, float pose6[6], const lisreg_lm_params* prm,
                        lisreg_lm_result* res, lisreg_lm_iter* iter_log) {
  lisreg_batch_item it;
  it.corner = corner; it.clabel = clabel; it.surf = surf; it.slabel = slabel; it.nc = nc; it.ns = ns; it.map_id = map_id; it.reserved = 0;
  return lisreg_scan2map_batch(ctx, 1, &it, pose6, prm, res, iter_log);
}

}  // extern "C"
