/* lisreg — Blackwell-native (sm_100a) scan-to-map registration engine.
 *
 * C-ABI drop-in boundary for the ONE hot path of QingzhiWang/LIS-SLAM named by
 * BASELINE.json:north_star.  The reference has no FFI; its boundary is the set of
 * C++ member functions the ROS callbacks invoke.  Each entry point
 * below cites the reference interface it replaces (paths relative to the reference
 * repo root).  Plain pointers and sizes only; no CUDA/torch types in signatures
 * (a CUDA stream is passed as void*).  Point clouds are packed float4
 * {x, y, z, intensity} (16 B), labels/rings are separate uint16 arrays.
 *
 * Conventions: every call returns int32 status — 0 ok; >0 soft conditions that
 * mirror the reference (LISREG_NOT_ENOUGH_FEATURES: pose untouched,
 * odomEstimationNode.cpp:598,623-625; LISREG_FEW_CORRESPONDENCES: some iteration
 * had < min_sel matches and was a no-op, :870-872); <0 errors, text via
 * lisreg_last_error().  No exceptions cross the ABI.  One context must not be used
 * from two threads at once (the reference's scratch buffers are not re-entrant
 * either, odomEstimationNode.cpp:40-48); distinct contexts are independent.
 * There is NO CPU fallback: without a CUDA device every compute call fails with
 * LISREG_ERR_CUDA.
 */
#ifndef LISREG_H
#define LISREG_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LISREG_OK 0
#define LISREG_NOT_ENOUGH_FEATURES 1
#define LISREG_FEW_CORRESPONDENCES 2
#define LISREG_ERR_ARG (-1)
#define LISREG_ERR_CUDA (-2)
#define LISREG_ERR_CAPACITY (-3)

#define LISREG_LUT_SIZE 64
#define LISREG_MAX_ITERS 32

typedef struct lisreg_ctx lisreg_ctx;

typedef struct lisreg_config {
  int32_t device;        /* CUDA device ordinal */
  void* stream;          /* cudaStream_t to run on; NULL = the legacy default stream */
  int32_t max_grid_cells;/* per-cloud uniform-grid capacity (0 = default 4M cells) */
  int32_t own_stream;    /* 1: ignore `stream`, create a private non-blocking stream */
  int32_t reserved[4];
} lisreg_config;

/* Parameters of the three copies of the loop (variant A odomEstimationNode.cpp:596-974,
 * B subMapOptmizationNode.cpp:1509-2001, C :4485-4976).  lisreg_lm_params_preset fills
 * the reference constants. */
typedef struct lisreg_lm_params {
  int32_t max_iters;        /* 15 / 20 / 30 */
  int32_t early_exit;       /* 1 = reference; 0 = run exactly max_iters (benchmark mode) */
  float sqdist_gate;        /* 1.0 (A) / 2.0 (B, C): 5th-NN squared-distance gate */
  float conv_rot_deg;       /* 0.005 / 0.003 / 0.002 */
  float conv_trans_cm;      /* 0.05 / 0.03 / 0.02 */
  int32_t edge_min_valid;   /* edgeFeatureMinValidNum (-1) */
  int32_t surf_min_valid;   /* surfFeatureMinValidNum (100) */
  int32_t min_sel;          /* 50 */
  float degenerate_eig;     /* 100 */
  int32_t use_label_weight; /* B/C: w = 2.0 - LabelSorce[label] */
  float label_score[LISREG_LUT_SIZE];
  int32_t degenerate_in;    /* persistent isDegenerate member carried in (quirk Q1) */
  float rot_tolerance;      /* transformUpdate clamps, <= 0 disables */
  float z_tolerance;
  int32_t want_iter_log;    /* 1: fill lisreg_lm_iter records (debug/parity) */
} lisreg_lm_params;

typedef struct lisreg_lm_iter {
  float AtA[36];
  float AtB[6];
  float X[6];
  float pose[6];
  int32_t n_sel, n_corner_sel, n_surf_sel, solved;
  float deltaR, deltaT;
} lisreg_lm_iter;

typedef struct lisreg_lm_result {
  int32_t status;
  int32_t iters;
  int32_t converged;
  int32_t is_degenerate;
  int32_t n_sel_last;
  float deltaR, deltaT;
  float pose[6];
  int32_t n_corner, n_surf;   /* query points fed to the loop (after voxel down-sampling in the frame pipeline) */
} lisreg_lm_result;

/* one registration of a batch: host OR device pointers depending on the call */
typedef struct lisreg_batch_item {
  const float* corner;      /* nc x float4 */
  const uint16_t* clabel;   /* nullable */
  const float* surf;        /* ns x float4 */
  const uint16_t* slabel;   /* nullable */
  int32_t nc, ns;
  int32_t map_id;
  int32_t reserved;
} lisreg_batch_item;

/* ---- lifecycle ---- */
int32_t lisreg_create(const lisreg_config* cfg, lisreg_ctx** out);
void lisreg_destroy(lisreg_ctx* ctx);
const char* lisreg_last_error(const lisreg_ctx* ctx);
const char* lisreg_version(void);
int32_t lisreg_sync(lisreg_ctx* ctx);
/* number of engine kernels launched since create (bench.py "gpu_launches") */
int64_t lisreg_launch_count(const lisreg_ctx* ctx);

void lisreg_lm_params_preset(lisreg_lm_params* p, char variant /* 'A','B','C' */);

/* ---- local map + spatial index ----
 * Replaces kdtreeCornerFromMap/SurfFromMap->setInputCloud (odomEstimationNode.cpp:602-603;
 * B subMapOptmizationNode.cpp:1516-1517): uploads the two map clouds and builds the
 * device index (uniform grid, exact within the sqdist gate).  gate_hint = largest
 * sqdist_gate that will be used against this map (sets the cell size). */
int32_t lisreg_map_create(lisreg_ctx* ctx, const float* corner, int32_t mc,
                          const float* surf, int32_t ms, float gate_hint, int32_t* map_id);
int32_t lisreg_map_create_dev(lisreg_ctx* ctx, const float* d_corner, int32_t mc,
                              const float* d_surf, int32_t ms, float gate_hint, int32_t* map_id);
int32_t lisreg_map_destroy(lisreg_ctx* ctx, int32_t map_id);

/* exact k-NN against one cloud of a map (which: 0 corner, 1 surf), host buffers.
 * Replaces KdTreeFLANN::nearestKSearch (odomEstimationNode.cpp:650, :766) for tests:
 * idx/sqd are nq x 5, sorted ascending; entries beyond the gate are idx=-1, sqd=FLT_MAX. */
int32_t lisreg_knn5(lisreg_ctx* ctx, int32_t map_id, int32_t which, const float* queries,
                    int32_t nq, float sqdist_gate, int32_t* idx, float* sqd);

/* map-based dynamic-object removal of the local-map update (SURVEY.md 8f "next" #2, kernel part): replaces
 * map_scan_feature_pts_distance_removal (subMap.h:1063-1098; call site update_local_map :886-905, defaults
 * center_radius 30, dyn_min 0.3, dyn_max 3.0, near 0.03).  feat = n x float4 in the map frame; keep[i] = 1 when the
 * point survives (outside the centre disc, or nearest-map-point distance in (near, dyn_min) or beyond dyn_max);
 * survivors keep their order.  The map is the cloud `which` (0 corner, 1 surf) of a lisreg_map_create'd map. */
int32_t lisreg_map_distance_filter(lisreg_ctx* ctx, int32_t map_id, int32_t which, const float* feat, int32_t n,
                                   float center_radius, float dyn_min, float dyn_max, float near_thre,
                                   uint8_t* keep, int32_t* n_kept);

/* ---- scan-to-map registration (B2) ----
 * Replaces OdomEstimationNode::scan2SubMapOptimization() (odomEstimationNode.cpp:596-626)
 * and its two copies.  pose6 = transformTobeMapped [roll,pitch,yaw,x,y,z] in: guess, out: result. */
int32_t lisreg_scan2map(lisreg_ctx* ctx, int32_t map_id,
                        const float* corner, const uint16_t* clabel, int32_t nc,
                        const float* surf, const uint16_t* slabel, int32_t ns,
                        float pose6[6], const lisreg_lm_params* prm,
                        lisreg_lm_result* res, lisreg_lm_iter* iter_log /* nullable, max_iters */);

/* throughput mode (BASELINE config 3): B independent registrations, host buffers;
 * H2D of all clouds, the solve, and D2H of results happen inside the call. */
int32_t lisreg_scan2map_batch(lisreg_ctx* ctx, int32_t B, const lisreg_batch_item* items,
                              float* pose6xB, const lisreg_lm_params* prm,
                              lisreg_lm_result* resxB, lisreg_lm_iter* iter_log /* nullable, B*max_iters */);

/* same with every buffer already resident in HBM (items[] itself is a host array of
 * device pointers; d_pose6xB and d_resxB are device pointers).  Asynchronous on the
 * context stream; call lisreg_sync() before reading results. */
int32_t lisreg_scan2map_batch_dev(lisreg_ctx* ctx, int32_t B, const lisreg_batch_item* items,
                                  float* d_pose6xB, const lisreg_lm_params* prm,
                                  lisreg_lm_result* d_resxB);

/* same, with the frame packets packed by the caller into ONE contiguous (ideally pinned) host
 * arena: the corner/clabel/surf/slabel members of items[] are BYTE OFFSETS into the arena
 * (cast to pointers; (size_t)-1 = absent label array; offsets must be 16-byte aligned).  One
 * H2D copy of the arena, the solve, one D2H copy of the results.  This is the analogue of the
 * reference's serialised lis_slam::cloud_info packet (msg/cloud_info.msg) arriving in one buffer. */
int32_t lisreg_scan2map_batch_arena(lisreg_ctx* ctx, int32_t B, const lisreg_batch_item* items,
                                    const void* host_arena, uint64_t arena_bytes,
                                    float* pose6xB, const lisreg_lm_params* prm, lisreg_lm_result* resxB);

/* ---- LOAM feature extraction (B1) ----
 * Replaces LaserProcessing::projectPointCloud + cloudExtraction (laserProcessing.cpp:467-539, with the
 * de-skew step as identity: motion de-skew is out of scope) and LaserProcessing::featureExtraction()
 * (:108-115 -> calculateSmoothness :544, markOccludedPoints :568, extractFeatures :610); the duplicate class
 * FeatureExtraction (featureExtraction.cpp:68-80) maps to the same call.  Index lists refer to the
 * extracted (ring-major compacted) cloud and follow the reference push order.  All output arrays are
 * caller-allocated and optional (NULL = not wanted); capacities: src_index/col_ind/range/surf_idx/
 * curvature/label n_scan*horizon, start_ring/end_ring n_scan, corner_idx n_scan*120, sharp_idx n_scan*24,
 * flat_idx n_scan*60. */
/* Memory layout of a raw sweep, sensor_msgs/PointCloud2 style (SURVEY.md 8f "next" #4, rows T1 / T3): the engine reads
 * the caller's records in place - the 32-byte PCL PointXYZIRT records of cloud_info.cloud_deskewed (common.h:12-23:
 * x 0, y 4, z 8, intensity 16, ring 20 (uint16), time 24; laserProcessing.cpp:729-747), a bare xyz stream, ... - so no
 * adapter has to repack them.  point_step == 0 (the zero-initialised default) = packed float4 {x, y, z, intensity}
 * records (16 B) with the ring ids in the separate uint16 array of the call.
 * off_ring: >= 0 byte offset of a uint16 ring field inside the record; -1 separate ring array; -2 no ring input:
 * scanID is synthesised from the elevation angle exactly as laserPretreatmentNode.cpp:95-126 does for N_SCAN 16 / 32 /
 * 64 (points it would drop are dropped).  Offsets and point_step must be multiples of 4 (ring: of 2). */
typedef struct lisreg_cloud_layout {
  int32_t point_step;
  int32_t off_x, off_y, off_z;
  int32_t off_intensity;                      /* < 0: absent (intensity = 0) */
  int32_t off_ring;
  int32_t off_time;                           /* >= 0: float32 time field (de-skew entry point); < 0: separate array / absent */
  int32_t reserved;
} lisreg_cloud_layout;

typedef struct lisreg_feat_params {
  int32_t n_scan, horizon, downsample_rate;   /* N_SCAN 64, Horizon_SCAN 1800, downsampleRate */
  float min_range, max_range;                 /* lidarMinRange, lidarMaxRange */
  float edge_thr, surf_thr;                   /* edgeThreshold 1.0, surfThreshold 0.1 */
  int32_t reserved;
  lisreg_cloud_layout layout;                 /* how `pts` is laid out (all-zero = packed float4 + ring array) */
} lisreg_feat_params;
typedef struct lisreg_deskew {
  const double* imu_time;
  const double* imu_rot;
  int32_t n_imu;
  int32_t reserved;
  double time_scan_cur;
} lisreg_deskew;
/* presets: 0 packed float4 + ring array (16 + 2 B), 1 xyz float3 + ring array (12 + 2 B), 2 PCL PointXYZIRT records
 * (32 B, ring and time inside), 3 xyz float3 only, ring synthesised (12 B) */
void lisreg_cloud_layout_preset(lisreg_cloud_layout* l, int32_t which);

typedef struct lisreg_feat_out {
  int32_t n_extracted, n_corner, n_sharp, n_flat, n_surf;
  int32_t* src_index; int32_t* col_ind; float* range;
  int32_t* start_ring; int32_t* end_ring;
  int32_t* corner_idx; int32_t* sharp_idx; int32_t* flat_idx; int32_t* surf_idx;
  float* curvature; int32_t* label;
} lisreg_feat_out;

void lisreg_feat_params_default(lisreg_feat_params* p);
int32_t lisreg_extract_features(lisreg_ctx* ctx, const float* pts, const uint16_t* ring, int32_t n,
                                const lisreg_feat_params* prm, lisreg_feat_out* out);

/* Same with the per-point motion de-skew of projectPointCloud (SURVEY.md 8f "next" #3): replaces
 * LaserProcessing::deskewPoint / findRotation / findPosition (laserProcessing.cpp:368-462, call site :501).
 * imu_time / imu_rot are the rotation table that imuDeskewInfo integrates over the sweep (:213-262: imuTime[],
 * imuRotX/Y/Z[] interleaved, n_imu = imuPointerCur + 1 entries), time_scan_cur = timeScanCur, time[i] =
 * PointXYZIRT::time of input point i.  Range, column and every feature decision come from the ORIGINAL points
 * (as upstream); ext_xyzi (nullable, n_scan*horizon x float4) receives the de-skewed extracted cloud the index
 * lists refer to.  dsk == NULL or n_imu <= 0 = deskewFlag -1 / IMU unavailable (:429): points pass through. */
/* (lisreg_deskew is declared with the feature parameters above) */
int32_t lisreg_extract_features_deskew(lisreg_ctx* ctx, const float* pts, const uint16_t* ring, const float* time, int32_t n,
                                       const lisreg_feat_params* prm, const lisreg_deskew* dsk, lisreg_feat_out* out,
                                       float* ext_xyzi);

/* ---- sweep pre-treatment in front of the feature extractor (SURVEY.md 8f "next" #3) ----
 * lisreg_pretreat replaces the body of LaserPretreatment's cloud handler (laserPretreatmentNode.cpp:60-230, dup
 * src/core/laserPretreatment.cpp:20-160) for sensors that deliver bare x, y, z, intensity: removeNaNFromPointCloud,
 * removeClosedPointCloud(lidarMinRange, lidarMaxRange) (:244-272), PointXYZIRT::ring from the elevation angle (N_SCAN 16 /
 * 32 / 64, :95-126) and PointXYZIRT::time = scanPeriod * relTime from the azimuth (:128-141; the sequential halfPassed
 * flag is resolved with one atomicMin).  Outputs (capacity n each) keep the input order; *n_out = points kept.
 * lisreg_deskew_constant_velocity replaces DistortionAdjust::AdjustCloud (distortionAdjust.cpp:419-479): every point but
 * the first is moved by R(angular_rate * t) p + velocity * t with t = time - scan_period / 2; out holds n - 1 points. */
int32_t lisreg_pretreat(lisreg_ctx* ctx, const float* pts, int32_t n, int32_t n_scan, double scan_period, float min_range, float max_range,
                        float* pts_out, uint16_t* ring_out, float* time_out, int32_t* n_out);
int32_t lisreg_deskew_constant_velocity(lisreg_ctx* ctx, const float* pts, const float* time, int32_t n, float scan_period,
                                        const float lin_vel[3], const float ang_vel[3], float* out);

/* ---- voxel-grid down-sampling (F6) ----
 * Replaces pcl::VoxelGrid<PointType>::filter as used by downSizeFilterCorner/Surf
 * (odomEstimationNode.cpp:110-111, :196-201, :272-277): centroid per occupied voxel, ascending voxel
 * index.  out: caller-allocated n x float4 (worst case one voxel per point); *m receives the count. */
int32_t lisreg_voxel_grid(lisreg_ctx* ctx, const float* pts, int32_t n, float leaf, float* out, int32_t* m);

/* ---- whole-frame pipeline (B1 + F6 + B2 fused on the device) ----
 * One call = for every frame: projectPointCloud/cloudExtraction/featureExtraction (laserProcessing.cpp:467-713)
 * -> currentCloudInit voxel down-sampling of the corner/surface clouds (odomEstimationNode.cpp:260-281)
 * -> scan2SubMapOptimization (:596-626) against the frame's local map.  Nothing returns to the host
 * between the stages.  Frames are independent (throughput mode). */
typedef struct lisreg_frame_params {
  lisreg_feat_params feat;
  float corner_leaf, surf_leaf;     /* mappingCornerLeafSize 0.2, mappingSurfLeafSize 0.4 */
  lisreg_lm_params lm;
  /* NULL, or one entry per frame of the batch: motion de-skew of projectPointCloud (laserProcessing.cpp:427-462, :501) with
   * the frame's IMU rotation table (host pointers), exactly as lisreg_extract_features_deskew; the per-point time is read from
   * the records (feat.layout.off_time >= 0, e.g. PointXYZIRT).  Not used by lisreg_odom_* (leave NULL there). */
  const lisreg_deskew* deskew;
} lisreg_frame_params;

typedef struct lisreg_frame_item {
  const float* pts;        /* n x float4 raw sweep (device pointer, or byte offset into the arena) */
  const uint16_t* ring;    /* n ring ids */
  int32_t n;
  int32_t map_id;
} lisreg_frame_item;

void lisreg_frame_params_default(lisreg_frame_params* p);
/* everything resident in HBM; asynchronous with respect to the host and ordered on the context stream (work queued on
 * that stream before the call is complete before the batch starts, work queued after it starts after the batch).  A batch of
 * >= 64 frames runs internally as up to four sub-batches on private streams (LISREG_DEV_SPLIT, 0 = off) so that kernels of
 * different pipeline stages overlap; the results do not depend on the split. */
int32_t lisreg_frames_batch_dev(lisreg_ctx* ctx, int32_t F, const lisreg_frame_item* items,
                                float* d_pose6xF, const lisreg_frame_params* prm, lisreg_lm_result* d_resxF);
/* raw sweeps packed in one (pinned) host arena; items hold byte offsets (pts 16-byte aligned, ring 2-byte).  Blocking.
 * Batches larger than 128 frames whose sweeps are packed in frame order are uploaded in chunks on a copy stream while
 * the earlier chunks already run (LISREG_E2E_CHUNK frames per chunk, 0 = one copy); results do not depend on it. */
int32_t lisreg_frames_batch_arena(lisreg_ctx* ctx, int32_t F, const lisreg_frame_item* items,
                                  const void* host_arena, uint64_t arena_bytes,
                                  float* pose6xF, const lisreg_frame_params* prm, lisreg_lm_result* resxF);

/* Pipelined form of lisreg_frames_batch_arena for a continuous stream of batches (a node that receives sweeps
 * while the previous ones are still being registered): submit() enqueues the upload of the arena, the whole
 * pipeline and the download of the results on a private stream and returns a ticket at once; wait() blocks
 * until that batch is done and hands out its poses / results.  Two tickets can be in flight, so the PCIe upload
 * of batch k+1 overlaps the compute of batch k.  host_arena must stay valid (and should be pinned) until wait();
 * pose6xF is consumed before submit() returns.  Results are bit-identical to lisreg_frames_batch_arena. */
int32_t lisreg_frames_batch_submit(lisreg_ctx* ctx, int32_t F, const lisreg_frame_item* items,
                                   const void* host_arena, uint64_t arena_bytes, const float* pose6xF,
                                   const lisreg_frame_params* prm, int32_t* ticket);
int32_t lisreg_frames_batch_wait(lisreg_ctx* ctx, int32_t ticket, float* pose6xF, lisreg_lm_result* resxF);

/* ---- streaming odometry: device-resident sliding-window local map (SURVEY.md 8f "next" #1; BASELINE configs[1], [4]) ----
 * One lisreg_odom = the state OdomEstimationNode keeps between frames (odomEstimationNode.cpp:56-71, 80-100):
 * transformTobeMapped, lastTransformTobeMapped, transformPriFrame, keyFrameId, deltaR / deltaT and the vectors
 * laserCloudCornerVec / laserCloudSurfVec of the last <= window key-frame clouds (already in the map frame) - the
 * latter resident in HBM.  lisreg_odom_push = laserCloudInfoHandler (:163-239) for one sweep without IMU / odometry
 * input (cloudInfo.odomAvailable == false): updateInitialGuess (constant-velocity model, :353-391) -> map = the
 * window concatenated newest first + VoxelGrid (:185-207; rebuilt only when a key frame was added - the reference
 * rebuilds the identical map every frame) -> projectPointCloud .. extractFeatures (laserProcessing.cpp:467-713) ->
 * currentCloudInit (:260-281) -> scan2SubMapOptimization (:596-626) -> key-frame rule (:216-229) -> saveKeyFrames
 * (:421-478).  Nothing but the sweep goes to the device and nothing but the pose / result record comes back.
 * use_graph = 1 replays the per-frame kernel sequence as ONE CUDA graph (needs a context with a non-default stream:
 * lisreg_config.own_stream = 1 or an explicit stream; ignored otherwise); results are identical.
 * Frame t needs the pose and the map of frame t-1: this mode does not shard (multi-GPU = independent replicas). */
typedef struct lisreg_odom_params {
  lisreg_frame_params frame;        /* feature extraction, leaf sizes 0.2 / 0.4, loop constants (variant A) */
  float keyframe_min_distance;      /* keyFrameMiniDistance 1.4 (config/params.yaml:140) */
  float keyframe_min_yaw;           /* keyFrameMiniYaw 0.5 (:141) */
  int32_t window;                   /* 19: while (laserCloudSurfVec.size() >= 20) erase(begin) (:463-467) */
  int32_t use_graph;
  int32_t use_imu_heading_initialization;   /* useImuHeadingInitialization (config/params.yaml:77): keep imuYawInit on the first frame */
  float imu_rpy_weight;             /* imuRPYWeight 0.01 (utility.h:405): slerp weight of transformUpdate (:982-997) */
} lisreg_odom_params;
/* the scalar hints of lis_slam::cloud_info (msg/cloud_info.msg:4-19) that updateInitialGuess (:297-419) and transformUpdate
 * (:976-999) read; the clouds of the message are what lisreg_odom_push extracts itself */
typedef struct lisreg_cloud_info {
  int32_t imu_available, odom_available;
  float imu_roll_init, imu_pitch_init, imu_yaw_init;
  float initial_guess[6];           /* initialGuessX, Y, Z, Roll, Pitch, Yaw (message order) */
} lisreg_cloud_info;
typedef struct lisreg_odom_result {
  lisreg_lm_result lm;              /* status / iterations / deltaR / deltaT of this frame's loop (zeros for the first frame) */
  int32_t frame_id, keyframe_id;    /* frames pushed so far, keyFrameId after this frame */
  int32_t keyframe_saved, map_rebuilt;
  int32_t n_map_corner, n_map_surf; /* laserCloudCornerFromMapDSNum / SurfFromMapDSNum of the map this frame was registered to */
  float guess[6];                   /* transformTobeMapped after updateInitialGuess */
} lisreg_odom_result;
void lisreg_odom_params_default(lisreg_odom_params* p);
int32_t lisreg_odom_create(lisreg_ctx* ctx, const lisreg_odom_params* prm, int32_t* odom_id);
int32_t lisreg_odom_destroy(lisreg_ctx* ctx, int32_t odom_id);
/* host buffers (pinned memory makes the upload asynchronous); init_pose6 (nullable) = the pose of the FIRST frame
 * (imuRollInit / PitchInit / YawInit + origin upstream), ignored afterwards */
int32_t lisreg_odom_push(lisreg_ctx* ctx, int32_t odom_id, const float* pts, const uint16_t* ring, int32_t n,
                         const float* init_pose6, float pose6[6], lisreg_odom_result* res);
/* the same frame step with the cloud_info hints (info != NULL): first frame = [imuRollInit, imuPitchInit, imuYawInit or 0];
 * odomAvailable: pose *= lastImuPreTransformation^-1 * initialGuess (:322-349); otherwise the constant-velocity guess
 * (:353-391); imuAvailable on the frame that first sees odomAvailable: rotation increment of the IMU attitude (:394-417);
 * after the loop, imuAvailable and |imuPitchInit| < 1.4 pull roll / pitch toward the IMU by a tf slerp (transformUpdate).
 * on_device != 0: pts / ring are device pointers.  info == NULL behaves as lisreg_odom_push with init_pose6 == NULL. */
int32_t lisreg_odom_push_info(lisreg_ctx* ctx, int32_t odom_id, const float* pts, const uint16_t* ring, int32_t n, int32_t on_device,
                              const lisreg_cloud_info* info, float pose6[6], lisreg_odom_result* res);
/* transformUpdate on its own (odomEstimationNode.cpp:976-1006; the same body at subMapOptmizationNode.cpp:1969-1996 and
 * :4968-4995 for a caller that drives lisreg_scan2map variant 'B' / 'C' itself): IMU roll / pitch slerp when
 * info->imu_available and |imu_pitch_init| < 1.4, then the roll / pitch / z clamps (tolerance <= 0 disables).  Host arithmetic
 * on six floats, no device work.  info may be NULL (clamps only). */
void lisreg_transform_update(const lisreg_cloud_info* info, float imu_rpy_weight, float rot_tolerance, float z_tolerance,
                             float pose6[6]);
/* same with the sweep already resident in HBM */
int32_t lisreg_odom_push_dev(lisreg_ctx* ctx, int32_t odom_id, const float* d_pts, const uint16_t* d_ring, int32_t n,
                             const float* init_pose6, float pose6[6], lisreg_odom_result* res);

/* ---- device-resident local map / submap (rows T4 + 8f "next" #2) ----
 * One lisreg_submap = the five class clouds of localMap_t / submap_t (subMap.h:435-777: submap_dynamic, _pole, _ground,
 * _building, _outlier; class index 0..4 in that order) kept in HBM together with their bounding box.
 * lisreg_submap_insert = SubMapManager::insert_local_map / insert_submap / fisrt_submap (subMap.h:785-1055): the key
 * frame's class clouds are moved by its pose (transformPointCloud, common.cpp:134-160: optimized_pose for the local map,
 * relative_pose for a submap), the DYNAMIC class optionally goes through the map-based dynamic-object removal against the
 * map's current dynamic cloud (map_scan_feature_pts_distance_removal, :1001-1017, only when feature_point_num >
 * max_num_pts / 5), everything is appended (append_feature :742-753) and the box of all five clouds is refreshed
 * (get_cloud_bbx_cpt :131-171).
 * lisreg_submap_extract = SubMapOdometryNode::extractSlidingCloud (subMapOptmizationNode.cpp:1369-1432; the submap copy
 * :3976-4081): every class is voxel-filtered IN PLACE (leaf 0.1 / 0.05 / 0.4 / 0.2 / 0.6), box-filtered IN PLACE with the
 * sensor box (+-70, +-70, -10..20 moved by cur_pose6) intersected with the map box padded by 2 m (strict inequalities,
 * subMap.h:1125-1150), then laserCloudCornerFromSubMap = pole and laserCloudSurfFromSubMap = ground + building + dynamic
 * become a registration map (*map_id, usable with lisreg_scan2map variant 'B' / 'C'; pass the previous id to rebuild
 * that map in place, -1 for a new one).  Map-side labels are not kept: the loop reads the label of the query point only. */
#define LISREG_SUBMAP_CLASSES 5
typedef struct lisreg_submap_insert_params {
  int32_t dynamic_removal_on;       /* map_based_dynamic_removal_on */
  int32_t max_num_pts;              /* 20000 */
  float center_radius, dist_min, dist_max, near_dist;   /* 30, 0.3, 3.0, 0.03 */
} lisreg_submap_insert_params;
typedef struct lisreg_submap_info {
  int32_t n[LISREG_SUBMAP_CLASSES]; /* points per class after the call */
  int32_t feature_point_num;
  double bound_min[3], bound_max[3];/* localMap->bound */
  int32_t n_map_corner, n_map_surf; /* extract: size of the registration map */
} lisreg_submap_info;
int32_t lisreg_submap_create(lisreg_ctx* ctx, int32_t* submap_id);
int32_t lisreg_submap_destroy(lisreg_ctx* ctx, int32_t submap_id);
int32_t lisreg_submap_clear(lisreg_ctx* ctx, int32_t submap_id);                     /* localMap_t::free() */
int32_t lisreg_submap_insert(lisreg_ctx* ctx, int32_t submap_id, const float* const pts[LISREG_SUBMAP_CLASSES],
                             const int32_t n[LISREG_SUBMAP_CLASSES], const float pose6[6],
                             const lisreg_submap_insert_params* prm /* NULL = no dynamic removal */, lisreg_submap_info* info);
int32_t lisreg_submap_extract(lisreg_ctx* ctx, int32_t submap_id, const float cur_pose6[6], const float leaf[LISREG_SUBMAP_CLASSES] /* NULL = reference */,
                              float gate_hint, int32_t* map_id, lisreg_submap_info* info);
/* copies class cloud `cls` to the host (parity tests / visualisation); *n = its size (may exceed cap: nothing is copied then) */
int32_t lisreg_submap_download(lisreg_ctx* ctx, int32_t submap_id, int32_t cls, float* out, int32_t cap, int32_t* n);

/* ---- EPSC loop-closure descriptors and scoring (B3 pieces) ----
 * lisreg_epsc_describe replaces EPSCGeneration::calculateEPSC / calculateSEPSC / calculateFEPSC
 * (epscGeneration.cpp:478-607) for n submaps/keyframes at once; using_map is the 256-entry label -> class
 * LUT of config/label.yaml using_label (SemanticLabelParam::UsingLableMap, utility.h:138).  Outputs are
 * n x 1600 bytes each (20 rings x 80 sectors, row-major), any may be NULL. */
typedef struct lisreg_epsc_cloud {
  const float* corner; const float* surf; const float* sem; const uint16_t* sem_label;
  int32_t nc, ns, nsem, reserved;
} lisreg_epsc_cloud;
int32_t lisreg_epsc_describe(lisreg_ctx* ctx, int32_t n, const lisreg_epsc_cloud* clouds, const uint8_t using_map[256],
                             uint8_t* epsc, uint8_t* sepsc, uint8_t* fepsc);
/* lisreg_epsc_score_all replaces the per-candidate EPSCGeneration::calculateDistance loop of loopDetection
 * (epscGeneration.cpp:633-660, :736-860): every descriptor q is scored against its history j < q over the
 * 20 column shifts; per query the topk (<= 8) candidates with score > DISTANCE_THRESHOLD (0.75) are returned,
 * best first (idx = -1 when fewer qualify); shift = winning i in [-10, 10) (yaw offset = i * 2*pi/80). */
int32_t lisreg_epsc_score_all(lisreg_ctx* ctx, const uint8_t* desc, int32_t N, int32_t topk,
                              int32_t* idx, float* score, int8_t* shift);
/* the shard of one rank of a multi-GPU run: only the query rows q = row_begin + r * row_stride are
 * scored (row q costs q pairs, so ranks take the cyclic rows rank, rank + world, ...); outputs are n_rows x topk with
 * n_rows = ceil((N - row_begin) / row_stride), row r describing query q.  (0, 1) == lisreg_epsc_score_all. */
int32_t lisreg_epsc_score_rows(lisreg_ctx* ctx, const uint8_t* desc, int32_t N, int32_t row_begin, int32_t row_stride,
                               int32_t topk, int32_t* idx, float* score, int8_t* shift);
/* same with descriptors and outputs resident in HBM; asynchronous */
int32_t lisreg_epsc_score_all_dev(lisreg_ctx* ctx, const uint8_t* d_desc, int32_t N, int32_t topk,
                                  int32_t* d_idx, float* d_score, int8_t* d_shift);

/* ---- EPSC loop detector (B3) ----
 * Replaces EPSCGeneration::loopDetection(corner, surf, semantic, odom) (epscGeneration.h:155-158,
 * epscGeneration.cpp:663-992) together with project() (:84-120) and globalICP() (:258-401): one stateful,
 * append-only detector per EPSCGeneration instance.  Every call gates the stored keyframes by travelled distance
 * (SKIP_NEIBOUR_DISTANCE, INFLATION_COVARIANCE), aligns the current 360-sector projection to every gated
 * candidate (shift search + default-parameter 2-D ICP), re-describes the moved clouds (EPSC / SEPSC / FEPSC),
 * scores them against the stored descriptors, then appends the current keyframe.  Clouds are in the sensor
 * frame; odom is the row-major 4x4 world pose.  The result mirrors the public members current_frame_id,
 * matched_frame_id[] and matched_frame_transform[] (:121-123) in the reference push order
 * (EPSC, SEPSC, FEPSC, POSE); ISC / SC / SSC descriptors are not part of this path. */
typedef struct lisreg_loop_params {
  int32_t use_epsc, use_sepsc, use_fepsc, use_pose;   /* UsingEPSCFlag .. UsingPoseFlag (config/params.yaml:22-28) */
  float skip_neighbour_distance;                      /* 20   */
  float inflation_covariance;                         /* 0.01 */
  float distance_threshold;                           /* 0.75 */
  int32_t reserved;
} lisreg_loop_params;
typedef struct lisreg_loop_match {
  int32_t kind;        /* 0 EPSC, 1 SEPSC, 2 FEPSC, 3 POSE */
  int32_t frame_id;    /* matched_frame_id */
  double score;        /* descriptor score (POSE: the position distance) */
  float T[16];         /* matched_frame_transform, row-major 4x4 */
} lisreg_loop_match;
typedef struct lisreg_loop_result {
  int32_t current_frame_id, n_candidates, n_matched, reserved;
  lisreg_loop_match match[4];
} lisreg_loop_result;
void lisreg_loop_params_default(lisreg_loop_params* p);
int32_t lisreg_loop_create(lisreg_ctx* ctx, const lisreg_loop_params* prm, const uint8_t using_map[256], int32_t* det_id);
int32_t lisreg_loop_destroy(lisreg_ctx* ctx, int32_t det_id);
int32_t lisreg_loop_detect(lisreg_ctx* ctx, int32_t det_id, const float* corner, int32_t nc, const float* surf, int32_t ns,
                           const float* sem, const uint16_t* sem_label, int32_t nsem, const float odom[16],
                           lisreg_loop_result* res);

/* ---- loop-closure ICP verification (B4) ----
 * Replaces the pcl::IterativeClosestPoint block of SubMapOdometryNode::detectLoopClosureForSubMap
 * (subMapOptmizationNode.cpp:2739-2916; settings :2763-2769): for each (keyframe cloud, candidate submap cloud)
 * pair, point-to-point ICP from the pre-transformed source (:2822-2824), then getFitnessScore(); the caller
 * keeps the best converged pair and accepts it iff fitness <= historyKeyframeFitnessScore (0.5, :2855).
 * The target cloud is registered once with lisreg_map_create(ctx, NULL, 0, target, n, ...) (its "surf" cloud).
 * T is the final 4x4 transformation (row-major) of the pre-transformed source, as getFinalTransformation(). */
typedef struct lisreg_icp_params {
  float max_corr_dist;     /* 10   */
  int32_t max_iters;       /* 30   */
  double trans_eps;        /* 1e-4 */
  double fitness_eps;      /* 1e-4 */
} lisreg_icp_params;
typedef struct lisreg_icp_pair { const float* src; int32_t ns; int32_t target_id; } lisreg_icp_pair;
typedef struct lisreg_icp_result { float T[16]; double fitness; int32_t converged, iters, n_corr_last, reserved; } lisreg_icp_result;
void lisreg_icp_params_default(lisreg_icp_params* p);
int32_t lisreg_icp_verify_batch(lisreg_ctx* ctx, int32_t P, const lisreg_icp_pair* pairs, const lisreg_icp_params* prm,
                                lisreg_icp_result* out);

/* ---- multi-GPU exchange: one NCCL all-gather of fixed-size result records ----
 * Frames and loop-closure candidate pairs are independent units sharded one process per GPU; the ONLY exchange step
 * of the path is an all-gather of the per-rank result blocks (6-DoF poses + status: 40 B per frame).  NCCL is bound
 * at run time (dlopen of libnccl.so.2 - the copy already loaded in the process when there is one), so liblisreg.so
 * has no link-time NCCL dependency.  Either adopt an existing ncclComm_t of the host program or let the context
 * create its own: rank 0 calls lisreg_comm_unique_id(), the 128 bytes travel out of band (ROS parameter, file,
 * MPI, torch.distributed ...), every rank calls lisreg_comm_init().
 * lisreg_allgather_results enqueues ncclAllGather(d_send -> d_recv, bytes_per_rank per rank) on a private stream
 * that first waits for everything enqueued so far on the context stream, and returns at once: the kernels of the
 * next batch do not wait for the collective.  lisreg_allgather_wait(ctx, back) blocks the host until the gather issued
 * `back` calls before the most recent one has landed (0 = the last one; with double-buffered send / receive blocks a
 * producer waits with back = 1 before reusing a buffer, without draining the step still in flight). */
#define LISREG_NCCL_ID_BYTES 128
int32_t lisreg_comm_unique_id(uint8_t id[LISREG_NCCL_ID_BYTES]);
int32_t lisreg_comm_init(lisreg_ctx* ctx, int32_t world, int32_t rank, const uint8_t id[LISREG_NCCL_ID_BYTES]);
int32_t lisreg_comm_adopt(lisreg_ctx* ctx, void* nccl_comm /* ncclComm_t */, int32_t world, int32_t rank);
int32_t lisreg_comm_destroy(lisreg_ctx* ctx);
int32_t lisreg_allgather_results(lisreg_ctx* ctx, const void* d_send, void* d_recv, uint64_t bytes_per_rank);
int32_t lisreg_allgather_wait(lisreg_ctx* ctx, int32_t back);

/* ---- loop-closure verification against candidate submaps (B4: the whole of detectLoopClosureForSubMap) ----
 * Replaces SubMapOdometryNode::detectLoopClosureForSubMap (subMapOptmizationNode.cpp:2739-2916) on top of the
 * device-resident submaps: for every candidate submap i the initial alignment key2PreSubMapTrans is composed
 * (EPSC: T(keyframe_poses_6D_map[loopKeyPreLast[i]]) * curKey2PreKeyInitTrans[i], :2797-2803; pose-based:
 * T(submap_pose_6D_optimized)^-1 * T(cur_keyframe->optimized_pose), :2807-2809), the key-frame cloud (dynamic + pole +
 * ground + building, sensor frame) is moved by it (transformPointCloud, :2822-2824) and ICP-aligned to the candidate's
 * dynamic + pole + ground + building cloud (:2785-2790, settings :2763-2769) - all candidates in one batch; the best
 * converged candidate by getFitnessScore wins (:2835-2842) and is accepted iff its score <= fitness_threshold
 * (historyKeyframeFitnessScore, :2855); the loop constraint tCorrect = correctionLidarFrame * key2PreSubMapTrans *
 * T(cur_keyframe->relative_pose)^-1 and its X, Y, Z, ROLL, PITCH, YAW (:2874-2878) are returned (poseFrom of the GTSAM
 * between-factor; GTSAM itself stays upstream).  The ICP target index of a submap is cached until the submap changes. */
typedef struct lisreg_loop_candidate {
  int32_t submap_id;            /* lisreg_submap of loopSubMapPre[i] */
  int32_t use_epsc_init;        /* 1: EPSC initial pose (:2797-2803); 0: pose-based (:2807-2809) */
  float prekey_pose6[6];        /* keyframe_poses_6D_map[loopKeyPreLast[i]]  [roll, pitch, yaw, x, y, z] */
  float epsc_T[16];             /* curKey2PreKeyInitTrans[i], row-major 4x4 (lisreg_loop_match.T) */
  float submap_pose6[6];        /* submap_pose_6D_optimized */
} lisreg_loop_candidate;
typedef struct lisreg_loop_verify_result {
  int32_t found;                /* the reference's return value */
  int32_t best;                 /* bestID: index into the candidate array, -1 = no candidate converged */
  double best_score;            /* bestScore */
  float correction[16];         /* correctionLidarFrame (getFinalTransformation of the best candidate) */
  float key2pre[16];            /* key2PreSubMapTrans of the best candidate */
  float t_correct[16];          /* tCorrect */
  float constraint6[6];         /* X, Y, Z, ROLL, PITCH, YAW of tCorrect */
} lisreg_loop_verify_result;
int32_t lisreg_loop_verify(lisreg_ctx* ctx, const float* key_cloud, int32_t n, const float key_pose6[6], const float key_rel_pose6[6],
                           int32_t P, const lisreg_loop_candidate* cand, float fitness_threshold, const lisreg_icp_params* prm,
                           lisreg_loop_verify_result* out, lisreg_icp_result* per_candidate /* nullable, P entries */);

/* device self-test of the small dense routines (cv::eigen / cv::solve(QR) / cv::Mat::inv restatements):
 * out98 = E[6], V[36] (eigenvectors in rows), X[6] (QR solve of A x = b), ok, Ainv[36] (LU), ok,
 * then W3[3], V3[9] of the register-only 3x3 Jacobi applied to the leading 3x3 block of A */
int32_t lisreg_selftest_smallmat(lisreg_ctx* ctx, const float* A36, const float* b6, float* out98);

/* integer-ALU roofline of lisreg_epsc_score_* (the all-pairs SAD is bound by VABSDIFF4 + funnel-shift
 * issue, not by HBM): measured rate of the kernel's inner-loop instruction mix on registers only, in 1e9 VABSDIFF4
 * instructions per second (one instruction = 4 byte abs-diff-accumulates). */
int32_t lisreg_selftest_alu_peak(lisreg_ctx* ctx, double* gsad_per_s);

/* ---- profiling (CUDA events on the context stream around the dominant kernels) ---- */
typedef struct lisreg_profile {
  double lm_iter_ms;        /* total device time of the Gauss-Newton iterations (kNN check / search, residual, solve kernels) */
  int64_t lm_iter_launches; /* iterations timed */
  double lm_alg_bytes;      /* algorithmic bytes of those launches: sum (nc+ns) * 96 B (SURVEY.md 8d A_iter) */
  double feat_ms;  int64_t feat_launches;  double feat_alg_bytes;
  double voxel_ms; int64_t voxel_launches; double voxel_alg_bytes;
  double index_ms; int64_t index_launches; double index_alg_bytes;
} lisreg_profile;
int32_t lisreg_profile_enable(lisreg_ctx* ctx, int32_t on);
/* synchronises the stream, accumulates all pending event pairs, returns totals since the last reset */
int32_t lisreg_profile_get(lisreg_ctx* ctx, lisreg_profile* out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* LISREG_H */
