/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header). PARITY UNPINNED.
 * Plain-C interface of the CPU restatement of the LIS-SLAM hot path, loaded with
 * ctypes by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline /
 * --impl reference).  Point clouds are packed float4 {x,y,z,intensity}. */
#ifndef ORC_API_H
#define ORC_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_LUT_SIZE 64

typedef struct orc_lm_params {
  int32_t max_iters;        /* 15 (A, odomEstimationNode.cpp:606) / 20 (B) / 30 (C) */
  int32_t early_exit;       /* 1 = reference behaviour; 0 = always run max_iters   */
  float sqdist_gate;        /* 1.0 (A :657,:776) / 2.0 (B/C subMapOptmizationNode.cpp:1610) */
  float conv_rot_deg;       /* 0.005 / 0.003 / 0.002 */
  float conv_trans_cm;      /* 0.05  / 0.03  / 0.02  */
  int32_t edge_min_valid;   /* edgeFeatureMinValidNum  (-1)  */
  int32_t surf_min_valid;   /* surfFeatureMinValidNum  (100) */
  int32_t min_sel;          /* 50, odomEstimationNode.cpp:870 */
  float degenerate_eig;     /* 100, :932 */
  int32_t use_label_weight; /* variants B/C: w = 2.0 - LabelSorce[label] */
  float label_score[ORC_LUT_SIZE];
  int32_t degenerate_in;    /* persistent isDegenerate member carried in (Q1) */
  float rot_tolerance;      /* transformUpdate clamps (:1001-1003); <=0 disables */
  float z_tolerance;
  int32_t n_threads;        /* 1 = as-built reference (OpenMP inert) */
} orc_lm_params;

typedef struct orc_lm_iter {
  float AtA[36];
  float AtB[6];
  float X[6];
  float pose[6];            /* after the update */
  int32_t n_sel, n_corner_sel, n_surf_sel, solved;
  float deltaR, deltaT;
} orc_lm_iter;

typedef struct orc_lm_result {
  int32_t status;           /* 0 ok, 1 not enough features (pose untouched), 2 some iter had <min_sel */
  int32_t iters;            /* LMOptimization calls made */
  int32_t converged;
  int32_t is_degenerate;
  int32_t n_sel_last;
  float deltaR, deltaT;
  double ms_build, ms_iters;
} orc_lm_result;

int orc_scan2map(const float* corner, const uint16_t* clabel, int32_t nc,
                 const float* surf, const uint16_t* slabel, int32_t ns,
                 const float* map_corner, int32_t mc, const float* map_surf, int32_t ms,
                 float pose6[6], const orc_lm_params* prm, orc_lm_result* res,
                 orc_lm_iter* iter_log /* nullable, max_iters entries */);

/* exact k-NN (FLANN KDTreeSingleIndex semantics: L2, sorted ascending) */
void* orc_kdtree_build(const float* pts4, int32_t n);
void orc_kdtree_free(void* t);
int32_t orc_kdtree_knn(const void* t, const float* q3, int32_t k, int32_t* idx, float* sqd);
void orc_knn_batch(const void* t, const float* q4, int32_t nq, int32_t k, int32_t* idx, float* sqd,
                   int32_t n_threads);

/* per-point coefficient functions exposed for unit tests */
int orc_corner_coeff(const float q[3], const float nb[15], float coeff[4]);
int orc_surf_coeff(const float q[3], const float nb[15], float coeff[4]);
void orc_pose_to_affine(const float pose6[6], float T12[12]);

/* small dense routines exposed so tests can pin them against cv2 */
void orc_jacobi_eigen_f32(const float* A, int32_t n, float* W, float* V);
int orc_qr_solve_f32(const float* A, int32_t n, const float* b, float* x);
int orc_lu_inv_f32(const float* A, int32_t n, float* Ainv);
void orc_plane_fit_5x3(const float* A15, float* x3);

/* ---- feature extraction (laserProcessing.cpp:467-713) ---- */
typedef struct orc_feat_params {
  int32_t n_scan, horizon, downsample_rate;
  float min_range, max_range;
  float edge_thr, surf_thr;
} orc_feat_params;

int32_t orc_project_scan(const float* pts4, const uint16_t* ring, int32_t n, const orc_feat_params* prm,
                         int32_t* src_index, int32_t* col_ind, float* range, int32_t* start_ring, int32_t* end_ring);
void orc_extract_features(const float* range, const int32_t* col_ind, int32_t M,
                          const int32_t* start_ring, const int32_t* end_ring, const orc_feat_params* prm,
                          int32_t* corner_idx, int32_t* n_corner, int32_t* sharp_idx, int32_t* n_sharp,
                          int32_t* flat_idx, int32_t* n_flat, int32_t* surf_idx, int32_t* n_surf,
                          float* curvature_out, int32_t* label_out);

/* ---- pcl::VoxelGrid<PointXYZI> centroid down-sampling (odomEstimationNode.cpp:196-201, :272-277) ---- */
int32_t orc_voxel_grid(const float* pts4, int32_t n, float leaf, float* out4, int32_t cap);

/* ---- EPSC descriptors + scoring (epscGeneration.cpp:478-607, :633-660) ---- */
void orc_epsc_describe(const float* corner4, int32_t nc, const float* surf4, int32_t ns,
                       const float* sem4, const uint16_t* sem_label, int32_t nsem, const uint8_t* using_map,
                       uint8_t* epsc, uint8_t* sepsc, uint8_t* fepsc);
double orc_epsc_distance(const uint8_t* d1, const uint8_t* d2, int32_t* best_shift, int32_t* min_sad);
void orc_epsc_score_all(const uint8_t* desc, int32_t N, int32_t topk, int32_t* idx, float* score, int8_t* shift, int32_t n_threads);

/* ---- loop-closure ICP verify (subMapOptmizationNode.cpp:2739-2916, pcl::IterativeClosestPoint) ---- */
typedef struct orc_icp_params { float max_corr_dist; int32_t max_iters; double trans_eps; double fitness_eps; } orc_icp_params;
typedef struct orc_icp_result { float T[16]; double fitness; int32_t converged, iters, n_corr_last, pad; } orc_icp_result;
int orc_icp(const float* src4, int32_t ns, const float* tgt4, int32_t nt, const orc_icp_params* prm, orc_icp_result* res);

/* ---- motion de-skew of the extracted points (laserProcessing.cpp:368-462, :501) ---- */
void orc_deskew(const float* pts4, const float* time, const int32_t* src_index, int32_t M,
                const double* imu_time, const double* imu_rot3, int32_t n_imu, double time_scan_cur, float* out4);

/* ---- map-based dynamic removal (subMap.h:1063-1098) ---- */
int32_t orc_map_distance_filter(const float* feat4, int32_t n, const float* map4, int32_t m, float center_radius,
                                float dyn_min, float dyn_max, float near_thre, uint8_t* keep);

/* ---- EPSC loop detector (epscGeneration.cpp:84-120 project, :258-401 globalICP, :663-992 loopDetection) ---- */
void* orc_loop_create(const uint8_t* using_map, int32_t use_epsc, int32_t use_sepsc, int32_t use_fepsc, int32_t use_pose);
void orc_loop_free(void* h);
int32_t orc_loop_detect(void* h, const float* corner4, int32_t nc, const float* surf4, int32_t ns,
                        const float* sem4, const uint16_t* sem_label, int32_t nsem, const float* odom16,
                        int32_t* current_id, int32_t* n_cand, int32_t* kinds, int32_t* ids, double* scores, float* T16s);
void orc_loop_project(const float* sem4, const uint16_t* label, int32_t n, float* out1440);
void orc_loop_global_icp(const float* proj1, const float* proj2, float yaw_diff, float* T16);

/* ---- sweep pre-treatment: ring / time synthesis (laserPretreatmentNode.cpp:60-230), constant-velocity de-skew (distortionAdjust.cpp:419-479) ---- */
int32_t orc_pretreat(const float* pts4, int32_t n, int32_t n_scan, double scan_period, float min_range, float max_range,
                     float* out4, uint16_t* ring_out, float* time_out);
int32_t orc_deskew_cv(const float* pts4, const float* time, int32_t n, float scan_period, const float* lin_vel3, const float* ang_vel3, float* out4);

/* ---- local-map / submap assembly (subMap.h:435-777, :957-1055; subMapOptmizationNode.cpp:1369-1432) ---- */
void* orc_submap_create();
void orc_submap_free(void* h);
void orc_submap_clear(void* h);
void orc_submap_insert(void* h, const float* const* pts, const int32_t* n, const float* pose6, int32_t dynrem_on, int32_t max_num_pts,
                       float center_radius, float dist_min, float dist_max, float near_dist, int32_t* counts, double* bound);
void orc_submap_extract(void* h, const float* cur_pose6, const float* leaf5, const double* map_bound, float* corner_out, int32_t* nc,
                        float* surf_out, int32_t* ns, int32_t* counts);
int32_t orc_submap_get(void* h, int32_t c, float* out, int32_t cap);
/* detectLoopClosureForSubMap (subMapOptmizationNode.cpp:2739-2916); orc_icp_params / orc_icp_result are declared above */
int32_t orc_loop_verify(const float* key4, int32_t n, const float* key_pose6, const float* key_rel_pose6, int32_t P, void* const* submaps,
                        const float* cand_pose, float fitness_threshold, const orc_icp_params* prm, int32_t* best, double* best_score,
                        float* correction16, float* key2pre16, float* t_correct16, float* constraint6, double* fitness_out, int32_t* conv_out);

/* ---- transformUpdate (odomEstimationNode.cpp:976-1006): IMU roll / pitch slerp (tf restated) + clamps ---- */
void orc_transform_update(float pose6[6], int32_t imu_available, float imu_roll, float imu_pitch, float imu_rpy_weight,
                          float rot_tol, float z_tol);

#ifdef __cplusplus
}
#endif
#endif
