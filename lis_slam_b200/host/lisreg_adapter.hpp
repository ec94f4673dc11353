// Header-only C++ adapter: keeps LIS-SLAM's own call surface (pcl::PointCloud<PointXYZI / PointXYZIL / PointXYZIRT>,
// float transformTobeMapped[6]) and forwards to the lisreg C-ABI (include/lisreg.h).
//
// It mirrors the member functions the reference's ROS callbacks invoke:
//   LaserProcessing::featureExtraction()          src/core/laserProcessing.cpp:108-115      -> Registrar::featureExtraction
//   OdomEstimationNode::scan2SubMapOptimization() src/node/odomEstimationNode.cpp:596-626   -> Registrar::scan2SubMapOptimization
//     (+ variants B/C src/node/subMapOptmizationNode.cpp:1509, :4485 via the `variant` argument)
//   kdtree*FromMap->setInputCloud                 odomEstimationNode.cpp:602-603            -> Registrar::setMap
//   pcl::VoxelGrid::filter                        odomEstimationNode.cpp:196-201, :272-277  -> Registrar::voxelGrid
//   EPSCGeneration::calculateFEPSC / calculateDistance  src/core/epscGeneration.cpp:591, :633 -> Registrar::describe / scoreAll
//
//   OdomEstimationNode::laserCloudInfoHandler     odomEstimationNode.cpp:163-239            -> Odometry::push (device-resident window)
//
// The only thing it needs from PCL is the record layout: PCL points are 32-byte records
// {x, y, z, 1.0f pad, intensity, ...} (common.h:9-35).  Raw sweeps - the big clouds - are handed to the engine IN
// PLACE through a lisreg_cloud_layout (PointXYZIRT records, or the `data` blob of a sensor_msgs/PointCloud2 such as
// cloud_info.cloud_deskewed with the offsets of its `fields`, msg/cloud_info.msg:17-25): no repacking.  The small
// down-sampled feature / map clouds of the scan2SubMapOptimization call are repacked to float4 + uint16.
// Define LISREG_ADAPTER_MOCK_PCL to compile it without PCL / ROS (unit tests in this repo do that).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <array>
#include <vector>

#include "../../include/lisreg.h"

#ifdef LISREG_ADAPTER_MOCK_PCL
namespace lisreg_mock {
struct alignas(16) PointXYZI { float x, y, z, pad; float intensity; float pad2[3]; };
struct alignas(16) PointXYZIL { float x, y, z, pad; float intensity; uint32_t label; float pad2[2]; };
struct alignas(16) PointXYZIRT { float x, y, z, pad; float intensity; uint16_t ring; float time; float pad2; };
template <typename P> struct PointCloud { std::vector<P> points; size_t size() const { return points.size(); } };
// sensor_msgs/PointField + PointCloud2, the members the adapter reads (datatype 7 = FLOAT32, 4 = UINT16)
struct PointField { std::string name; uint32_t offset; uint8_t datatype; uint32_t count; };
struct PointCloud2 { uint32_t height = 1, width = 0; std::vector<PointField> fields; uint32_t point_step = 0, row_step = 0; std::vector<uint8_t> data; };
}  // namespace lisreg_mock
#define LISREG_POINTCLOUD2 lisreg_mock::PointCloud2
#define LISREG_CLOUD(P) lisreg_mock::PointCloud<lisreg_mock::P>
#else
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
#define LISREG_CLOUD(P) pcl::PointCloud<P>
#define LISREG_POINTCLOUD2 sensor_msgs::PointCloud2
#endif

namespace lisreg_host {

struct Packed {
  std::vector<float> xyzi;        // n x 4
  std::vector<uint16_t> aux;      // label or ring
  int32_t n = 0;
};

template <typename CloudT> inline Packed pack_xyzi(const CloudT& c) {
  Packed p; p.n = (int32_t)c.points.size(); p.xyzi.resize(4 * (size_t)p.n);
  for (int32_t i = 0; i < p.n; i++) { const auto& q = c.points[i]; float* o = &p.xyzi[4 * (size_t)i]; o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.intensity; }
  return p;
}
template <typename CloudT> inline Packed pack_xyzil(const CloudT& c) {
  Packed p = pack_xyzi(c); p.aux.resize(p.n);
  for (int32_t i = 0; i < p.n; i++) p.aux[i] = (uint16_t)c.points[i].label;
  return p;
}
template <typename CloudT> inline Packed pack_xyzirt(const CloudT& c) {
  Packed p = pack_xyzi(c); p.aux.resize(p.n);
  for (int32_t i = 0; i < p.n; i++) p.aux[i] = (uint16_t)c.points[i].ring;
  return p;
}

// lisreg_cloud_layout of an in-memory PCL PointXYZIRT cloud (32-byte records: x 0, y 4, z 8, intensity 16, ring 20, time 24)
inline lisreg_cloud_layout layout_xyzirt() { lisreg_cloud_layout l; lisreg_cloud_layout_preset(&l, 2); return l; }

// lisreg_cloud_layout of a sensor_msgs/PointCloud2 (e.g. cloud_info.cloud_deskewed, laserProcessing.cpp:729-747) from its
// `fields`: x / y / z / intensity must be FLOAT32, ring UINT16, time FLOAT32.  The engine then reads msg.data in place.
template <typename Msg> inline lisreg_cloud_layout layout_of(const Msg& msg) {
  lisreg_cloud_layout l; std::memset(&l, 0, sizeof(l));
  l.point_step = (int32_t)msg.point_step; l.off_x = l.off_y = l.off_z = -1; l.off_intensity = -1; l.off_ring = -2; l.off_time = -1;
  for (const auto& f : msg.fields) {
    const bool f32 = f.datatype == 7, u16 = f.datatype == 4;
    if (f.name == "x" && f32) l.off_x = (int32_t)f.offset;
    else if (f.name == "y" && f32) l.off_y = (int32_t)f.offset;
    else if (f.name == "z" && f32) l.off_z = (int32_t)f.offset;
    else if ((f.name == "intensity" || f.name == "i") && f32) l.off_intensity = (int32_t)f.offset;
    else if (f.name == "ring" && u16) l.off_ring = (int32_t)f.offset;
    else if ((f.name == "time" || f.name == "t") && f32) l.off_time = (int32_t)f.offset;
  }
  if (l.off_x < 0 || l.off_y < 0 || l.off_z < 0) throw std::runtime_error("PointCloud2 without float32 x / y / z fields");
  return l;   // off_ring == -2 (no ring field): scanID is synthesised from the elevation angle, like laserPretreatmentNode.cpp:95-126
}

// One engine context per node process (not re-entrant, exactly like the member scratch buffers it replaces).
class Registrar {
 public:
  explicit Registrar(int device = 0) {
    lisreg_config cfg; std::memset(&cfg, 0, sizeof(cfg)); cfg.device = device;
    if (lisreg_create(&cfg, &ctx_) != LISREG_OK) throw std::runtime_error("lisreg_create failed: no CUDA device (there is no CPU fallback)");
  }
  ~Registrar() { if (ctx_) lisreg_destroy(ctx_); }
  Registrar(const Registrar&) = delete;
  Registrar& operator=(const Registrar&) = delete;

  // kdtreeCornerFromMap->setInputCloud(cornerMapDS); kdtreeSurfFromMap->setInputCloud(surfMapDS)
  template <typename CloudT> void setMap(const CloudT& cornerMapDS, const CloudT& surfMapDS, float sqdist_gate = 1.0f) {
    if (map_id_ >= 0) { lisreg_map_destroy(ctx_, map_id_); map_id_ = -1; }
    Packed c = pack_xyzi(cornerMapDS), s = pack_xyzi(surfMapDS);
    check(lisreg_map_create(ctx_, c.xyzi.data(), c.n, s.xyzi.data(), s.n, sqdist_gate, &map_id_));
  }

  // scan2SubMapOptimization(): transformTobeMapped is [roll, pitch, yaw, x, y, z] in/out.
  // Returns the reference's soft conditions: 0 ok, 1 "Not enough features" (pose untouched), 2 some iteration had < 50 matches.
  // Variant A (odometry, PointXYZI clouds):
  template <typename CloudT>
  int scan2SubMapOptimization(const CloudT& cornerLast, const CloudT& surfLast, float transformTobeMapped[6], lisreg_lm_result* out = nullptr) {
    return solve(pack_xyzi(cornerLast), pack_xyzi(surfLast), transformTobeMapped, 'A', out);
  }
  // Variants B (scan-to-submap) / C (submap-to-submap) on PointXYZIL clouds: w = 2.0 - LabelSorce[label]
  template <typename CloudT>
  int scan2SubMapOptimizationLabelled(const CloudT& cornerLastDS, const CloudT& surfLastDS, float transformTobeSubMapped[6], char variant,
                                      lisreg_lm_result* out = nullptr) {
    return solve(pack_xyzil(cornerLastDS), pack_xyzil(surfLastDS), transformTobeSubMapped, variant, out);
  }

  // featureExtraction() on a raw sweep (ring ids from PointXYZIRT); index lists refer to `extracted`.
  template <typename CloudT>
  void featureExtraction(const CloudT& laserCloudIn, const lisreg_feat_params& prm, std::vector<int32_t>& extracted_src,
                         std::vector<int32_t>& corner, std::vector<int32_t>& surface, std::vector<int32_t>& sharpCorner, std::vector<int32_t>& sharpSurface) {
    static_assert(sizeof(laserCloudIn.points[0]) == 32, "PointXYZIRT is a 32-byte record");
    lisreg_feat_params q = prm; q.layout = layout_xyzirt();            // the records are read where they are: no repack
    featureExtractionRaw(laserCloudIn.points.data(), (int32_t)laserCloudIn.points.size(), q, extracted_src, corner, surface, sharpCorner, sharpSurface);
  }
  // the same on the `data` blob of a sensor_msgs/PointCloud2 (cloud_info.cloud_deskewed): zero-copy
  template <typename Msg>
  void featureExtractionMsg(const Msg& msg, const lisreg_feat_params& prm, std::vector<int32_t>& extracted_src,
                            std::vector<int32_t>& corner, std::vector<int32_t>& surface, std::vector<int32_t>& sharpCorner, std::vector<int32_t>& sharpSurface) {
    lisreg_feat_params q = prm; q.layout = layout_of(msg);
    featureExtractionRaw(msg.data.data(), (int32_t)(msg.point_step ? msg.data.size() / msg.point_step : 0), q, extracted_src, corner, surface, sharpCorner, sharpSurface);
  }
  void featureExtractionRaw(const void* records, int32_t n, const lisreg_feat_params& prm, std::vector<int32_t>& extracted_src,
                            std::vector<int32_t>& corner, std::vector<int32_t>& surface, std::vector<int32_t>& sharpCorner, std::vector<int32_t>& sharpSurface) {
    const size_t cells = (size_t)prm.n_scan * prm.horizon;
    extracted_src.assign(cells, 0); corner.assign((size_t)prm.n_scan * 120, 0); surface.assign(cells, 0);
    sharpCorner.assign((size_t)prm.n_scan * 24, 0); sharpSurface.assign((size_t)prm.n_scan * 60, 0);
    lisreg_feat_out o; std::memset(&o, 0, sizeof(o));
    o.src_index = extracted_src.data(); o.corner_idx = corner.data(); o.surf_idx = surface.data(); o.sharp_idx = sharpCorner.data(); o.flat_idx = sharpSurface.data();
    check(lisreg_extract_features(ctx_, (const float*)records, nullptr, n, &prm, &o));
    extracted_src.resize(o.n_extracted); corner.resize(o.n_corner); surface.resize(o.n_surf); sharpCorner.resize(o.n_sharp); sharpSurface.resize(o.n_flat);
  }

  // downSizeFilter.setInputCloud(in); downSizeFilter.filter(out)   (xyz + intensity centroids, ascending voxel index)
  template <typename CloudT> void voxelGrid(const CloudT& in, float leaf, std::vector<float>& out_xyzi) {
    Packed p = pack_xyzi(in); out_xyzi.assign(4 * (size_t)p.n + 4, 0.f); int32_t m = 0;
    check(lisreg_voxel_grid(ctx_, p.xyzi.data(), p.n, leaf, out_xyzi.data(), &m));
    out_xyzi.resize(4 * (size_t)m);
  }

  // calculateFEPSC for one keyframe (20 x 80 bytes)
  template <typename CloudI, typename CloudL>
  void describe(const CloudI& corner, const CloudI& surf, const CloudL& semantic, const uint8_t using_map[256], uint8_t fepsc[1600]) {
    Packed c = pack_xyzi(corner), s = pack_xyzi(surf), m = pack_xyzil(semantic);
    lisreg_epsc_cloud cl; std::memset(&cl, 0, sizeof(cl));
    cl.corner = c.xyzi.data(); cl.nc = c.n; cl.surf = s.xyzi.data(); cl.ns = s.n; cl.sem = m.xyzi.data(); cl.sem_label = m.aux.data(); cl.nsem = m.n;
    check(lisreg_epsc_describe(ctx_, 1, &cl, using_map, nullptr, nullptr, fepsc));
  }

  // calculateDistance of every descriptor against its history (top-k loop candidates per keyframe)
  void scoreAll(const uint8_t* desc, int32_t n, int32_t topk, std::vector<int32_t>& idx, std::vector<float>& score, std::vector<int8_t>& shift) {
    idx.assign((size_t)n * topk, -1); score.assign((size_t)n * topk, 0.f); shift.assign((size_t)n * topk, 0);
    check(lisreg_epsc_score_all(ctx_, desc, n, topk, idx.data(), score.data(), shift.data()));
  }

  // pcl::IterativeClosestPoint block of detectLoopClosureForSubMap (subMapOptmizationNode.cpp:2763-2769, :2822-2855):
  // `source` = current keyframe cloud already moved by the initial guess, `target` = candidate submap cloud.
  // Returns hasConverged(); fitness = getFitnessScore(), T = getFinalTransformation() (row-major 4x4).
  template <typename CloudS, typename CloudT>
  bool icpVerify(const CloudS& source, const CloudT& target, float T[16], double& fitness) {
    Packed s = pack_xyzi(source), t = pack_xyzi(target);
    int32_t tid = -1;
    check(lisreg_map_create(ctx_, nullptr, 0, t.xyzi.data(), t.n, 4.0f, &tid));
    lisreg_icp_params ip; lisreg_icp_params_default(&ip);
    lisreg_icp_pair pr; pr.src = s.xyzi.data(); pr.ns = s.n; pr.target_id = tid;
    lisreg_icp_result r;
    int rc = lisreg_icp_verify_batch(ctx_, 1, &pr, &ip, &r);
    lisreg_map_destroy(ctx_, tid);
    check(rc);
    std::memcpy(T, r.T, sizeof(float) * 16); fitness = r.fitness;
    return r.converged != 0;
  }

  lisreg_ctx* ctx() { return ctx_; }
  float deltaR = 100.f, deltaT = 100.f;   // same members as the reference node (odomEstimationNode.cpp:70-71)
  bool isDegenerate() const { return is_degenerate_; }

 private:
  int solve(const Packed& c, const Packed& s, float pose6[6], char variant, lisreg_lm_result* out) {
    if (map_id_ < 0) throw std::runtime_error("Registrar::setMap must be called first");
    lisreg_lm_params prm; lisreg_lm_params_preset(&prm, variant);
    prm.degenerate_in = is_degenerate_ ? 1 : 0;   // the reference's isDegenerate member persists across frames (quirk Q1)
    lisreg_lm_result res;
    int rc = check(lisreg_scan2map(ctx_, map_id_, c.xyzi.data(), c.aux.empty() ? nullptr : c.aux.data(), c.n,
                                   s.xyzi.data(), s.aux.empty() ? nullptr : s.aux.data(), s.n, pose6, &prm, &res, nullptr));
    is_degenerate_ = res.is_degenerate != 0; deltaR = res.deltaR; deltaT = res.deltaT;
    if (out) *out = res;
    return rc;
  }
  int check(int rc) { if (rc < 0) throw std::runtime_error(std::string("lisreg: ") + lisreg_last_error(ctx_)); return rc; }
  lisreg_ctx* ctx_ = nullptr;
  int32_t map_id_ = -1;
  bool is_degenerate_ = false;
};

// Stand-in for the per-frame body of OdomEstimationNode::laserCloudInfoHandler (odomEstimationNode.cpp:163-239) when
// laserProcessing and odomEstimation run in one process: the sweep goes in (in place, any PointCloud2 layout), the
// refined transformTobeMapped comes out; the key-frame window, the local map and its index stay in HBM.
class Odometry {
 public:
  float transformTobeMapped[6] = {0, 0, 0, 0, 0, 0};
  int keyFrameId = 0;
  float deltaR = 100.f, deltaT = 100.f;
  lisreg_odom_result last;

  Odometry(Registrar& reg, const lisreg_odom_params& prm) : ctx_(reg.ctx()), prm_(prm) {
    if (lisreg_odom_create(ctx_, &prm_, &id_) != LISREG_OK) throw std::runtime_error(std::string("lisreg: ") + lisreg_last_error(ctx_));
  }
  ~Odometry() { if (id_ >= 0) lisreg_odom_destroy(ctx_, id_); }
  Odometry(const Odometry&) = delete;
  Odometry& operator=(const Odometry&) = delete;

  // one sweep as 32-byte PointXYZIRT records (the layout must be the one given in prm.frame.feat.layout at construction)
  template <typename CloudT> int push(const CloudT& laserCloudIn, const float* init_pose6 = nullptr) {
    return pushRaw(laserCloudIn.points.data(), (int32_t)laserCloudIn.points.size(), init_pose6);
  }
  int pushRaw(const void* records, int32_t n, const float* init_pose6 = nullptr) {
    const int rc = lisreg_odom_push(ctx_, id_, (const float*)records, nullptr, n, init_pose6, transformTobeMapped, &last);
    if (rc < 0) throw std::runtime_error(std::string("lisreg: ") + lisreg_last_error(ctx_));
    keyFrameId = last.keyframe_id;
    if (last.frame_id > 1 && !(last.lm.deltaR == 100.f && last.lm.deltaT == 100.f)) { deltaR = last.lm.deltaR; deltaT = last.lm.deltaT; }
    return rc;
  }
  // laserCloudInfoHandler (odomEstimationNode.cpp:163-239) with the message's hints: any type with the scalar members of
  // lis_slam::cloud_info (imuAvailable, odomAvailable, imuRollInit .. initialGuessYaw) - the ROS message itself qualifies
  template <typename InfoT> static lisreg_cloud_info hintsOf(const InfoT& m) {
    lisreg_cloud_info ci;
    ci.imu_available = m.imuAvailable ? 1 : 0; ci.odom_available = m.odomAvailable ? 1 : 0;
    ci.imu_roll_init = m.imuRollInit; ci.imu_pitch_init = m.imuPitchInit; ci.imu_yaw_init = m.imuYawInit;
    ci.initial_guess[0] = m.initialGuessX; ci.initial_guess[1] = m.initialGuessY; ci.initial_guess[2] = m.initialGuessZ;
    ci.initial_guess[3] = m.initialGuessRoll; ci.initial_guess[4] = m.initialGuessPitch; ci.initial_guess[5] = m.initialGuessYaw;
    return ci;
  }
  template <typename CloudT, typename InfoT> int pushInfo(const CloudT& laserCloudIn, const InfoT& cloudInfo) {
    return pushRawInfo(laserCloudIn.points.data(), (int32_t)laserCloudIn.points.size(), cloudInfo);
  }
  template <typename InfoT> int pushRawInfo(const void* records, int32_t n, const InfoT& cloudInfo) {
    const lisreg_cloud_info ci = hintsOf(cloudInfo);
    const int rc = lisreg_odom_push_info(ctx_, id_, (const float*)records, nullptr, n, 0, &ci, transformTobeMapped, &last);
    if (rc < 0) throw std::runtime_error(std::string("lisreg: ") + lisreg_last_error(ctx_));
    keyFrameId = last.keyframe_id;
    if (last.frame_id > 1 && !(last.lm.deltaR == 100.f && last.lm.deltaT == 100.f)) { deltaR = last.lm.deltaR; deltaT = last.lm.deltaT; }
    return rc;
  }

 private:
  lisreg_ctx* ctx_ = nullptr;
  lisreg_odom_params prm_;
  int32_t id_ = -1;
};

// Stand-in for the EPSCGeneration instance of loopClosureThread (subMapOptmizationNode.cpp:2332): same call, same
// public members (epscGeneration.h:121-123, :155-158).  The history lives on the device.
class LoopDetector {
 public:
  int current_frame_id = -1;
  std::vector<int> matched_frame_id;
  std::vector<std::array<float, 16>> matched_frame_transform;   // row-major 4x4 (Eigen::Affine3f::matrix() transposed)

  LoopDetector(Registrar& reg, const uint8_t using_map[256], bool usingEPSC = false, bool usingSEPSC = false,
               bool usingFEPSC = true, bool usingPose = false) : ctx_(reg.ctx()) {
    lisreg_loop_params p; lisreg_loop_params_default(&p);
    p.use_epsc = usingEPSC; p.use_sepsc = usingSEPSC; p.use_fepsc = usingFEPSC; p.use_pose = usingPose;
    if (lisreg_loop_create(ctx_, &p, using_map, &id_) != LISREG_OK) throw std::runtime_error(std::string("lisreg: ") + lisreg_last_error(ctx_));
  }
  ~LoopDetector() { if (id_ >= 0) lisreg_loop_destroy(ctx_, id_); }
  LoopDetector(const LoopDetector&) = delete;
  LoopDetector& operator=(const LoopDetector&) = delete;

  // void EPSCGeneration::loopDetection(corner_pc, surf_pc, semantic_pc, odom)  — odom = row-major 4x4 world pose
  template <typename CloudI, typename CloudL>
  void loopDetection(const CloudI& corner_pc, const CloudI& surf_pc, const CloudL& semantic_pc, const float odom[16]) {
    Packed c = pack_xyzi(corner_pc), s = pack_xyzi(surf_pc), m = pack_xyzil(semantic_pc);
    lisreg_loop_result r;
    if (lisreg_loop_detect(ctx_, id_, c.xyzi.data(), c.n, s.xyzi.data(), s.n, m.xyzi.data(), m.aux.data(), m.n, odom, &r) < 0)
      throw std::runtime_error(std::string("lisreg: ") + lisreg_last_error(ctx_));
    current_frame_id = r.current_frame_id;
    matched_frame_id.clear(); matched_frame_transform.clear();
    for (int i = 0; i < r.n_matched; i++) {
      matched_frame_id.push_back(r.match[i].frame_id);
      std::array<float, 16> T; std::memcpy(T.data(), r.match[i].T, sizeof(float) * 16);
      matched_frame_transform.push_back(T);
    }
  }

 private:
  lisreg_ctx* ctx_ = nullptr;
  int32_t id_ = -1;
};

}  // namespace lisreg_host
