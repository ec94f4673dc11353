// Compile-and-link check of the C++ adapter against the PCL / ROS-free mock types (no GPU needed to build).
//   adapter_check                 : no device -> verifies that a context cannot be created (there is no CPU path)
//   adapter_check run             : GPU smoke of every adapter entry point on trivial clouds
//   adapter_check data <dir>      : GPU run on REAL data written by tests/test_abi_and_host.py (a synthetic sweep, a local map,
//                                   an initial guess): Registrar::setMap + scan2SubMapOptimization, featureExtraction on
//                                   PointXYZIRT records in place and on a PointCloud2 blob, Odometry::push for a short stream;
//                                   results go to <dir>/out_*.bin and are compared with the CPU oracle by the test
#define LISREG_ADAPTER_MOCK_PCL
#include "../lis_slam_b200/host/lisreg_adapter.hpp"
#include <cstdio>
#include <fstream>

template <typename T> static std::vector<T> rd(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + path);
  const size_t bytes = (size_t)f.tellg(); f.seekg(0);
  std::vector<T> v(bytes / sizeof(T));
  f.read((char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <typename T> static void wr(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary); f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}
static LISREG_CLOUD(PointXYZI) cloud_of(const std::vector<float>& xyzi) {
  LISREG_CLOUD(PointXYZI) c; c.points.resize(xyzi.size() / 4);
  for (size_t i = 0; i < c.points.size(); i++) { auto& p = c.points[i]; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.pad = 1.f; p.intensity = xyzi[4 * i + 3]; }
  return c;
}
static LISREG_CLOUD(PointXYZIRT) sweep_of(const std::vector<float>& xyzi, const std::vector<uint16_t>& ring) {
  LISREG_CLOUD(PointXYZIRT) c; c.points.resize(ring.size());
  for (size_t i = 0; i < c.points.size(); i++) { auto& p = c.points[i]; p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.pad = 1.f; p.intensity = xyzi[4 * i + 3]; p.ring = ring[i]; p.time = 0.f; }
  return c;
}

static int run_data(const std::string& dir) {
  using namespace lisreg_host;
  Registrar r(0);
  // ---- B2: scan2SubMapOptimization through the adapter on real clouds ----
  auto mapc = cloud_of(rd<float>(dir + "/map_corner.bin")), maps = cloud_of(rd<float>(dir + "/map_surf.bin"));
  auto sc = cloud_of(rd<float>(dir + "/scan_corner.bin")), ss = cloud_of(rd<float>(dir + "/scan_surf.bin"));
  std::vector<float> pose = rd<float>(dir + "/guess.bin");
  r.setMap(mapc, maps, 1.0f);
  lisreg_lm_result res;
  const int rc = r.scan2SubMapOptimization(sc, ss, pose.data(), &res);
  std::vector<float> out(pose.begin(), pose.end());
  out.push_back((float)res.iters); out.push_back((float)rc); out.push_back(r.deltaR); out.push_back(r.deltaT);
  wr(dir + "/out_pose.bin", out);
  // ---- B1: featureExtraction on PointXYZIRT records in place, and on a PointCloud2 blob with another field order ----
  auto xyzi = rd<float>(dir + "/sweep_pts.bin"); auto ring = rd<uint16_t>(dir + "/sweep_ring.bin");
  auto sweep = sweep_of(xyzi, ring);
  lisreg_feat_params fp; lisreg_feat_params_default(&fp); fp.n_scan = 16;
  std::vector<int32_t> src, corner, surf, sharp, flat;
  r.featureExtraction(sweep, fp, src, corner, surf, sharp, flat);
  wr(dir + "/out_corner_idx.bin", corner); wr(dir + "/out_surf_idx.bin", surf); wr(dir + "/out_src.bin", src);
  LISREG_POINTCLOUD2 msg;                                        // ring first, then xyz, then intensity: 20-byte records
  msg.point_step = 20; msg.width = (uint32_t)ring.size();
  msg.fields = {{"ring", 0, 4, 1}, {"x", 4, 7, 1}, {"y", 8, 7, 1}, {"z", 12, 7, 1}, {"intensity", 16, 7, 1}};
  msg.data.resize((size_t)20 * ring.size());
  for (size_t i = 0; i < ring.size(); i++) {
    uint8_t* q = &msg.data[20 * i];
    std::memcpy(q, &ring[i], 2); std::memcpy(q + 4, &xyzi[4 * i], 12); std::memcpy(q + 16, &xyzi[4 * i + 3], 4);
  }
  std::vector<int32_t> src2, corner2, surf2, sharp2, flat2;
  r.featureExtractionMsg(msg, fp, src2, corner2, surf2, sharp2, flat2);
  if (corner2 != corner || surf2 != surf || src2 != src) { std::printf("PointCloud2 path differs from the PCL-record path\n"); return 6; }
  // ---- streaming odometry over the sweeps sweep_000.bin ... (PointXYZIRT records read in place) ----
  lisreg_odom_params op; lisreg_odom_params_default(&op); op.frame.feat.n_scan = 16; op.frame.feat.layout = layout_xyzirt(); op.use_graph = 0;
  Odometry od(r, op);
  std::vector<float> traj;
  std::vector<float> init = rd<float>(dir + "/stream_init.bin");
  for (int t = 0;; t++) {
    char name[64]; std::snprintf(name, sizeof(name), "/stream_%03d", t);
    std::ifstream probe(dir + name + "_pts.bin", std::ios::binary);
    if (!probe) break;
    auto sw = sweep_of(rd<float>(dir + name + "_pts.bin"), rd<uint16_t>(dir + name + "_ring.bin"));
    od.push(sw, init.data());
    for (int k = 0; k < 6; k++) traj.push_back(od.transformTobeMapped[k]);
    traj.push_back((float)od.keyFrameId);
  }
  wr(dir + "/out_traj.bin", traj);
  {   // the cloud_info overload: a message-shaped struct with no hints available behaves like the plain push (first pose = IMU attitude)
    struct { bool imuAvailable = false, odomAvailable = false; float imuRollInit = 0, imuPitchInit = 0, imuYawInit = 0, initialGuessX = 0,
             initialGuessY = 0, initialGuessZ = 0, initialGuessRoll = 0, initialGuessPitch = 0, initialGuessYaw = 0; } info;
    info.imuRollInit = init[0]; info.imuPitchInit = init[1]; info.imuYawInit = init[2];
    Odometry od2(r, op);
    auto sw0 = sweep_of(rd<float>(dir + "/stream_000_pts.bin"), rd<uint16_t>(dir + "/stream_000_ring.bin"));
    od2.pushInfo(sw0, info);
    if (od2.keyFrameId != 1 || od2.transformTobeMapped[0] != init[0] || od2.transformTobeMapped[2] != 0.f) { std::printf("cloud_info push differs\n"); return 8; }
  }
  std::printf("rc=%d iters=%d corner=%zu surf=%zu frames=%zu\n", rc, res.iters, corner.size(), surf.size(), traj.size() / 7);
  return 0;
}

int main(int argc, char** argv) {
  using namespace lisreg_host;
  static_assert(sizeof(lisreg_mock::PointXYZI) == 32, "PCL PointXYZI is a 32-byte record");
  static_assert(sizeof(lisreg_mock::PointXYZIRT) == 32 && offsetof(lisreg_mock::PointXYZIRT, ring) == 20 && offsetof(lisreg_mock::PointXYZIRT, time) == 24,
                "PCL PointXYZIRT layout (common.h:12-23)");
  LISREG_CLOUD(PointXYZI) c; c.points.resize(3); c.points[1].x = 1.f; c.points[1].intensity = 7.f;
  Packed p = pack_xyzi(c);
  if (p.n != 3 || p.xyzi[4] != 1.f || p.xyzi[7] != 7.f) return 2;
  if (argc > 2 && std::string(argv[1]) == "data") {
    try { return run_data(argv[2]); } catch (const std::exception& e) { std::printf("error: %s\n", e.what()); return 7; }
  }
  if (argc > 1) {   // "run": needs a GPU
    Registrar r(0);
    LISREG_CLOUD(PointXYZI) mapc, maps, sc, ss; float pose[6] = {0, 0, 0, 0, 0, 0};
    mapc.points.resize(10); maps.points.resize(10); sc.points.resize(10); ss.points.resize(10);
    r.setMap(mapc, maps);
    LISREG_CLOUD(PointXYZIL) lc, ls; lc.points.resize(4); ls.points.resize(4);
    r.scan2SubMapOptimizationLabelled(lc, ls, pose, 'B');
    int rc = r.scan2SubMapOptimization(sc, ss, pose);   // 10 surf points <= surfFeatureMinValidNum => "Not enough features"
    // loop detector stand-in: two empty keyframes, ids count up and nothing matches
    uint8_t lut[256] = {0};
    LoopDetector ld(r, lut);
    float odom[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    ld.loopDetection(sc, ss, lc, odom); ld.loopDetection(sc, ss, lc, odom);
    if (ld.current_frame_id != 1 || !ld.matched_frame_id.empty()) return 4;
    float T[16]; double fit = 0;
    LISREG_CLOUD(PointXYZI) a, b; a.points.resize(50); b.points.resize(60);
    for (int i = 0; i < 50; i++) { a.points[i].x = 0.1f * i; a.points[i].y = 0.01f * (i % 7); }
    for (int i = 0; i < 60; i++) { b.points[i].x = 0.1f * i + 0.02f; b.points[i].y = 0.01f * (i % 7); }
    bool conv = r.icpVerify(a, b, T, fit);
    std::printf("rc=%d icp conv=%d fitness=%g\n", rc, (int)conv, fit);
    if (!conv || !(fit < 0.01)) return 5;
    return rc == LISREG_NOT_ENOUGH_FEATURES ? 0 : 3;
  }
  try { Registrar r(0); } catch (const std::exception& e) { std::printf("no device: %s\n", e.what()); return 0; }
  return 0;
}
