// Compile-and-link check of the C++ adapter against the PCL-free mock point types (no GPU needed to build;
// at run time it only verifies that a context cannot be created without a device, i.e. that there is no CPU path).
#define LISREG_ADAPTER_MOCK_PCL
#include "../lis_slam_b200/host/lisreg_adapter.hpp"
#include <cstdio>
int main(int argc, char** argv) {
  using namespace lisreg_host;
  static_assert(sizeof(lisreg_mock::PointXYZI) == 32, "PCL PointXYZI is a 32-byte record");
  LISREG_CLOUD(PointXYZI) c; c.points.resize(3); c.points[1].x = 1.f; c.points[1].intensity = 7.f;
  Packed p = pack_xyzi(c);
  if (p.n != 3 || p.xyzi[4] != 1.f || p.xyzi[7] != 7.f) return 2;
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
