"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/lisreg.h declares (no compute calls without a GPU), the ctypes structs match the C layout, the C++
adapter compiles and links against PCL-free mock point types, and there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lisreg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lisreg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lis_slam_b200 import engine as E
    lib = C.CDLL(E.build())
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "liblisreg.so does not export %s" % s
    lib.lisreg_version.restype = C.c_char_p
    assert b"sm_100a" in lib.lisreg_version()


def test_struct_layouts_match_header():
    """sizeof() of the ctypes mirrors == sizeof() in C (a tiny C program is compiled against the header)."""
    from lis_slam_b200 import engine as E
    prog = r'''
    #include <stdio.h>
    #include "lisreg.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(lisreg_config), sizeof(lisreg_lm_params), sizeof(lisreg_lm_iter),
             sizeof(lisreg_lm_result), sizeof(lisreg_batch_item), sizeof(lisreg_feat_params), sizeof(lisreg_feat_out),
             sizeof(lisreg_frame_params), sizeof(lisreg_frame_item), sizeof(lisreg_epsc_cloud), sizeof(lisreg_profile),
             sizeof(lisreg_icp_params), sizeof(lisreg_icp_pair), sizeof(lisreg_icp_result), sizeof(lisreg_odom_params),
             sizeof(lisreg_odom_result), sizeof(lisreg_loop_params), sizeof(lisreg_loop_result), sizeof(lisreg_deskew), sizeof(lisreg_cloud_info));
      return 0;
    }'''
    exe = "/tmp/lisreg_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [E.Config, E.LmParams, E.LmIter, E.LmResult, E.BatchItem, E.FeatParams, E.FeatOut, E.FrameParams, E.FrameItem,
               E.EpscCloud, E.Profile, E.IcpParams, E.IcpPair, E.IcpResult, E.OdomParams, E.OdomResult, E.LoopParams, E.LoopResult,
               E.Deskew, E.CloudInfo]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_transform_update_matches_oracle_and_scipy():
    """lisreg_transform_update (transformUpdate, odomEstimationNode.cpp:976-1006; host arithmetic, no device): bit-equal to the
    oracle's restatement of the tf slerp, and both agree with scipy's Slerp (an independent quaternion implementation) to
    float rounding; the clamps are constraintTransformation (common.cpp:286-292)."""
    from lis_slam_b200 import engine as E
    from oracle import orc
    from scipy.spatial.transform import Rotation, Slerp
    rng = np.random.default_rng(5)
    for k in range(200):
        pose = rng.uniform(-0.6, 0.6, 6).astype(np.float32)
        imu = rng.uniform(-0.6, 0.6, 3)
        if k % 7 == 0:
            imu[1] = 1.45                                   # |imuPitchInit| >= 1.4: no slerp
        if k % 11 == 0:
            imu[0] = float(pose[0])                         # theta == 0 branch of slerp
        w = float(rng.choice([0.01, 0.1, 0.5]))
        tol = (0.0, 0.0) if k % 3 else (0.2, 0.3)
        info = E.cloud_info(imu_available=(k % 5 != 0), imu_rpy=imu)
        got = E.transform_update(pose, info, w, tol[0], tol[1])
        ref = orc.transform_update(pose, k % 5 != 0, np.float32(imu[0]), np.float32(imu[1]), w, tol[0], tol[1])
        assert np.array_equal(got, ref), (k, got, ref)
        exp = pose.astype(np.float64).copy()
        if k % 5 != 0 and abs(np.float32(imu[1])) < 1.4:
            for ax, name in ((0, "x"), (1, "y")):
                r = Rotation.from_euler(name, [[float(pose[ax])], [float(np.float32(imu[ax]))]])
                exp[ax] = Slerp([0.0, 1.0], r)([w]).as_euler("xyz")[0][ax]
        if tol[0] > 0:
            exp[0] = np.clip(exp[0], -tol[0], tol[0]); exp[1] = np.clip(exp[1], -tol[0], tol[0]); exp[5] = np.clip(exp[5], -tol[1], tol[1])
        assert np.allclose(got, exp, atol=2e-7, rtol=0), (k, got, exp)
    # info = NULL: clamps only
    p = E.transform_update(np.array([0.5, -0.5, 0.1, 1, 2, 3], np.float32), None, 0.01, 0.25, 2.0)
    assert np.array_equal(p, np.array([0.25, -0.25, 0.1, 1, 2, 2.0], np.float32))


def test_presets_match_reference_constants():
    from lis_slam_b200 import engine as E
    a, b, c = E.lm_params("A"), E.lm_params("B"), E.lm_params("C")
    assert (a.max_iters, b.max_iters, c.max_iters) == (15, 20, 30)                      # odomEstimationNode.cpp:606; subMapOptmizationNode.cpp:1520, :4500
    assert (a.sqdist_gate, b.sqdist_gate) == (1.0, 2.0)                                # :657 / :1610
    assert abs(a.conv_rot_deg - 0.005) < 1e-9 and abs(b.conv_rot_deg - 0.003) < 1e-9 and abs(c.conv_rot_deg - 0.002) < 1e-9
    assert (a.use_label_weight, b.use_label_weight) == (0, 1)
    assert abs(b.label_score[18] - 1.5) < 1e-7 and abs(b.label_score[9] - 1.2) < 1e-7   # config/label.yaml:214-234
    assert (a.min_sel, a.edge_min_valid, a.surf_min_valid) == (50, -1, 100)
    f = E.feat_params()
    assert (f.n_scan, f.horizon, f.edge_thr, abs(f.surf_thr - 0.1) < 1e-7) == (64, 1800, 1.0, True)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lis_slam_b200 import engine as E
    with pytest.raises(E.LisregError):
        E.Engine(device=0)


def test_cpp_adapter_compiles_and_links():
    from lis_slam_b200 import engine as E
    E.build()
    exe = "/tmp/lisreg_adapter_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "adapter_check.cpp"),
                           "-L" + os.path.join(ROOT, "lis_slam_b200"), "-llisreg", "-Wl,-rpath," + os.path.join(ROOT, "lis_slam_b200"),
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"])
    assert subprocess.call([exe], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL) == 0


@pytest.mark.gpu
def test_cpp_adapter_runs_on_gpu():
    exe = "/tmp/lisreg_adapter_check"
    if not os.path.exists(exe):
        test_cpp_adapter_compiles_and_links()
    assert subprocess.call([exe, "run"]) == 0


@pytest.mark.gpu
def test_cpp_adapter_on_real_data_matches_oracle(tmp_path):
    """The header-only C++ adapter (host/lisreg_adapter.hpp) driven with REAL clouds: Registrar::setMap +
    scan2SubMapOptimization (pose vs the CPU oracle, 1e-4 rad / 1e-3 m), featureExtraction on 32-byte PCL PointXYZIRT
    records read in place and on a sensor_msgs/PointCloud2 blob with its own field order (index lists == oracle),
    Odometry::push over a short stream (trajectory vs the oracle flow)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import local_map, reg_case, scene
    from lis_slam_b200 import stream, synth
    from oracle import orc
    from test_stream_parity import OracleBackend, _stream_traj
    d = str(tmp_path)
    m = local_map(); f, truth, guess = reg_case(1, n_corner=2000, n_surf=6000)
    for name, arr in (("map_corner", m["corner"]), ("map_surf", m["surf"]), ("scan_corner", f["corner"]), ("scan_surf", f["surf"]),
                      ("guess", np.asarray(guess, np.float32))):
        np.ascontiguousarray(arr, np.float32).tofile(os.path.join(d, name + ".bin"))
    sw = scene().scan(np.array([0.0, 0.0, 0.3, 2.0, 0.4, 0.0], np.float32), sensor="vlp16", seed=321, fast=True)
    np.ascontiguousarray(sw["pts"], np.float32).tofile(os.path.join(d, "sweep_pts.bin")); np.ascontiguousarray(sw["ring"], np.uint16).tofile(os.path.join(d, "sweep_ring.bin"))
    sweeps = [scene().scan(_stream_traj(t), sensor="vlp16", seed=7000 + t, fast=True) for t in range(8)]
    for t, s in enumerate(sweeps):
        np.ascontiguousarray(s["pts"], np.float32).tofile(os.path.join(d, "stream_%03d_pts.bin" % t))
        np.ascontiguousarray(s["ring"], np.uint16).tofile(os.path.join(d, "stream_%03d_ring.bin" % t))
    np.asarray(_stream_traj(0), np.float32).tofile(os.path.join(d, "stream_init.bin"))
    exe = "/tmp/lisreg_adapter_check"
    test_cpp_adapter_compiles_and_links()
    assert subprocess.call([exe, "data", d]) == 0
    out = np.fromfile(os.path.join(d, "out_pose.bin"), np.float32)
    pose_o, res_o, _ = orc.scan2map(f["corner"], f["surf"], m["corner"], m["surf"], guess, orc.lm_params("A"), log=False)
    er, et = synth.pose_error(pose_o, out[:6])
    assert er <= 1e-4 and et <= 1e-3 and int(out[6]) == res_o.iters and int(out[7]) == res_o.status
    fo = orc.extract_features(sw["pts"], sw["ring"], orc.feat_params(n_scan=16))
    assert np.array_equal(np.fromfile(os.path.join(d, "out_corner_idx.bin"), np.int32), fo["corner_idx"])
    assert np.array_equal(np.fromfile(os.path.join(d, "out_surf_idx.bin"), np.int32), fo["surf_idx"])
    assert np.array_equal(np.fromfile(os.path.join(d, "out_src.bin"), np.int32), fo["src_index"])
    traj = np.fromfile(os.path.join(d, "out_traj.bin"), np.float32).reshape(-1, 7)
    so = stream.OdometryStream(OracleBackend(), orc.lm_params("A"), orc.feat_params(n_scan=16))
    for t, s in enumerate(sweeps):
        po = so.push(s["pts"], s["ring"], initial_pose=_stream_traj(0))
        er, et = synth.pose_error(po, traj[t, :6])
        assert er <= 1e-4 and et <= 1e-3, (t, er, et)
    assert int(traj[-1, 6]) == so.keyframe_id
