"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/lisreg.h declares (no compute calls without a GPU), the ctypes structs match the C layout, the C++
adapter compiles and links against PCL-free mock point types, and there is no CPU fallback."""
import ctypes as C
import os
import re
import subprocess

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
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(lisreg_config), sizeof(lisreg_lm_params), sizeof(lisreg_lm_iter),
             sizeof(lisreg_lm_result), sizeof(lisreg_batch_item), sizeof(lisreg_feat_params), sizeof(lisreg_feat_out),
             sizeof(lisreg_frame_params), sizeof(lisreg_frame_item), sizeof(lisreg_epsc_cloud), sizeof(lisreg_profile),
             sizeof(lisreg_icp_params), sizeof(lisreg_icp_pair), sizeof(lisreg_icp_result), sizeof(lisreg_odom_params),
             sizeof(lisreg_odom_result), sizeof(lisreg_loop_params), sizeof(lisreg_loop_result), sizeof(lisreg_deskew));
      return 0;
    }'''
    exe = "/tmp/lisreg_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    mirrors = [E.Config, E.LmParams, E.LmIter, E.LmResult, E.BatchItem, E.FeatParams, E.FeatOut, E.FrameParams, E.FrameItem,
               E.EpscCloud, E.Profile, E.IcpParams, E.IcpPair, E.IcpResult, E.OdomParams, E.OdomResult, E.LoopParams, E.LoopResult,
               E.Deskew]
    assert sizes == [C.sizeof(m) for m in mirrors]


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
