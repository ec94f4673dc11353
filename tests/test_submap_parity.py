"""GPU parity of the device-resident local map / submap (lisreg_submap_*: insert_local_map subMap.h:957-1055 with the
map-based dynamic removal, extractSlidingCloud subMapOptmizationNode.cpp:1369-1432) against the CPU oracle
(oracle/orc_submap.cpp).  Everything is index / compaction / bit-copied float work: the class clouds, their counts, the
bounding box and the registration map must be IDENTICAL.  The extracted map then serves a variant-B registration."""
import numpy as np
import pytest

from lis_slam_b200 import engine as E
from lis_slam_b200 import synth
from oracle import orc

from common import local_map, reg_case

pytestmark = pytest.mark.gpu


def _keyframe_clouds(m, rng, sizes, pose):
    """Five class clouds of one key frame in ITS frame (so that moving them by `pose` puts them back on the map)."""
    T = synth.pose_to_T(pose)
    out = []
    for n, src in zip(sizes, (m["surf"], m["corner"], m["surf"], m["surf"], m["surf"])):
        w = src[rng.choice(len(src), n, replace=False)].copy()
        w[:, :3] += rng.normal(0, 0.02, (n, 3)).astype(np.float32)
        q = w.copy(); q[:, :3] = ((w[:, :3].astype(np.float64) - T[:3, 3]) @ T[:3, :3]).astype(np.float32)
        out.append(q)
    return out


@pytest.mark.parametrize("dynrem", [None, (30.0, 0.3, 3.0, 0.03)])
def test_submap_insert_extract_matches_oracle(engine, dynrem):
    m = local_map()
    rng = np.random.default_rng(17)
    so = orc.Submap(); sid = engine.submap_create()
    poses = [np.array([0.01 * k, -0.005 * k, 0.1 * k, 2.0 * k, 0.3 * k, 0.02 * k], np.float32) for k in range(4)]
    for k, pose in enumerate(poses):
        clouds = _keyframe_clouds(m, rng, (3000, 1500, 20000, 12000, 800) if k else (3000, 0, 20000, 12000, 5), pose)
        co = so.insert(clouds, pose, dynrem=dynrem)
        info = engine.submap_insert(sid, clouds, pose, dynrem=dynrem)
        assert list(info.n) == co, (k, list(info.n), co)
        assert np.array_equal(np.array(list(info.bound_min) + list(info.bound_max)), so.bound)
        for c in range(5):
            assert np.array_equal(engine.submap_download(sid, c), so.get(c)), (k, c)
    if dynrem:
        assert co[0] < 4 * 3000                      # the dynamic class really lost points to the distance test
    cur = np.array([0.0, 0.0, 0.25, 5.0, 0.8, 0.0], np.float32)
    corner_o, surf_o, cnt_o = so.extract(cur)
    mid, info = engine.submap_extract(sid, cur, gate_hint=2.0)
    assert list(info.n) == cnt_o and (info.n_map_corner, info.n_map_surf) == (len(corner_o), len(surf_o))
    for c in range(5):
        assert np.array_equal(engine.submap_download(sid, c), so.get(c)), c
    assert 0 < len(surf_o) < sum(cnt_o) and len(corner_o) > 100
    # the extracted map is a registration map: variant B against it == the oracle on the oracle's extracted clouds
    f, truth, guess = reg_case(4, n_corner=1500, n_surf=5000)
    pose_o, res_o, _ = orc.scan2map(f["corner"], f["surf"], corner_o, surf_o, guess, orc.lm_params("B"), clabel=f["corner_label"], slabel=f["surf_label"], log=False)
    pose_g, res_g, _ = engine.scan2map(mid, f["corner"], f["surf"], guess, E.lm_params("B"), clabel=f["corner_label"], slabel=f["surf_label"])
    er, et = synth.pose_error(pose_o, pose_g)
    assert er <= 1e-4 and et <= 1e-3 and res_g.iters == res_o.iters
    # a second insert + extract re-uses the map slot in place
    clouds = _keyframe_clouds(m, rng, (1000, 500, 5000, 4000, 100), poses[1])
    so.insert(clouds, poses[1], dynrem=dynrem); engine.submap_insert(sid, clouds, poses[1], dynrem=dynrem)
    corner_o, surf_o, cnt_o = so.extract(cur)
    mid2, info = engine.submap_extract(sid, cur, gate_hint=2.0, map_id=mid)
    assert mid2 == mid and list(info.n) == cnt_o
    for c in range(5):
        assert np.array_equal(engine.submap_download(sid, c), so.get(c)), c
    engine.map_destroy(mid); engine.submap_destroy(sid); so.close()


def test_submap_empty_and_clear(engine):
    sid = engine.submap_create(); so = orc.Submap()
    empty = [np.zeros((0, 4), np.float32)] * 5
    info = engine.submap_insert(sid, empty, np.zeros(6, np.float32)); so.insert(empty, np.zeros(6, np.float32))
    assert list(info.n) == [0] * 5 and np.array_equal(np.array(list(info.bound_min) + list(info.bound_max)), so.bound)
    pts = np.zeros((50, 4), np.float32); pts[:, 0] = np.arange(50) * 0.3
    clouds = [pts, pts[:5], pts, pts[:0], pts[:1]]
    info = engine.submap_insert(sid, clouds, np.array([0, 0, 0.5, 1, 1, 0], np.float32), dynrem=(30.0, 0.3, 3.0, 0.03), max_num_pts=10)
    co = so.insert(clouds, np.array([0, 0, 0.5, 1, 1, 0], np.float32), dynrem=(30.0, 0.3, 3.0, 0.03), max_num_pts=10)
    assert list(info.n) == co
    mid, info = engine.submap_extract(sid, np.zeros(6, np.float32)); c, s, cnt = so.extract(np.zeros(6, np.float32))
    assert list(info.n) == cnt
    engine.submap_clear(sid)
    assert engine.submap_download(sid, 2).shape == (0, 4)
    engine.map_destroy(mid); engine.submap_destroy(sid); so.close()


def test_loop_verify_on_submaps_matches_oracle(engine):
    """B4 glue: detectLoopClosureForSubMap (subMapOptmizationNode.cpp:2739-2916) on device-resident submaps - initial
    alignment composed from the EPSC transform or from the poses, key-frame cloud moved and ICP-aligned to every candidate
    submap in one batch, best converged candidate, acceptance threshold, loop constraint tCorrect and its 6-DoF.  Same
    winner and decision as the oracle; transforms / scores within the ICP tolerance (fp64 sums are reduced per block on
    the device, sequentially in the oracle: tests/test_icp_parity.py)."""
    m = local_map()
    rng = np.random.default_rng(23)
    sub_pose = [np.array([0.0, 0.0, 0.3 * k, 4.0 * k, 1.0 * k, 0.0], np.float32) for k in range(3)]
    subs_o, subs_g = [], []
    for k in range(3):
        so = orc.Submap(); sid = engine.submap_create()
        # submap clouds live in the SUBMAP frame: map points expressed relative to the submap pose
        T = synth.pose_to_T(sub_pose[k])
        def local(src, n):
            w = src[rng.choice(len(src), n, replace=False)].copy()
            w[:, :3] = ((w[:, :3].astype(np.float64) - T[:3, 3]) @ T[:3, :3]).astype(np.float32)
            return w
        clouds = [local(m["surf"], 2000), local(m["corner"], 3000), local(m["surf"], 30000), local(m["surf"], 20000), local(m["surf"], 500)]
        so.insert(clouds, np.zeros(6, np.float32)); engine.submap_insert(sid, clouds, np.zeros(6, np.float32))
        subs_o.append(so); subs_g.append(sid)
    # current key frame: a world pose close to submap 1, cloud in its own frame
    key_pose = np.array([0.01, -0.01, 0.32, 4.3, 1.2, 0.02], np.float32)
    Tk = synth.pose_to_T(key_pose)
    kc = np.concatenate([m["surf"][rng.choice(len(m["surf"]), 8000, replace=False)], m["corner"][rng.choice(len(m["corner"]), 1500, replace=False)]]).copy()
    kc[:, :3] = ((kc[:, :3].astype(np.float64) - Tk[:3, 3]) @ Tk[:3, :3]).astype(np.float32)
    kc[:, :3] += rng.normal(0, 0.01, (len(kc), 3)).astype(np.float32)
    key_rel = np.array([0.0, 0.0, 0.05, 0.8, 0.1, 0.0], np.float32)
    # candidate 1 gets an EPSC-style initial pose (prekey pose * planar transform), the others the pose-based one
    prekey = np.array([0.0, 0.0, 0.02, 0.2, 0.1, 0.0], np.float32)
    want = np.linalg.inv(synth.pose_to_T(sub_pose[1])) @ Tk                     # true key -> submap 1
    epsc_T = (np.linalg.inv(synth.pose_to_T(prekey)) @ want @ synth.pose_to_T(np.array([0, 0, 0.02, 0.15, -0.1, 0.0]))).astype(np.float32)
    cands_o, cands_g = [], []
    for k in range(3):
        c = dict(use_epsc=(k == 1), prekey_pose6=prekey, epsc_T=epsc_T, submap_pose6=sub_pose[k])
        cands_o.append(dict(c, submap=subs_o[k])); cands_g.append(dict(c, submap_id=subs_g[k]))
    ro = orc.loop_verify(kc, key_pose, key_rel, cands_o)
    rg, per = engine.loop_verify(kc, key_pose, key_rel, cands_g)
    assert rg.found == ro["found"] == 1 and rg.best == ro["best"]
    assert [p.converged for p in per] == list(ro["converged"])
    for p, fo in zip(per, ro["fitness"]):
        assert abs(p.fitness - fo) <= 1e-4 * max(1.0, fo)
    assert abs(rg.best_score - ro["best_score"]) <= 1e-5
    assert np.array_equal(np.array(rg.key2pre).reshape(4, 4), ro["key2pre"])            # pure fp32 pose arithmetic: identical
    assert np.abs(np.array(rg.correction).reshape(4, 4) - ro["correction"]).max() <= 1e-4
    assert np.abs(np.array(rg.t_correct).reshape(4, 4) - ro["t_correct"]).max() <= 1e-4
    assert np.abs(np.array(rg.constraint6) - ro["constraint6"]).max() <= 1e-4
    # a threshold nobody passes -> "loop not found", best candidate still reported
    rg2, _ = engine.loop_verify(kc, key_pose, key_rel, cands_g, fitness_threshold=1e-9)
    ro2 = orc.loop_verify(kc, key_pose, key_rel, cands_o, fitness_threshold=1e-9)
    assert rg2.found == ro2["found"] == 0 and rg2.best == ro2["best"]
    # the cached target index is rebuilt after the submap changes
    extra = [m["surf"][:500].copy()] * 5
    subs_o[0].insert(extra, np.zeros(6, np.float32)); engine.submap_insert(subs_g[0], extra, np.zeros(6, np.float32))
    rg3, per3 = engine.loop_verify(kc, key_pose, key_rel, cands_g); ro3 = orc.loop_verify(kc, key_pose, key_rel, cands_o)
    assert rg3.best == ro3["best"] and abs(per3[0].fitness - ro3["fitness"][0]) <= 1e-4 * max(1.0, ro3["fitness"][0])
    for so, sid in zip(subs_o, subs_g):
        so.close(); engine.submap_destroy(sid)
