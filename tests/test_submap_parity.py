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
