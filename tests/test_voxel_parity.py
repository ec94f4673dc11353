"""GPU parity: pcl::VoxelGrid restatement on the device vs the CPU oracle — bit-exact (same voxel order,
same fp32 accumulation order) — plus idempotence / count properties at larger sizes."""
import numpy as np
import pytest

from oracle import orc

from common import local_map, scene

pytestmark = pytest.mark.gpu


def _cloud(n, seed, spread=40.0):
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(-spread, spread, (n, 3)) * np.array([1, 1, 0.1])
    p[:, 3] = rng.uniform(0, 255, n)
    return p


@pytest.mark.parametrize("n,leaf", [(1, 0.4), (7, 0.2), (2047, 0.4), (2049, 0.2), (50000, 0.4), (120000, 0.2)])
def test_voxel_grid_bit_exact(engine, n, leaf):
    p = _cloud(n, n)
    assert np.array_equal(orc.voxel_grid(p, leaf), engine.voxel_grid(p, leaf))


def test_scan_features_and_map(engine):
    s = scene().scan(np.array([0, 0, 0.3, 5, 0.5, 0], np.float32))
    f = orc.extract_features(s["pts"], s["ring"])
    ext = s["pts"][f["src_index"]]
    for idx, leaf in ((f["corner_idx"], 0.2), (f["surf_idx"], 0.4)):
        c = np.ascontiguousarray(ext[idx])
        o = orc.voxel_grid(c, leaf); g = engine.voxel_grid(c, leaf)
        assert np.array_equal(o, g) and 0 < len(g) < len(c)
    m = local_map()
    assert np.array_equal(orc.voxel_grid(m["surf"], 0.4), engine.voxel_grid(m["surf"], 0.4))


def test_duplicates_collisions_and_empty(engine):
    assert len(engine.voxel_grid(np.zeros((0, 4), np.float32), 0.4)) == 0
    p = np.tile(np.array([[1.0, 2.0, 3.0, 4.0]], np.float32), (5000, 1))     # everything in one voxel
    g = engine.voxel_grid(p, 0.4)
    assert len(g) == 1 and np.array_equal(g, orc.voxel_grid(p, 0.4))
    q = _cloud(30000, 5, spread=3.0)                                          # many points per voxel
    assert np.array_equal(orc.voxel_grid(q, 0.4), engine.voxel_grid(q, 0.4))


def test_properties_at_full_size(engine):
    """Size-independent properties on a 2M-point cloud: each output lies in its own voxel (so a second
    pass keeps the count), ascending voxel order, count <= n."""
    p = _cloud(2_000_000, 9, spread=70.0)
    g = engine.voxel_grid(p, 0.4)
    g2 = engine.voxel_grid(g, 0.4)
    assert len(g2) == len(g) <= len(p)
    inv = np.float32(1.0) / np.float32(0.4)
    ijk = np.floor(g[:, :3] * inv).astype(np.int64)
    mn = np.floor(p[:, :3].min(0) * inv).astype(np.int64); mx = np.floor(p[:, :3].max(0) * inv).astype(np.int64)
    div = mx - mn + 1
    lin = (ijk[:, 0] - mn[0]) + (ijk[:, 1] - mn[1]) * div[0] + (ijk[:, 2] - mn[2]) * div[0] * div[1]
    assert np.all(np.diff(lin) >= 0)


@pytest.mark.parametrize("spread,leaf", [(0.5, 0.2), (3.0, 0.2), (20.0, 0.2), (400.0, 0.2)])
def test_every_radix_pass_count(engine, spread, leaf):
    """The sort only runs the 8-bit passes the largest voxel index needs (1 .. 4): tiny, small, medium and very large
    extents, with voxels that straddle the 2048-entry chunks of the centroid kernel."""
    rng = np.random.default_rng(int(spread * 10))
    p = np.zeros((30000, 4), np.float32)
    p[:, :3] = rng.uniform(-spread, spread, (30000, 3))
    p[:5000, :3] = rng.uniform(-0.05, 0.05, (5000, 3))            # one crowded voxel neighbourhood (> 2048 points in a voxel)
    p[:, 3] = rng.uniform(0, 255, 30000)
    o = orc.voxel_grid(p, leaf)
    g = engine.voxel_grid(p, leaf)
    assert np.array_equal(o, g)


def test_block_kernel_equals_multi_kernel_path(monkeypatch):
    """k_vox_block (one block per cloud, the sort in shared memory; clouds of up to 131072 points) against the multi-kernel
    path (LISREG_VOX_UNFUSED=1) on a feature cloud (runs -> shared-memory sort), a shuffled one (every point its own run ->
    the in-kernel global-memory sort) and sizes around the 22528-run capacity."""
    from lis_slam_b200 import engine as E
    s = scene().scan(np.array([0, 0, 0.3, 5, 0.5, 0], np.float32))
    f = orc.extract_features(s["pts"], s["ring"])
    surf = np.ascontiguousarray(s["pts"][f["src_index"]][f["surf_idx"]])
    rng = np.random.default_rng(3)
    clouds = [(surf, 0.4), (surf[rng.permutation(len(surf))], 0.4), (_cloud(22528, 1), 0.2), (_cloud(22529, 2), 0.2), (_cloud(131072, 3), 0.4),
              (surf[:1], 0.4)]
    outs = {}
    for mode, batch in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):     # batch form: centroids inside the block kernel / thread per voxel
        monkeypatch.setenv("LISREG_VOX_UNFUSED", mode)
        monkeypatch.setenv("LISREG_VOX_BATCH_FORM", batch)
        eng = E.Engine(device=0)
        outs[(mode, batch)] = [eng.voxel_grid(c, leaf) for c, leaf in clouds]
        eng.close()
    ref = [orc.voxel_grid(c, leaf) for c, leaf in clouds]
    for k, v in outs.items():
        for a, b in zip(ref, v):
            assert np.array_equal(a, b), k
