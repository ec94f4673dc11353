"""GPU parity: EPSC/SEPSC/FEPSC descriptors (bit-exact, incl. the u8 wrap quirk Q4) and the shifted-SAD
loop-closure scoring + top-k (bit-exact indices/shifts, scores from identical integers) vs the CPU oracle."""
import functools

import numpy as np
import pytest

from oracle import orc

from common import scene

pytestmark = pytest.mark.gpu


@functools.lru_cache(maxsize=None)
def keyframe(seed):
    rng = np.random.default_rng(900 + seed)
    pose = np.array([0, 0, rng.uniform(-3, 3), rng.uniform(-40, 40), rng.uniform(-2, 2), 0], np.float32)
    s = scene().scan(pose, seed=2000 + seed)
    f = orc.extract_features(s["pts"], s["ring"])
    ext = s["pts"][f["src_index"]]; lab = s["label"][f["src_index"]]
    return (np.ascontiguousarray(ext[f["corner_idx"]]), np.ascontiguousarray(ext[f["surf_idx"]]), ext, lab)


def test_descriptors_bit_exact(engine):
    clouds = [keyframe(i) for i in range(3)]
    # u8-wrap case (Q4): > 256 surf points in one bin, plus labels outside the LUT and empty clouds
    rng = np.random.default_rng(1)
    dense = np.zeros((3000, 4), np.float32); dense[:, 0] = 10 + rng.uniform(0, 0.5, 3000); dense[:, 1] = rng.uniform(0, 0.3, 3000)
    clouds.append((dense[:10].copy(), dense, dense, np.full(3000, 9, np.uint16)))
    clouds.append((np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32), dense[:50], np.full(50, 300, np.uint16)))
    lut = orc.using_map_lut()
    g = engine.epsc_describe(clouds, lut)
    for i, (c, s, m, l) in enumerate(clouds):
        o = orc.epsc_describe(c, s, m, l, lut)
        for k in ("epsc", "sepsc", "fepsc"):
            assert np.array_equal(o[k], g[k][i]), (i, k)
    assert g["epsc"][3].max() > 0    # the wrapped bin is populated


def test_rotation_is_a_column_shift(engine):
    c, s, m, l = keyframe(0)
    yaw = np.deg2rad(18.0)
    R = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]], np.float32)

    def rot(p):
        q = p.copy(); q[:, :2] = p[:, :2] @ R.T; return q
    d = engine.epsc_describe([(c, s, m, l), (rot(c), rot(s), rot(m), l)], orc.using_map_lut())["fepsc"]
    idx, score, shift = engine.epsc_score_all(d, topk=5)
    assert idx[1, 0] == 0 and shift[1, 0] == 4 and score[1, 0] > 0.9      # 18 deg / 4.5 deg per sector


def test_score_all_matches_oracle(engine):
    rng = np.random.default_rng(7)
    base = [rng.integers(0, 256, (20, 80), dtype=np.uint8) for _ in range(12)]
    desc = []
    for i in range(200):
        b = base[rng.integers(0, len(base))].copy()
        b = np.roll(b, int(rng.integers(-12, 13)), axis=1)                   # revisit under a yaw offset
        noise = rng.integers(0, 256, b.shape, dtype=np.uint8)
        mask = rng.random(b.shape) < rng.choice([0.02, 0.2, 0.6])
        desc.append(np.where(mask, noise, b).astype(np.uint8))
    desc = np.stack(desc)
    desc[50] = desc[10]                                                     # identical descriptor => score 1.0
    desc[60] = 0; desc[61] = 0                                              # empty descriptors => SAD 0
    io, so, ho = orc.epsc_score_all(desc, topk=5, n_threads=4)
    ig, sg, hg = engine.epsc_score_all(desc, topk=5)
    assert np.array_equal(io, ig) and np.array_equal(ho, hg) and np.array_equal(so, sg)
    assert sg[50, 0] == 1.0 and ig[50, 0] == 10
    assert (ig >= 0).sum() > 100


def test_score_all_tile_edges(engine):
    rng = np.random.default_rng(8)
    for n in (1, 2, 7, 8, 9, 17, 33):
        base = rng.integers(0, 256, (20, 80), dtype=np.uint8)
        desc = np.stack([np.roll(base, int(rng.integers(-9, 10)), axis=1) for _ in range(n)])
        io, so, ho = orc.epsc_score_all(desc, topk=3)
        ig, sg, hg = engine.epsc_score_all(desc, topk=3)
        assert np.array_equal(io, ig) and np.array_equal(ho, hg) and np.array_equal(so, sg), n


def test_cyclic_row_shards_reassemble_the_full_result(engine):
    """Multi-GPU sharding of the loop-closure scoring: every rank scores the cyclic query rows rank,
    rank + world, ...; interleaving the shards gives exactly the single-GPU all-pairs result (ragged shards included)."""
    from lis_slam_b200 import shard
    rng = np.random.default_rng(12)
    base = [rng.integers(0, 256, (20, 80), dtype=np.uint8) for _ in range(6)]
    desc = np.stack([np.roll(base[rng.integers(0, 6)], int(rng.integers(-9, 10)), axis=1) for _ in range(101)])
    full = engine.epsc_score_all(desc, topk=4)
    for world in (2, 3, 8):
        got = [np.zeros_like(a) for a in full]
        for rank in range(world):
            rows = shard.cyclic_rows(len(desc), rank, world)
            part = engine.epsc_score_rows(desc, rank, world, topk=4)
            assert len(part[0]) == len(rows)
            for a, b in zip(got, part):
                a[rows] = b
        for a, b in zip(got, full):
            assert np.array_equal(a, b), world
    assert engine.epsc_score_rows(desc, 500, 2, topk=4)[0].shape == (0, 4)      # a rank beyond the rows owns nothing
