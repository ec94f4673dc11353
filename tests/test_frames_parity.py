"""GPU parity: the whole-frame pipeline (features -> voxel grid -> scan-to-map LM, all on the device,
one C-ABI call) against the same chain run with the CPU oracle.  Index/byte stages are bit-exact, so
the query clouds are identical and the final pose must agree to the LM tolerance (1e-4 rad / 1e-3 m)."""
import functools

import numpy as np
import pytest

from lis_slam_b200 import engine as E
from lis_slam_b200 import synth
from oracle import orc

from common import local_map, scene

pytestmark = pytest.mark.gpu


@functools.lru_cache(maxsize=None)
def frame(seed):
    rng = np.random.default_rng(500 + seed)
    truth = synth.random_pose(rng)
    guess = synth.perturb_pose(truth, rng)
    s = scene().scan(truth, seed=2000 + seed)
    return s, truth, guess


def oracle_chain(s, m, guess, variant="A", **kw):
    f = orc.extract_features(s["pts"], s["ring"])
    ext = s["pts"][f["src_index"]]
    corner = orc.voxel_grid(np.ascontiguousarray(ext[f["corner_idx"]]), 0.2)
    surf = orc.voxel_grid(np.ascontiguousarray(ext[f["surf_idx"]]), 0.4)
    pose, res, _ = orc.scan2map(corner, surf, m["corner"], m["surf"], guess, orc.lm_params(variant, **kw), log=False)
    return pose, res, len(corner), len(surf)


def test_frames_batch_matches_oracle_chain(engine):
    m = local_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=1.0)
    frames = [frame(i) for i in range(3)]
    poses, res = engine.frames_batch([(mid, s["pts"], s["ring"]) for s, _, _ in frames], [g for _, _, g in frames], E.frame_params("A"))
    for i, (s, truth, guess) in enumerate(frames):
        po, ro, nc, ns = oracle_chain(s, m, guess)
        assert (res[i].n_corner, res[i].n_surf) == (nc, ns)
        assert res[i].iters == ro.iters and res[i].status == ro.status
        er, et = synth.pose_error(po, poses[i])
        assert er <= 1e-4 and et <= 1e-3, (er, et)
        er, et = synth.pose_error(truth, poses[i])
        assert er < 3e-3 and et < 3e-2, (er, et)      # and the registration recovers the ground truth
    engine.map_destroy(mid)


def test_frames_fixed_iterations_and_ragged_batch(engine):
    m = local_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=1.0)
    s0, t0, g0 = frame(0)
    s1, t1, g1 = frame(1)
    half = {"pts": s1["pts"][: len(s1["pts"]) // 3], "ring": s1["ring"][: len(s1["pts"]) // 3]}   # a partial sweep
    empty = {"pts": np.zeros((0, 4), np.float32), "ring": np.zeros(0, np.uint16)}
    prm = E.frame_params("A", early_exit=0, max_iters=10)
    poses, res = engine.frames_batch([(mid, s0["pts"], s0["ring"]), (mid, half["pts"], half["ring"]), (mid, empty["pts"], empty["ring"])],
                                     [g0, g1, g0], prm)
    po, ro, nc, ns = oracle_chain(s0, m, g0, early_exit=0, max_iters=10)
    assert res[0].iters == 10 == ro.iters
    er, et = synth.pose_error(po, poses[0]); assert er <= 1e-4 and et <= 1e-3
    po, ro, nc, ns = oracle_chain(half, m, g1, early_exit=0, max_iters=10)
    assert (res[1].n_corner, res[1].n_surf) == (nc, ns) and res[1].status == ro.status
    er, et = synth.pose_error(po, poses[1]); assert er <= 1e-4 and et <= 1e-3
    assert res[2].status == E.NOT_ENOUGH_FEATURES and np.array_equal(poses[2], np.asarray(g0, np.float32))
    engine.map_destroy(mid)


def test_chunked_e2e_pipeline_is_bit_identical(monkeypatch):
    """The arena entry point uploads the sweeps in chunks on a copy stream and starts each chunk's pipeline as soon
    as its bytes have landed; the chunking must not change a single bit of the result (ragged + empty frames incl.)."""
    m = local_map()
    s0, _, g0 = frame(0)
    s1, _, g1 = frame(1)
    third = {"pts": s1["pts"][: len(s1["pts"]) // 3], "ring": s1["ring"][: len(s1["pts"]) // 3]}
    empty = {"pts": np.zeros((0, 4), np.float32), "ring": np.zeros(0, np.uint16)}
    sweeps = [s0, third, empty, s1, s0]
    guesses = [g0, g1, g0, g1, g1]
    prm = E.frame_params("A", early_exit=0, max_iters=6)
    out = {}
    for chunk in ("0", "2"):
        monkeypatch.setenv("LISREG_E2E_CHUNK", chunk)
        eng = E.Engine(device=0)
        mid = eng.map_create(m["corner"], m["surf"], gate_hint=1.0)
        poses, res = eng.frames_batch([(mid, s["pts"], s["ring"]) for s in sweeps], guesses, prm)
        out[chunk] = (poses.copy(), [(r.status, r.iters, r.n_corner, r.n_surf, r.n_sel_last) for r in res])
        eng.close()
    assert np.array_equal(out["0"][0], out["2"][0])
    assert out["0"][1] == out["2"][1]
    assert out["0"][1][2][0] == E.NOT_ENOUGH_FEATURES


def test_submit_wait_pipeline_matches_blocking_call(engine):
    """lisreg_frames_batch_submit / _wait (two batches in flight on private streams and work sets) must return exactly
    what the blocking arena call returns, in ticket order, and refuse a third submit."""
    import ctypes as C
    m = local_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=1.0)
    s0, _, g0 = frame(0)
    s1, _, g1 = frame(1)
    prm = E.frame_params("A", early_exit=0, max_iters=5)
    batches = [([s0, s1], [g0, g1]), ([s1, s0, s1], [g1, g0, g0])]
    ref = [engine.frames_batch([(mid, s["pts"], s["ring"]) for s in sw], gs, prm) for sw, gs in batches]
    packed = []
    for sw, gs in batches:
        chunks, off = [], 0
        items = (E.FrameItem * len(sw))()
        for i, s in enumerate(sw):
            p = np.ascontiguousarray(s["pts"], np.float32); r = np.ascontiguousarray(s["ring"], np.uint16)
            op = off; chunks.append(p.view(np.uint8).reshape(-1)); off += p.nbytes
            orr = off; chunks.append(r.view(np.uint8).reshape(-1)); off += r.nbytes
            pad = (-off) % 16
            if pad:
                chunks.append(np.zeros(pad, np.uint8)); off += pad
            items[i] = E.FrameItem(op, orr, len(p), mid)
        packed.append((items, np.concatenate(chunks), np.asarray(gs, np.float32).reshape(-1, 6).copy()))
    tickets = [engine.frames_batch_submit(it, len(it), ar.ctypes.data, ar.nbytes, gs, prm) for it, ar, gs in packed]
    assert sorted(tickets) == [0, 1]
    with pytest.raises(Exception):
        engine.frames_batch_submit(packed[0][0], 2, packed[0][1].ctypes.data, packed[0][1].nbytes, packed[0][2], prm)
    for t, (it, ar, gs), (pose_ref, res_ref) in zip(tickets, packed, ref):
        pose = np.zeros_like(gs); res = (E.LmResult * len(it))()
        engine.frames_batch_wait(t, pose, res)
        assert np.array_equal(pose, pose_ref)
        assert [(r.status, r.iters, r.n_sel_last) for r in res] == [(r.status, r.iters, r.n_sel_last) for r in res_ref]
    engine.map_destroy(mid)


def test_device_resident_sub_batches_are_bit_identical(monkeypatch):
    """lisreg_frames_batch_dev runs a large batch as concurrent sub-batches on private streams (kernels of different
    pipeline stages overlap); splitting must not change a bit: 70 frames (ragged and empty ones included) as 2 sub-batches
    of 35 against the same batch kept whole (LISREG_DEV_SPLIT=0), twice in a row without a host sync in between."""
    import ctypes as C
    import torch
    m = local_map()
    s0, _, g0 = frame(0)
    s1, _, g1 = frame(1)
    third = {"pts": s1["pts"][: len(s1["pts"]) // 3], "ring": s1["ring"][: len(s1["pts"]) // 3]}
    empty = {"pts": np.zeros((0, 4), np.float32), "ring": np.zeros(0, np.uint16)}
    base = [s0, third, s1, empty, s0, s1, s0]
    F = 70
    rng = np.random.default_rng(8)
    sweeps = [base[i % len(base)] for i in range(F)]
    guesses = np.stack([(g0 if i % 2 else g1) + rng.normal(0, [0.002, 0.002, 0.005, 0.05, 0.05, 0.02]).astype(np.float32) for i in range(F)]).astype(np.float32)
    prm = E.frame_params("A", early_exit=0, max_iters=5)
    dev = torch.device("cuda", 0)
    d_sw = {id(s): (torch.from_numpy(np.ascontiguousarray(s["pts"], np.float32)).to(dev), torch.from_numpy(s["ring"].astype(np.int16)).to(dev)) for s in base}
    out = {}
    # fused: k_feat_front with 8-ring groups (opt-in); plain: curvature / occlusion marks by k_feat_curv_occl instead of inside
    # the selection kernel
    for split, fused, plain in (("0", "0", "0"), ("4", "0", "0"), ("0", "1", "0"), ("4", "1", "0"), ("4", "0", "1")):
        monkeypatch.setenv("LISREG_DEV_SPLIT", split)
        monkeypatch.setenv("LISREG_FEAT_FUSED", fused)
        monkeypatch.setenv("LISREG_FEAT_SEG_UNFUSED", plain)
        eng = E.Engine(device=0)
        mid = eng.map_create(m["corner"], m["surf"], gate_hint=1.0)
        items = (E.FrameItem * F)()
        for i, s in enumerate(sweeps):
            p, r = d_sw[id(s)]
            items[i] = E.FrameItem(p.data_ptr() if len(s["pts"]) else None, r.data_ptr() if len(s["pts"]) else None, len(s["pts"]), mid)
        runs = []
        pose = [torch.from_numpy(guesses).to(dev) for _ in range(2)]
        res = [torch.zeros(F * C.sizeof(E.LmResult), dtype=torch.uint8, device=dev) for _ in range(2)]
        torch.cuda.synchronize()
        for k in range(2):
            eng.frames_batch_dev(items, F, pose[k].data_ptr(), prm, res[k].data_ptr())
        eng.sync()
        for k in range(2):
            rr = (E.LmResult * F).from_buffer_copy(res[k].cpu().numpy().tobytes())
            runs.append((pose[k].cpu().numpy().copy(), [(r.status, r.iters, r.n_corner, r.n_surf, r.n_sel_last) for r in rr]))
        eng.close()
        assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]
        out[(split, fused, plain)] = runs[0]
    ref = out[("0", "0", "0")]
    for k, v in out.items():
        assert np.array_equal(ref[0], v[0]), k
        assert ref[1] == v[1], k
    assert not np.array_equal(ref[0], guesses)
