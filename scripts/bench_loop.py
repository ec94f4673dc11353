#!/usr/bin/env python
"""Loop-closure side of the hot path (BASELINE.json configs[3] shape, one GPU's share): timings for DESIGN.md / profiles/.
  * EPSC all-pairs shifted-SAD + top-k over N descriptors (lisreg_epsc_score_all): achieved byte-absdiff/s
  * submap ICP verification of P (keyframe 50k pts, submap 200k pts) pairs (lisreg_icp_verify_batch)
  * the online loop detector (lisreg_loop_detect) per keyframe with a gated history
and the CPU restatement of each on a bounded sample.  Prints one JSON object.  usage: bench_loop.py [--n 5000] [--pairs 16]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--cpu-rows", type=int, default=40)
    args = ap.parse_args()
    from lis_slam_b200 import engine as E, synth
    from oracle import orc
    eng = E.Engine(device=0)
    out = {}
    rng = np.random.default_rng(5001)
    # ---- F16: all-pairs scoring ----
    base = [rng.integers(0, 256, (20, 80), dtype=np.uint8) for _ in range(400)]
    desc = np.stack([np.roll(base[rng.integers(0, 400)], int(rng.integers(-9, 10)), axis=1) for _ in range(args.n)])
    eng.epsc_score_all(desc, topk=5)                  # warm-up at full size (scratch allocation)
    t0 = time.perf_counter(); idx, score, shift = eng.epsc_score_all(desc, topk=5); dt = time.perf_counter() - t0
    pairs = args.n * (args.n - 1) // 2
    out["epsc_score_all"] = {"N": args.n, "pairs": pairs, "seconds_incl_h2d_d2h": dt, "pairs_per_s": pairs / dt,
                             "byte_absdiff_per_s": pairs * 20 * 1600 / dt, "candidates_found": int((idx >= 0).sum())}
    t0 = time.perf_counter(); orc.epsc_score_all(desc[:args.cpu_rows * 10], topk=5, n_threads=os.cpu_count()); dtc = time.perf_counter() - t0
    pc = (args.cpu_rows * 10) * (args.cpu_rows * 10 - 1) // 2
    out["epsc_score_all"]["cpu_pairs_per_s"] = pc / dtc; out["epsc_score_all"]["cpu_threads"] = os.cpu_count()
    # ---- F19: ICP verify ----
    sc = synth.Scene(seed=1001)
    m = sc.sample_map(n_edge=0, n_surf=200000, seed=3001)
    tid = eng.target_create(m["surf"])
    srcs = []
    for k in range(args.pairs):
        sel = m["surf"][rng.choice(len(m["surf"]), 50000, replace=False)].copy()
        yaw = rng.uniform(-0.03, 0.03); c, s = np.cos(yaw), np.sin(yaw)
        xy = sel[:, :2] @ np.array([[c, -s], [s, c]], np.float32).T + rng.uniform(-0.4, 0.4, 2).astype(np.float32)
        sel[:, :2] = xy; sel[:, :3] += rng.normal(0, 0.01, (len(sel), 3)).astype(np.float32)
        srcs.append(sel)
    eng.icp_verify_batch([(s, tid) for s in srcs])    # warm-up at full size
    t0 = time.perf_counter(); res = eng.icp_verify_batch([(s, tid) for s in srcs]); dt = time.perf_counter() - t0
    out["icp_verify"] = {"pairs": args.pairs, "src_pts": 50000, "tgt_pts": 200000, "seconds_incl_h2d": dt, "pairs_per_s": args.pairs / dt,
                         "mean_iters": float(np.mean([r.iters for r in res])), "converged": int(sum(r.converged for r in res)),
                         "max_fitness": float(max(r.fitness for r in res))}
    t0 = time.perf_counter(); orc.icp(srcs[0], m["surf"]); dtc = time.perf_counter() - t0
    out["icp_verify"]["cpu_pairs_per_s_1thread"] = 1.0 / dtc
    # ---- F17/F18: online detector ----
    from common import loop_keyframes
    kfs = loop_keyframes()
    warm = eng.loop_create(orc.using_map_lut(), use_fepsc=True)      # first pass: staging buffers, history allocation
    for (c, s, sem, lab, od) in kfs:
        eng.loop_detect(warm, c, s, sem, lab, od)
    eng.loop_destroy(warm)
    det = eng.loop_create(orc.using_map_lut(), use_fepsc=True)
    odet = orc.LoopDetector(use_fepsc=True)
    tg, tc, ncand = 0.0, 0.0, 0
    for (c, s, sem, lab, od) in kfs:
        t0 = time.perf_counter(); _, n1, _ = eng.loop_detect(det, c, s, sem, lab, od); tg += time.perf_counter() - t0
        t0 = time.perf_counter(); odet.detect(c, s, sem, lab, od); tc += time.perf_counter() - t0
        ncand += n1
    out["loop_detect"] = {"keyframes": len(kfs), "candidates": ncand, "gpu_ms_per_keyframe": 1e3 * tg / len(kfs), "cpu_ms_per_keyframe": 1e3 * tc / len(kfs),
                          "points_per_keyframe": int(np.mean([len(k[2]) for k in kfs]))}
    # many gated candidates per keyframe (a long mission re-visiting one place): history of H copies of a keyframe
    det2 = eng.loop_create(orc.using_map_lut(), use_fepsc=True)
    c, s, sem, lab, od = kfs[0]
    H = 200
    for k in range(H):
        T = od.copy(); T[0, 3] = 1000.0 + 0.001 * k
        eng.loop_detect(det2, c, s, sem, lab, T)
    T = od.copy(); T[0, 3] = 0.0
    eng.loop_detect(det2, c, s, sem, lab, T)          # jump far away: travel grows by 1000 m
    T = od.copy(); T[0, 3] = 1000.0
    eng.loop_detect(det2, c, s, sem, lab, T)          # and back: the next keyframe sees the whole history gated
    t0 = time.perf_counter(); _, n2, mm = eng.loop_detect(det2, c, s, sem, lab, T); dt = time.perf_counter() - t0
    out["loop_detect_many"] = {"candidates": n2, "gpu_ms": 1e3 * dt, "matched": len(mm)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
