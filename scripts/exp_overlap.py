"""Experiment: one 256-frame step as ONE batch on one stream vs TWO 128-frame half-batches on two streams (two contexts).
Prints ms per 256 frames for both.  Not a bench line."""
import ctypes as C
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from lis_slam_b200 import engine as E
from lis_slam_b200 import workload

def main():
    nsplit = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda", 0)
    F = 256
    wl = workload.frame_batch(F=F, n_maps=8, n_sweeps=8, seed=0)
    arena_np, offs = workload.pack_frame_arena(wl)
    arena_dev = torch.from_numpy(arena_np).to(dev)
    guess = torch.from_numpy(np.stack([r["guess"] for r in wl["regs"]]).astype(np.float32)).to(dev)
    base = arena_dev.data_ptr()
    prm = E.frame_params("A", early_exit=0, max_iters=10)
    def make(n_eng):
        engs = []
        for k in range(n_eng):
            st = torch.cuda.Stream(dev)
            eng = E.Engine(device=0, stream=st.cuda_stream)
            mids = [eng.map_create(m["corner"], m["surf"], gate_hint=1.0) for m in wl["maps"]]
            engs.append((eng, st, mids))
        return engs
    def run(engs, steps):
        n_eng = len(engs)
        per = F // n_eng
        items = []
        poses = []
        ress = []
        for k, (eng, st, mids) in enumerate(engs):
            it = (E.FrameItem * per)()
            for b in range(per):
                r = wl["regs"][k * per + b]; op, og, n = offs[k * per + b]
                it[b] = E.FrameItem(base + op, base + og, n, mids[r["map"]])
            items.append(it)
            poses.append(torch.empty(per, 6, dtype=torch.float32, device=dev))
            ress.append(torch.empty(per * C.sizeof(E.LmResult), dtype=torch.uint8, device=dev))
        def one():
            for k, (eng, st, mids) in enumerate(engs):
                with torch.cuda.stream(st):
                    poses[k].copy_(guess[k * per:(k + 1) * per])
                eng.frames_batch_dev(items[k], per, poses[k].data_ptr(), prm, ress[k].data_ptr())
        for _ in range(3): one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps): one()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / steps
        return ms, torch.cat(poses).cpu().numpy()
    e1 = make(1)
    ms1, p1 = run(e1, 10)
    e2 = make(nsplit)
    ms2, p2 = run(e2, 10)
    print("one stream: %.3f ms / 256 frames;  %d streams: %.3f ms;  poses identical: %s" % (ms1, nsplit, ms2, np.array_equal(p1, p2)))

main()
