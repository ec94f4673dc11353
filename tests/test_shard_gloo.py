"""world_size-2 gloo test (CPU) of the multi-rank path: frame sharding, cyclic loop-closure rows and the single
all-gather of result records reproduce the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lis_slam_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.frame_shard(n_units, rank, world)
    # stand-in for the per-frame registration result: a deterministic function of the unit id (7 floats: pose6 + status)
    ids = torch.arange(lo, hi, dtype=torch.float32)
    local = torch.stack([ids * (k + 1) + 0.5 * k for k in range(7)], 1)
    full = shard.gather_results(local, n_units, world, dist=dist)
    rows = shard.cyclic_rows(n_units, rank, world)
    work = torch.tensor([float(rows.sum())])          # triangular load: row q costs q
    dist.all_reduce(work, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((full.numpy(), float(work[0])))
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process():
    n_units, world = 37, 2                      # ragged shards (19 + 18)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, max_work = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(n_units, dtype=np.float32)
    ref = np.stack([ids * (k + 1) + 0.5 * k for k in range(7)], 1)
    assert np.array_equal(full, ref)
    total = n_units * (n_units - 1) / 2
    assert max_work <= 0.55 * total               # cyclic rows balance the triangular loop-closure load


def test_shard_covers_all_units():
    for n in (0, 1, 5, 4096):
        for w in (1, 2, 4, 8):
            spans = [shard.frame_shard(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            rows = np.concatenate([shard.cyclic_rows(n, r, w) for r in range(w)])
            assert sorted(rows.tolist()) == list(range(n))


def test_candidate_deal_is_balanced_and_identical_on_every_rank():
    """Loop closure at N > 1: every rank turns the gathered candidate table into the same canonical list (by query row, then
    slot) and verifies every world-th entry: the shares cover the list exactly once and differ by at most one pair, however
    unevenly the candidates fell on the ranks' own rows."""
    from lis_slam_b200 import shard
    rng = np.random.default_rng(4)
    world, N, topk = 3, 64, 5
    blocks = []
    for r in range(world):
        rows = shard.cyclic_rows(N, r, world)
        cap = (N + world - 1) // world
        t = np.full((cap * topk, 4), -1.0, np.float32)
        idx = rng.integers(-1, 12, (len(rows), topk)) * (rng.random((len(rows), topk)) < (0.2 + 0.3 * r))   # rank 2 has many more
        idx[idx == 0] = -1
        t[: len(rows) * topk, 0] = np.repeat(rows, topk); t[: len(rows) * topk, 1] = idx.reshape(-1)
        t[: len(rows) * topk, 2] = rng.random(len(rows) * topk)
        blocks.append(t)
    table = np.concatenate(blocks)
    g = shard.canonical_candidates(table)
    assert len(g) == int((table[:, 1] >= 0).sum()) and np.all(np.diff(g[:, 0]) >= 0)
    shares = [g[shard.deal_round_robin(len(g), r, world)] for r in range(world)]
    assert sum(len(s) for s in shares) == len(g) and max(map(len, shares)) - min(map(len, shares)) <= 1
    merged = np.concatenate(shares)
    assert np.array_equal(merged[np.lexsort((merged[:, 2], merged[:, 0]))], g[np.lexsort((g[:, 2], g[:, 0]))])
