"""Multi-GPU sharding of the hot path.  Frames and loop-closure query rows are independent units:
each rank owns a shard, there is no data-path collective, and the single exchange step is one all-gather of the
fixed-size result records (6-DoF pose + status) over the process group (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def frame_shard(n_units, rank, world):
    """Contiguous block [lo, hi) of `n_units` independent frames owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def cyclic_rows(n_rows, rank, world):
    """Loop-closure query rows owned by `rank`: row q scores q history entries (triangular load), so rows are dealt
    cyclically to balance the work."""
    return np.arange(rank, n_rows, world, dtype=np.int64)


def deal_round_robin(n_items, rank, world):
    """Items (loop-closure candidates in canonical order, known to every rank after the candidate exchange) verified by
    `rank`: every world-th one, so that the ICP work differs by at most one pair between ranks whatever rows produced them."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def canonical_candidates(table):
    """table: (n, 4) float32 rows [query q, history j, score, shift] gathered from all ranks, j < 0 = empty slot, each rank's
    block in (row, slot) order.  Returns the valid rows ordered by (q, slot) - the same list on every rank."""
    t = np.asarray(table)
    pos = np.arange(len(t))
    keep = t[:, 1] >= 0
    t, pos = t[keep], pos[keep]
    return t[np.lexsort((pos, t[:, 0]))]


def gather_results(local, n_units, world, dist=None, device=None):
    """All-gathers per-rank result blocks (equal record width; ragged shard sizes are padded to the largest shard)
    into the global array ordered by unit id.  `local`: (n_local, width) float32 torch tensor."""
    import torch
    if world == 1:
        return local
    width = local.shape[1]
    cap = (n_units + world - 1) // world
    pad = torch.zeros(cap, width, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * cap, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        lo, hi = frame_shard(n_units, r, world)
        parts.append(out[r * cap: r * cap + (hi - lo)])
    return torch.cat(parts, 0)
