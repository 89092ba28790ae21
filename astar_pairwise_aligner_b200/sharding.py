"""Multi-GPU sharding of a batch of independent pairs: one process per GPU, contiguous shards balanced by
bases, no data-path collective; torch.distributed (NCCL on GPUs, gloo in the CPU tests) only gathers the results.

The reference has no parallelism of its own; pairs are independent because a fresh aligner is built per call
(astarpa2/src/lib.rs:50-53)."""
import ctypes as C

import numpy as np


def shard_bounds(a_off, b_off, world):
    """Contiguous [start, end) pair ranges, one per rank, balanced by |a| + |b|."""
    a_off = np.asarray(a_off, dtype=np.int64)
    b_off = np.asarray(b_off, dtype=np.int64)
    n = len(a_off) - 1
    work = (a_off[1:] - a_off[:-1]) + (b_off[1:] - b_off[:-1])
    cum = np.concatenate([[0], np.cumsum(work)])
    total = int(cum[-1])
    bounds, start = [], 0
    for r in range(world):
        if r == world - 1:
            end = n
        else:
            target = total * (r + 1) / world
            end = int(np.searchsorted(cum, target, side="left"))
            end = min(max(end, start), n)
        bounds.append((start, end))
        start = end
    return bounds


def slice_batch(a_all, a_off, b_all, b_off, start, end):
    a_off = np.asarray(a_off, dtype=np.int64)
    b_off = np.asarray(b_off, dtype=np.int64)
    return (a_all[a_off[start]:a_off[end]], a_off[start:end + 1] - a_off[start],
            b_all[b_off[start]:b_off[end]], b_off[start:end + 1] - b_off[start])


def align_batch_sharded(a_all, a_off, b_all, b_off, preset, trace, align_fn, dist=None):
    """Every rank aligns its shard with `align_fn(a_all, a_off, b_all, b_off, preset, trace) -> (costs, cigars)`;
    rank 0 returns the results of the whole batch in input order (other ranks return None).
    `dist` is an initialised torch.distributed module (or None for a single process)."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    bounds = shard_bounds(a_off, b_off, world)
    s, e = bounds[rank]
    costs, cigars = align_fn(*slice_batch(a_all, a_off, b_all, b_off, s, e), preset, trace)
    if dist is None:
        return costs, cigars
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((s, e, np.asarray(costs), cigars), gathered, dst=0)
    if rank != 0:
        return None
    n = len(a_off) - 1
    out_costs = np.zeros(n, dtype=np.int64)
    out_cigars = [None] * n if trace else None
    for (gs, ge, gc, gcig) in gathered:
        out_costs[gs:ge] = gc
        if trace:
            out_cigars[gs:ge] = gcig
    return out_costs, out_cigars


def gpu_align_fn(device):
    """The product aligner for `align_batch_sharded`: this rank's shard through the multi-GPU C-ABI entry
    (apa_align_batch_multi) on its own device - (costs, list of CIGAR strings or None)."""
    import astar_pairwise_aligner_b200 as A

    def fn(a_all, a_off, b_all, b_off, preset, trace):
        costs, pool, off, ln, _ = A.align_batch_multi([device], a_all, a_off, b_all, b_off, preset, trace)
        cigars = None
        if trace and pool.value:
            cigars = [C.string_at(pool.value + int(off[p]), int(ln[p])).decode() for p in range(len(costs))]
        elif trace:
            cigars = []
        A.free_pool(pool)
        return costs, cigars
    return fn
