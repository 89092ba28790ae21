"""CPU tests of the N > 1 path: world_size-2 gloo processes shard a batch, align their shard (with the oracle standing
in for the GPU aligner — this is host logic only) and gather to rank 0; the result must equal the single-process one."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np
import torch.distributed as dist
import astar_pairwise_aligner_b200 as A
from astar_pairwise_aligner_b200.sharding import align_batch_sharded
import oracle_lib as O

def oracle_align(a_all, a_off, b_all, b_off, preset, trace):
    costs, cigs = [], []
    for p in range(len(a_off) - 1):
        c, cg, _ = O.align(bytes(a_all[a_off[p]:a_off[p+1]]), bytes(b_all[b_off[p]:b_off[p+1]]), preset, trace)
        costs.append(c); cigs.append(cg)
    return np.array(costs, dtype=np.int64), cigs

dist.init_process_group("gloo")
lens = [50, 3000, 0, 700, 1200, 10, 2500, 64]
pairs = [A.generate_pair(n, 0.08, k % 4, 100 + k) for k, n in enumerate(lens)]
args = A._concat(pairs)
res = align_batch_sharded(*args, 1, True, oracle_align, dist)
if dist.get_rank() == 0:
    ref = oracle_align(*args, 1, True)
    assert (res[0] == ref[0]).all() and res[1] == ref[1]
    print("SHARD_OK", len(lens))
dist.destroy_process_group()
'''


def test_shard_bounds_cover_and_balance():
    from astar_pairwise_aligner_b200.sharding import shard_bounds
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 1000):
        la = rng.integers(0, 5000, size=n)
        a_off = np.concatenate([[0], np.cumsum(la)])
        b_off = np.concatenate([[0], np.cumsum(la + rng.integers(0, 50, size=n))])
        for world in (1, 2, 4, 8):
            b = shard_bounds(a_off, b_off, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[k][1] == b[k + 1][0] for k in range(world - 1))
            if n == 1000:
                w = [(a_off[e] - a_off[s]) + (b_off[e] - b_off[s]) for s, e in b]
                assert max(w) - min(w) <= 2 * 10100  # within one pair (|a|+|b| <= ~10 050) per boundary


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29513", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, er) in zip(procs, outs):
        assert p.returncode == 0, er[-2000:]
    assert "SHARD_OK 8" in outs[0][0]
