"""N>1 host logic on CPU: world_size-2 gloo processes run the shard partition, the packed all-gather and the merge
rule; the merged answer must equal the single-collection oracle answer.  (The merge itself is a CUDA kernel in the
product; here its rule is restated in numpy - tests/helpers.py - because no GPU exists in this container.)"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p_ in (str(ROOT), str(ROOT / "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

from code_rag_b200.sharded import allgather_packed, shard_bounds  # noqa: E402


def test_shard_bounds_partition():
    for n, w, a in [(10, 3, 1), (10_000_000, 8, 250_000), (7, 8, 1), (0, 2, 1), (1_000_001, 4, 1000)]:
        b = shard_bounds(n, w, a)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 2 * max(a, 1) or n < w * a
    with pytest.raises(ValueError):
        shard_bounds(10, 0)


def _worker(rank, world, port, n, dim, Q, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lvs_synth as synth
    from helpers import merge_lists_numpy
    from oracle.qdrant_local import OracleCollection
    x, q = synth.unit_rows(n, dim, seed=5678, n_queries=Q)
    lo, hi = shard_bounds(n, world)[rank]
    ora = OracleCollection(dim)
    ora.upsert_rows_f32(lo, x[lo:hi], [None] * (hi - lo))
    local = torch.zeros((3, Q, k), dtype=torch.int64)
    local[1].fill_(-1)
    for i in range(Q):
        rows, scores = ora.search_topk_rows(q[i].astype(np.float64), k)
        m = len(rows)
        local[0, i, :m] = torch.from_numpy(scores.view(np.int64).copy())
        local[1, i, :m] = torch.from_numpy(rows + lo)           # global rows
        local[2, i, :m] = torch.from_numpy(rows + lo)           # tie key = global row
    gathered = torch.zeros((world, 3, Q, k), dtype=torch.int64)
    allgather_packed(local, gathered)
    g = gathered.numpy()
    s, r, t = merge_lists_numpy(g[:, 0].view(np.float64), g[:, 1], g[:, 2].view(np.uint64), k)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), s=s, r=r)
    dist.destroy_process_group()


def test_world2_allgather_merge_equals_single(tmp_path):
    import lvs_synth as synth
    from oracle.qdrant_local import OracleCollection
    n, dim, Q, k, world = 4001, 64, 3, 10, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, dim, Q, k, str(tmp_path)), nprocs=world, join=True)
    x, q = synth.unit_rows(n, dim, seed=5678, n_queries=Q)
    ora = OracleCollection(dim)
    ora.upsert_rows_f32(0, x, [None] * n)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for i in range(Q):
        rows, scores = ora.search_topk_rows(q[i].astype(np.float64), k)
        for r in range(world):
            assert np.array_equal(res[r]["r"][i], rows)
            assert np.allclose(res[r]["s"][i], scores, rtol=0, atol=1e-12)
