#!/usr/bin/env python
"""Adapter-level timings over N GPUs (torchrun, one rank per GPU): what a lattice process on rank 0 sees when it talks to
``ShardedB200VectorStore`` like a ``QdrantManager`` - upsert throughput through the control plane, then searches (one query, a batch,
a payload filter, the filter-only lookup) as wall time per call on rank 0, next to the same calls on a one-GPU ``B200VectorStore``
when N = 1.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 benchmarks/sharded_store_bench.py [rows] [dim] [storage]

One JSON line per measurement.  These are host-API numbers (Python lists in, dicts out), not kernel numbers: bench.py holds the
headline metric and the roofline.  Not yet run (written after round 1's GPU budget was spent)."""
from __future__ import annotations

import asyncio
import json
import statistics
import sys
import time
import uuid
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from code_rag_b200.sharded_store import ShardedB200VectorStore, ShardPlane  # noqa: E402
from code_rag_b200.sharded_store import run as run_sharded  # noqa: E402


async def run(plane: ShardPlane, rows: int, dim: int, storage: str) -> None:
    rng = np.random.default_rng(1)
    store = ShardedB200VectorStore(dimensions=dim, storage=storage, plane=plane)
    await store.connect()
    await store.create_collections()
    batch = 20_000
    t0 = time.perf_counter()
    for lo in range(0, rows, batch):
        n = min(batch, rows - lo)
        x = rng.standard_normal((n, dim), dtype=np.float32)
        ids = [str(uuid.UUID(int=int(rng.integers(1 << 62)) << 64 | (lo + i))) for i in range(n)]
        pl = [{"file_path": f"src/f{(lo + i) // 20}.py", "project_name": f"p{(lo + i) % 8}", "language": "python", "entity_type": "function",
               "entity_name": f"fn{lo + i}", "content": "x" * 40, "start_line": 1, "end_line": 9} for i in range(n)]
        await store.upsert("code_chunks", ids, x, pl)
    dt = time.perf_counter() - t0
    info = await store.get_collection_info("code_chunks")
    print(json.dumps({"what": "upsert", "gpus": plane.world, "rows": rows, "dim": dim, "storage": storage, "seconds": dt,
                      "points_per_s": rows / dt, "shard_points": info.shard_points}), flush=True)
    q = rng.standard_normal((64, dim))

    async def timed(label, fn, reps=30, warm=3):
        ts = []
        for i in range(warm + reps):
            t = time.perf_counter()
            await fn(i)
            if i >= warm:
                ts.append((time.perf_counter() - t) * 1e3)
        print(json.dumps({"what": label, "gpus": plane.world, "rows": rows, "ms_median": statistics.median(ts), "ms_min": min(ts)}), flush=True)

    await timed("search limit=10", lambda i: store.search("code_chunks", q[i % 64].tolist(), 10))
    await timed("search limit=10 filter project", lambda i: store.search("code_chunks", q[i % 64].tolist(), 10, {"project_name": "p3"}))
    await timed("search_batch 64 x limit=10", lambda i: store.search_batch("code_chunks", q.tolist(), 10), reps=10)
    await timed("filter-only lookup limit=1", lambda i: store.search("code_chunks", None, 1, {"entity_name": f"fn{i}", "file_path": f"src/f{i // 20}.py"}))
    await timed("delete one file + re-upsert 20 chunks", lambda i: _reindex(store, rng, dim, i))
    await store.close()


async def _reindex(store, rng, dim, i):
    await store.delete("code_chunks", {"file_path": f"src/f{i}.py"})
    x = rng.standard_normal((20, dim), dtype=np.float32)
    pl = [{"file_path": f"src/f{i}.py", "project_name": "p0", "language": "python", "entity_type": "function", "entity_name": f"re{i}_{j}",
           "content": "y" * 30, "start_line": j, "end_line": j + 3} for j in range(20)]
    await store.upsert("code_chunks", [str(uuid.uuid4()) for _ in range(20)], x, pl)


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    storage = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    run_sharded(lambda plane: run(plane, rows, dim, storage))


if __name__ == "__main__":
    main()
