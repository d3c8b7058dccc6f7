#!/usr/bin/env python
"""Multi-GPU correctness check of the sharded path (run under torchrun, one rank per GPU):
    torchrun --nproc-per-node N tests/sharded_check_torchrun.py
Every rank holds a contiguous row shard; results of ShardedSearcher (device async, sync and host submit/wait paths, Q = 1 and
batched) must equal the CPU oracle over the whole corpus, for both exchange modes."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]   # a checker (it uses the oracle), hence under tests/; not collected by pytest

import lvs_synth as synth  # noqa: E402
from code_rag_b200.collection import DeviceCollection  # noqa: E402
from code_rag_b200.sharded import ShardedSearcher, shard_bounds  # noqa: E402
from oracle.qdrant_local import OracleCollection  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, dim, k = 40_000, 768, 10
    x, q = synth.unit_rows(n, dim, seed=5678, n_queries=24)
    xb = synth.bf16_round(x)
    lo, hi = shard_bounds(n, world)[rank]
    ok = True
    for mode in ("p2p", "nccl"):
        os.environ["LATTICE_B200_EXCHANGE"] = mode
        ora = OracleCollection(dim)
        ora.upsert_rows_f32(0, xb, [None] * n)
        shard = DeviceCollection(f"chk_{mode}_{rank}", dim, storage="bf16", row_base=lo, device=local)
        shard.upsert(xb[lo:hi], rows=np.arange(lo, hi))
        ss = ShardedSearcher(shard)
        assert ss.exchange_mode == mode
        qi = 0
        # host path, one at a time
        for _ in range(4):
            s, r, t, c, f = ss.search(q[qi], k)
            rows_o, scores_o = ora.search_topk_rows(q[qi], k)
            ok &= bool(np.array_equal(r[0], rows_o) and np.abs(s[0] - scores_o).max() < 1e-12 and f.sum() == 0)
            qi += 1
        # host path, pipelined (3 in flight)
        hs = [ss.submit(q[qi + j], k) for j in range(3)]
        for j, h in enumerate(hs):
            s, r, t, c, f = ss.wait(h)
            rows_o, scores_o = ora.search_topk_rows(q[qi + j], k)
            ok &= bool(np.array_equal(r[0], rows_o) and np.abs(s[0] - scores_o).max() < 1e-12)
        qi += 3
        # device path, batched (K1 passes for 4 queries, then K2 for 13)
        for nb in (4, 13):
            dq = torch.from_numpy(q[qi:qi + nb].astype(np.float64)).cuda()
            torch.cuda.synchronize()
            s, r, t, c, f = ss.search_device(dq, k)
            s, r = s.cpu().numpy(), r.cpu().numpy()
            for j in range(nb):
                rows_o, scores_o = ora.search_topk_rows(q[qi + j], k)
                ok &= bool(np.array_equal(r[j], rows_o) and np.abs(s[j] - scores_o).max() < 1e-12)
            qi += nb
        ss.close()
        shard.close()
        if rank == 0:
            print(f"mode {mode}: {'OK' if ok else 'MISMATCH'} ({qi} queries, world {world})", flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(t.item()) != 1:
        sys.exit(1)
    if rank == 0:
        print("sharded check OK", flush=True)


if __name__ == "__main__":
    main()
