#!/usr/bin/env python
"""Run on a box with >= 2 GPUs (not collected by pytest): the adapter bound to cuda:1 must do ALL of its work there, although its
C calls come from asyncio worker threads whose CUDA current device starts at 0 (the library binds the calling thread at every entry
point).  Checks results against the oracle and that cuda:0's free memory did not move while cuda:1's did."""
import asyncio
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

import adapter_scenarios as S  # noqa: E402
import lvs_synth as synth  # noqa: E402
from code_rag_b200.client import B200VectorStore  # noqa: E402
from oracle.qdrant_local import OracleManager  # noqa: E402


async def main():
    assert torch.cuda.device_count() >= 2
    free0, free1 = torch.cuda.mem_get_info(0)[0], torch.cuda.mem_get_info(1)[0]
    n, dim = 200_000, 256
    x, q = synth.unixcoder_like(n, dim, seed=1, n_queries=6)
    pl = synth.payloads(n, seed=2)
    ids = synth.random_uuids(n, seed=3)
    st = B200VectorStore(dimensions=dim, device=1, rank_attrs=True)
    ora = OracleManager(dim)
    await st.connect(); await st.create_collections(); ora.create_collections()
    for lo in range(0, n, 20_000):
        v = x[lo:lo + 20_000].astype(np.float64).tolist()
        await st.upsert("code_chunks", ids[lo:lo + 20_000], v, pl[lo:lo + 20_000])
        ora.upsert("code_chunks", ids[lo:lo + 20_000], v, pl[lo:lo + 20_000])
    # concurrent awaits land on different worker threads
    for rnd in range(3):
        got = await asyncio.gather(*[st.search("code_chunks", q[i].tolist(), 10, None if i % 2 else {"language": pl[0]["language"]})
                                     for i in range(6)])
        # gather runs the coroutines concurrently but the collection lock serialises them in submission order
        for i in range(6):
            S._same_hits(got[i], ora.search("code_chunks", q[i].tolist(), 10, None if i % 2 else {"language": pl[0]["language"]}), what=f"r{rnd} q{i}")
    used0, used1 = free0 - torch.cuda.mem_get_info(0)[0], free1 - torch.cuda.mem_get_info(1)[0]
    await st.close()
    print(f"cuda:0 delta {used0 / 2**20:.1f} MiB, cuda:1 delta {used1 / 2**20:.1f} MiB")
    assert used1 > 150 * 2**20, "the shard (200k x 256 fp32 = 195 MiB) must live on cuda:1"
    assert used0 < 64 * 2**20, "nothing of the store may land on cuda:0"
    print("device binding OK")


if __name__ == "__main__":
    asyncio.run(main())
