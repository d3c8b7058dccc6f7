#!/usr/bin/env python
"""Multi-GPU check of the sharded ADAPTER (run under torchrun, one rank per GPU; not collected by pytest):
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tests/sharded_store_check_torchrun.py
Rank 0 drives a ``ShardedB200VectorStore`` over the real shards through the scenarios of tests/adapter_scenarios.py (the reference's
database test, parity with the oracle's QdrantManager over every filter shape, errors, the ``.client`` shim) and the sharded
specifics of tests/test_sharded_store_cpu.py (placement, overwrite in place, per-shard compaction, scroll order, a worker-side
failure); the other ranks serve.  The same file runs on CPU (gloo, oracle-backed shards) as tests/test_sharded_store_cpu.py."""
import sys
from pathlib import Path
from types import SimpleNamespace as NS

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

from code_rag_b200.sharded_store import ShardedB200VectorStore, run  # noqa: E402


async def checks(plane):
    import adapter_scenarios as S
    from test_sharded_store_cpu import _sharded_specifics
    for storage in ("f32", "bf16"):
        factory = NS(make_store=lambda **kw: ShardedB200VectorStore(plane=plane, **{"storage": storage, **kw}))
        await S.scenario_test_database(factory)
        if storage == "f32":          # bf16 storage rounds the inputs: the 1e-5 bar of the scenario is the fp32 one
            await S.scenario_parity_with_oracle(factory, n=3000, dim=256)
            await _sharded_specifics(factory.make_store, plane.world)
            await S.scenario_exact_ties_follow_the_id(factory)
        await S.scenario_random_ops(factory, 4, storage=storage)
        await S.scenario_errors(factory)
        if storage == "f32":          # compares with the oracle on unrounded inputs at the fp32 bar
            await S.scenario_edge_cases(factory)
        await S.scenario_client_shim(factory)
        if storage == "f32":
            # embedding on the GPUs that will search (SURVEY section 8f row 4 over N GPUs): every rank builds the same small encoder,
            # chunks are embedded on the rank that owns them; expected vectors from the numpy oracle of the encoder
            from oracle import roberta_encoder as R
            from test_sharded_store_cpu import _sharded_upsert_tokens
            r = {"vocab": 600, "hidden": 128, "layers": 2, "intermediate": 256, "max_pos": 64, "seed": 3}
            sd = R.random_state_dict(r["vocab"], r["hidden"], r["layers"], r["intermediate"], r["max_pos"], seed=r["seed"])
            await _sharded_upsert_tokens(factory.make_store, plane.world, {"random": r, "n_heads": 2, "pad_id": 1},
                                         lambda t: R.encode(sd, t, n_layers=2, n_heads=2, pad_id=1)[1])
            print(f"upsert_tokens over {plane.world} GPU(s) (one encoder per rank): OK", flush=True)
        print(f"sharded store [{storage}] over {plane.world} GPU(s): OK", flush=True)


if __name__ == "__main__":
    run(checks)
