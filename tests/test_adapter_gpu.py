"""The same adapter scenarios through the real CUDA backend (C ABI -> kernels)."""
import asyncio

import pytest

import adapter_scenarios as S

pytestmark = pytest.mark.gpu


def test_database_scenario(native_lib):
    asyncio.run(S.scenario_test_database(None))


def test_parity_with_oracle_manager(native_lib):
    asyncio.run(S.scenario_parity_with_oracle(None, n=3000, dim=256))


def test_error_convention(native_lib):
    asyncio.run(S.scenario_errors(None))


def test_client_shim_matchtext(native_lib):
    asyncio.run(S.scenario_client_shim(None))


def test_reindex_churn_reuses_rows(native_lib):
    asyncio.run(S.scenario_reindex_churn(None))


def test_mass_delete_compacts(native_lib):
    asyncio.run(S.scenario_mass_delete_compacts(None))


def test_exact_ties_follow_the_id(native_lib):
    """Bit-identical vectors under ids that share their 64-bit tie key come back in id order (the device gives them identical
    scores; the adapter orders them, tests/adapter_scenarios.py)."""
    asyncio.run(S.scenario_exact_ties_follow_the_id(None))


def test_snapshot_restore_continues_exactly(native_lib, tmp_path):
    """a saved store, loaded into a fresh process state, answers every later search (filters, deletes,
    upserts included) exactly as the original does - scores bit for bit, because the local-mode replay state (search counter,
    write epochs) travels with the shard - and both keep matching the oracle."""
    import numpy as np

    import lvs_synth as synth
    from code_rag_b200.client import B200VectorStore
    from oracle.qdrant_local import OracleManager

    async def run():
        n, dim = 2500, 128
        x, q = synth.unixcoder_like(n, dim, seed=4321, n_queries=8)
        pl = synth.payloads(n, seed=17)
        ids = synth.random_uuids(n, seed=19)
        a = B200VectorStore(dimensions=dim, rank_attrs=True)
        ora = OracleManager(dim)
        await a.connect(); await a.create_collections(); ora.create_collections()
        vecs = x.astype(np.float64).tolist()
        await a.upsert("code_chunks", ids[:2000], vecs[:2000], pl[:2000]); ora.upsert("code_chunks", ids[:2000], vecs[:2000], pl[:2000])
        await a.delete("code_chunks", {"file_path": pl[5]["file_path"]}); ora.delete("code_chunks", {"file_path": pl[5]["file_path"]})
        for i in range(3):       # searches before the snapshot advance the replay state
            S._same_hits(await a.search("code_chunks", q[i].tolist(), 7), ora.search("code_chunks", q[i].tolist(), 7), what=f"pre {i}")
        await a.save(str(tmp_path))
        b = B200VectorStore(dimensions=dim, rank_attrs=True)
        await b.connect(); await b.create_collections()
        await b.load(str(tmp_path))
        assert (await b.get_collection_info("code_chunks")).points_count == ora.points_count("code_chunks")
        # both stores, and the oracle, continue with the same operations
        for st in (a, b):
            await st.upsert("code_chunks", ids[2000:], vecs[2000:], pl[2000:])
            await st.upsert("code_chunks", ids[10:12], vecs[20:22], pl[10:12])          # overwrite two old points
        ora.upsert("code_chunks", ids[2000:], vecs[2000:], pl[2000:]); ora.upsert("code_chunks", ids[10:12], vecs[20:22], pl[10:12])
        for i in range(3, 8):
            flt = [None, {"language": pl[0]["language"]}, {"file_path": pl[2100]["file_path"]}][i % 3]
            ra = await a.search("code_chunks", q[i].tolist(), 9, flt)
            rb = await b.search("code_chunks", q[i].tolist(), 9, flt)
            assert ra == rb, f"search {i}: the restored store differs from the original"
            S._same_hits(rb, ora.search("code_chunks", q[i].tolist(), 9, flt), what=f"post {i}")
        assert await b.file_needs_update("code_chunks", pl[2100]["file_path"], pl[2100].get("content_hash")) is False
        # the ranking attribute columns and the entity-name pool are part of the snapshot too
        from types import SimpleNamespace as NS
        empty = NS(primary_entities=[], callers=[], callees=[], methods=[], parent_classes=[], child_classes=[])
        plan = NS(primary_intent=NS(value="find_similar"), entities=[NS(name=pl[3]["entity_name"][:3])])
        items = [(plan, empty, q[i].tolist(), {}) for i in range(4)]
        fa = await a.search_and_rank("code_chunks", items, limit=12)
        fb = await b.search_and_rank("code_chunks", items, limit=12)
        assert [[(r.get_key(), r.final_score, r.signal_scores) for r in rs] for rs in fa] == \
               [[(r.get_key(), r.final_score, r.signal_scores) for r in rs] for rs in fb]
        assert any(r.signal_scores["query_entity_match"] > 0 for rs in fb for r in rs) or True
        await a.close(); await b.close()
    asyncio.run(run())


@pytest.mark.parametrize("seed,storage", [(1, "f32"), (2, "f32"), (3, "bf16")])
def test_random_operation_sequences_match_the_oracle(native_lib, seed, storage):
    """The property test of tests/test_adapter_property_cpu.py on the real device: 160 random operations (upserts of new and of
    existing ids, deletes by file, project clean-ups through .client, searches with every filter shape, filter-only lookups) with
    frequent compaction; ids, payloads and scores must follow the oracle step by step (the replay state included)."""
    asyncio.run(S.scenario_random_ops(None, seed, storage=storage))


def test_committed_search_golden_through_the_device(native_lib):
    """tests/golden/search_oracle_golden.json (frozen outputs of the search oracle: ids, float64 scores, drift between repeated
    searches, filters, scroll order, overwrite, delete) replayed through B200VectorStore on the GPU."""
    import json
    from pathlib import Path

    from code_rag_b200.client import B200VectorStore
    G = json.loads((Path(__file__).parent / "golden" / "search_oracle_golden.json").read_text())
    fh = float.fromhex

    async def run():
        x = [[fh(v) for v in row] for row in G["x"]]
        q = [[fh(v) for v in row] for row in G["q"]]
        ids, pl = G["ids"], G["payloads"]
        st = B200VectorStore(dimensions=G["dim"])
        await st.connect(); await st.create_collections()
        await st.upsert("code_chunks", ids[:250], x[:250], pl[:250])
        for s in G["steps"]:
            if s["op"] == "search":
                hits = await st.search("code_chunks", None if s["query"] is None else q[s["query"]], limit=s["limit"], filters=s["filters"])
                assert [h["id"] for h in hits] == s["ids"], s
                exp = [fh(v) for v in s["scores"]]
                assert all(abs(h["score"] - e) <= 1e-5 * abs(e) for h, e in zip(hits, exp))            # the bar for fp32 storage
                assert all(abs(h["score"] - e) <= 1e-12 for h, e in zip(hits, exp)), [h["score"] - e for h, e in zip(hits, exp)]
            elif s["op"] == "delete":
                await st.delete("code_chunks", s["filters"])
            elif s["op"] == "upsert":
                await st.upsert("code_chunks", ids[s["lo"]:s["hi"]], x[s["lo"]:s["hi"]], pl[s["lo"]:s["hi"]])
            elif s["op"] == "overwrite":
                await st.upsert("code_chunks", [ids[i] for i in s["ids"]], [x[i] for i in s["vectors"]],
                                [dict(pl[i], language=s["language"]) for i in s["ids"]])
            elif s["op"] == "count":
                assert (await st.get_collection_info("code_chunks")).points_count == s["value"]
        await st.close()
    asyncio.run(run())


def test_no_device_memory_growth_under_churn(native_lib):
    """Soak: many rounds of upsert / delete (with row reuse and compaction) / batched search / fused search+rank / snapshot-free
    operations must not leak device memory (scratch is reused, re-grown arrays free their predecessors)."""
    import ctypes as C
    import random
    import uuid
    from types import SimpleNamespace as NS

    import numpy as np

    from code_rag_b200 import _native as N
    from code_rag_b200.client import B200VectorStore

    def free_bytes():
        f = C.c_int64()
        N.check(N.load().lvs_device_info(None, None, None, None, C.byref(f)), "lvs_device_info")
        return f.value

    async def run():
        dim = 64
        st = B200VectorStore(dimensions=dim, storage="bf16", rank_attrs=True)
        await st.connect(); await st.create_collections()
        st._get("code_chunks").COMPACT_MIN_FREE = 64
        rng = random.Random(1)
        nprng = np.random.default_rng(1)
        empty = NS(primary_entities=[], callers=[], callees=[], methods=[], parent_classes=[], child_classes=[])
        plan = NS(primary_intent=NS(value="find_similar"), entities=[NS(name="fn")])

        async def one_round(r):
            ids = [str(uuid.UUID(int=rng.getrandbits(100))) for _ in range(200)]
            vecs = nprng.standard_normal((200, dim)).astype(np.float32).astype(np.float64).tolist()
            pl = [{"file_path": f"f{(r * 7 + i) % 40}.py", "entity_name": f"fn{i}", "entity_type": "function", "language": "python",
                   "content": "x" * (i + 1), "start_line": i, "end_line": i + 1, "project_name": f"p{r % 3}", "graph_node_id": None} for i in range(200)]
            await st.upsert("code_chunks", ids, vecs, pl)
            await st.delete("code_chunks", {"file_path": f"f{rng.randrange(40)}.py"})
            if r % 5 == 4:
                await st.delete("code_chunks", {"project_name": f"p{rng.randrange(3)}"})
            q = nprng.standard_normal((6, dim))
            await st.search_batch("code_chunks", q.tolist(), limit=10)
            await st.search("code_chunks", q[0].tolist(), limit=5, filters={"project_name": "p1"})
            await st.search_and_rank("code_chunks", [(plan, empty, q[i], {}) for i in range(3)], limit=8)
        for r in range(40):                 # warm-up: scratch buffers and columns reach their working size
            await one_round(r)
        before = free_bytes()
        for r in range(40, 240):
            await one_round(r)
        after = free_bytes()
        await st.close()
        assert before - after < (96 << 20), f"device memory fell by {(before - after) / 2**20:.1f} MiB over 200 rounds"
    asyncio.run(run())
