"""Host logic of the QdrantManager-compatible adapter, on CPU: the device is replaced by an oracle-backed fake."""
import asyncio

import adapter_scenarios as S
from helpers import ExactTieDevice, FakeDevice


def test_database_scenario():
    asyncio.run(S.scenario_test_database(FakeDevice))


def test_parity_with_oracle_manager():
    asyncio.run(S.scenario_parity_with_oracle(FakeDevice, n=1200, dim=64))


def test_error_convention():
    asyncio.run(S.scenario_errors(FakeDevice))


def test_client_shim_matchtext():
    asyncio.run(S.scenario_client_shim(FakeDevice))


def test_reindex_churn_reuses_rows():
    asyncio.run(S.scenario_reindex_churn(FakeDevice))


def test_mass_delete_compacts():
    asyncio.run(S.scenario_mass_delete_compacts(FakeDevice))


def test_random_operation_sequences_match_the_oracle():
    for seed in (1, 2):
        asyncio.run(S.scenario_random_ops(FakeDevice, seed))


def test_exact_ties_follow_the_id():
    asyncio.run(S.scenario_exact_ties_follow_the_id(ExactTieDevice))


def test_snapshot_host_half_round_trip(tmp_path, monkeypatch):
    """B200VectorStore.save / load on CPU: the host half (ids, payloads, dictionaries, free rows, tie-key counts) survives; the
    device half is the fake's pickle here (the raw .lvs format is covered on the GPU and in test_abi.py)."""
    import numpy as np

    import lvs_synth as synth
    from code_rag_b200 import client
    monkeypatch.setattr(client, "DeviceCollection", type("D", (), {"load_snapshot": staticmethod(
        lambda path, name=None, device=0: FakeDevice.load_snapshot(path, name=name))}))

    async def run():
        st = client.B200VectorStore(dimensions=16, _device_factory=FakeDevice)
        await st.connect(); await st.create_collections()
        x, _ = synth.unit_rows(30, 16, seed=1)
        ids = synth.random_uuids(30, 3)
        pl = [{"file_path": f"f{i % 3}.py", "entity_name": f"e{i}"} for i in range(30)]
        await st.upsert("code_chunks", ids, x.astype(np.float64).tolist(), pl)
        await st.delete("code_chunks", {"file_path": "f1.py"})
        await st.save(str(tmp_path))
        before = await st.search("code_chunks", x[0].tolist(), 5, {"file_path": "f0.py"})
        await st.delete("code_chunks", {"file_path": "f0.py"})
        await st.load(str(tmp_path))
        coll = st._get("code_chunks")
        assert (await st.get_collection_info("code_chunks")).points_count == 20 == len(coll.tie_counts) and not coll.dup_keys
        assert len(coll.free_rows) == 10
        after = await st.search("code_chunks", x[0].tolist(), 5, {"file_path": "f0.py"})
        assert [(h["id"], h["payload"]) for h in after] == [(h["id"], h["payload"]) for h in before]
        await st.upsert("code_chunks", ids[:3], x[:3].astype(np.float64).tolist(), pl[:3])       # reuse of freed rows after a load
        assert coll.dev.rows == 30 and (await st.get_collection_info("code_chunks")).points_count == 21
        await st.close()
    asyncio.run(run())


def test_edge_cases():
    asyncio.run(S.scenario_edge_cases(FakeDevice))
