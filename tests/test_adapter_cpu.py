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


def test_scenarios_through_the_polled_search_path(monkeypatch):
    """Large collections are searched by submit + poll on the event loop (no worker thread); force that path on the small ones."""
    from code_rag_b200 import client
    monkeypatch.setattr(client, "_INLINE_SEARCH_BYTES", -1)
    asyncio.run(S.scenario_parity_with_oracle(FakeDevice, n=600, dim=32))
    asyncio.run(S.scenario_reindex_churn(FakeDevice))
    asyncio.run(S.scenario_errors(FakeDevice))
    for seed in (3,):
        asyncio.run(S.scenario_random_ops(FakeDevice, seed))


def test_gathered_searches_pipeline_and_writers_wait(monkeypatch):
    """asyncio.gather of searches on one collection (query/engine.py:142-146): every search sees the collection as it was when it
    was submitted, up to four are in flight, and a write issued meanwhile goes in between two searches, never inside one."""
    import threading

    import numpy as np

    import lvs_synth as synth
    from code_rag_b200 import client
    monkeypatch.setattr(client, "_INLINE_SEARCH_BYTES", -1)

    peak = [0]

    class Counting(FakeDevice):
        def search_submit(self, queries, k, want=None):
            t = super().search_submit(queries, k, want)
            peak[0] = max(peak[0], len(self._tickets))
            return t

    async def run():
        st = client.B200VectorStore(dimensions=16, _device_factory=Counting)
        await st.connect(); await st.create_collections()
        x, _ = synth.unit_rows(200, 16, seed=5)
        ids = synth.random_uuids(200, 9)
        pl = [{"file_path": f"f{i % 7}.py", "entity_name": f"e{i}"} for i in range(200)]
        await st.upsert("code_chunks", ids, x.astype(np.float64).tolist(), pl)
        qs = [x[i].tolist() for i in range(24)]
        one_by_one = [await st.search("code_chunks", q, 5) for q in qs]
        together = await asyncio.gather(*[st.search("code_chunks", q, 5) for q in qs])
        ids_of = lambda hits: [h["id"] for h in hits]  # noqa: E731  (scores drift with local mode's in-place re-normalisation)
        assert [ids_of(h) for h in together] == [ids_of(h) for h in one_by_one]
        assert 2 <= peak[0] <= 4
        # a delete in the middle of a gather: each search answers from before or from after it, as a whole
        before = {i: ids_of(r) for i, r in enumerate(one_by_one)}
        res = await asyncio.gather(*[st.search("code_chunks", q, 5) for q in qs[:12]], st.delete("code_chunks", {"file_path": "f3.py"}),
                                   *[st.search("code_chunks", q, 5) for q in qs[12:]])
        after = [ids_of(await st.search("code_chunks", q, 5)) for q in qs]
        hits = res[:12] + res[13:]
        assert any(before[i] != after[i] for i in range(24))
        for i, h in enumerate(hits):
            assert ids_of(h) == before[i] or ids_of(h) == after[i]
            assert all(p["payload"]["entity_name"] == f"e{ids.index(p['id'])}" for p in h)
        assert st._collections["code_chunks"].lock.readers == 0
        await st.close()
    asyncio.run(run())

    # the lock itself: an exclusive acquire waits for the readers in flight and keeps new ones out meanwhile
    lk = client._CollectionLock()
    assert lk.try_enter(4)
    lk.entered()
    got = []
    th = threading.Thread(target=lambda: (lk.acquire(), got.append(1), lk.release()))
    th.start()
    th.join(0.2)
    assert th.is_alive() and not got
    for _ in range(100):                          # the writer is waiting by now (or will be): readers are refused once it is
        if not lk.try_enter(4):
            break
        lk.release()
        th.join(0.01)
    assert lk.try_reenter()
    lk.reader_done()
    th.join(5)
    assert got == [1] and lk.readers == 0


def test_polled_searches_under_concurrent_writes(monkeypatch):
    """Searches on the event loop (submit + poll), upserts and deletes in worker threads, all on one collection for a while: no
    exception, no reader left behind, and every hit pairs an id with ITS payload (a write never lands between a search's
    submission and the moment its rows become payloads)."""
    import random
    import time

    import numpy as np

    import lvs_synth as synth
    from code_rag_b200 import client
    monkeypatch.setattr(client, "_INLINE_SEARCH_BYTES", -1)

    async def run():
        st = client.B200VectorStore(dimensions=16, _device_factory=FakeDevice)
        await st.connect(); await st.create_collections()
        n = 300
        x, _ = synth.unit_rows(n, 16, seed=8)
        ids = synth.random_uuids(n, 10)
        tag = {pid: i for i, pid in enumerate(ids)}
        pl = lambda i, gen: {"file_path": f"f{i % 9}.py", "entity_name": f"e{i}", "gen": gen}  # noqa: E731
        await st.upsert("code_chunks", ids[:200], x[:200].astype(np.float64).tolist(), [pl(i, 0) for i in range(200)])
        stop = time.perf_counter() + 1.5
        rnd = random.Random(3)
        counts = {"search": 0, "write": 0}

        async def searcher(seed):
            r = random.Random(seed)
            while time.perf_counter() < stop:
                hits = await st.search("code_chunks", x[r.randrange(n)].tolist(), 6, {"file_path": f"f{r.randrange(9)}.py"} if r.random() < 0.3 else None)
                for h in hits:
                    assert h["payload"]["entity_name"] == f"e{tag[h['id']]}", h
                counts["search"] += 1

        async def writer():
            gen = 1
            while time.perf_counter() < stop:
                lo = rnd.randrange(0, n - 20)
                sel = list(range(lo, lo + 20))
                if rnd.random() < 0.5:
                    await st.upsert("code_chunks", [ids[i] for i in sel], x[sel].astype(np.float64).tolist(), [pl(i, gen) for i in sel])
                else:
                    await st.delete("code_chunks", {"file_path": f"f{rnd.randrange(9)}.py"})
                gen += 1
                counts["write"] += 1
                await asyncio.sleep(0)

        await asyncio.gather(*[searcher(s) for s in range(6)], writer(), writer())
        lock = st._collections["code_chunks"].lock
        assert lock.readers == 0 and lock.acquire(blocking=False)
        lock.release()
        assert counts["search"] > 50 and counts["write"] > 5, counts
        await st.close()
    asyncio.run(run())
