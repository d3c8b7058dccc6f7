"""Property test of the adapter's host-side bookkeeping (ids <-> rows, dictionary codes, reuse of deleted rows, compaction):
random sequences of the operations the reference issues (upsert of new and of existing ids, delete by filter, cleanup through
``manager.client.delete``, search with and without filters, filter-only lookups, counts) must leave B200VectorStore and the
oracle's QdrantManager restatement in agreement after every step.  The device is the FakeDevice (arithmetic by the oracle), so
any disagreement is a host-logic bug."""
import asyncio
import uuid
from types import SimpleNamespace as NS

import numpy as np
from hypothesis import HealthCheck, example, given, settings
from hypothesis import strategies as st

from code_rag_b200.client import B200VectorStore
from helpers import FakeDevice
from oracle.qdrant_local import OracleManager

DIM = 8
FILES = [f"src/f{i}.py" for i in range(5)]
PROJECTS = ["p0", "p1"]
CODE = "code_chunks"

op = st.one_of(
    st.tuples(st.just("upsert"), st.lists(st.integers(0, 39), min_size=1, max_size=6), st.integers(0, 10_000)),
    st.tuples(st.just("delete_file"), st.sampled_from(FILES)),
    st.tuples(st.just("cleanup_project"), st.sampled_from(PROJECTS)),
    st.tuples(st.just("search"), st.integers(0, 10_000), st.sampled_from([None, "file", "project", "both", "lang"]), st.integers(1, 12)),
    st.tuples(st.just("scroll"), st.sampled_from(FILES)),
)


def _payload(slot: int, salt: int) -> dict:
    return {"file_path": FILES[(slot + salt) % len(FILES)], "project_name": PROJECTS[(slot * 7 + salt) % 2], "language": ("python", "go")[salt % 2],
            "entity_type": "function", "entity_name": f"fn{slot}_{salt % 3}", "content": "x" * (slot + 1), "start_line": slot, "end_line": slot + 1,
            "content_hash": f"h{salt % 4}"}


# derandomize: the gate explores the same 100 sequences on every run (a find made by a random run is kept as an @example)
@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow], derandomize=True)
@given(st.lists(op, min_size=1, max_size=25))
# identical vectors under two ids that share their 64-bit tie key: the hits must still come back in id order (client.in_id_order)
@example([("upsert", [0, 3, 2], 1), ("upsert", [0, 1, 3], 1), ("search", 0, None, 1)])
def test_store_and_oracle_agree_after_every_operation(ops):
    async def run():
        store = B200VectorStore(dimensions=DIM, _device_factory=FakeDevice)
        ora = OracleManager(DIM)
        await store.connect(); await store.create_collections(); ora.create_collections()
        coll = store._get(CODE)
        coll.COMPACT_MIN_FREE = 2                            # compaction whenever a quarter of the shard is free
        ids = [str(uuid.UUID(int=1000 + i)) for i in range(40)]
        for o in ops:
            if o[0] == "upsert":
                slots, salt = list(dict.fromkeys(o[1])), o[2]
                rng = np.random.default_rng(salt)
                vecs = rng.standard_normal((len(slots), DIM)).tolist()
                pls = [_payload(s, salt) for s in slots]
                await store.upsert(CODE, [ids[s] for s in slots], vecs, pls)
                ora.upsert(CODE, [ids[s] for s in slots], vecs, pls)
            elif o[0] == "delete_file":
                await store.delete(CODE, {"file_path": o[1]})
                ora.delete(CODE, {"file_path": o[1]})
            elif o[0] == "cleanup_project":
                flt = NS(must=[NS(key="project_name", match=NS(value=o[1]))])
                await store.client.delete(collection_name=CODE, points_selector=NS(filter=flt))
                ora.delete(CODE, {"project_name": o[1]})
            elif o[0] == "search":
                q = np.random.default_rng(o[1]).standard_normal(DIM).tolist()
                flt = {None: None, "file": {"file_path": FILES[o[1] % 5]}, "project": {"project_name": PROJECTS[o[1] % 2]},
                       "both": {"file_path": FILES[o[1] % 5], "project_name": PROJECTS[o[1] % 2]}, "lang": {"language": "go"}}[o[2]]
                got, exp = await store.search(CODE, q, o[3], flt), ora.search(CODE, q, o[3], flt)
                assert [h["id"] for h in got] == [h["id"] for h in exp]
                assert [h["payload"] for h in got] == [h["payload"] for h in exp]
                assert all(abs(a["score"] - b["score"]) < 1e-12 for a, b in zip(got, exp))
            else:
                got, exp = await store.search(CODE, None, 50, {"file_path": o[1]}), ora.search(CODE, None, 50, {"file_path": o[1]})
                assert [h["id"] for h in got] == [h["id"] for h in exp]
            # invariants of the host half
            assert (await store.get_collection_info(CODE)).points_count == ora.points_count(CODE)
            live = [r for r, pid in enumerate(coll.ids) if pid is not None]
            assert len(live) == ora.points_count(CODE) and len(coll.ids) == coll.dev.rows == len(coll.payloads)
            assert sorted(coll.free_rows) == [r for r, pid in enumerate(coll.ids) if pid is None]
            assert all(coll.id_to_row[coll.ids[r]] == r for r in live) and len(coll.id_to_row) == len(live)
        await store.close()
    asyncio.run(run())
