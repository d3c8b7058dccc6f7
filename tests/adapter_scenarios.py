"""Adapter scenarios shared by the CPU run (FakeDevice: host logic only) and the GPU run (real DeviceCollection).

They read like the reference's own tests of this seam: call shapes from ``tests/test_embeddings.py:553-565,620-625``
(keyword arguments ``collection=, query_vector=, limit=, filters=``; result dicts ``{"id","score","payload"}``) and the
live-database scenario of ``tests/test_database.py:76-124`` (create both collections; upsert ``[0.1]*1536`` then search
the same vector with ``limit=1``; delete by ``file_path``).  Results are compared with the CPU oracle's QdrantManager
restatement (``oracle.qdrant_local.OracleManager``) on the same inputs.
"""
from __future__ import annotations

import uuid

import numpy as np

import lvs_synth as synth
from code_rag_b200.client import B200VectorStore, CollectionName
from code_rag_b200.errors import VectorStoreError
from oracle.qdrant_local import OracleManager

CODE = CollectionName.CODE_CHUNKS.value
SUMM = CollectionName.SUMMARIES.value


def _store(factory, **kw):
    """`factory` is a device class (single-GPU adapter over it) or carries `make_store` (tests/test_sharded_store_cpu.py: the
    sharded adapter on rank 0 of a gloo job) - the scenarios are the same either way."""
    make = getattr(factory, "make_store", None)
    return make(**kw) if make is not None else B200VectorStore(_device_factory=factory, **kw)


def _same_hits(got, exp, rel=1e-5, what=""):
    assert [h["id"] for h in got] == [h["id"] for h in exp], f"{what}: ids differ\n got {[h['id'] for h in got]}\n exp {[h['id'] for h in exp]}"
    for g, e in zip(got, exp):
        assert abs(g["score"] - e["score"]) <= rel * max(abs(e["score"]), 1e-30) + 1e-300, what
        assert g["payload"] == e["payload"], what


async def scenario_test_database(factory):
    """reference tests/test_database.py:62-124 against the new backend."""
    manager = _store(factory, dimensions=1536)
    await manager.connect()
    assert await manager.health_check() is True
    await manager.create_collections()
    await manager.create_collections()          # idempotent (client.py:72-91)
    names = [c.name for c in (await manager.client.get_collections()).collections]
    assert "code_chunks" in names and "summaries" in names
    test_id = str(uuid.uuid4())
    test_vector = [0.1] * 1536
    test_payload = {"file_path": "/test/file.py", "entity_type": "function", "entity_name": "test_func",
                    "language": "python", "content": "def test_func(): pass", "start_line": 1, "end_line": 1}
    await manager.upsert(collection="code_chunks", ids=[test_id], vectors=[test_vector], payloads=[test_payload])
    results = await manager.search(collection="code_chunks", query_vector=test_vector, limit=1)
    assert len(results) >= 1
    assert results[0]["payload"]["entity_name"] == "test_func"
    assert results[0]["id"] == test_id
    assert abs(results[0]["score"] - 1.0) < 1e-6
    assert (await manager.get_collection_info("code_chunks")).points_count == 1
    await manager.delete(collection="code_chunks", filters={"file_path": "/test/file.py"})
    assert await manager.search(collection="code_chunks", query_vector=test_vector, limit=1) == []
    assert (await manager.get_collection_info("code_chunks")).points_count == 0
    await manager.close()
    assert await manager.health_check() is False
    try:
        manager.client
        raise AssertionError("client must raise before connect()")
    except VectorStoreError as e:
        assert "Client not connected" in str(e)


async def scenario_parity_with_oracle(factory, n=3000, dim=256):
    """Random uuid4 ids, CodeChunk payloads, every filter shape the reference issues; ids, scores, payloads vs oracle."""
    x, q = synth.unixcoder_like(n, dim, seed=1234, n_queries=6)
    pl = synth.payloads(n, seed=7)
    ids = synth.random_uuids(n, seed=9)
    store = _store(factory, dimensions=dim)
    ora = OracleManager(dim)
    await store.connect()
    await store.create_collections()
    ora.create_collections()
    # index file by file like VectorIndexer.index_file (embeddings/indexer.py:46-94): batches of a few chunks
    step = 257
    for s in range(0, n, step):
        sl = slice(s, min(n, s + step))
        vecs = x[sl].astype(np.float64).tolist()
        await store.upsert(collection=CODE, ids=ids[sl], vectors=vecs, payloads=pl[sl])
        ora.upsert(CODE, ids[sl], vecs, pl[sl])
    assert (await store.get_collection_info(CODE)).points_count == n == ora.points_count(CODE)
    some_file = pl[11]["file_path"]
    filter_sets = [None, {"language": "python"}, {"language": "typescript", "entity_type": "class", "project_name": "proj1"},
                   {"project_name": "proj0"}, {"file_path": some_file}, {"entity_type": "method"}, {"project_name": "nope"},
                   {"entity_name": pl[5]["entity_name"], "file_path": pl[5]["file_path"]}]
    for qi in range(len(q)):
        qv = q[qi].astype(np.float64).tolist()
        for flt in filter_sets:
            for limit in (1, 5, 10, 20):
                if qi > 1 and limit != 10:
                    continue
                got = await store.search(collection=CODE, query_vector=qv, limit=limit, filters=flt)
                exp = ora.search(CODE, qv, limit=limit, filters=flt)
                _same_hits(got, exp, what=f"q{qi} filters={flt} limit={limit}")
    # query_vector=None: filter-only lookup used by ContextBuilder (query/context/builder.py:111-119)
    got = await store.search(collection=CODE, query_vector=None, limit=1,
                             filters={"entity_name": pl[5]["entity_name"], "file_path": pl[5]["file_path"]})
    exp = ora.search(CODE, None, limit=1, filters={"entity_name": pl[5]["entity_name"], "file_path": pl[5]["file_path"]})
    _same_hits(got, exp, what="filter-only")
    got = await store.search(collection=CODE, query_vector=None, limit=7, filters={"project_name": "proj2"})
    exp = ora.search(CODE, None, limit=7, filters={"project_name": "proj2"})
    _same_hits(got, exp, what="filter-only scroll order")
    # incremental indexing (client.py:178-202)
    assert await store.file_needs_update(CODE, some_file, pl[11]["content_hash"]) is False
    assert await store.file_needs_update(CODE, some_file, "other-hash") is True
    assert await store.file_needs_update(CODE, "missing.py", "h") is True
    # re-index one file: delete by file_path then upsert new chunks with new ids (indexer.py:61-86)
    await store.delete(collection=CODE, filters={"file_path": some_file})
    ora.delete(CODE, {"file_path": some_file})
    assert (await store.get_collection_info(CODE)).points_count == ora.points_count(CODE)
    x2, _ = synth.unixcoder_like(5, dim, seed=99)
    ids2 = synth.random_uuids(5, seed=100)
    pl2 = [dict(pl[11], entity_name=f"new_{i}", content_hash="h2") for i in range(5)]
    await store.upsert(collection=CODE, ids=ids2, vectors=x2.astype(np.float64).tolist(), payloads=pl2)
    ora.upsert(CODE, ids2, x2.astype(np.float64).tolist(), pl2)
    # overwrite an existing id with a new vector and payload (Qdrant upsert semantics)
    await store.upsert(collection=CODE, ids=[ids[3]], vectors=[x2[0].astype(np.float64).tolist()], payloads=[dict(pl[3], language="go")])
    ora.upsert(CODE, [ids[3]], [x2[0].astype(np.float64).tolist()], [dict(pl[3], language="go")])
    for qi in range(3):
        qv = q[qi].astype(np.float64).tolist()
        for flt in (None, {"file_path": some_file}, {"language": "go"}):
            _same_hits(await store.search(collection=CODE, query_vector=qv, limit=10, filters=flt),
                       ora.search(CODE, qv, limit=10, filters=flt), what=f"after reindex q{qi} {flt}")
    # batched additive API == consecutive single searches
    qb = q[3:6].astype(np.float64)
    got_b = await store.search_batch(collection=CODE, query_vectors=qb.tolist(), limit=5)
    for i in range(3):
        _same_hits(got_b[i], ora.search(CODE, qb[i].tolist(), limit=5), what=f"batch {i}")
    # concurrent awaits, as QueryEngine does with asyncio.gather (query/engine.py:142-146): calls are serialised per collection
    import asyncio
    qv = q[0].astype(np.float64).tolist()
    many = await asyncio.gather(*[store.search(collection=CODE, query_vector=qv, limit=10) for _ in range(6)])
    exp = [ora.search(CODE, qv, limit=10) for _ in range(6)]
    for g in many:
        assert [h["id"] for h in g] == [h["id"] for h in exp[0]]
        assert max(abs(a["score"] - b["score"]) for a, b in zip(g, exp[0])) < 1e-9
    # summaries collection is independent
    assert await store.search(collection=SUMM, query_vector=q[0].astype(np.float64).tolist(), limit=3) == []
    # clear_collections resets both (client.py:212-221)
    await store.clear_collections()
    assert (await store.get_collection_info(CODE)).points_count == 0
    await store.close()


async def scenario_errors(factory):
    store = _store(factory, dimensions=8)
    for coro in (store.search(collection=CODE, query_vector=[0.0] * 8), store.upsert(CODE, [], [], []),
                 store.delete(CODE, {"a": 1}), store.create_collections(), store.get_collection_info(CODE)):
        try:
            await coro
            raise AssertionError("expected VectorStoreError before connect()")
        except VectorStoreError:
            pass
    assert await store.file_needs_update(CODE, "f", "h") is True      # never raises (client.py:200-202)
    await store.connect()
    await store.create_collections()
    bad = [("not-a-uuid", [0.1] * 8), (str(uuid.uuid4()), [0.1] * 7), (str(uuid.uuid4()), [float("nan")] * 8)]
    for pid, vec in bad:
        try:
            await store.upsert(collection=CODE, ids=[pid], vectors=[vec], payloads=[{}])
            raise AssertionError(f"expected VectorStoreError for {pid!r}")
        except VectorStoreError as e:
            assert e.cause is not None and "Failed to upsert vectors to code_chunks" in str(e)
    for kw in ({"collection": "nope", "query_vector": [0.1] * 8}, {"collection": CODE, "query_vector": [0.1] * 5},
               {"collection": CODE, "query_vector": [0.1] * 8, "filters": {"language": None}}):
        try:
            await store.search(**kw)
            raise AssertionError(f"expected VectorStoreError for {kw}")
        except VectorStoreError as e:
            assert "Failed to search" in str(e)
    await store.close()


async def scenario_client_shim(factory):
    """projects/cleanup.py:38-73: count + delete with a MatchText (substring) filter through manager.client."""
    from types import SimpleNamespace as NS
    store = _store(factory, dimensions=16)
    await store.connect()
    await store.create_collections()
    x, _ = synth.unit_rows(40, 16, seed=1)
    pl = [{"file_path": f"/repos/{'alpha' if i % 2 else 'beta'}/f{i}.py", "entity_type": "function", "project_name": "p"} for i in range(40)]
    await store.upsert(CODE, synth.random_uuids(40, 3), x.astype(np.float64).tolist(), pl)
    flt = NS(must=[NS(key="file_path", match=NS(text="/repos/alpha/"))])
    assert (await store.client.count(collection_name=CODE, count_filter=flt, exact=True)).count == 20
    await store.client.delete(collection_name=CODE, points_selector=NS(filter=flt))
    assert (await store.client.count(collection_name=CODE, count_filter=flt)).count == 0
    assert (await store.get_collection_info(CODE)).points_count == 20
    flt2 = NS(must=[NS(key="project_name", match=NS(value="p"))])
    assert (await store.client.count(collection_name=CODE, count_filter=flt2)).count == 20
    # scroll (what QdrantManager.file_needs_update issues, client.py:182-190): id order, pages chained by next_offset
    seen, offset = [], None
    while True:
        recs, offset = await store.client.scroll(collection_name=CODE, scroll_filter=flt2, limit=7, offset=offset, with_payload=True,
                                                 with_vectors=False)
        seen += [(r.id, r.payload["file_path"]) for r in recs]
        if offset is None:
            break
    everything = await store.search(collection=CODE, query_vector=None, limit=100, filters={"project_name": "p"})
    assert len(seen) == 20 and seen == [(h["id"], h["payload"]["file_path"]) for h in everything]
    one, nxt = await store.client.scroll(collection_name=CODE, scroll_filter=NS(must=[NS(key="file_path", match=NS(value="/repos/beta/f4.py"))]), limit=1)
    assert len(one) == 1 and nxt is None and one[0].payload["file_path"] == "/repos/beta/f4.py"
    # query_points (what QdrantManager.search issues, client.py:142-148)
    res = await store.client.query_points(collection_name=CODE, query=x[4].astype(np.float64).tolist(), limit=3, query_filter=flt2, with_payload=True)
    via = await store.search(collection=CODE, query_vector=x[4].astype(np.float64).tolist(), limit=3, filters={"project_name": "p"})
    tol = 1e-4 if str(getattr(store, "_storage", "f32")).startswith("bf") else 1e-6      # a bf16 shard stores the rounded vector
    assert [(p.id, p.payload) for p in res.points] == [(h["id"], h["payload"]) for h in via] and abs(res.points[0].score - 1.0) < tol
    await store.close()


async def scenario_reindex_churn(factory, n_files=40, chunks=6, dim=48, rounds=12):
    """VectorIndexer.index_file (embeddings/indexer.py:61-86) over and over: delete a file's chunks, insert new ones under fresh
    uuid4 ids.  The shard must not grow (deleted rows are reused) and must keep answering like the oracle."""
    import random
    rng = random.Random(5)
    store = _store(factory, dimensions=dim)
    ora = OracleManager(dim)
    await store.connect()
    await store.create_collections()
    ora.create_collections()
    seed = [0]

    async def index_file(f: int):
        seed[0] += 1
        x, _ = synth.unixcoder_like(chunks, dim, seed=1000 + seed[0])
        ids = synth.random_uuids(chunks, seed=2000 + seed[0])
        pl = [{"file_path": f"src/f{f}.py", "entity_type": "function", "entity_name": f"fn_{f}_{i}_{seed[0]}", "language": "python",
               "content": "x" * (20 + i), "start_line": i, "end_line": i + 2, "content_hash": f"h{seed[0]}", "project_name": "p"}
              for i in range(chunks)]
        await store.delete(collection=CODE, filters={"file_path": f"src/f{f}.py"})
        ora.delete(CODE, {"file_path": f"src/f{f}.py"})
        vecs = x.astype(np.float64).tolist()
        await store.upsert(collection=CODE, ids=ids, vectors=vecs, payloads=pl)
        ora.upsert(CODE, ids, vecs, pl)
        return ids
    last_ids = {}
    for f in range(n_files):
        last_ids[f] = await index_file(f)
    coll = store._get(CODE)
    rows_after_first_pass = coll.dev.rows
    assert rows_after_first_pass == n_files * chunks
    _, q = synth.unixcoder_like(1, dim, seed=77, n_queries=4)
    for r in range(rounds):
        for f in rng.sample(range(n_files), 10):
            last_ids[f] = await index_file(f)
        qv = q[r % 4].astype(np.float64).tolist()
        for flt in (None, {"file_path": f"src/f{rng.randrange(n_files)}.py"}):
            _same_hits(await store.search(collection=CODE, query_vector=qv, limit=8, filters=flt),
                       ora.search(CODE, qv, limit=8, filters=flt), what=f"churn round {r} {flt}")
    assert coll.dev.rows == rows_after_first_pass, "re-indexing must reuse the rows of deleted points"
    assert (await store.get_collection_info(CODE)).points_count == n_files * chunks == ora.points_count(CODE)
    # a deleted id can come back (new row), and the filter-only lookup still orders by id
    victim = last_ids[3][0]
    await store.delete(collection=CODE, filters={"file_path": "src/f3.py"}); ora.delete(CODE, {"file_path": "src/f3.py"})
    x, _ = synth.unixcoder_like(1, dim, seed=4040)
    p = {"file_path": "src/f3.py", "entity_type": "class", "entity_name": "Back", "language": "python", "content": "c", "start_line": 1,
         "end_line": 2, "content_hash": "hb", "project_name": "p"}
    await store.upsert(collection=CODE, ids=[victim], vectors=x.astype(np.float64).tolist(), payloads=[p])
    ora.upsert(CODE, [victim], x.astype(np.float64).tolist(), [p])
    _same_hits(await store.search(collection=CODE, query_vector=None, limit=50, filters={"project_name": "p"}),
               ora.search(CODE, None, limit=50, filters={"project_name": "p"}), what="scroll after churn")
    await store.close()


async def scenario_mass_delete_compacts(factory, n=2400, dim=48):
    """projects/cleanup.py:38-73 removes a whole project: the shard is compacted (live rows of the tail move into the holes, the
    rest is truncated) and keeps answering like the oracle - with filters, after further upserts, and for ids that come back."""
    from types import SimpleNamespace as NS
    x, q = synth.unixcoder_like(n, dim, seed=606, n_queries=6)
    pl = synth.payloads(n, seed=607)
    for i, p in enumerate(pl):
        p["project_name"] = ("alpha", "beta", "gamma")[i % 3]
    ids = synth.random_uuids(n, seed=608)
    store = _store(factory, dimensions=dim)
    ora = OracleManager(dim)
    await store.connect(); await store.create_collections(); ora.create_collections()
    vecs = x.astype(np.float64).tolist()
    await store.upsert(collection=CODE, ids=ids[:2000], vectors=vecs[:2000], payloads=pl[:2000])
    ora.upsert(CODE, ids[:2000], vecs[:2000], pl[:2000])
    coll = store._get(CODE)
    coll.COMPACT_MIN_FREE = 100                        # small shard: let the quarter rule decide
    _same_hits(await store.search(collection=CODE, query_vector=q[0].tolist(), limit=8), ora.search(CODE, q[0].tolist(), limit=8), what="before")
    # the reference's cleanup goes through manager.client.delete with a models.Filter (duck-typed here)
    flt = NS(must=[NS(key="project_name", match=NS(value="beta"))])
    await store.client.delete(collection_name=CODE, points_selector=NS(filter=flt))
    ora.delete(CODE, {"project_name": "beta"})
    n_left = ora.points_count(CODE)
    assert coll.dev.rows == n_left == len(coll.ids), "a third of the shard was deleted: it must have been compacted"
    assert not coll.free_rows and all(i is not None for i in coll.ids)
    for qi in range(1, 4):
        for f in (None, {"project_name": "gamma"}, {"project_name": "beta"}, {"language": pl[1]["language"]}):
            _same_hits(await store.search(collection=CODE, query_vector=q[qi].tolist(), limit=8, filters=f),
                       ora.search(CODE, q[qi].tolist(), limit=8, filters=f), what=f"after compaction q{qi} {f}")
    # new points, a deleted id that comes back, an overwrite of a moved point
    await store.upsert(collection=CODE, ids=ids[2000:], vectors=vecs[2000:], payloads=pl[2000:]); ora.upsert(CODE, ids[2000:], vecs[2000:], pl[2000:])
    back = next(i for i in range(2000) if pl[i]["project_name"] == "beta")
    moved = max(range(2000), key=lambda i: coll.id_to_row.get(str(__import__("uuid").UUID(ids[i])), -1) if pl[i]["project_name"] != "beta" else -1)
    for i in (back, moved):
        await store.upsert(collection=CODE, ids=[ids[i]], vectors=[vecs[(i + 7) % n]], payloads=[dict(pl[i], language="go")])
        ora.upsert(CODE, [ids[i]], [vecs[(i + 7) % n]], [dict(pl[i], language="go")])
    for qi in range(4, 6):
        for f in (None, {"language": "go"}, {"project_name": "alpha"}):
            _same_hits(await store.search(collection=CODE, query_vector=q[qi].tolist(), limit=8, filters=f),
                       ora.search(CODE, q[qi].tolist(), limit=8, filters=f), what=f"after upserts q{qi} {f}")
    assert (await store.get_collection_info(CODE)).points_count == ora.points_count(CODE)
    await store.close()


async def scenario_random_ops(factory, seed, storage="f32", steps=160, rel=1e-9):
    """Random operation sequence: upserts of new and of existing ids, deletes by file, project clean-ups through .client, searches
    with every filter shape, filter-only lookups, with frequent compaction; ids, payloads and scores must follow the oracle step
    by step (the replay state included)."""
    import random
    from types import SimpleNamespace as NS

    dim, n_ids = 32, 120
    files = [f"src/f{i}.py" for i in range(7)]
    projects = ["p0", "p1", "p2"]
    rng = random.Random(seed)

    def bf(v):
        if storage != "bf16":
            return v
        u = np.asarray(v, dtype=np.float32).view(np.uint32).astype(np.uint64)
        return ((((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32)).view(np.float32).astype(np.float64)

    store = _store(factory, dimensions=dim, storage=storage)
    ora = OracleManager(dim)
    await store.connect(); await store.create_collections(); ora.create_collections()
    coll = store._get("code_chunks")
    shards = getattr(coll, "shards", [coll])            # the multi-GPU adapter keeps one bookkeeping object per shard
    for sh in shards:
        sh.COMPACT_MIN_FREE = 4
    ids = [str(uuid.UUID(int=5000 + i)) for i in range(n_ids)]
    for step in range(steps):
        r = rng.random()
        if r < 0.35 or step < 3:
            slots = rng.sample(range(n_ids), rng.randint(1, 25))
            salt = rng.randrange(10_000)
            vecs = bf(np.random.default_rng(salt).standard_normal((len(slots), dim))).tolist()
            pls = [{"file_path": files[(s + salt) % 7], "project_name": projects[(s * 5 + salt) % 3], "language": ("python", "go")[salt % 2],
                    "entity_type": "function", "entity_name": f"fn{s}", "content": "x" * (s + 1), "start_line": s, "end_line": s + 1,
                    "content_hash": f"h{salt % 4}"} for s in slots]
            await store.upsert("code_chunks", [ids[s] for s in slots], vecs, pls)
            ora.upsert("code_chunks", [ids[s] for s in slots], vecs, pls)
        elif r < 0.45:
            f = rng.choice(files)
            await store.delete("code_chunks", {"file_path": f}); ora.delete("code_chunks", {"file_path": f})
        elif r < 0.50:
            p = rng.choice(projects)
            flt = NS(must=[NS(key="project_name", match=NS(value=p))])
            await store.client.delete(collection_name="code_chunks", points_selector=NS(filter=flt))
            ora.delete("code_chunks", {"project_name": p})
        elif r < 0.92:
            q = np.random.default_rng(step * 31 + seed).standard_normal(dim).tolist()
            flt = rng.choice([None, None, {"file_path": rng.choice(files)}, {"project_name": rng.choice(projects)},
                              {"file_path": rng.choice(files), "language": "go"}, {"language": "rust"}])
            k = rng.choice([1, 5, 10, 20])
            _same_hits(await store.search("code_chunks", q, k, flt), ora.search("code_chunks", q, k, flt), rel=rel, what=f"step {step} {flt}")
        else:
            f = rng.choice(files)
            got, exp = await store.search("code_chunks", None, 40, {"file_path": f}), ora.search("code_chunks", None, 40, {"file_path": f})
            assert [h["id"] for h in got] == [h["id"] for h in exp], f"step {step}"
        assert (await store.get_collection_info("code_chunks")).points_count == ora.points_count("code_chunks"), f"step {step}"
    assert sum(len(sh.ids) for sh in shards) <= n_ids
    if len(shards) == 1:
        assert coll.dev.rows == len(coll.ids)
    await store.close()




async def scenario_exact_ties_follow_the_id(factory, dim=16):
    """Identical vectors under ids that share their leading 64 bits (uuid.UUID(int=small)): the device can only order such hits
    by row, the adapter must return them in id order (BASELINE.json: score desc, id asc) - also when the run of equal hits
    straddles `limit`, with a filter, after deletes, and next to ordinary uuid4 ids.  Checked against the rule itself (the
    oracle's BLAS product gives identical rows position-dependent scores); needs a device with position-independent scores
    (helpers.ExactTieDevice on CPU, the GPU)."""
    rng = np.random.default_rng(5)
    v = rng.standard_normal((3, dim))
    store = _store(factory, dimensions=dim)
    await store.connect(); await store.create_collections()
    crafted = [str(uuid.UUID(int=1000 + i)) for i in range(12)]
    order = [7, 2, 11, 0, 5, 9, 1, 3, 10, 4, 8, 6]                       # rows are assigned in this order, ids must win
    ids = [crafted[i] for i in order] + synth.random_uuids(6, seed=2)
    which = [0] * 8 + [1] * 4 + [2] * 3 + [0] * 3
    pl = [{"file_path": f"f{i % 2}.py", "entity_name": f"e{i}"} for i in range(18)]
    for lo, hi in ((0, 5), (5, 18)):
        await store.upsert(CODE, ids[lo:hi], [v[w].tolist() for w in which[lo:hi]], pl[lo:hi])
    live = set(range(18))

    def expected(q, limit, flt):
        cos = [float(np.dot(v[w], q) / np.linalg.norm(v[w]) / np.linalg.norm(q)) for w in which]
        pts = [i for i in sorted(live) if flt is None or all(pl[i].get(k) == val for k, val in flt.items())]
        return [str(uuid.UUID(ids[i])) for i in sorted(pts, key=lambda i: (-round(cos[i], 9), str(uuid.UUID(ids[i]))))[:limit]]

    async def check(what):
        for q in (v[0], v[1], v[0] + v[1]):
            for limit in (1, 3, 5, 8, 11, 18):
                for flt in (None, {"file_path": "f1.py"}):
                    got = await store.search(CODE, q.tolist(), limit, flt)
                    assert [h["id"] for h in got] == expected(q, limit, flt), f"{what}: limit={limit} {flt}"
                    assert all(a["score"] >= b["score"] for a, b in zip(got, got[1:]))
    await check("ties")
    await store.delete(CODE, {"entity_name": "e3"}); live.discard(3)
    await store.delete(CODE, {"entity_name": "e16"}); live.discard(16)
    await check("ties after deletes")
    got = await store.search_batch(CODE, [v[0].tolist(), v[1].tolist()], limit=4)
    assert [[h["id"] for h in g] for g in got] == [expected(v[0], 4, None), expected(v[1], 4, None)]
    await store.close()


async def scenario_edge_cases(factory, dim=24):
    """Empty and ragged inputs, limits at and past their bounds, empty collections - against the oracle's QdrantManager where it
    defines the answer, against the reference's documented behaviour otherwise."""
    from code_rag_b200 import _native as N
    x, q = synth.unixcoder_like(40, dim, seed=71, n_queries=2)
    vecs = x.astype(np.float64).tolist()
    ids = synth.random_uuids(40, seed=72)
    pl = [{"file_path": f"f{i % 4}.py", "entity_name": f"e{i}", "language": "python"} for i in range(40)]
    store, ora = _store(factory, dimensions=dim), OracleManager(dim)
    await store.connect(); await store.create_collections(); ora.create_collections()
    qv = q[0].astype(np.float64).tolist()
    # empty collection: every read answers, nothing raises
    assert await store.search(collection=CODE, query_vector=qv, limit=5) == []
    assert await store.search(collection=CODE, query_vector=None, limit=5, filters={"file_path": "f0.py"}) == []
    assert await store.search_batch(collection=CODE, query_vectors=[qv, qv], limit=3) == [[], []]
    assert (await store.get_collection_info(CODE)).points_count == 0
    await store.delete(collection=CODE, filters={"file_path": "f0.py"})
    assert await store.file_needs_update(CODE, "f0.py", "h") is True
    # empty and ragged upserts: the reference zips ids, vectors and payloads (client.py:123-126)
    await store.upsert(collection=CODE, ids=[], vectors=[], payloads=[])
    await store.upsert(collection=CODE, ids=ids[:10], vectors=vecs[:7], payloads=pl[:9]); ora.upsert(CODE, ids[:7], vecs[:7], pl[:7])
    assert (await store.get_collection_info(CODE)).points_count == 7
    await store.upsert(collection=CODE, ids=ids[7:], vectors=vecs[7:], payloads=pl[7:]); ora.upsert(CODE, ids[7:], vecs[7:], pl[7:])
    # limit: 0 and negative give nothing, more than there is gives everything, the largest supported, one past it
    assert await store.search(collection=CODE, query_vector=qv, limit=0) == []
    assert await store.search(collection=CODE, query_vector=qv, limit=-3) == []
    for limit in (40, 41, 200, N.MAX_K):
        _same_hits(await store.search(collection=CODE, query_vector=qv, limit=limit), ora.search(CODE, qv, limit=limit), what=f"limit {limit}")
    _same_hits(await store.search(collection=CODE, query_vector=qv, limit=N.MAX_K, filters={"file_path": "f1.py"}),
               ora.search(CODE, qv, limit=N.MAX_K, filters={"file_path": "f1.py"}), what="filtered, limit past the matches")
    try:
        await store.search(collection=CODE, query_vector=qv, limit=N.MAX_K + 1)
        raise AssertionError("limit past MAX_K must raise")
    except VectorStoreError as e:
        assert "exceeds the largest supported top-k" in str(e.cause)
    # a batch with one query, an empty batch
    got = await store.search_batch(collection=CODE, query_vectors=[qv], limit=4)
    _same_hits(got[0], ora.search(CODE, qv, limit=4), what="batch of one")
    assert await store.search_batch(collection=CODE, query_vectors=np.zeros((0, dim)).tolist() or np.zeros((0, dim)), limit=4) == []
    # a zero query vector: local mode divides by EPSILON-guarded norm -> all scores 0.0, order by id
    zero = await store.search(collection=CODE, query_vector=[0.0] * dim, limit=5)
    assert len(zero) == 5 and all(h["score"] == 0.0 for h in zero)
    # delete everything, then the collection behaves as empty again and can be refilled
    for f in range(4):
        await store.delete(collection=CODE, filters={"file_path": f"f{f}.py"}); ora.delete(CODE, {"file_path": f"f{f}.py"})
    assert (await store.get_collection_info(CODE)).points_count == 0 and await store.search(collection=CODE, query_vector=qv, limit=5) == []
    await store.upsert(collection=CODE, ids=ids[:5], vectors=vecs[:5], payloads=pl[:5]); ora.upsert(CODE, ids[:5], vecs[:5], pl[:5])
    _same_hits(await store.search(collection=CODE, query_vector=qv, limit=10), ora.search(CODE, qv, limit=10), what="refilled")
    await store.close()
