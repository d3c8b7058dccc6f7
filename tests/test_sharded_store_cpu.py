"""The multi-GPU adapter's host logic on CPU: a gloo job (world size 2 and 3) in which rank 0 drives a ``ShardedB200VectorStore``
through the SAME scenarios as the single-GPU adapter (tests/adapter_scenarios.py: the reference's database test, parity with the
oracle's QdrantManager over every filter shape, the error convention, the ``.client`` shim) while the other ranks sit in
``ShardPlane.serve()``.  Each rank's shard is the oracle-backed FakeDevice (its position-independent form, helpers.ExactTieDevice); the exchange step is the real packed all-gather.
Placement (least-full shard), overwrite in place, per-shard row reuse and compaction are checked on top."""
import os
import sys
import traceback
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p_ in (str(ROOT), str(ROOT / "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

from code_rag_b200.sharded_store import least_full, split_row  # noqa: E402


def test_row_coding_and_placement():
    assert split_row((5 << 32) | 77) == (5, 77) and split_row(3) == (0, 3)
    assert least_full([4, 2, 2]) == 1 and least_full([0]) == 0 and least_full([3, 3, 3]) == 0


async def _sharded_specifics(make_store, world):
    """What only exists with several shards: balance, overwrite stays put, per-shard reuse and compaction, scroll order."""
    import lvs_synth as synth
    from adapter_scenarios import CODE, _same_hits
    from oracle.qdrant_local import OracleManager
    n, dim = 1800, 48
    x, q = synth.unixcoder_like(n, dim, seed=31, n_queries=4)
    pl = synth.payloads(n, seed=32)
    for i, p in enumerate(pl):
        p["project_name"] = ("alpha", "beta", "gamma")[i % 3]
    ids = synth.random_uuids(n, seed=33)
    vecs = x.astype(np.float64).tolist()
    store, ora = make_store(dimensions=dim), OracleManager(dim)
    await store.connect(); await store.create_collections(); ora.create_collections()
    for s in range(0, 1500, 211):
        sl = slice(s, min(1500, s + 211))
        await store.upsert(collection=CODE, ids=ids[sl], vectors=vecs[sl], payloads=pl[sl]); ora.upsert(CODE, ids[sl], vecs[sl], pl[sl])
    coll = store._get(CODE)
    info = await store.get_collection_info(CODE)
    assert info.points_count == 1500 and info.shards == world and sum(info.shard_points) == 1500
    assert max(info.shard_points) - min(info.shard_points) <= 1, info.shard_points
    # an overwrite stays in its shard and row
    where = coll.shard_of(store_id := str(__import__("uuid").UUID(ids[7])))
    row = coll.shards[where].id_to_row[store_id]
    await store.upsert(collection=CODE, ids=[ids[7]], vectors=[vecs[99]], payloads=[dict(pl[7], language="go")])
    ora.upsert(CODE, [ids[7]], [vecs[99]], [dict(pl[7], language="go")])
    assert coll.shard_of(store_id) == where and coll.shards[where].id_to_row[store_id] == row
    # delete a third everywhere: every shard compacts on its own
    for sh in coll.shards:
        sh.COMPACT_MIN_FREE = 50
    await store.delete(collection=CODE, filters={"project_name": "beta"}); ora.delete(CODE, {"project_name": "beta"})
    assert (await store.get_collection_info(CODE)).points_count == ora.points_count(CODE)
    assert all(not sh.free_rows and all(i is not None for i in sh.ids) for sh in coll.shards), "each shard must have been compacted"
    # rebalancing: with 3 shards the deleted project sat in ONE shard (points were placed round-robin); half of the gap moves from
    # the fullest to the emptiest shard, ids / payloads / filters keep working, scores stay within a float32 ulp of the oracle
    before = (await store.get_collection_info(CODE)).shard_points
    moved = await store.rebalance(CODE)
    pts = (await store.get_collection_info(CODE)).shard_points
    assert sum(pts) == sum(before) == ora.points_count(CODE) and max(pts) - min(pts) <= 1, (before, pts)
    assert (moved > 0) == (max(before) - min(before) > 1)
    assert all(not sh.free_rows and all(i is not None for i in sh.ids) for sh in coll.shards), "the source shard must have been compacted"
    for qi in range(2):
        for f in (None, {"project_name": "gamma"}, {"language": "go"}):
            _same_hits(await store.search(collection=CODE, query_vector=q[qi].tolist(), limit=9, filters=f),
                       ora.search(CODE, q[qi].tolist(), limit=9, filters=f), rel=1e-6, what=f"after rebalancing q{qi} {f}")
    # the emptier shards fill up first; a filter on a key that was never indexed (new column over all shards)
    expect = (await store.get_collection_info(CODE)).shard_points
    for _ in range(1500, n):
        expect[least_full(expect)] += 1
    await store.upsert(collection=CODE, ids=ids[1500:], vectors=vecs[1500:], payloads=pl[1500:]); ora.upsert(CODE, ids[1500:], vecs[1500:], pl[1500:])
    assert (await store.get_collection_info(CODE)).shard_points == expect
    for qi in range(4):
        for f in (None, {"project_name": "gamma"}, {"language": "go"}, {"entity_name": pl[1600]["entity_name"]}, {"project_name": "beta"}):
            _same_hits(await store.search(collection=CODE, query_vector=q[qi].tolist(), limit=9, filters=f),
                       ora.search(CODE, q[qi].tolist(), limit=9, filters=f), what=f"sharded q{qi} {f}")
    got = await store.search_batch(collection=CODE, query_vectors=q.tolist(), limit=6, filters={"project_name": "alpha"})
    for qi in range(4):
        _same_hits(got[qi], ora.search(CODE, q[qi].tolist(), limit=6, filters={"project_name": "alpha"}), what=f"sharded batch {qi}")
    # concurrent awaits (query/engine.py:142-146): one search at a time crosses the shards, the others queue - all answer correctly
    import asyncio
    many = await asyncio.gather(*[store.search(collection=CODE, query_vector=q[i % 4].tolist(), limit=5) for i in range(8)])
    for i, h in enumerate(many):
        _same_hits(h, ora.search(CODE, q[i % 4].tolist(), limit=5), rel=1e-6, what=f"gathered {i}")
    # the filter-only lookup orders by id across shards
    _same_hits(await store.search(collection=CODE, query_vector=None, limit=40, filters={"project_name": "gamma"}),
               ora.search(CODE, None, limit=40, filters={"project_name": "gamma"}), what="sharded scroll")
    # search_and_rank (two-step over shards): the ranker receives, per query, exactly what VectorSearcher would hand it - code hits
    # of the batched search, then limit // 2 summary hits for the intents QueryEngine extends (query/engine.py:331-344)
    from types import SimpleNamespace as NS
    from adapter_scenarios import SUMM
    from code_rag_b200.client import vector_result_from_payload as shape
    sx, _ = synth.unixcoder_like(30, dim, seed=35)
    spl = [{"file_path": f"src/s{i}.py", "entity_type": "file", "entity_name": f"s{i}", "summary": f"summary {i}", "graph_node_id": f"m.s{i}"}
           for i in range(30)]
    sids = synth.random_uuids(30, seed=36)
    await store.upsert(collection=SUMM, ids=sids, vectors=sx.astype(np.float64).tolist(), payloads=spl)
    ora.upsert(SUMM, sids, sx.astype(np.float64).tolist(), spl)
    seen = []
    ranker = NS(rank_batch=lambda items: seen.append(items) or [f"ranked{i}" for i in range(len(items))])
    intents = ["find_callers", "explain_architecture", "search_functionality", "find_similar"]
    items = [(NS(primary_intent=NS(value=intents[i]), entities=[]), f"ctx{i}", q[i].tolist(), {"c": i}) for i in range(4)]
    out = await store.search_and_rank(CODE, items, limit=7, filters={"project_name": "alpha"}, ranker=ranker, summaries=True)
    assert out == ["ranked0", "ranked1", "ranked2", "ranked3"] and len(seen) == 1
    for i, (plan, ctx, vr, cen) in enumerate(seen[0]):
        assert plan is items[i][0] and ctx == f"ctx{i}" and cen == {"c": i}
        exp = [shape(h["payload"], h["score"], "code") for h in ora.search(CODE, q[i].tolist(), limit=7, filters={"project_name": "alpha"})]
        if intents[i] in ("explain_architecture", "search_functionality"):
            exp += [shape(h["payload"], h["score"], "summary") for h in ora.search(SUMM, q[i].tolist(), limit=3)]
        assert [{k: v for k, v in d.items() if k != "score"} for d in vr] == [{k: v for k, v in d.items() if k != "score"} for d in exp], i
        assert np.allclose([d["score"] for d in vr], [d["score"] for d in exp], rtol=1e-5)
    # snapshot: every rank writes / reads its own shard, the restored store answers identically and keeps working
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        await store.save(d)
        before = await store.search(collection=CODE, query_vector=q[1].tolist(), limit=9, filters={"language": "go"})
        await store.delete(collection=CODE, filters={"project_name": "gamma"})            # diverge, then restore
        assert (await store.get_collection_info(CODE)).points_count < ora.points_count(CODE)
        await store.load(d)
    coll = store._get(CODE)
    assert (await store.get_collection_info(CODE)).points_count == ora.points_count(CODE)
    after = await store.search(collection=CODE, query_vector=q[1].tolist(), limit=9, filters={"language": "go"})
    assert [h["id"] for h in after] == [h["id"] for h in before] and [h["payload"] for h in after] == [h["payload"] for h in before]
    # a worker-side failure reaches the caller as VectorStoreError and the plane keeps working
    from code_rag_b200.errors import VectorStoreError
    store.plane.queue(world - 1, CODE, "no_such_write")
    try:
        await store.search(collection=CODE, query_vector=q[0].tolist(), limit=3)
        raise AssertionError("expected VectorStoreError")
    except VectorStoreError as e:
        assert f"rank {world - 1}" in str(e.cause)
    _same_hits(await store.search(collection=CODE, query_vector=q[0].tolist(), limit=3), ora.search(CODE, q[0].tolist(), limit=3), what="after failure")
    await store.close()


async def _sharded_upsert_tokens(make_store, world, spec, embed):
    """``upsert_tokens`` over the shards (SURVEY section 8f row 4 at N GPUs): token ids go to the rank that owns the point, which
    embeds them itself.  `embed(token_ids)` = the vectors the ranks' encoders are expected to produce (the oracle's)."""
    import lvs_synth as synth
    from adapter_scenarios import CODE
    from code_rag_b200.errors import VectorStoreError
    hidden = spec["random"]["hidden"]
    store = make_store(dimensions=hidden)
    await store.connect(); await store.create_collections()
    rng = np.random.default_rng(41)
    n, L = 48, 24
    tok = rng.integers(3, spec["random"]["vocab"], size=(n, L)).astype(np.int32)
    for i in range(1, n):
        tok[i, int(rng.integers(L // 3, L + 1)):] = 1                      # ragged: pad id 1
    ids = synth.random_uuids(n, seed=42)
    pl = [{"file_path": f"src/t{i % 5}.py", "entity_name": f"tok{i}", "project_name": ("alpha", "beta")[i % 2]} for i in range(n)]
    try:
        await store.upsert_tokens(CODE, ids[:4], tok[:4], pl[:4])
        raise AssertionError("upsert_tokens without an attached encoder must fail")
    except VectorStoreError:
        pass
    assert await store.attach_encoder(spec) == hidden
    await store.upsert_tokens(CODE, ids[:30], tok[:30], pl[:30])
    await store.upsert_tokens(CODE, ids[25:], tok[25:], pl[25:])                # five overwrites among them
    info = await store.get_collection_info(CODE)
    assert info.points_count == n and max(info.shard_points) - min(info.shard_points) <= 1, info.shard_points
    want = embed(tok)
    for i in (0, 7, 26, 29, 47):
        hits = await store.search(collection=CODE, query_vector=want[i].astype(np.float64).tolist(), limit=3)
        assert hits[0]["id"] == ids[i] and hits[0]["payload"]["entity_name"] == f"tok{i}" and abs(hits[0]["score"] - 1.0) < 2e-3, (i, hits[0])
    hits = await store.search(collection=CODE, query_vector=want[3].astype(np.float64).tolist(), limit=5, filters={"project_name": "beta"})
    assert hits and hits[0]["id"] == ids[3] and all(h["payload"]["project_name"] == "beta" for h in hits)
    # vectors and tokens may be mixed in one collection: the same point re-written from its vector stays where it is
    await store.upsert(collection=CODE, ids=[ids[7]], vectors=[want[9].astype(np.float64).tolist()], payloads=[dict(pl[7], entity_name="swapped")])
    hits = await store.search(collection=CODE, query_vector=want[9].astype(np.float64).tolist(), limit=2)
    assert {h["payload"]["entity_name"] for h in hits} == {"swapped", "tok9"}
    await store.attach_encoder(None)
    await store.close()


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    status = "ok"
    try:
        from helpers import ExactTieDevice, FakeEncoder, FakeShardSearcher, fake_embed
        from code_rag_b200.sharded_store import ShardedB200VectorStore, run

        async def main(plane):                               # rank 0 only; the other ranks serve inside run()
            import adapter_scenarios as S
            from types import SimpleNamespace as NS
            factory = NS(make_store=lambda **kw: ShardedB200VectorStore(plane=plane, **kw))
            await S.scenario_test_database(factory)
            await S.scenario_parity_with_oracle(factory, n=900, dim=48)
            await S.scenario_errors(factory)
            await S.scenario_client_shim(factory)
            await _sharded_specifics(factory.make_store, world)
            for seed in (1, 2):
                await S.scenario_random_ops(factory, seed)
            await S.scenario_exact_ties_follow_the_id(factory)
            await S.scenario_edge_cases(factory)
            spec = {"random": {"vocab": 500, "hidden": 32, "layers": 1, "intermediate": 64, "max_pos": 64, "seed": 3}}
            await _sharded_upsert_tokens(factory.make_store, world, spec, lambda t: fake_embed(t, 32))
            assert plane.polled_searches > 0 and plane._polled is None, "searches must take the event-loop path (submit + poll)"
            return "done"
        got = run(main, device_factory=ExactTieDevice, searcher_factory=FakeShardSearcher, encoder_factory=FakeEncoder)
        assert got == ("done" if rank == 0 else None)
    except BaseException:  # noqa: BLE001
        status = traceback.format_exc()
    Path(out_dir, f"rank{rank}.txt").write_text(status)


def _run(world, tmp_path):
    port = 31000 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        status = Path(tmp_path, f"rank{r}.txt").read_text()
        if status != "ok":
            raise AssertionError(f"rank {r}:\n{status}")


def test_sharded_store_world2(tmp_path):
    _run(2, tmp_path)


def test_sharded_store_world3(tmp_path):
    _run(3, tmp_path)
