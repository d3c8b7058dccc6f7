"""Fused search -> rank (lvs_search_rank) vs the two-step route and the ranking oracle.

The fused call must return exactly what the reference's flow returns (query/engine.py:315-346 then 176-181):
VectorSearcher-shaped hits of ``search`` fed to ``HybridRanker.rank_results``.  Checked three ways: against our own two-step
``search`` + ``rank_batch`` (bit-exact K3 pinned on goldens made by the reference's code), against the CPU ranking oracle
fed with the hits of the CPU search oracle, and on the edge cases (filters, empty graph context, no entities, no hits).
"""
import asyncio
import random
from types import SimpleNamespace as NS

import numpy as np
import pytest

import lvs_synth as synth

pytestmark = pytest.mark.gpu

DIM = 96
N_ROWS = 6000
INTENTS = ["find_callers", "find_similar", "search_functionality", "explain_architecture", "locate_entity", "find_hierarchy"]


def _payloads(rng: random.Random, n: int):
    out = []
    for i in range(n):
        name = rng.choice(["Parse", "parse_file", "load", "UserService", "save", "handle_request", "Größe", "λ_fn"]) + (f"_{i % 97}" if i % 3 else "")
        p = {"file_path": f"src/m{rng.randrange(60)}.py", "entity_type": rng.choice(["function", "method", "class"]),
             "entity_name": name, "language": "python", "start_line": rng.randrange(1, 400), "end_line": 500,
             "content": "x" * rng.choice([0, 10, 60, 150, 1999, 2500, 3500]), "project_name": rng.choice(["a", "b"]),
             "graph_node_id": rng.choice([None, f"pkg.{name}"])}
        if p["content"] == "":
            p["content"] = None
        out.append(p)
    return out


def _case(rng: random.Random, payloads, q: int, qvec):
    nodes = []
    for i in range(rng.choice([0, 7, 30])):
        if rng.random() < 0.4:
            p = rng.choice(payloads)      # same key as a stored row: merges when that row is a hit
            nm, fp, sl = p["entity_name"], p["file_path"], p["start_line"]
        else:
            nm, fp, sl = f"g{i}", f"src/m{rng.randrange(60)}.py", 1000 + i
        nodes.append(NS(node_type="Function", name=nm, qualified_name=rng.choice([None, f"pkg.{nm}"]), file_path=fp,
                        signature=rng.choice([None, "s"]), docstring=rng.choice([None, "d"]), summary=rng.choice([None, "x"]),
                        start_line=sl, end_line=sl + 1, metadata={"depth": rng.choice([0, 1, 2, 3, 5])}))
    cut = [0, len(nodes) // 6, len(nodes) // 2, (2 * len(nodes)) // 3, (5 * len(nodes)) // 6, (11 * len(nodes)) // 12, len(nodes)]
    names = ("primary_entities", "callers", "callees", "methods", "parent_classes", "child_classes")
    ctx = NS(**{nm: nodes[cut[i]:cut[i + 1]] for i, nm in enumerate(names)})
    ents = rng.choice([[], ["parse"], ["Parse_5", "save"], ["größe"], [""], ["userservice", "load_12", "zzz"]])
    plan = NS(primary_intent=NS(value=INTENTS[q % len(INTENTS)]), entities=[NS(name=e) for e in ents])
    cent = {}
    for p in rng.sample(payloads, 8):
        cent[p["graph_node_id"] or p["entity_name"]] = {"total_degree": rng.choice([-3, 0, 5, 12, 49, 50, 80])}
    cent["not.in.the.index"] = {"total_degree": 7}
    return plan, ctx, qvec, cent


@pytest.fixture(scope="module")
def store(native_lib):
    """Two identical stores: local-mode scores depend on how many searches ran since a row was written,
    so the fused route runs on one and the two-step route on the other, search for search."""
    from code_rag_b200.client import B200VectorStore
    rng = random.Random(77)
    x, q = synth.unit_rows(N_ROWS, DIM, seed=5, n_queries=24)
    payloads = _payloads(rng, N_ROWS)
    ids = [str(__import__("uuid").UUID(int=i + 1)) for i in range(N_ROWS)]
    rewritten = {i: dict(payloads[i], entity_name="Rewritten", start_line=7, content="y" * 120) for i in (3, 500, 4242)}

    async def setup(st):
        await st.connect()
        await st.create_collections()
        for lo in range(0, N_ROWS, 1500):
            await st.upsert("code_chunks", ids[lo:lo + 1500], x[lo:lo + 1500].astype(np.float64).tolist(), payloads[lo:lo + 1500])
        # overwrite a few rows (new payloads, same ids) and delete one file's chunks: the attribute columns must follow
        await st.upsert("code_chunks", [ids[i] for i in rewritten], x[list(rewritten)].astype(np.float64).tolist(), list(rewritten.values()))
        await st.delete("code_chunks", {"file_path": "src/m7.py"})
    st_a = B200VectorStore(dimensions=DIM, storage="f32", rank_attrs=True)
    st_b = B200VectorStore(dimensions=DIM, storage="f32", rank_attrs=True)
    asyncio.run(setup(st_a))
    asyncio.run(setup(st_b))
    for i, p in rewritten.items():
        payloads[i] = p
    yield st_a, st_b, payloads, q.astype(np.float64), x
    asyncio.run(st_a.close())
    asyncio.run(st_b.close())


def _same(a, b):
    assert [r.get_key() for r in a] == [r.get_key() for r in b]
    for ra, rb in zip(a, b):
        assert ra.final_score == rb.final_score and ra.source == rb.source and ra.signal_scores == rb.signal_scores, ra.get_key()
        for f in ("file_path", "entity_name", "entity_type", "qualified_name", "content", "summary", "signature", "docstring",
                  "start_line", "end_line", "graph_node_id", "relationship_path", "depth_from_query"):
            assert getattr(ra, f) == getattr(rb, f), (ra.get_key(), f)


@pytest.mark.parametrize("filters", [None, {"project_name": "a"}, {"file_path": "src/m3.py"}, {"project_name": "nobody"}])
def test_fused_equals_two_step(store, filters):
    from code_rag_b200.ranking import HybridRanker
    st, st_b, payloads, q, _ = store
    rng = random.Random(11)
    items = [_case(rng, payloads, i, q[i]) for i in range(len(q))]
    coll = st_b._get("code_chunks")
    ranker = HybridRanker()
    fused = asyncio.run(st.search_and_rank("code_chunks", items, limit=20, filters=filters, ranker=ranker))
    assert ranker.last_search_ms > 0 and ranker.last_device_ms > 0
    hits = asyncio.run(st_b.search_batch("code_chunks", [it[2] for it in items], limit=20, filters=filters))
    two_step = ranker.rank_batch([(it[0], it[1], [coll.vector_result_from_hit(h) for h in hits[i]], it[3]) for i, it in enumerate(items)])
    assert len(fused) == len(two_step) == len(items)
    n_hybrid = 0
    for a, b in zip(fused, two_step):
        _same(a, b)
        n_hybrid += sum(r.source == "hybrid" for r in a)
    if filters is None:
        assert n_hybrid > 0, "the cases are meant to produce graph/vector merges"


def test_fused_equals_cpu_oracles(store):
    """End to end against the CPU restatements: search oracle (qdrant local mode) -> ranking oracle (reference ranker)."""
    from oracle import ranking as R
    from oracle.qdrant_local import OracleCollection
    st, _, payloads, q, x = store
    coll = st._get("code_chunks")
    # the oracle replays the store's history: same rows, same overwrites and deletions, same number of earlier searches
    ora = OracleCollection(DIM)
    ora.upsert_rows_f32(0, x.astype(np.float32), [None] * len(x))
    dead = [r for r, p in enumerate(payloads) if p.get("file_path") == "src/m7.py"]     # rows: ids were upserted in order
    ora.deleted[dead] = True
    for _ in range(coll.dev.search_counter):
        ora.search_topk_rows(q[0], 1)           # a local-mode search re-normalises the matrix in place, whatever the query
    rng = random.Random(23)
    items = [_case(rng, payloads, i, q[i]) for i in range(8)]
    fused = asyncio.run(st.search_and_rank("code_chunks", items, limit=15))
    for i, (plan, ctx, qv, cent) in enumerate(items):
        rows, scores = ora.search_topk_rows(qv, 15)
        vec = [coll.vector_result(int(r), float(s)) for r, s in zip(rows, scores)]
        node = lambda n: {"node_type": n.node_type, "name": n.name, "qualified_name": n.qualified_name, "file_path": n.file_path,
                          "signature": n.signature, "docstring": n.docstring, "summary": n.summary, "start_line": n.start_line,
                          "end_line": n.end_line, "metadata": n.metadata}
        case = {"intent": plan.primary_intent.value, "entities": [e.name for e in plan.entities], "vector": vec, "centrality": cent,
                "graph": {k: [node(n) for n in getattr(ctx, k)] for k in
                          ("primary_entities", "callers", "callees", "methods", "parent_classes", "child_classes")}}
        exp = R.hybrid_rank(case)
        got = fused[i]
        assert [r.get_key() for r in got] == [e["key"] for e in exp]
        assert [r.source for r in got] == [e["source"] for e in exp]
        for r, e in zip(got, exp):
            assert abs(r.final_score - e["final_score"]) <= 1e-9 * max(1.0, abs(e["final_score"]))


def test_fused_needs_rank_attrs(native_lib):
    from code_rag_b200.client import B200VectorStore
    from code_rag_b200.errors import VectorStoreError
    st = B200VectorStore(dimensions=8)

    async def run():
        await st.connect()
        await st.create_collections()
        with pytest.raises(VectorStoreError):
            await st.search_and_rank("code_chunks", [(None, None, [0.0] * 8, None)])
        await st.close()
    asyncio.run(run())


def test_fused_on_summaries_collection(native_lib):
    """Summary hits carry no start_line and no content (query/vector_search.py:243-260): keys end in ':None', code_quality is 0."""
    from code_rag_b200.client import B200VectorStore
    from code_rag_b200.ranking import HybridRanker
    rng = random.Random(3)
    n, dim = 800, 64
    x, q = synth.unit_rows(n, dim, seed=8, n_queries=5)
    pl = [{"file_path": f"src/m{i % 30}.py", "entity_type": "file", "entity_name": f"m{i % 30}_{i}", "summary": rng.choice([None, "", "does things"]),
           "graph_node_id": None, "content_hash": "h"} for i in range(n)]
    ids = [str(__import__("uuid").UUID(int=i + 10)) for i in range(n)]

    async def run():
        a = B200VectorStore(dimensions=dim, rank_attrs=True)
        b = B200VectorStore(dimensions=dim, rank_attrs=True)
        for st in (a, b):
            await st.connect(); await st.create_collections()
            await st.upsert("summaries", ids, x.astype(np.float64).tolist(), pl)
        node = NS(node_type="File", name=pl[5]["entity_name"], qualified_name=None, file_path=pl[5]["file_path"], signature=None, docstring=None,
                  summary="s", start_line=None, end_line=None, metadata={})
        ctx = NS(primary_entities=[node], callers=[], callees=[], methods=[], parent_classes=[], child_classes=[])
        plan = NS(primary_intent=NS(value="explain_architecture"), entities=[NS(name="m5")])
        items = [(plan, ctx, q[i].astype(np.float64), {pl[5]["entity_name"]: {"total_degree": 30}}) for i in range(5)]
        ranker = HybridRanker()
        fused = await a.search_and_rank("summaries", items, limit=25, ranker=ranker)
        coll = b._get("summaries")
        hits = await b.search_batch("summaries", [it[2] for it in items], limit=25)
        two = ranker.rank_batch([(it[0], it[1], [coll.vector_result_from_hit(h) for h in hits[i]], it[3]) for i, it in enumerate(items)])
        for fa, fb in zip(fused, two):
            _same(fa, fb)
            assert all(r.get_key().endswith(":None") for r in fa)
        await a.close(); await b.close()
    asyncio.run(run())


def test_fused_with_summaries_extension(native_lib):
    """QueryEngine._execute_vector_search (query/engine.py:315-346): five intents also search `summaries` (limit // 2) and append
    those hits behind the code hits.  One fused call over both collections == the two-step route, search for search."""
    from code_rag_b200.client import B200VectorStore
    from code_rag_b200.ranking import SUMMARY_INTENTS, HybridRanker
    rng = random.Random(41)
    n, ns, dim = 3000, 600, 64
    x, q = synth.unit_rows(n + ns, dim, seed=12, n_queries=12)
    pl = _payloads(rng, n)
    spl = [{"file_path": pl[rng.randrange(n)]["file_path"], "entity_type": "file", "entity_name": rng.choice(["parse", "Loader", "m"]) + str(i),
            "summary": rng.choice([None, "does things"]), "project_name": rng.choice(["a", "b"]), "graph_node_id": None} for i in range(ns)]
    ids = [str(__import__("uuid").UUID(int=i + 1)) for i in range(n + ns)]
    intents = ["explain_architecture", "find_callers", "search_functionality", "find_similar", "explain_data_flow", "locate_entity"]

    async def run():
        a = B200VectorStore(dimensions=dim, rank_attrs=True)
        b = B200VectorStore(dimensions=dim, rank_attrs=True)
        for st in (a, b):
            await st.connect(); await st.create_collections()
            await st.upsert("summaries", ids[n:n + 200], x[n:n + 200].astype(np.float64).tolist(), spl[:200])   # interleaved on purpose:
            await st.upsert("code_chunks", ids[:n], x[:n].astype(np.float64).tolist(), pl)                     # one id space for both
            await st.upsert("summaries", ids[n + 200:], x[n + 200:].astype(np.float64).tolist(), spl[200:])
        items = []
        for i in range(12):
            plan, ctx, qv, cent = _case(rng, pl, i, q[i].astype(np.float64))
            plan.primary_intent = NS(value=intents[i % len(intents)])
            items.append((plan, ctx, qv, cent))
        assert any(it[0].primary_intent.value in SUMMARY_INTENTS for it in items)
        ranker = HybridRanker()
        for flt, sflt in ((None, None), ({"project_name": "a"}, {"project_name": "a"})):
            fused = await a.search_and_rank("code_chunks", items, limit=12, filters=flt, ranker=ranker, summaries=True, summaries_filters=sflt)
            code, summ = b._get("code_chunks"), b._get("summaries")
            hits = await b.search_batch("code_chunks", [it[2] for it in items], limit=12, filters=flt)
            vrs = [[code.vector_result_from_hit(h) for h in hits[i]] for i in range(12)]
            sel = [i for i, it in enumerate(items) if it[0].primary_intent.value in SUMMARY_INTENTS]
            hits2 = await b.search_batch("summaries", [items[i][2] for i in sel], limit=6, filters=sflt)
            for j, i in enumerate(sel):
                vrs[i].extend(summ.vector_result_from_hit(h) for h in hits2[j])
            two = ranker.rank_batch([(it[0], it[1], vrs[i], it[3]) for i, it in enumerate(items)])
            n_summary = 0
            for fa, fb in zip(fused, two):
                _same(fa, fb)
                n_summary += sum(r.get_key().endswith(":None") for r in fa)
            assert n_summary > 0, "summary hits should reach the ranked lists"
        await a.close(); await b.close()
    asyncio.run(run())


def test_fused_after_compaction(native_lib):
    """A mass delete compacts the shard (rows move): the per-row ranking attributes and the name ids must move with them."""
    from code_rag_b200.client import B200VectorStore
    from code_rag_b200.ranking import HybridRanker
    rng = random.Random(9)
    n, dim = 2100, 64
    x, q = synth.unit_rows(n, dim, seed=44, n_queries=8)
    pl = _payloads(rng, n)
    for i, p in enumerate(pl):
        p["project_name"] = ("a", "b", "c")[i % 3]
    ids = [str(__import__("uuid").UUID(int=i + 1)) for i in range(n)]

    async def run():
        stores = [B200VectorStore(dimensions=dim, rank_attrs=True) for _ in range(2)]
        for st in stores:
            await st.connect(); await st.create_collections()
            st._get("code_chunks").COMPACT_MIN_FREE = 100
            await st.upsert("code_chunks", ids, x.astype(np.float64).tolist(), pl)
            await st.delete("code_chunks", {"project_name": "b"})
            coll = st._get("code_chunks")
            assert coll.dev.rows == len(coll.ids) == n - n // 3 and not coll.free_rows
        a, b = stores
        items = [_case(rng, [p for p in pl if p["project_name"] != "b"], i, q[i].astype(np.float64)) for i in range(8)]
        ranker = HybridRanker()
        fused = await a.search_and_rank("code_chunks", items, limit=15, ranker=ranker)
        coll = b._get("code_chunks")
        hits = await b.search_batch("code_chunks", [it[2] for it in items], limit=15)
        assert all(h["payload"]["project_name"] != "b" for hs in hits for h in hs)
        two = ranker.rank_batch([(it[0], it[1], [coll.vector_result_from_hit(h) for h in hits[i]], it[3]) for i, it in enumerate(items)])
        for fa, fb in zip(fused, two):
            _same(fa, fb)
        for st in stores:
            await st.close()
    asyncio.run(run())


def test_fused_with_an_empty_summaries_collection(native_lib):
    """summaries=True while nothing has been indexed into `summaries` yet: same result as without the extension."""
    from code_rag_b200.client import B200VectorStore
    rng = random.Random(5)
    n, dim = 900, 64
    x, q = synth.unit_rows(n, dim, seed=3, n_queries=4)
    pl = _payloads(rng, n)
    ids = [str(__import__("uuid").UUID(int=i + 1)) for i in range(n)]

    async def run():
        a = B200VectorStore(dimensions=dim, rank_attrs=True)
        b = B200VectorStore(dimensions=dim, rank_attrs=True)
        for st in (a, b):
            await st.connect(); await st.create_collections()
            await st.upsert("code_chunks", ids, x.astype(np.float64).tolist(), pl)
        items = []
        for i in range(4):
            plan, ctx, qv, cent = _case(rng, pl, i, q[i].astype(np.float64))
            plan.primary_intent = NS(value="explain_architecture")
            items.append((plan, ctx, qv, cent))
        fa = await a.search_and_rank("code_chunks", items, limit=9, summaries=True)
        fb = await b.search_and_rank("code_chunks", items, limit=9, summaries=False)
        for ra, rb in zip(fa, fb):
            _same(ra, rb)
        await a.close(); await b.close()
    asyncio.run(run())
