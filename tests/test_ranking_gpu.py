"""K3 (fused hybrid ranking) through the reference-shaped API vs golden vectors made by the reference's own code."""
import gzip
import json
from pathlib import Path
from types import SimpleNamespace as NS

import pytest

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden" / "ranking_golden.json.gz"


@pytest.fixture(scope="module")
def golden(native_lib):
    from code_rag_b200 import _native
    _native.init(0)
    return json.loads(gzip.decompress(GOLDEN.read_bytes()))["cases"]


def _inputs(case):
    node = lambda d: NS(node_type=d["node_type"], name=d["name"], qualified_name=d["qualified_name"], file_path=d["file_path"],
                        signature=d["signature"], docstring=d["docstring"], summary=d["summary"], start_line=d["start_line"],
                        end_line=d["end_line"], metadata=dict(d["metadata"]))
    g = case["graph"]
    ctx = NS(**{k: [node(n) for n in g[k]] for k in ("primary_entities", "callers", "callees", "methods", "parent_classes", "child_classes")})
    plan = NS(primary_intent=NS(value=case["intent"]), entities=[NS(name=e) for e in case["entities"]])
    return plan, ctx, [dict(v) for v in case["vector"]], dict(case["centrality"])


def _check(case, got):
    exp = case["expected"]["hybrid"]
    assert [r.get_key() for r in got] == [e["key"] for e in exp], f"case {case['id']}: order differs"
    for r, e in zip(got, exp):
        assert r.final_score == e["final_score"], (case["id"], r.get_key(), r.final_score, e["final_score"])
        assert r.source == e["source"] and r.signal_scores == e["signal_scores"], (case["id"], r.get_key(), r.signal_scores, e["signal_scores"])
        for f in ("content", "summary", "signature", "docstring", "relationship_path", "depth_from_query"):
            assert getattr(r, f) == e[f], (case["id"], r.get_key(), f)


def test_hybrid_ranker_single_calls_bit_exact(golden):
    from code_rag_b200.ranking import HybridRanker
    ranker = HybridRanker()
    for case in golden[:20]:
        _check(case, ranker.rank_results(*_inputs(case)))


def test_hybrid_ranker_batch_bit_exact(golden):
    from code_rag_b200.ranking import HybridRanker
    ranker = HybridRanker()
    out = ranker.rank_batch([_inputs(c) for c in golden])
    assert len(out) == len(golden)
    for case, got in zip(golden, out):
        _check(case, got)
    assert ranker.last_device_ms > 0


def test_reranker_and_normalize_bit_exact(golden):
    from code_rag_b200.ranking import ResultReranker, normalize_scores
    rr = ResultReranker()
    for case in golden[:40]:
        exp = case["expected"]
        fused = rr.fuse_results(exp["graph_rows"], [dict(v) for v in case["vector"]])
        tup = lambda rs: [(r.get_key(), r.score, r.source, r.content, r.summary) for r in rs]
        want = lambda rs: [(e["key"], e["score"], e["source"], e["content"], e["summary"]) for e in rs]
        assert tup(fused) == want(exp["fused"]), case["id"]
        dedup = rr.deduplicate(fused)
        assert tup(dedup) == want(exp["dedup"]), case["id"]
        assert tup(normalize_scores(dedup)) == want(exp["normalized"]), case["id"]
    assert normalize_scores([]) == []


def test_c4_shaped_batch(golden):
    """configs[3] shape: 64 queries x (100 vector hits + 64 graph candidates, 25 % sharing a key), all 17 intents."""
    import random
    from code_rag_b200.ranking import HybridRanker
    from oracle import ranking as R
    rng = random.Random(4567)
    intents = sorted({c["intent"] for c in golden})
    cases = []
    for q in range(64):
        vec = [{"score": rng.uniform(0.2, 0.9), "file_path": f"f{rng.randrange(40)}.py", "entity_type": "function", "entity_name": f"e{i}",
                "content": "c" * rng.choice([10, 60, 150, 2500]), "start_line": i, "end_line": i + 5, "graph_node_id": f"m.e{i}"}
               for i in range(100)]
        nodes = []
        for i in range(64):
            if rng.random() < 0.25:
                v = rng.choice(vec)
                nm, fp, sl = v["entity_name"], v["file_path"], v["start_line"]
            else:
                nm, fp, sl = f"g{i}", f"f{rng.randrange(40)}.py", 1000 + i
            nodes.append({"node_type": "Function", "name": nm, "qualified_name": f"m.{nm}", "file_path": fp, "signature": rng.choice([None, "s"]),
                          "docstring": rng.choice([None, "d"]), "summary": rng.choice([None, "x"]), "start_line": sl, "end_line": sl + 1,
                          "metadata": {"depth": rng.choice([1, 2, 3])}})
        graph = {"primary_entities": nodes[:4], "callers": nodes[4:24], "callees": nodes[24:44], "methods": nodes[44:54],
                 "parent_classes": nodes[54:59], "child_classes": nodes[59:]}
        cent = {f"m.e{i}": {"total_degree": rng.choice([0, 5, 12, 30, 80])} for i in rng.sample(range(100), 10)}
        cases.append({"id": q, "intent": intents[q % len(intents)], "entities": [f"e{rng.randrange(100)}", "g1"], "graph": graph,
                      "vector": vec, "centrality": cent})
    ranker = HybridRanker()
    out = ranker.rank_batch([_inputs(c) for c in cases])
    for c, got in zip(cases, out):
        exp = R.hybrid_rank(c)
        assert [r.get_key() for r in got] == [e["key"] for e in exp]
        assert [r.final_score for r in got] == [e["final_score"] for e in exp]
        assert [r.source for r in got] == [e["source"] for e in exp]
    print(f"K3: 64 queries x 164 candidates ranked in {ranker.last_device_ms * 1e3:.1f} us of device time")
