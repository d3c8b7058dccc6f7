"""The ranking oracle (oracle/ranking.py) against golden vectors produced by the reference's OWN code
(tests/golden/make_ranking_golden.py) and the hand-checked known answers of SURVEY.md Appendix B."""
import gzip
import json
from pathlib import Path

import pytest

from oracle import ranking as R

GOLDEN = Path(__file__).parent / "golden" / "ranking_golden.json.gz"


@pytest.fixture(scope="module")
def golden():
    return json.loads(gzip.decompress(GOLDEN.read_bytes()))


def test_golden_covers_every_intent_and_merges(golden):
    cases = golden["cases"]
    assert len(cases) == 85 and len({c["intent"] for c in cases}) == 17
    assert sum(r["source"] == "hybrid" for c in cases for r in c["expected"]["hybrid"]) > 50


def test_hybrid_rank_equals_reference(golden):
    for c in golden["cases"]:
        got = R.hybrid_rank(c)
        exp = c["expected"]["hybrid"]
        assert [g["key"] for g in got] == [e["key"] for e in exp], f"case {c['id']}: order differs"
        for g, e in zip(got, exp):
            assert g["final_score"] == e["final_score"], (c["id"], g["key"], g["final_score"], e["final_score"])
            assert g["source"] == e["source"] and g["signal_scores"] == e["signal_scores"], (c["id"], g["key"])
            for f in ("content", "summary", "signature", "docstring", "relationship_path", "depth_from_query"):
                assert g[f] == e[f], (c["id"], g["key"], f)


def test_reranker_equals_reference(golden):
    for c in golden["cases"]:
        exp = c["expected"]
        fused = R.rerank_fuse(exp["graph_rows"], c["vector"])
        dedup = R.rerank_dedup(fused)
        norm = R.rerank_normalize(dedup)
        for got, want in ((fused, exp["fused"]), (dedup, exp["dedup"]), (norm, exp["normalized"])):
            assert [(g["key"], g["score"], g["source"], g["content"], g["summary"]) for g in got] == \
                   [(e["key"], e["score"], e["source"], e["content"], e["summary"]) for e in want], c["id"]


def _node(name, qn, fp, sl, **kw):
    return {"node_type": "Function", "name": name, "qualified_name": qn, "file_path": fp, "signature": kw.get("signature"),
            "docstring": kw.get("docstring"), "summary": kw.get("summary"), "start_line": sl, "end_line": sl + 5,
            "metadata": kw.get("metadata", {})}


def _case(intent, entities, graph=None, vector=(), centrality=None):
    g = {k: [] for k in ("primary_entities", "callers", "callees", "methods", "parent_classes", "child_classes")}
    g.update(graph or {})
    return {"intent": intent, "entities": list(entities), "graph": g, "vector": list(vector), "centrality": centrality or {}}


def test_survey_appendix_b_known_answers():
    # B1 vector-only
    hits = [{"score": 0.9, "file_path": "a.py", "entity_name": "f1", "start_line": 1, "content": "x" * 150, "graph_node_id": "m.f1"},
            {"score": 0.8, "file_path": "a.py", "entity_name": "f2", "start_line": 10, "content": "x" * 60, "graph_node_id": "m.f2"},
            {"score": 0.7, "file_path": "b.py", "entity_name": "f3", "start_line": 5, "content": "x" * 10, "graph_node_id": None},
            {"score": 0.6, "file_path": "c.py", "entity_name": "f4", "start_line": 7, "content": None, "graph_node_id": None}]
    out = R.hybrid_rank(_case("find_similar", [], vector=hits, centrality={"m.f1": {"total_degree": 25}, "m.f2": {"total_degree": 80}}))
    assert [(r["key"], r["final_score"]) for r in out] == [("a.py:f1:1", 0.9000000000000001), ("a.py:f2:10", 0.8900000000000001),
                                                          ("b.py:f3:5", 0.59), ("c.py:f4:7", 0.48)]
    # B2 graph-only
    g = {"primary_entities": [_node("save", "m.User.save", "u.py", 3, signature="def save()", docstring="d")],
         "callers": [_node("create_user", "m.create_user", "api.py", 9, summary="s", metadata={"depth": 1}),
                     _node("bulk_save_all", "m.bulk_save_all", "api.py", 30, metadata={"depth": 3}),
                     _node("deep", "m.deep", "x.py", 1, metadata={"depth": 5})],
         "methods": [_node("other", "m.User.other", "u.py", 20)]}
    out = R.hybrid_rank(_case("find_callers", ["save"], graph=g, centrality={"m.User.save": {"total_degree": 100}}))
    assert [(r["key"], r["final_score"]) for r in out] == [("u.py:save:3", 1.49), ("api.py:create_user:9", 0.9500000000000001),
                                                          ("u.py:other:20", 0.875), ("api.py:bulk_save_all:30", 0.75), ("x.py:deep:1", 0.36)]
    # B3 order-dependent triple merge
    g = {"primary_entities": [_node("f", "m.f", "a.py", 1)], "callees": [_node("f", "m.f", "a.py", 1, metadata={"depth": 2})]}
    out = R.hybrid_rank(_case("explain_architecture", ["zzz"], graph=g,
                              vector=[{"score": 0.5, "file_path": "a.py", "entity_name": "f", "start_line": 1, "content": "c" * 500}]))
    assert len(out) == 1 and out[0]["source"] == "hybrid" and out[0]["final_score"] == 0.5308875000000002
    assert out[0]["signal_scores"]["graph_match"] == 1.0 and out[0]["signal_scores"]["vector_similarity"] == 0.5
    assert out[0]["signal_scores"]["code_quality"] == 0.8 and out[0]["signal_scores"]["relationship_relevance"] == 1.0
    # B4 stable sort + per-file cap
    hits = [{"score": 0.5, "file_path": "same.py", "entity_name": f"e{i}", "start_line": i} for i in range(7)]
    assert [r["key"].split(":")[1] for r in R.hybrid_rank(_case("search_pattern", [], vector=hits))] == ["e0", "e1", "e2", "e3", "e4"]
    # B5 ResultReranker
    fused = R.rerank_fuse([{"file_path": "a.py", "name": "f", "start_line": 1}, {"file_path": "z.py", "name": "g", "start_line": 2}],
                          [{"score": 0.9, "file_path": "a.py", "entity_name": "f", "start_line": 1},
                           {"score": 0.5, "file_path": "b.py", "entity_name": "h", "start_line": 3}])
    assert [(r["key"], r["source"], r["score"]) for r in fused] == [("a.py:f:1", "hybrid", 0.9400000000000001), ("z.py:g:2", "graph", 0.4),
                                                                    ("b.py:h:3", "vector", 0.3)]
    assert [r["score"] for r in R.rerank_normalize(fused)] == [1.0, 0.15625000000000003, 0.0]
