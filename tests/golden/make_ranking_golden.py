#!/usr/bin/env python
"""Generates tests/golden/ranking_golden.json.gz by running the REFERENCE's own ranking code
(/root/reference/src/lattice/query/ranking/*, query/reranker.py) on seeded synthetic inputs.

Run in the build container only (the reference is not present on the GPU box):
    python tests/golden/make_ranking_golden.py
The reference package's __init__ chains import qdrant_client / tree_sitter / neo4j, which are not installed here, so
the namespace packages are stubbed exactly as SURVEY.md Appendix A describes; the ranking modules themselves are the
unmodified reference files.
"""
from __future__ import annotations

import json
import random
import sys
import types
from pathlib import Path

SRC = "/root/reference/src"


def _ns(name, path=None):
    m = types.ModuleType(name)
    if path:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, SRC)
    _ns("lattice", SRC + "/lattice"); _ns("lattice.query", SRC + "/lattice/query"); _ns("lattice.graph", SRC + "/lattice/graph")
    n = _ns("neo4j"); n.AsyncGraphDatabase = n.AsyncDriver = n.AsyncSession = object
    e = _ns("neo4j.exceptions"); e.ServiceUnavailable = e.Neo4jError = e.AuthError = Exception
    from lattice.query.graph_reasoning import GraphContext, GraphNode
    from lattice.query.query_planner import ExtractedEntity, QueryIntent, QueryPlan
    from lattice.query.ranking import HybridRanker
    from lattice.query.reranker import ResultReranker, normalize_scores
    return dict(GraphContext=GraphContext, GraphNode=GraphNode, ExtractedEntity=ExtractedEntity, QueryIntent=QueryIntent,
                QueryPlan=QueryPlan, HybridRanker=HybridRanker, ResultReranker=ResultReranker, normalize_scores=normalize_scores)


NAMES = ["save", "load", "User", "create_user", "bulk_save_all", "parse", "Parser", "handle_request", "connect", "close",
         "Session", "query", "rank_results", "tokenize", "embed", "index_file", "Config", "get_settings", "main", "run"]
FILES = [f"src/pkg{i % 4}/mod{i}.py" for i in range(12)]
TEXT = [None, "", "short", "x" * 60, "y" * 150, "z" * 2500, "w" * 3100]


def make_case(rng: random.Random, intents, case_id: int) -> dict:
    intent = intents[case_id % len(intents)]
    n_ent = rng.choice([0, 1, 1, 2, 3])
    entities = [rng.choice(NAMES + ["sav", "", "user"]) for _ in range(n_ent)]

    def node(depth=None):
        nm = rng.choice(NAMES)
        f = rng.choice(FILES)
        md = {} if depth is None else {"depth": depth}
        return {"node_type": rng.choice(["Function", "Method", "Class"]), "name": nm, "qualified_name": rng.choice([f"m.{nm}", f"pkg.{nm}", ""]),
                "file_path": f, "signature": rng.choice([None, f"def {nm}()"]), "docstring": rng.choice([None, "", "doc"]),
                "summary": rng.choice([None, "sum"]), "start_line": rng.choice([1, 3, 10, 20, None]), "end_line": 99, "metadata": md}

    graph = {
        "primary_entities": [node() for _ in range(rng.randint(0, 4))],
        "callers": [node(rng.choice([None, 0, 1, 2, 3, 5])) for _ in range(rng.randint(0, 12))],
        "callees": [node(rng.choice([None, 1, 2, 4])) for _ in range(rng.randint(0, 12))],
        "methods": [node() for _ in range(rng.randint(0, 8))],
        "parent_classes": [node() for _ in range(rng.randint(0, 3))],
        "child_classes": [node() for _ in range(rng.randint(0, 3))],
    }
    all_nodes = [n for v in graph.values() for n in v]
    vector = []
    for _ in range(rng.randint(0, 30)):
        if all_nodes and rng.random() < 0.35:          # share a key with a graph node -> hybrid merge
            g = rng.choice(all_nodes)
            nm, f, sl, gid = g["name"], g["file_path"], g["start_line"], g["qualified_name"] or None
        else:
            nm, f, sl, gid = rng.choice(NAMES), rng.choice(FILES), rng.choice([1, 3, 10, 20, 44, None]), rng.choice([None, "m.x", f"m.{rng.choice(NAMES)}"])
        vector.append({"score": round(rng.uniform(-0.2, 1.0), 6), "file_path": f, "entity_type": "function", "entity_name": nm,
                       "language": "python", "content": rng.choice(TEXT), "start_line": sl, "end_line": 50, "graph_node_id": gid,
                       "summary": rng.choice([None, "vs"])})
    cent = {}
    for nm in rng.sample(NAMES, rng.randint(0, 8)):
        cent[rng.choice([f"m.{nm}", f"pkg.{nm}", nm])] = {"in_degree": 1, "out_degree": 2, "total_degree": rng.choice([0, 3, 12, 25, 49, 50, 80, 100])}
    return {"id": case_id, "intent": intent, "entities": entities, "graph": graph, "vector": vector, "centrality": cent}


def run_case(ref, case: dict) -> dict:
    GN, GC = ref["GraphNode"], ref["GraphContext"]
    intent = ref["QueryIntent"](case["intent"])
    plan = ref["QueryPlan"](original_query="q", primary_intent=intent, sub_queries=[],
                            entities=[ref["ExtractedEntity"](name=e, entity_type="function") for e in case["entities"]], relationships=[])
    mk = lambda d: GN(node_type=d["node_type"], name=d["name"], qualified_name=d["qualified_name"], file_path=d["file_path"],
                      signature=d["signature"], docstring=d["docstring"], summary=d["summary"], start_line=d["start_line"],
                      end_line=d["end_line"], metadata=dict(d["metadata"]))
    g = case["graph"]
    ctx = GC(primary_entities=[mk(n) for n in g["primary_entities"]], callers=[mk(n) for n in g["callers"]],
             callees=[mk(n) for n in g["callees"]], parent_classes=[mk(n) for n in g["parent_classes"]],
             child_classes=[mk(n) for n in g["child_classes"]], methods=[mk(n) for n in g["methods"]], containing_class=None,
             file_context=[], dependencies=[], dependents=[], call_chains=[], inheritance_chains=[])
    ranked = ref["HybridRanker"]().rank_results(plan, ctx, [dict(v) for v in case["vector"]], dict(case["centrality"]))
    hybrid = [{"key": r.get_key(), "final_score": r.final_score, "source": r.source, "signal_scores": dict(r.signal_scores),
               "content": r.content, "summary": r.summary, "signature": r.signature, "docstring": r.docstring,
               "relationship_path": r.relationship_path, "depth_from_query": r.depth_from_query} for r in ranked]
    # the older fusion (query/reranker.py): graph rows are plain dicts
    graph_rows = [{"file_path": n["file_path"], "name": n["name"], "type": n["node_type"], "summary": n["summary"],
                   "start_line": n["start_line"], "end_line": n["end_line"], "qualified_name": n["qualified_name"]}
                  for k in ("primary_entities", "callers", "callees") for n in g[k]]
    rr = ref["ResultReranker"]()
    fused = rr.fuse_results(graph_rows, [dict(v) for v in case["vector"]])
    dedup = rr.deduplicate(fused)
    norm = ref["normalize_scores"](dedup)
    ser = lambda rs: [{"key": r.get_key(), "score": r.score, "source": r.source, "content": r.content, "summary": r.summary} for r in rs]
    return {"hybrid": hybrid, "fused": ser(fused), "dedup": ser(dedup), "normalized": ser(norm), "graph_rows": graph_rows}


def main():
    ref = import_reference()
    intents = [i.value for i in ref["QueryIntent"]]
    rng = random.Random(4567)
    cases = []
    for cid in range(85):
        c = make_case(rng, intents, cid)
        c["expected"] = run_case(ref, c)
        cases.append(c)
    import gzip
    out = Path(__file__).with_name("ranking_golden.json.gz")
    blob = json.dumps({"generator": "tests/golden/make_ranking_golden.py", "reference": "lattice.query.ranking + lattice.query.reranker (unmodified)",
                       "seed": 4567, "cases": cases}, indent=None, separators=(",", ":")).encode()
    with gzip.GzipFile(filename=str(out), mode="wb", mtime=0) as f:   # mtime=0: byte-reproducible
        f.write(blob)
    print(f"wrote {out} ({out.stat().st_size} bytes, {len(cases)} cases)")


if __name__ == "__main__":
    main()
