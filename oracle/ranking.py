"""CPU oracle for the hybrid score fusion (regime 3).  TEST INFRASTRUCTURE ONLY.

Restates, over plain records, what the reference computes in
``lattice.query.ranking`` (``HybridRanker.rank_results`` ranker.py:18-54, ``ResultScorer`` scorer.py:9-126,
``RankingConfig`` models.py:59-91) and in the older ``lattice.query.reranker`` (``fuse_results`` reranker.py:84-120,
``deduplicate`` :122-145, ``normalize_scores`` :29-70).  PINNED: ``tests/golden/ranking_golden.json.gz`` was produced by
running the reference's own, unmodified modules (``tests/golden/make_ranking_golden.py``) and
``tests/test_ranking_oracle.py`` checks this restatement against it value for value (float64 ``==``), together with the
hand-checked known-answer vectors of SURVEY.md Appendix B.

A *candidate* is a flat record (the numeric view the CUDA kernel consumes, built by ``flatten_case``):
    kind          0 primary, 1 caller, 2 callee, 3 method / parent_class / child_class, 4 vector hit
    key           f"{file_path}:{entity_name}:{start_line}"            (models.py:55-56)
    file          file_path
    depth         ``metadata.get("depth", 1)`` for callers / callees    (ranker.py:88, 102)
    entity_match  1.0 exact (lower-cased) name in the query entities, 0.5 substring, else 0.0   (scorer.py:31-35, 91-96)
    degree        total_degree of ``qualified_name or entity_name`` in the centrality dict, or None (scorer.py:48-53)
    has_*         truthiness of summary / docstring / signature / content                   (scorer.py:56-65)
    content_len   len(content) or None                                                        (scorer.py:106-115)
    vscore        the vector hit's score
"""
from __future__ import annotations

from typing import Any

GRAPH_SIGNALS = ("graph_match", "query_entity_match", "relationship_relevance", "centrality", "context_richness")
VECTOR_SIGNALS = ("vector_similarity", "query_entity_match", "centrality", "code_quality")

# models.py:6-13, 59-91
DEFAULT_WEIGHTS = {"graph_weight": 0.5, "vector_weight": 0.5, "centrality_weight": 0.2, "context_weight": 0.1}
ENTITY_MATCH_BONUS = 0.3
RELATIONSHIP_BONUS = 0.15
MAX_PER_FILE = 5
MAX_TOTAL = 50
INTENT_WEIGHTS = {
    "find_callers": (0.8, 0.2), "find_callees": (0.8, 0.2), "find_call_chain": (0.9, 0.1), "find_hierarchy": (0.85, 0.15),
    "find_usages": (0.7, 0.3), "find_dependencies": (0.75, 0.25), "locate_entity": (0.6, 0.4), "locate_file": (0.5, 0.5),
    "explain_implementation": (0.5, 0.5), "explain_relationship": (0.6, 0.4), "explain_data_flow": (0.65, 0.35),
    "find_similar": (0.2, 0.8), "search_functionality": (0.3, 0.7), "search_pattern": (0.25, 0.75),
}


def weights_for_intent(intent: str) -> dict[str, float]:
    """ranker.py:56-68: defaults, overridden per intent (3 of the 17 intents have no override)."""
    w = dict(DEFAULT_WEIGHTS)
    if intent in INTENT_WEIGHTS:
        w["graph_weight"], w["vector_weight"] = INTENT_WEIGHTS[intent]
    return w


def entity_match(name: str, query_entities: set[str]) -> float:
    low = name.lower()
    if low in query_entities:
        return 1.0
    if any(qe in low for qe in query_entities):
        return 0.5
    return 0.0


def flatten_case(case: dict) -> list[dict[str, Any]]:
    """Insertion order of ranker.py:70-169: primary, callers, callees, methods, parent_classes, child_classes, vector."""
    qents = {e.lower() for e in case["entities"]}
    cent = case["centrality"]
    out: list[dict[str, Any]] = []
    order = (("primary_entities", 0, None), ("callers", 1, "caller"), ("callees", 2, "callee"), ("methods", 3, "method"),
             ("parent_classes", 3, "parent_class"), ("child_classes", 3, "child_class"))
    for field, kind, rel in order:
        for n in case["graph"][field]:
            ck = n["qualified_name"] or n["name"]
            depth = None
            if kind in (1, 2):
                depth = n["metadata"].get("depth", 1) if n["metadata"] else 1
            out.append({
                "kind": kind, "key": f"{n['file_path']}:{n['name']}:{n['start_line']}", "file": n["file_path"],
                "depth": depth, "entity_match": entity_match(n["name"], qents),
                "degree": cent[ck].get("total_degree", 0) if ck in cent else None,
                "has_summary": bool(n["summary"]), "has_docstring": bool(n["docstring"]), "has_signature": bool(n["signature"]),
                "has_content": False, "content_len": None, "vscore": None, "relationship_path": rel,
                "content": None, "summary": n["summary"], "signature": n["signature"], "docstring": n["docstring"],
            })
    for v in case["vector"]:
        name = v.get("entity_name", "")
        ck = v.get("graph_node_id") or name
        content = v.get("content")
        out.append({
            "kind": 4, "key": f"{v.get('file_path', '')}:{name}:{v.get('start_line')}", "file": v.get("file_path", ""),
            "depth": None, "entity_match": entity_match(name, qents),
            "degree": cent[ck].get("total_degree", 0) if ck in cent else None,
            "has_summary": bool(v.get("summary")), "has_docstring": False, "has_signature": False,
            "has_content": bool(content), "content_len": len(content) if content else None, "vscore": v.get("score", 0.0),
            "relationship_path": None, "content": content, "summary": v.get("summary"), "signature": None, "docstring": None,
        })
    return out


def score_candidate(c: dict, w: dict[str, float]) -> tuple[float, dict[str, float]]:
    cen = 0.0
    if c["degree"] is not None:
        cen = min(1.0, c["degree"] / 50)
    if c["kind"] == 4:                                           # scorer.py:79-126
        q = 0.0
        if c["content_len"] is not None:
            n = c["content_len"]
            q = 0.8 if 100 < n < 2000 else 0.5 if 50 < n < 3000 else 0.3
        sig = {"vector_similarity": c["vscore"], "query_entity_match": c["entity_match"], "centrality": cen, "code_quality": q}
        final = c["vscore"] * w["vector_weight"] + c["entity_match"] * ENTITY_MATCH_BONUS + cen * w["centrality_weight"] + q * 0.1
        return final, sig
    base = 1.0                                                   # scorer.py:9-77
    if c["kind"] in (1, 2):
        depth = c["depth"] or 1
        base = max(0.3, 1.0 - (depth - 1) * 0.2)
    rel = {0: 1.0, 1: 0.8, 2: 0.7}.get(c["kind"], 0.5)
    ctx = 0.0
    if c["has_summary"]:
        ctx += 0.3
    if c["has_docstring"]:
        ctx += 0.2
    if c["has_signature"]:
        ctx += 0.2
    if c["has_content"]:
        ctx += 0.3
    sig = {"graph_match": base, "query_entity_match": c["entity_match"], "relationship_relevance": rel, "centrality": cen,
           "context_richness": ctx}
    final = (base * w["graph_weight"] + c["entity_match"] * ENTITY_MATCH_BONUS + rel * RELATIONSHIP_BONUS
             + cen * w["centrality_weight"] + ctx * w["context_weight"])
    return final, sig


def hybrid_rank(case: dict) -> list[dict[str, Any]]:
    """HybridRanker.rank_results (ranker.py:18-54) on a golden-format case; returns the golden-format result list."""
    w = weights_for_intent(case["intent"])
    merged: dict[str, dict[str, Any]] = {}
    for c in flatten_case(case):
        final, sig = score_candidate(c, w)
        cur = merged.get(c["key"])
        if cur is None:                                          # ranker.py:171-178
            merged[c["key"]] = {"key": c["key"], "file": c["file"], "final_score": final, "signal_scores": sig,
                                "source": "vector" if c["kind"] == 4 else "graph", "content": c["content"], "summary": c["summary"],
                                "signature": c["signature"], "docstring": c["docstring"], "relationship_path": c["relationship_path"],
                                "depth_from_query": c["depth"]}
            continue
        combined = (cur["final_score"] + final) / 2              # ranker.py:179-202 (order-dependent fold)
        combined *= 1.1
        for f in ("content", "summary", "signature", "docstring"):
            if not cur[f] and c[f]:
                cur[f] = c[f]
        for s, v in sig.items():
            cur["signal_scores"][s] = max(cur["signal_scores"][s], v) if s in cur["signal_scores"] else v
        cur["final_score"] = combined
        cur["source"] = "hybrid"
    ranked = sorted(merged.values(), key=lambda r: r["final_score"], reverse=True)   # stable
    out, per_file = [], {}
    for r in ranked:                                             # ranker.py:204-226
        n = per_file.get(r["file"], 0)
        if n >= MAX_PER_FILE:
            continue
        per_file[r["file"]] = n + 1
        out.append({k: r[k] for k in ("key", "final_score", "source", "signal_scores", "content", "summary", "signature",
                                      "docstring", "relationship_path", "depth_from_query")})
        if len(out) >= MAX_TOTAL:
            break
    return out


# --------------------------------------------------------------------------------------------------------------
# the older fusion: query/reranker.py
# --------------------------------------------------------------------------------------------------------------
RERANK_GRAPH_WEIGHT = 0.4
RERANK_VECTOR_WEIGHT = 0.6
RERANK_MAX_PER_FILE = 3


def rerank_fuse(graph_rows: list[dict], vector_rows: list[dict]) -> list[dict[str, Any]]:
    """ResultReranker.fuse_results (reranker.py:84-120): graph rows score 0.4 (a repeated key REPLACES the entry but keeps
    its position), vector rows score 0.6*s and ADD to an existing key (source hybrid); stable sort by score desc."""
    m: dict[str, dict[str, Any]] = {}
    for r in graph_rows:
        key = f"{r.get('file_path', '')}:{r.get('name', r.get('entity_name', ''))}:{r.get('start_line')}"
        m[key] = {"key": key, "file": r.get("file_path", ""), "score": RERANK_GRAPH_WEIGHT, "source": "graph", "content": None,
                  "summary": r.get("summary")}
    for v in vector_rows:
        key = f"{v.get('file_path', '')}:{v.get('entity_name', '')}:{v.get('start_line')}"
        s = v.get("score", 0) * RERANK_VECTOR_WEIGHT
        if key in m:
            e = m[key]
            m[key] = {"key": key, "file": e["file"], "score": e["score"] + s, "source": "hybrid",
                      "content": v.get("content") or e["content"], "summary": e["summary"] or v.get("summary")}
        else:
            m[key] = {"key": key, "file": v.get("file_path", ""), "score": s, "source": "vector", "content": v.get("content"),
                      "summary": v.get("summary")}
    return sorted(m.values(), key=lambda r: r["score"], reverse=True)


def rerank_dedup(results: list[dict], max_per_file: int = RERANK_MAX_PER_FILE) -> list[dict]:
    out, per_file = [], {}
    for r in results:                                            # reranker.py:122-145
        n = per_file.get(r["file"], 0)
        if n >= max_per_file:
            continue
        per_file[r["file"]] = n + 1
        out.append(r)
    return out


def rerank_normalize(results: list[dict]) -> list[dict]:
    if not results:                                              # reranker.py:29-70
        return results
    hi = max(r["score"] for r in results)
    lo = min(r["score"] for r in results)
    rng = hi - lo
    return [dict(r, score=1.0 if rng == 0 else (r["score"] - lo) / rng) for r in results]
