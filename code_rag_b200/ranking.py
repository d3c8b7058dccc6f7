"""Hybrid score fusion on the GPU behind the reference's ranking interface.

Mirrors ``lattice.query.ranking.HybridRanker`` (reference ``src/lattice/query/ranking/ranker.py:13-54``),
``RankingConfig`` / ``RankedResult`` (``models.py:27-91``) and the older ``lattice.query.reranker``
(``ResultReranker.fuse_results`` / ``deduplicate`` ``reranker.py:73-145``, ``normalize_scores`` ``:29-70``): same names,
arguments and results, so ``QueryEngine`` (``query/engine.py:176-181, 246-251``) can use it unchanged.  Inputs are
duck-typed (``plan.primary_intent``, ``plan.entities[i].name``, ``graph_context.primary_entities`` ... lists of nodes with
``name / qualified_name / file_path / ...``).

The host only interns strings (keys, file paths), evaluates the entity-name match (string containment) and gathers the
numeric signals; scoring, the order-dependent merge, the stable sort and the per-file / total caps run in the K3 kernel
(``csrc/rank_kernel.cuh``) for a whole batch of queries at once.  Additive API: ``rank_batch``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

from . import _native as N

DEFAULT_GRAPH_WEIGHT = 0.5
DEFAULT_VECTOR_WEIGHT = 0.5
DEFAULT_CENTRALITY_WEIGHT = 0.2
DEFAULT_CONTEXT_WEIGHT = 0.1
MAX_RESULTS_PER_FILE = 5
MAX_TOTAL_RESULTS = 50

# RankingSignal values in the kernel's signal order (models.py:17-24)
SIGNAL_NAMES = ("graph_match", "vector_similarity", "centrality", "query_entity_match", "relationship_relevance",
                "code_quality", "context_richness")
# insertion order of the per-result signal dicts (scorer.py:19-65 for graph results, :88-115 for vector results)
_GRAPH_ORDER = ("graph_match", "query_entity_match", "relationship_relevance", "centrality", "context_richness")
_VECTOR_ORDER = ("vector_similarity", "query_entity_match", "centrality", "code_quality")
_SOURCES = ("graph", "vector", "hybrid")

_INTENT_ADJUSTMENTS = {          # models.py:75-91, keyed by QueryIntent.value
    "find_callers": (0.8, 0.2), "find_callees": (0.8, 0.2), "find_call_chain": (0.9, 0.1), "find_hierarchy": (0.85, 0.15),
    "find_usages": (0.7, 0.3), "find_dependencies": (0.75, 0.25), "locate_entity": (0.6, 0.4), "locate_file": (0.5, 0.5),
    "explain_implementation": (0.5, 0.5), "explain_relationship": (0.6, 0.4), "explain_data_flow": (0.65, 0.35),
    "find_similar": (0.2, 0.8), "search_functionality": (0.3, 0.7), "search_pattern": (0.25, 0.75),
}


# intents for which QueryEngine._execute_vector_search also searches the summaries collection (query/engine.py:331-337)
SUMMARY_INTENTS = frozenset({"explain_implementation", "explain_relationship", "explain_data_flow", "explain_architecture",
                             "search_functionality"})


def _intent_value(intent: Any) -> str:
    return str(getattr(intent, "value", intent))


@dataclass
class RankedResult:
    """Field for field ``lattice.query.ranking.models.RankedResult`` (models.py:27-56)."""
    file_path: str
    entity_name: str
    entity_type: str
    qualified_name: str | None = None
    content: str | None = None
    summary: str | None = None
    signature: str | None = None
    docstring: str | None = None
    start_line: int | None = None
    end_line: int | None = None
    source: str = "hybrid"
    graph_node_id: str | None = None
    final_score: float = 0.0
    signal_scores: dict[str, float] = field(default_factory=dict)
    callers: list[str] = field(default_factory=list)
    callees: list[str] = field(default_factory=list)
    depth_from_query: int | None = None
    relationship_path: str | None = None
    metadata: dict[str, Any] = field(default_factory=dict)

    def get_key(self) -> str:
        return f"{self.file_path}:{self.entity_name}:{self.start_line}"


@dataclass
class RankingConfig:
    """``lattice.query.ranking.models.RankingConfig`` (models.py:59-91); adjustments are keyed by the intent's value."""
    graph_weight: float = DEFAULT_GRAPH_WEIGHT
    vector_weight: float = DEFAULT_VECTOR_WEIGHT
    centrality_weight: float = DEFAULT_CENTRALITY_WEIGHT
    context_weight: float = DEFAULT_CONTEXT_WEIGHT
    entity_match_bonus: float = 0.3
    relationship_bonus: float = 0.15
    query_type_adjustments: dict[Any, dict[str, float]] = field(default_factory=dict)
    max_per_file: int = MAX_RESULTS_PER_FILE
    max_total: int = MAX_TOTAL_RESULTS

    def __post_init__(self):
        if not self.query_type_adjustments:
            self.query_type_adjustments = {k: {"graph_weight": g, "vector_weight": v} for k, (g, v) in _INTENT_ADJUSTMENTS.items()}
        else:
            self.query_type_adjustments = {_intent_value(k): dict(v) for k, v in self.query_type_adjustments.items()}

    def weights_for(self, intent: Any) -> list[float]:
        w = {"graph_weight": self.graph_weight, "vector_weight": self.vector_weight,
             "centrality_weight": self.centrality_weight, "context_weight": self.context_weight}
        w.update(self.query_type_adjustments.get(_intent_value(intent), {}))       # ranker.py:56-68
        return [w["graph_weight"], w["vector_weight"], w["centrality_weight"], w["context_weight"]]


class _Batch:
    """Flat candidate arrays of a batch of queries + the objects needed to rebuild results."""

    def __init__(self):
        self.offsets = [0]
        self.kind: list[int] = []
        self.key: list[int] = []
        self.file: list[int] = []
        self.depth: list[int] = []
        self.em: list[float] = []
        self.degree: list[int] = []
        self.flags: list[int] = []
        self.clen: list[int] = []
        self.vs: list[float] = []
        self.weights: list[list[float]] = []
        self.records: list[tuple] = []       # per candidate: (kind, source object, relationship_path, depth)

    def add_query(self, weights: list[float]):
        self.weights.append(weights)
        self._keys: dict[str, int] = {}
        self._files: dict[str, int] = {}

    def end_query(self):
        self.offsets.append(len(self.kind))

    def add(self, kind, key, file_path, depth, em, degree, flags, clen, vs, record):
        self.kind.append(kind)
        self.key.append(self._keys.setdefault(key, len(self._keys)))
        self.file.append(self._files.setdefault(file_path, len(self._files)))
        self.depth.append(int(depth) if depth else 0)
        self.em.append(em)
        self.degree.append(degree)
        self.flags.append(flags)
        self.clen.append(clen)
        self.vs.append(vs)
        self.records.append(record)


def _entity_match(name: str, query_entities: set[str]) -> float:
    low = name.lower()                                              # scorer.py:31-35 / 91-96
    if low in query_entities:
        return 1.0
    if any(qe in low for qe in query_entities):
        return 0.5
    return 0.0


def _degree(key: Any, centrality: dict) -> int:
    if key in centrality:                                           # scorer.py:48-53 / 98-104
        return max(int(centrality[key].get("total_degree", 0)), 0)
    return -1


def _run(batch: _Batch, mode: int, max_per_file: int, max_total: int, entity_bonus: float, rel_bonus: float):
    lib = N.load()
    if not N.is_initialised():
        import os
        N.init(int(os.environ.get("LOCAL_RANK", "0")))       # the same default the vector store uses (one process per GPU)
    nq = len(batch.weights)
    arr = {
        "offsets": np.asarray(batch.offsets, dtype=np.int32), "kind": np.asarray(batch.kind, dtype=np.uint8),
        "key_id": np.asarray(batch.key, dtype=np.uint32), "file_id": np.asarray(batch.file, dtype=np.uint32),
        "depth": np.asarray(batch.depth, dtype=np.int32), "entity_match": np.asarray(batch.em, dtype=np.float64),
        "degree": np.asarray(batch.degree, dtype=np.int32), "flags": np.asarray(batch.flags, dtype=np.uint8),
        "content_len": np.asarray(batch.clen, dtype=np.int32), "vscore": np.asarray(batch.vs, dtype=np.float64),
        "weights": np.asarray(batch.weights, dtype=np.float64).reshape(nq, 4),
    }
    rb = N.RankBatch()
    rb.n_queries = nq
    for name, a in arr.items():
        setattr(rb, name, a.ctypes.data_as(C.c_void_p))
    nc = len(batch.kind)
    out = {
        "count": np.zeros(nq, dtype=np.int32), "index": np.zeros((nq, max_total), dtype=np.int32),
        "score": np.zeros((nq, max_total), dtype=np.float64), "norm": np.zeros((nq, max_total), dtype=np.float64),
        "signals": np.zeros((nq, max_total, len(SIGNAL_NAMES)), dtype=np.float64), "mask": np.zeros((nq, max_total), dtype=np.uint8),
        "source": np.zeros((nq, max_total), dtype=np.uint8), "leader": np.zeros(max(nc, 1), dtype=np.int32),
    }
    ms = C.c_float()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    N.check(lib.lvs_rank_fuse(C.byref(rb), mode, int(max_per_file), int(max_total), float(entity_bonus), float(rel_bonus),
                              p(out["count"]), p(out["index"]), p(out["score"]), p(out["norm"]), p(out["signals"]), p(out["mask"]),
                              p(out["source"]), p(out["leader"]), C.byref(ms)), "lvs_rank_fuse")
    out["device_ms"] = ms.value
    return out


class HybridRanker:
    def __init__(self, config: RankingConfig | None = None):
        self.config = config or RankingConfig()
        self.last_device_ms = 0.0
        self.last_search_ms = 0.0

    def rank_results(self, plan, graph_context, vector_results: list[dict[str, Any]],
                     centrality_scores: dict[str, dict[str, int]] | None = None) -> list[RankedResult]:
        """ranker.py:18-54."""
        return self.rank_batch([(plan, graph_context, vector_results, centrality_scores)])[0]

    def rank_batch(self, items: Sequence[tuple]) -> list[list[RankedResult]]:
        """Each item is the argument tuple of ``rank_results``; one kernel launch ranks all of them."""
        cfg = self.config
        b = _Batch()
        for plan, ctx, vector_results, centrality in items:
            centrality = centrality or {}
            qents = {e.name.lower() for e in plan.entities}                     # ranker.py:78, 158
            b.add_query(cfg.weights_for(plan.primary_intent))
            groups = (("primary_entities", 0, None), ("callers", 1, "caller"), ("callees", 2, "callee"), ("methods", 3, "method"),
                      ("parent_classes", 3, "parent_class"), ("child_classes", 3, "child_class"))
            for attr, kind, rel in groups:                                      # ranker.py:80-148, in this order
                for node in getattr(ctx, attr):
                    depth = None
                    if kind in (1, 2):
                        md = getattr(node, "metadata", None)
                        depth = md.get("depth", 1) if md else 1
                    flags = (1 if node.summary else 0) | (2 if node.docstring else 0) | (4 if node.signature else 0)
                    b.add(kind, f"{node.file_path}:{node.name}:{node.start_line}", node.file_path, depth,
                          _entity_match(node.name, qents), _degree(node.qualified_name or node.name, centrality), flags, -1, 0.0,
                          (kind, node, rel, depth))
            for vr in vector_results:                                           # ranker.py:150-169
                name = vr.get("entity_name", "")
                content = vr.get("content")
                b.add(4, f"{vr.get('file_path', '')}:{name}:{vr.get('start_line')}", vr.get("file_path", ""), None,
                      _entity_match(name, qents), _degree(vr.get("graph_node_id") or name, centrality),
                      (1 if vr.get("summary") else 0) | (8 if content else 0), len(content) if content else -1,
                      float(vr.get("score", 0.0)), (4, vr, None, None))
            b.end_query()
        out = _run(b, 0, cfg.max_per_file, cfg.max_total, cfg.entity_match_bonus, cfg.relationship_bonus)
        self.last_device_ms = out["device_ms"]
        results = []
        for q in range(len(items)):
            lo, hi = b.offsets[q], b.offsets[q + 1]
            members: dict[int, list[int]] = {}
            for i in range(lo, hi):
                members.setdefault(int(out["leader"][i]), []).append(i)
            ranked = []
            for s in range(int(out["count"][q])):
                li = int(out["index"][q, s])
                r = self._to_result(b.records[lo + li])
                for mi in members[li][1:]:                                      # fill missing text fields, ranker.py:186-193
                    o = self._to_result(b.records[mi])
                    for f in ("content", "summary", "signature", "docstring"):
                        if not getattr(r, f) and getattr(o, f):
                            setattr(r, f, getattr(o, f))
                mask = int(out["mask"][q, s])
                order = _VECTOR_ORDER if b.records[lo + li][0] == 4 else _GRAPH_ORDER
                sig = {n: float(out["signals"][q, s, SIGNAL_NAMES.index(n)]) for n in order}
                for n_i, n in enumerate(SIGNAL_NAMES):
                    if mask & (1 << n_i) and n not in sig:
                        sig[n] = float(out["signals"][q, s, n_i])
                r.signal_scores = sig
                r.final_score = float(out["score"][q, s])
                r.source = _SOURCES[int(out["source"][q, s])]
                ranked.append(r)
            results.append(ranked)
        return results

    def rank_batch_fused(self, coll, items: Sequence[tuple], limit: int, filters=None, summaries=None) -> list[list[RankedResult]]:
        """Search + rank in one device pass (``lvs_search_rank``).  ``coll`` is the adapter's host collection
        (``client._HostCollection`` with ranking attributes); each item is ``(plan, graph_context, query_vector, centrality)``.
        Same results as ``rank_batch`` fed with the VectorSearcher-shaped hits of ``coll.search`` - the vector-hit candidates
        (key / file ids, entity match, centrality, quality inputs) are built on the GPU from per-row columns.
        ``summaries = (summaries collection, limit, filters)``: queries with one of ``SUMMARY_INTENTS`` also get that many
        summary hits behind their code hits, as ``QueryEngine._execute_vector_search`` does (query/engine.py:331-344)."""
        cfg = self.config
        nq = len(items)
        if nq == 0:
            return []
        want = coll.want_codes(filters)
        coll2, k2, want2, sel2 = None, 0, None, []
        if summaries is not None and summaries[1] > 0:
            coll2, k2 = summaries[0], int(summaries[1])
            want2 = coll2.want_codes(summaries[2])
            sel2 = [qi for qi, it in enumerate(items) if _intent_value(it[0].primary_intent) in SUMMARY_INTENTS]
            if not sel2:
                coll2, k2 = None, 0
        b = _Batch()
        ent_off, ent_str_off, ent_blob = [0], [0], []
        cen_off, cen_id, cen_deg = [0], [], []
        tmp_keys: dict[str, int] = {}
        tmp_files: dict[Any, int] = {}
        queries = np.empty((nq, coll.dim), dtype=np.float64)
        groups = (("primary_entities", 0, None), ("callers", 1, "caller"), ("callees", 2, "callee"), ("methods", 3, "method"),
                  ("parent_classes", 3, "parent_class"), ("child_classes", 3, "child_class"))
        for qi, (plan, ctx, qvec, centrality) in enumerate(items):
            centrality = centrality or {}
            queries[qi] = np.asarray(qvec, dtype=np.float64)
            qents = {e.name.lower() for e in plan.entities}
            b.weights.append(cfg.weights_for(plan.primary_intent))
            for attr, kind, rel in groups:
                for node in getattr(ctx, attr):
                    depth = None
                    if kind in (1, 2):
                        md = getattr(node, "metadata", None)
                        depth = md.get("depth", 1) if md else 1
                    flags = (1 if node.summary else 0) | (2 if node.docstring else 0) | (4 if node.signature else 0)
                    ks = f"{node.file_path}:{node.name}:{node.start_line}"
                    kid = coll.rk_keys.get(ks)
                    if kid is None:        # not a key of any stored row: ids above 2^31 cannot collide with row keys
                        kid = 0x80000000 + tmp_keys.setdefault(ks, len(tmp_keys))
                    fid = coll.rk_files.get(node.file_path)
                    if fid is None:
                        fid = 0x80000000 + tmp_files.setdefault(node.file_path, len(tmp_files))
                    b.kind.append(kind); b.key.append(kid); b.file.append(fid); b.depth.append(int(depth) if depth else 0)
                    b.em.append(_entity_match(node.name, qents)); b.degree.append(_degree(node.qualified_name or node.name, centrality))
                    b.flags.append(flags)
                    b.records.append((kind, node, rel, depth))
            b.offsets.append(len(b.kind))
            for e in qents:
                eb = e.encode("utf-8")
                ent_blob.append(eb)
                ent_str_off.append(ent_str_off[-1] + len(eb))
            ent_off.append(len(ent_str_off) - 1)
            for name, d in centrality.items():
                cid = coll.rk_cent.get(name)
                if cid is not None:
                    cen_id.append(cid)
                    cen_deg.append(max(int(d.get("total_degree", 0)), 0))
            cen_off.append(len(cen_id))
        ng_total = len(b.kind)
        arr = {
            "offsets": np.asarray(b.offsets, dtype=np.int32), "kind": np.asarray(b.kind, dtype=np.uint8),
            "key_id": np.asarray(b.key, dtype=np.uint32), "file_id": np.asarray(b.file, dtype=np.uint32),
            "depth": np.asarray(b.depth, dtype=np.int32), "entity_match": np.asarray(b.em, dtype=np.float64),
            "degree": np.asarray(b.degree, dtype=np.int32), "flags": np.asarray(b.flags, dtype=np.uint8),
            "weights": np.asarray(b.weights, dtype=np.float64).reshape(nq, 4),
        }
        rb = N.RankBatch()
        rb.n_queries = nq
        for name, a in arr.items():
            setattr(rb, name, a.ctypes.data_as(C.c_void_p))
        cx_arr = {
            "ent_off": np.asarray(ent_off, dtype=np.int32), "ent_str_off": np.asarray(ent_str_off, dtype=np.uint32),
            "ent_bytes": np.frombuffer(b"".join(ent_blob) or b"\0", dtype=np.uint8),
            "cen_off": np.asarray(cen_off, dtype=np.int32), "cen_id": np.asarray(cen_id or [0], dtype=np.uint32),
            "cen_deg": np.asarray(cen_deg or [0], dtype=np.int32),
        }
        cx = N.RankQueryCtx()
        for name, a in cx_arr.items():
            setattr(cx, name, a.ctypes.data_as(C.c_void_p))
        k = int(limit)
        out = coll.dev.search_rank(queries, k, want, rb, cx, ng_total, cfg.max_per_file, cfg.max_total, cfg.entity_match_bonus,
                                   cfg.relationship_bonus, second=coll2.dev if coll2 is not None else None, k2=k2, want2=want2, sel2=sel2)
        self.last_device_ms = out["rank_ms"]
        self.last_search_ms = out["search_ms"]
        k2 = out["k2"]
        pos2 = {qi: j for j, qi in enumerate(sel2)} if k2 else {}
        if (out["flags"] & 1).any() or (k2 and (out["flags2"][:len(sel2)] & 1).any()):
            # rare: a search could not prove exactness with the default candidate set - take the two-step route
            hits = coll.search(queries, k, filters)
            vrs = [[coll.vector_result_from_hit(h) for h in hits[i]] for i in range(nq)]
            if k2:
                hits2 = coll2.search(queries[sel2], k2, summaries[2])
                for j, qi in enumerate(sel2):
                    vrs[qi].extend(coll2.vector_result_from_hit(h) for h in hits2[j])
            return self.rank_batch([(it[0], it[1], vrs[i], it[3]) for i, it in enumerate(items)])
        results = []
        for q in range(nq):
            g0, g1 = b.offsets[q], b.offsets[q + 1]
            ng = g1 - g0
            base = g0 + q * (k + k2)                # combined candidate index of this query's first candidate
            nh1 = int(out["hit_counts"][q])
            j2 = pos2.get(q)
            nh = nh1 + (int(out["hit_counts2"][j2]) if j2 is not None else 0)

            def record(ci, q=q, g0=g0, ng=ng, nh1=nh1, j2=j2):
                if ci < ng:
                    return b.records[g0 + ci]
                s_ = ci - ng
                if s_ < nh1:
                    return (4, coll.vector_result(int(out["hit_rows"][q, s_]), float(out["hit_scores"][q, s_])), None, None)
                s_ -= nh1
                return (4, coll2.vector_result(int(out["hit_rows2"][j2, s_]), float(out["hit_scores2"][j2, s_])), None, None)
            members: dict[int, list[int]] = {}
            for ci in range(ng + nh):
                members.setdefault(int(out["leader"][base + ci]), []).append(ci)
            ranked = []
            for s_ in range(int(out["count"][q])):
                li = int(out["index"][q, s_])
                rec = record(li)
                r = self._to_result(rec)
                for mi in members[li][1:]:
                    o = self._to_result(record(mi))
                    for f in ("content", "summary", "signature", "docstring"):
                        if not getattr(r, f) and getattr(o, f):
                            setattr(r, f, getattr(o, f))
                mask = int(out["mask"][q, s_])
                order = _VECTOR_ORDER if rec[0] == 4 else _GRAPH_ORDER
                sig = {n: float(out["signals"][q, s_, SIGNAL_NAMES.index(n)]) for n in order}
                for n_i, n in enumerate(SIGNAL_NAMES):
                    if mask & (1 << n_i) and n not in sig:
                        sig[n] = float(out["signals"][q, s_, n_i])
                r.signal_scores = sig
                r.final_score = float(out["score"][q, s_])
                r.source = _SOURCES[int(out["source"][q, s_])]
                ranked.append(r)
            results.append(ranked)
        return results

    @staticmethod
    def _to_result(record) -> RankedResult:
        kind, obj, rel, depth = record
        if kind == 4:                                                           # ranker.py:243-254
            vr = obj
            return RankedResult(file_path=vr.get("file_path", ""), entity_name=vr.get("entity_name", ""),
                                entity_type=vr.get("entity_type", ""), qualified_name=vr.get("graph_node_id"),
                                content=vr.get("content"), summary=vr.get("summary"), start_line=vr.get("start_line"),
                                end_line=vr.get("end_line"), graph_node_id=vr.get("graph_node_id"))
        n = obj                                                                 # ranker.py:228-241
        r = RankedResult(file_path=n.file_path, entity_name=n.name, entity_type=n.node_type, qualified_name=n.qualified_name,
                         summary=n.summary, signature=n.signature, docstring=n.docstring, start_line=n.start_line,
                         end_line=n.end_line, graph_node_id=n.qualified_name, metadata=getattr(n, "metadata", {}))
        r.relationship_path = rel
        r.depth_from_query = depth
        return r


# ------------------------------------------------------------------------------------------------------------------
# the older fusion (lattice.query.reranker)
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class SearchResult:
    """``lattice.query.reranker.SearchResult`` (reranker.py:11-26)."""
    source: str
    score: float
    file_path: str
    entity_type: str
    entity_name: str
    content: str | None = None
    summary: str | None = None
    start_line: int | None = None
    end_line: int | None = None
    graph_node_id: str | None = None
    metadata: dict[str, Any] | None = None

    def get_key(self) -> str:
        return f"{self.file_path}:{self.entity_name}:{self.start_line}"


class ResultReranker:
    def __init__(self, graph_weight: float = 0.4, vector_weight: float = 0.6):
        self.graph_weight = graph_weight
        self.vector_weight = vector_weight

    def fuse_batch(self, items: Sequence[tuple[list[dict], list[dict]]], max_per_file: int = 0):
        """[(graph_results, vector_results)] -> per query (fused results, normalised scores of that list)."""
        b = _Batch()
        for graph_results, vector_results in items:
            b.add_query([self.graph_weight, self.vector_weight, 0.0, 0.0])
            for r in graph_results:
                name = r.get("name", r.get("entity_name", ""))
                b.add(0, f"{r.get('file_path', '')}:{name}:{r.get('start_line')}", r.get("file_path", ""), None, 0.0, -1, 0, -1, 0.0,
                      (0, r, None, None))
            for r in vector_results:
                b.add(4, f"{r.get('file_path', '')}:{r.get('entity_name', '')}:{r.get('start_line')}", r.get("file_path", ""), None,
                      0.0, -1, 0, -1, float(r.get("score", 0)), (4, r, None, None))
            b.end_query()
        max_total = max(1, max((b.offsets[i + 1] - b.offsets[i] for i in range(len(items))), default=1))
        out = _run(b, 1, max_per_file, max_total, 0.0, 0.0)
        fused = []
        for q in range(len(items)):
            lo, hi = b.offsets[q], b.offsets[q + 1]
            members: dict[int, list[int]] = {}
            for i in range(lo, hi):
                members.setdefault(int(out["leader"][i]), []).append(i)
            rows = []
            for s in range(int(out["count"][q])):
                li = int(out["index"][q, s])
                rows.append(self._assemble([b.records[i] for i in members[li]], float(out["score"][q, s]),
                                           _SOURCES[int(out["source"][q, s])]))
            fused.append((rows, [float(v) for v in out["norm"][q, :int(out["count"][q])]]))
        return fused

    def fuse_results(self, graph_results: list[dict], vector_results: list[dict]) -> list[SearchResult]:
        """reranker.py:84-120."""
        return self.fuse_batch([(graph_results, vector_results)])[0][0]

    def deduplicate(self, results: list[SearchResult], max_per_file: int = 3) -> list[SearchResult]:
        """reranker.py:122-145: order-preserving filter of an already ranked list (no arithmetic)."""
        seen, counts, out = set(), {}, []
        for r in results:
            k = r.get_key()
            if k in seen or counts.get(r.file_path, 0) >= max_per_file:
                continue
            seen.add(k)
            counts[r.file_path] = counts.get(r.file_path, 0) + 1
            out.append(r)
        return out

    def _assemble(self, recs: list[tuple], score: float, source: str) -> SearchResult:
        cur = None
        for kind, r, _, _ in recs:                                              # reranker.py:92-115, 147-172
            if kind != 4:
                cur = SearchResult(source="graph", score=self.graph_weight, file_path=r.get("file_path", ""),
                                   entity_type=r.get("type", r.get("entity_type", "")), entity_name=r.get("name", r.get("entity_name", "")),
                                   summary=r.get("summary"), start_line=r.get("start_line"), end_line=r.get("end_line"),
                                   graph_node_id=r.get("qualified_name"))
            elif cur is None:
                cur = SearchResult(source="vector", score=0.0, file_path=r.get("file_path", ""), entity_type=r.get("entity_type", ""),
                                   entity_name=r.get("entity_name", ""), content=r.get("content"), summary=r.get("summary"),
                                   start_line=r.get("start_line"), end_line=r.get("end_line"), graph_node_id=r.get("graph_node_id"))
            else:
                cur = SearchResult(source="hybrid", score=0.0, file_path=cur.file_path, entity_type=cur.entity_type,
                                   entity_name=cur.entity_name, content=r.get("content") or cur.content,
                                   summary=cur.summary or r.get("summary"), start_line=cur.start_line, end_line=cur.end_line,
                                   graph_node_id=cur.graph_node_id, metadata=cur.metadata)
        cur.score = score
        cur.source = source
        return cur


def normalize_scores(results: list[SearchResult]) -> list[SearchResult]:
    """reranker.py:29-70 (min-max to [0, 1]; all-equal -> 1.0).  Runs the K3 kernel's normalisation on the given scores."""
    if not results:
        return results
    b = _Batch()
    b.add_query([0.0, 1.0, 0.0, 0.0])
    for i, r in enumerate(results):       # distinct keys, score passed through as a unit-weight vector hit
        b.add(4, str(i), "", None, 0.0, -1, 0, -1, float(r.score), (4, r, None, None))
    b.end_query()
    out = _run(b, 1, 0, len(results), 0.0, 0.0)
    order = [int(i) for i in out["index"][0, :len(results)]]
    norm = {order[s]: float(out["norm"][0, s]) for s in range(len(results))}
    return [SearchResult(source=r.source, score=norm[i], file_path=r.file_path, entity_type=r.entity_type, entity_name=r.entity_name,
                         content=r.content, summary=r.summary, start_line=r.start_line, end_line=r.end_line,
                         graph_node_id=r.graph_node_id, metadata=r.metadata) for i, r in enumerate(results)]
