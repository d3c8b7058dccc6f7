"""Host half of the ranking mirror (code_rag_b200/ranking.py) on CPU: candidate packing (string interning, entity match,
centrality lookup, flags, insertion order) and result assembly (leader groups, filling of missing text fields, signal order,
sources) are exercised against the goldens made by the reference's own code, with the one native call (``ranking._run`` ->
``lvs_rank_fuse``) replaced by a TESTS-ONLY emulation over the SAME packed arrays.  The emulation scores with the ranking
oracle (``oracle.ranking.score_candidate``); the product path itself has no such fallback (tests/test_abi.py pins that)."""
import gzip
import json
from pathlib import Path
from types import SimpleNamespace as NS

import numpy as np
import pytest

from code_rag_b200 import ranking as RK
from oracle import ranking as R

GOLDEN = Path(__file__).parent / "golden" / "ranking_golden.json.gz"
SIG = RK.SIGNAL_NAMES


def _emulated_run(batch, mode, max_per_file, max_total, entity_bonus, rel_bonus):
    """What lvs_rank_fuse returns (csrc/rank_kernel.cuh), computed in Python from the packed candidate arrays."""
    nq = len(batch.weights)
    out = {"count": np.zeros(nq, np.int32), "index": np.zeros((nq, max_total), np.int32), "score": np.zeros((nq, max_total)),
           "norm": np.zeros((nq, max_total)), "signals": np.zeros((nq, max_total, 7)), "mask": np.zeros((nq, max_total), np.uint8),
           "source": np.zeros((nq, max_total), np.uint8), "leader": np.zeros(max(len(batch.kind), 1), np.int32), "device_ms": 0.0}
    for q in range(nq):
        lo, hi = batch.offsets[q], batch.offsets[q + 1]
        w = dict(zip(("graph_weight", "vector_weight", "centrality_weight", "context_weight"), batch.weights[q]))
        cands = []
        for i in range(lo, hi):
            f = batch.flags[i]
            c = {"kind": batch.kind[i], "depth": batch.depth[i], "entity_match": batch.em[i], "degree": None if batch.degree[i] < 0 else batch.degree[i],
                 "has_summary": bool(f & 1), "has_docstring": bool(f & 2), "has_signature": bool(f & 4), "has_content": bool(f & 8),
                 "content_len": None if batch.clen[i] < 0 else batch.clen[i], "vscore": batch.vs[i]}
            if mode == 0:
                sc, sig = R.score_candidate(c, w)
            else:
                sc, sig = (w["graph_weight"] if c["kind"] < 4 else c["vscore"] * w["vector_weight"]), {}
            cands.append((sc, sig))
        C = hi - lo
        lead = [next(j for j in range(i + 1) if batch.key[lo + j] == batch.key[lo + i]) for i in range(C)]
        for i in range(C):
            out["leader"][lo + i] = lead[i]
        groups = {}
        for i in range(C):
            groups.setdefault(lead[i], []).append(i)
        merged = []
        for l, members in groups.items():
            fin, sig = cands[l][0], dict(cands[l][1])
            any_vec = batch.kind[lo + l] == 4
            for j in members[1:]:
                if mode == 0:
                    fin = (fin + cands[j][0]) / 2 * 1.1
                else:
                    fin = cands[j][0] if batch.kind[lo + j] < 4 else fin + cands[j][0]
                for s, v in cands[j][1].items():
                    sig[s] = max(sig[s], v) if s in sig else v
                any_vec = any_vec or batch.kind[lo + j] == 4
            if mode == 0:
                src = 2 if len(members) > 1 else (1 if batch.kind[lo + l] == 4 else 0)
            else:
                src = 2 if (any_vec and len(members) > 1) else (1 if any_vec else 0)
            merged.append((l, fin, sig, src))
        merged.sort(key=lambda r: -r[1])                     # stable: leaders in insertion order among equal scores
        per_file, n = {}, 0
        for l, fin, sig, src in merged:
            fid = batch.file[lo + l]
            if max_per_file > 0 and per_file.get(fid, 0) >= max_per_file:
                continue
            per_file[fid] = per_file.get(fid, 0) + 1
            if n < max_total:
                out["index"][q, n], out["score"][q, n], out["source"][q, n] = l, fin, src
                for s, v in sig.items():
                    out["signals"][q, n, SIG.index(s)] = v
                    out["mask"][q, n] |= 1 << SIG.index(s)
                n += 1
        out["count"][q] = n
        sc = out["score"][q, :n]
        if n:
            rng = sc.max() - sc.min()
            out["norm"][q, :n] = 1.0 if rng == 0 else (sc - sc.min()) / rng
    return out


@pytest.fixture()
def cases(monkeypatch):
    monkeypatch.setattr(RK, "_run", _emulated_run)
    return json.loads(gzip.decompress(GOLDEN.read_bytes()))["cases"]


def _inputs(case):
    node = lambda d: NS(node_type=d["node_type"], name=d["name"], qualified_name=d["qualified_name"], file_path=d["file_path"],
                        signature=d["signature"], docstring=d["docstring"], summary=d["summary"], start_line=d["start_line"],
                        end_line=d["end_line"], metadata=dict(d["metadata"]))
    g = case["graph"]
    ctx = NS(**{k: [node(n) for n in g[k]] for k in ("primary_entities", "callers", "callees", "methods", "parent_classes", "child_classes")})
    plan = NS(primary_intent=NS(value=case["intent"]), entities=[NS(name=e) for e in case["entities"]])
    return plan, ctx, [dict(v) for v in case["vector"]], dict(case["centrality"])


def test_hybrid_ranker_host_half_against_reference_goldens(cases):
    out = RK.HybridRanker().rank_batch([_inputs(c) for c in cases])
    for case, got in zip(cases, out):
        exp = case["expected"]["hybrid"]
        assert [r.get_key() for r in got] == [e["key"] for e in exp], case["id"]
        for r, e in zip(got, exp):
            assert r.final_score == e["final_score"] and r.source == e["source"], (case["id"], r.get_key())
            assert r.signal_scores == e["signal_scores"] and list(r.signal_scores) == list(e["signal_scores"]), (case["id"], r.get_key())
            for f in ("content", "summary", "signature", "docstring", "relationship_path", "depth_from_query"):
                assert getattr(r, f) == e[f], (case["id"], r.get_key(), f)


def test_reranker_host_half_against_reference_goldens(cases):
    rr = RK.ResultReranker()
    for case in cases[:40]:
        exp = case["expected"]
        fused = rr.fuse_results(exp["graph_rows"], [dict(v) for v in case["vector"]])
        tup = lambda rs: [(r.get_key(), r.score, r.source, r.content, r.summary) for r in rs]
        want = lambda rs: [(e["key"], e["score"], e["source"], e["content"], e["summary"]) for e in rs]
        assert tup(fused) == want(exp["fused"]), case["id"]
        dedup = rr.deduplicate(fused)
        assert tup(dedup) == want(exp["dedup"]), case["id"]
        assert tup(RK.normalize_scores(dedup)) == want(exp["normalized"]), case["id"]


def test_ranking_config_matches_the_reference_defaults():
    cfg = RK.RankingConfig()
    assert (cfg.graph_weight, cfg.vector_weight, cfg.centrality_weight, cfg.context_weight) == (0.5, 0.5, 0.2, 0.1)
    assert cfg.weights_for(NS(value="find_similar")) == [0.2, 0.8, 0.2, 0.1]
    assert cfg.weights_for(NS(value="explain_architecture")) == [0.5, 0.5, 0.2, 0.1]      # no override for this intent
    assert cfg.weights_for("find_call_chain") == [0.9, 0.1, 0.2, 0.1]
    assert (cfg.max_per_file, cfg.max_total, cfg.entity_match_bonus, cfg.relationship_bonus) == (5, 50, 0.3, 0.15)
