"""The search oracle against its own frozen outputs (tests/golden/search_oracle_golden.json, made by
tests/golden/make_search_oracle_golden.py).  Self-generated: it guards the restatement against drift, it does not pin it to
qdrant-client (which cannot be installed here; the oracle's header and DESIGN.md section 2 say "parity unpinned")."""
import json
from pathlib import Path

from oracle.qdrant_local import OracleManager

G = json.loads((Path(__file__).parent / "golden" / "search_oracle_golden.json").read_text())
fh = float.fromhex


def test_search_oracle_reproduces_its_frozen_outputs():
    x = [[fh(v) for v in row] for row in G["x"]]
    q = [[fh(v) for v in row] for row in G["q"]]
    ids, pl = G["ids"], G["payloads"]
    m = OracleManager(G["dim"])
    m.create_collections()
    m.upsert("code_chunks", ids[:250], x[:250], pl[:250])
    drift = {}
    for s in G["steps"]:
        if s["op"] == "search":
            hits = m.search("code_chunks", None if s["query"] is None else q[s["query"]], limit=s["limit"], filters=s["filters"])
            assert [h["id"] for h in hits] == s["ids"]
            assert [float(h["score"]) for h in hits] == [fh(v) for v in s["scores"]]          # float64, bit for bit
            if s["filters"] is None and s["query"] is not None:
                drift.setdefault(s["query"], []).append([fh(v) for v in s["scores"]])
        elif s["op"] == "delete":
            m.delete("code_chunks", s["filters"])
        elif s["op"] == "upsert":
            m.upsert("code_chunks", ids[s["lo"]:s["hi"]], x[s["lo"]:s["hi"]], pl[s["lo"]:s["hi"]])
        elif s["op"] == "overwrite":
            m.upsert("code_chunks", [ids[i] for i in s["ids"]], [x[i] for i in s["vectors"]], [dict(pl[i], language=s["language"]) for i in s["ids"]])
        elif s["op"] == "count":
            assert m.points_count("code_chunks") == s["value"]
    # the same query asked twice: same ids, scores equal to within a few float32 ulps but not necessarily identical
    a, b = drift[0]
    assert all(abs(u - v) < 1e-6 for u, v in zip(a, b))
