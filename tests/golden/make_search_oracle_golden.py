#!/usr/bin/env python
"""Generates tests/golden/search_oracle_golden.json from oracle/qdrant_local.py ITSELF.

These are NOT reference vectors (qdrant-client cannot be installed here - parity of the search oracle stays unpinned, see the
oracle's header and DESIGN.md section 2).  They freeze the restatement's behaviour - ids, float64 scores, the drift that the
in-place re-normalisation produces from search to search, filter semantics, scroll order, overwrite and delete - so that an
accidental edit of the oracle cannot silently move the target the CUDA path is tested against.
    python tests/golden/make_search_oracle_golden.py
"""
from __future__ import annotations

import json
import sys
import uuid
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

from oracle.qdrant_local import OracleManager  # noqa: E402


def build():
    dim, n = 24, 300
    rng = np.random.default_rng(20261018)
    mu = 0.5 * rng.standard_normal(dim)
    x = (mu + rng.standard_normal((n, dim))).astype(np.float32)
    q = mu + rng.standard_normal((8, dim))
    ids = [str(uuid.UUID(int=int(v))) for v in rng.integers(1, 2**62, size=n)]
    pl = [{"file_path": f"f{i % 11}.py", "language": ("python", "go", "rust")[i % 3], "project_name": f"p{i % 2}", "entity_name": f"e{i}",
           "content_hash": f"h{i % 5}"} for i in range(n)]
    m = OracleManager(dim)
    m.create_collections()
    vec = x.astype(np.float64).tolist()
    m.upsert("code_chunks", ids[:250], vec[:250], pl[:250])
    steps = []

    def search(qi, k, flt):
        hits = m.search("code_chunks", None if qi is None else q[qi].tolist(), limit=k, filters=flt)
        steps.append({"op": "search", "query": qi, "limit": k, "filters": flt, "ids": [h["id"] for h in hits],
                      "scores": [float(h["score"]).hex() for h in hits]})
    search(0, 5, None); search(1, 5, {"language": "go"}); search(0, 5, None)            # same query twice: the scores drift
    m.delete("code_chunks", {"file_path": "f3.py"}); steps.append({"op": "delete", "filters": {"file_path": "f3.py"}})
    search(2, 7, {"project_name": "p1"}); search(None, 6, {"file_path": "f4.py"})
    m.upsert("code_chunks", ids[250:], vec[250:], pl[250:]); steps.append({"op": "upsert", "lo": 250, "hi": n})
    m.upsert("code_chunks", ids[10:12], vec[20:22], [dict(pl[10], language="zig"), dict(pl[11], language="zig")])
    steps.append({"op": "overwrite", "ids": [10, 11], "vectors": [20, 21], "language": "zig"})
    search(3, 10, None); search(4, 3, {"language": "zig"}); search(5, 5, {"language": "cobol"}); search(3, 10, None)
    steps.append({"op": "count", "value": m.points_count("code_chunks")})
    return {"dim": dim, "n": n, "x": [[float(v).hex() for v in row] for row in x.astype(np.float64)], "q": [[float(v).hex() for v in row] for row in q],
            "ids": ids, "payloads": pl, "steps": steps}


if __name__ == "__main__":
    out = ROOT / "tests" / "golden" / "search_oracle_golden.json"
    out.write_text(json.dumps(build()))
    print(f"wrote {out} ({out.stat().st_size} bytes)")
