"""Probe for the REAL engine behind the reference's search path.  TEST INFRASTRUCTURE ONLY.

The reference's ``QdrantManager`` (reference ``src/lattice/embeddings/client.py:18-228``) forwards to the third-party
package ``qdrant-client`` (``>=1.12.0``, reference ``pyproject.toml:10``).  That package is not in this image and cannot be
installed (no wheel in ``/opt/wheelhouse``, no network), so ``oracle/qdrant_local.py`` restates its local mode and says
"parity unpinned".  This module is what flips that the day the package is present: :func:`find` looks for an importable
``qdrant_client`` (site-packages, then ``baseline/_ref`` and ``oracle/_ref``, where a driver or maintainer may drop it) and
:class:`RealManager` drives ``AsyncQdrantClient(":memory:")`` through exactly the calls the reference makes:

* ``create_collection(collection_name, vectors_config=VectorParams(size, distance=COSINE))``   client.py:93-103
* ``create_payload_index(collection_name, field_name, field_schema=KEYWORD)``                  client.py:105-113
* ``upsert(collection_name, points=[PointStruct(id, vector, payload)...])``                    client.py:115-130
* ``query_points(collection_name, query, limit, query_filter, with_payload=True)``             client.py:132-157
* ``delete(collection_name, points_selector=FilterSelector(filter=...))``                      client.py:159-169
* ``scroll(collection_name, scroll_filter, limit, with_payload=True)``                         client.py:178-202
* ``get_collection(collection_name).points_count``                                             client.py:204-210
* ``count(collection_name, count_filter=Filter(must=[FieldCondition(key, MatchText)]))``       projects/cleanup.py:41-46

Users: ``tests/test_real_qdrant_parity.py`` (oracle == real package; device == real package) and the ``--impl reference`` /
``cpu_baseline`` legs of ``bench.py`` (which then report ``kind = "qdrant-client <version>"`` instead of ``"port"``).
"""
from __future__ import annotations

import asyncio
import importlib
import importlib.util
import sys
from pathlib import Path
from typing import Any

ROOT = Path(__file__).resolve().parent.parent
_EXTRA = (ROOT / "baseline" / "_ref", ROOT / "oracle" / "_ref")


def find():
    """Returns the imported ``qdrant_client`` module or None.  Never raises."""
    try:
        if importlib.util.find_spec("qdrant_client") is None:
            for p in _EXTRA:
                if (p / "qdrant_client").is_dir() and str(p) not in sys.path:
                    sys.path.append(str(p))
            importlib.invalidate_caches()
            if importlib.util.find_spec("qdrant_client") is None:
                return None
        return importlib.import_module("qdrant_client")
    except Exception:  # noqa: BLE001  (a broken drop-in must read as "absent", not as a test error)
        return None


def version() -> str | None:
    if find() is None:
        return None
    try:
        from importlib.metadata import version as _v
        return _v("qdrant-client")
    except Exception:  # noqa: BLE001
        return "unknown"


class RealManager:
    """``QdrantManager`` (client.py:18-228) over the real ``AsyncQdrantClient(":memory:")``, synchronous for the tests."""

    CODE_CHUNKS = "code_chunks"
    SUMMARIES = "summaries"

    def __init__(self, dim: int):
        qc = find()
        if qc is None:
            raise RuntimeError("qdrant_client is not importable")
        from qdrant_client import models  # type: ignore
        self.models = models
        self.dim = dim
        self.loop = asyncio.new_event_loop()
        self.client = qc.AsyncQdrantClient(":memory:")

    def _run(self, coro):
        return self.loop.run_until_complete(coro)

    def close(self) -> None:
        try:
            self._run(self.client.close())
        finally:
            self.loop.close()

    def create_collections(self) -> None:
        m = self.models
        existing = {c.name for c in self._run(self.client.get_collections()).collections}
        for name, fields in ((self.CODE_CHUNKS, ["file_path", "entity_type", "language", "content_hash", "project_name"]),
                             (self.SUMMARIES, ["file_path", "entity_type"])):
            if name in existing:
                continue
            self._run(self.client.create_collection(collection_name=name,
                                                    vectors_config=m.VectorParams(size=self.dim, distance=m.Distance.COSINE)))
            for f in fields:
                self._run(self.client.create_payload_index(collection_name=name, field_name=f,
                                                           field_schema=m.PayloadSchemaType.KEYWORD))

    def _filter(self, conditions: dict[str, Any] | None):
        if not conditions:
            return None
        m = self.models
        return m.Filter(must=[m.FieldCondition(key=k, match=m.MatchValue(value=v)) for k, v in conditions.items()])

    def upsert(self, collection, ids, vectors, payloads) -> None:
        m = self.models
        pts = [m.PointStruct(id=i, vector=list(map(float, v)), payload=p) for i, v, p in zip(ids, vectors, payloads)]
        self._run(self.client.upsert(collection_name=collection, points=pts))

    def search(self, collection, query_vector, limit=10, filters=None):
        r = self._run(self.client.query_points(collection_name=collection,
                                               query=None if query_vector is None else list(map(float, query_vector)),
                                               limit=limit, query_filter=self._filter(filters), with_payload=True))
        return [{"id": str(p.id), "score": p.score, "payload": p.payload} for p in r.points]

    def delete(self, collection, filters) -> None:
        m = self.models
        self._run(self.client.delete(collection_name=collection, points_selector=m.FilterSelector(filter=self._filter(filters))))

    def file_needs_update(self, collection, file_path, content_hash) -> bool:
        try:
            pts, _ = self._run(self.client.scroll(collection_name=collection, scroll_filter=self._filter({"file_path": file_path}),
                                                  limit=1, with_payload=True))
            if not pts:
                return True
            return pts[0].payload.get("content_hash") != content_hash
        except Exception:  # noqa: BLE001  (client.py:200-202 returns True on any error)
            return True

    def points_count(self, collection) -> int:
        return int(self._run(self.client.get_collection(collection_name=collection)).points_count)

    def count_text(self, collection, key: str, text: str) -> int:
        m = self.models
        flt = m.Filter(must=[m.FieldCondition(key=key, match=m.MatchText(text=text))])
        return int(self._run(self.client.count(collection_name=collection, count_filter=flt)).count)
