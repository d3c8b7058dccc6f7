"""CPU oracle for the vector-search half of the hot path.  TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of qdrant-client *local mode* (``AsyncQdrantClient(":memory:")``), the
engine behind the reference's ``QdrantManager`` (reference ``src/lattice/embeddings/client.py:18-228``).
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product path (``code_rag_b200``) never does.

PARITY UNPINNED.  The arithmetic lives in the third-party package ``qdrant-client`` (constraint
``>=1.12.0``, un-pinned: reference ``pyproject.toml:10``), which is neither vendored under
``/root/reference`` nor installable in this image (no wheel, no network), and the reference's own tests
hold no golden vector, score or id list for this path (``tests/test_embeddings.py`` mocks Qdrant;
``tests/test_database.py:88-119`` needs a live server and asserts only ``len(results) >= 1``).  What is
restated below is the published 1.12-era algorithm of ``qdrant_client/local/{local_collection,distances,
payload_filters}.py`` anchored on the reference's call sites:

* create: ``VectorParams(size=D, distance=COSINE)``                       -> client.py:93-103
* upsert: ``PointStruct(id, vector, payload)`` overwrite-by-id            -> client.py:115-130
* search: ``query_points(query=list|None, limit, query_filter, payload)`` -> client.py:132-157
* filter: ``Filter(must=[FieldCondition(key, MatchValue(value))...])``     -> client.py:171-176
* delete by filter / scroll(limit=1) / count / get_collection             -> client.py:159-210,
                                                                             projects/cleanup.py:41-73

Local-mode arithmetic (what the device path must reproduce):

1. upsert, COSINE: ``v = np.array(vector)`` (float64) -> ``v / ||v||`` if the norm is > 0 -> stored in a
   float32 matrix row.  Existing id => the row and payload are overwritten in place.
2. search with a vector: ``q = np.array(query)`` (float64).  ``cosine_similarity`` RE-NORMALISES THE STORED
   MATRIX IN PLACE on every call (``vectors /= where(norm != 0, norm, EPSILON)`` on a float32 view, the
   norm being numpy's float32 pair-wise row reduction), normalises ``q`` in float64 and returns
   ``np.dot(vectors, q)`` - a float64 score per row.  ``argsort(scores)[::-1]`` orders the rows; rows that
   are deleted or fail the payload filter are skipped; the first ``limit`` survivors are returned.
   numpy's argsort leaves the order of exactly equal scores unspecified; BASELINE.json fixes it as
   (score desc, id asc) and so does this restatement.
3. search with ``query=None``: degenerate scroll - matching live points in ascending id order, score 0.0.
4. ``MatchValue`` is ``==`` on the payload value (any element, if the value is a list); ``MatchText`` is a
   substring test; a missing key never matches.  ``must`` is a conjunction.

The in-place re-normalisation of (2) is observable: about 9 % of float32 unit rows move by one ulp on the
first search and a few per 10^4 end in a 2-cycle, so scores depend on how many searches ran since the
row was written.  The device path replays that chain exactly on its candidates ("exact
rescoring"), which is why parity tests can demand identical id lists.
"""
from __future__ import annotations

import copy
import uuid
from typing import Any, Iterable, Sequence

import numpy as np

EPSILON = 1.1920929e-7  # qdrant_client/local/distances.py
_CHUNK_ROWS = 65536     # rows up-cast to float64 per np.dot call (bounds the temp to 65536*D*8 bytes)


class MatchValue:
    def __init__(self, value: Any):
        self.value = value


class MatchText:
    def __init__(self, text: str):
        self.text = text


class FieldCondition:
    def __init__(self, key: str, match: Any):
        self.key = key
        self.match = match


class Filter:
    def __init__(self, must: Sequence[FieldCondition] | None = None):
        self.must = list(must or [])


def build_filter(conditions: dict[str, Any] | None) -> Filter | None:
    """reference client.py:171-176 (``QdrantManager._build_filter``)."""
    if not conditions:
        return None
    return Filter(must=[FieldCondition(k, MatchValue(v)) for k, v in conditions.items()])


def _id_sort_key(point_id: Any):
    # local mode sorts scroll results by id; ints before uuid strings, each in natural order
    if isinstance(point_id, int):
        return (0, point_id, "")
    return (1, 0, str(point_id))


def _check_condition(cond: FieldCondition, payload: dict[str, Any] | None) -> bool:
    if payload is None or cond.key not in payload:
        return False
    value = payload[cond.key]
    m = cond.match
    if isinstance(m, MatchValue):
        if isinstance(value, (list, tuple)):
            return any(v == m.value and type(v) is type(m.value) or v == m.value for v in value)
        return value == m.value
    if isinstance(m, MatchText):
        return isinstance(value, str) and m.text in value
    raise TypeError(f"unsupported match {type(m).__name__}")


def check_filter(flt: Filter | None, payload: dict[str, Any] | None) -> bool:
    if flt is None:
        return True
    return all(_check_condition(c, payload) for c in flt.must)


class OracleCollection:
    """One local-mode collection: float32 row matrix + payload list + tombstones + id map."""

    def __init__(self, dim: int, distance: str = "cosine", capacity: int = 1024):
        if distance not in ("cosine", "dot"):
            raise ValueError(distance)
        self.dim = int(dim)
        self.distance = distance
        self.vectors = np.zeros((max(1, capacity), self.dim), dtype=np.float32)
        self.payload: list[dict[str, Any] | None] = []
        self.deleted = np.zeros(max(1, capacity), dtype=bool)
        self.ids: dict[Any, int] = {}
        self.ids_inv: list[Any] = []

    # ---- storage -------------------------------------------------------------------------------
    def _grow(self, need: int) -> None:
        cap = self.vectors.shape[0]
        if need <= cap:
            return
        new_cap = max(need, cap * 2)
        v = np.zeros((new_cap, self.dim), dtype=np.float32)
        v[:cap] = self.vectors
        self.vectors = v
        d = np.zeros(new_cap, dtype=bool)
        d[:cap] = self.deleted
        self.deleted = d

    @staticmethod
    def _check_id(point_id: Any) -> Any:
        if isinstance(point_id, bool):
            raise ValueError("point id must be an unsigned int or a UUID string")
        if isinstance(point_id, int):
            if point_id < 0:
                raise ValueError("point id must be an unsigned int or a UUID string")
            return point_id
        if isinstance(point_id, str):
            return str(uuid.UUID(point_id))  # raises ValueError on a malformed id, like local mode
        raise ValueError("point id must be an unsigned int or a UUID string")

    def upsert(self, ids: Sequence[Any], vectors: Iterable[Sequence[float]],
               payloads: Sequence[dict[str, Any] | None]) -> None:
        for point_id, vector, payload in zip(ids, vectors, payloads):
            point_id = self._check_id(point_id)
            v = np.array(vector, dtype=np.float64)
            if v.shape != (self.dim,):
                raise ValueError(f"vector has shape {v.shape}, collection expects ({self.dim},)")
            if self.distance == "cosine":
                norm = np.linalg.norm(v)
                v = v / norm if norm > 0 else v
            if point_id in self.ids:
                row = self.ids[point_id]
            else:
                row = len(self.ids_inv)
                self._grow(row + 1)
                self.ids[point_id] = row
                self.ids_inv.append(point_id)
                self.payload.append(None)
            self.vectors[row] = v  # float64 -> float32 rounding happens here
            self.payload[row] = copy.deepcopy(payload)
            self.deleted[row] = False

    def upsert_rows_f32(self, first_id: int, rows: np.ndarray, payloads: Sequence[dict[str, Any] | None]) -> None:
        """Bulk form of :meth:`upsert` for synthetic corpora: integer ids ``first_id..``; same arithmetic."""
        rows = np.asarray(rows)
        n = rows.shape[0]
        base = len(self.ids_inv)
        self._grow(base + n)
        for s in range(0, n, _CHUNK_ROWS):
            x = rows[s:s + _CHUNK_ROWS].astype(np.float64)
            if self.distance == "cosine":
                # per-row np.linalg.norm(v) is sqrt(dot(v, v)); einsum keeps the same float64 arithmetic
                nrm = np.sqrt(np.einsum("ij,ij->i", x, x))
                nz = nrm > 0
                x[nz] = x[nz] / nrm[nz, None]
            self.vectors[base + s: base + s + x.shape[0]] = x
        for i in range(n):
            pid = first_id + i
            if pid in self.ids:
                raise ValueError("upsert_rows_f32 is append-only")
            self.ids[pid] = base + i
            self.ids_inv.append(pid)
            self.payload.append(copy.copy(payloads[i]) if payloads[i] is not None else None)

    # ---- queries -------------------------------------------------------------------------------
    def _live_mask(self, flt: Filter | None) -> np.ndarray:
        n = len(self.payload)
        mask = ~self.deleted[:n]
        if flt is not None:
            for i in range(n):
                if mask[i] and not check_filter(flt, self.payload[i]):
                    mask[i] = False
        return mask

    def _scores(self, query: Sequence[float]) -> np.ndarray:
        n = len(self.payload)
        q = np.array(query, dtype=np.float64)
        if q.shape != (self.dim,):
            raise ValueError(f"query has shape {q.shape}, collection expects ({self.dim},)")
        assert not np.isnan(q).any(), "Query vector must not contain NaN"
        scores = np.empty(n, dtype=np.float64)
        if self.distance == "cosine":
            qn = np.linalg.norm(q)
            q = q / np.where(qn != 0.0, qn, EPSILON)
        for s in range(0, n, _CHUNK_ROWS):
            v = self.vectors[s:min(n, s + _CHUNK_ROWS)]  # a view: the division below mutates storage
            if self.distance == "cosine":
                vn = np.linalg.norm(v, axis=-1)[:, np.newaxis]          # float32, pair-wise per row
                v /= np.where(vn != 0.0, vn, np.float32(EPSILON))        # in place, float32
            scores[s:s + v.shape[0]] = np.dot(v.astype(np.float64), q)  # np.dot(f32 matrix, f64 q) up-casts
        return scores

    def search(self, query: Sequence[float] | None, limit: int = 10, flt: Filter | None = None,
               with_payload: bool = True) -> list[dict[str, Any]]:
        """``query_points`` as the reference calls it (client.py:142-148)."""
        if query is None:
            return self.scroll(flt, limit, with_payload=with_payload, as_scored=True)
        scores = self._scores(query)
        mask = self._live_mask(flt)
        cand = np.nonzero(mask)[0]
        if cand.size == 0 or limit <= 0:
            return []
        # order = (score desc, id asc); ids that are ints sort numerically, uuid strings lexically
        id_rank = {r: k for k, r in enumerate(sorted(cand.tolist(), key=lambda r: _id_sort_key(self.ids_inv[r])))}
        order = sorted(cand.tolist(), key=lambda r: (-scores[r], id_rank[r]))[:limit]
        return [
            {"id": self.ids_inv[r], "score": float(scores[r]),
             "payload": copy.deepcopy(self.payload[r]) if with_payload else None, "row": r}
            for r in order
        ]

    def search_topk_rows(self, query: Sequence[float], limit: int, mask: np.ndarray | None = None):
        """Vectorised variant for big synthetic corpora whose ids are monotone in the row number
        (so ``id asc`` == ``row asc``).  Returns ``(rows int64[k], scores float64[k])``."""
        scores = self._scores(query)
        n = scores.shape[0]
        live = ~self.deleted[:n]
        if mask is not None:
            live &= mask[:n]
        cand = np.nonzero(live)[0]
        if cand.size == 0 or limit <= 0:
            return np.empty(0, np.int64), np.empty(0, np.float64)
        s = scores[cand]
        order = np.lexsort((cand, -s))[:limit]  # primary: -score asc, secondary: row asc
        return cand[order].astype(np.int64), s[order]

    def scroll(self, flt: Filter | None, limit: int, with_payload: bool = True, as_scored: bool = False):
        n = len(self.payload)
        rows = [r for r in range(n) if not self.deleted[r] and check_filter(flt, self.payload[r])]
        rows.sort(key=lambda r: _id_sort_key(self.ids_inv[r]))
        rows = rows[:max(0, limit)]
        out = []
        for r in rows:
            rec = {"id": self.ids_inv[r], "payload": copy.deepcopy(self.payload[r]) if with_payload else None, "row": r}
            if as_scored:
                rec["score"] = 0.0
            out.append(rec)
        return out

    def delete(self, flt: Filter) -> int:
        n = len(self.payload)
        cnt = 0
        for r in range(n):
            if not self.deleted[r] and check_filter(flt, self.payload[r]):
                self.deleted[r] = True
                cnt += 1
        return cnt

    def count(self, flt: Filter | None = None) -> int:
        return int(self._live_mask(flt).sum())


class OracleClient:
    """The slice of ``AsyncQdrantClient(':memory:')`` that ``QdrantManager`` touches (synchronous here)."""

    def __init__(self):
        self.collections: dict[str, OracleCollection] = {}

    def create_collection(self, name: str, dim: int, distance: str = "cosine") -> None:
        if name in self.collections:
            raise ValueError(f"Collection {name} already exists")
        self.collections[name] = OracleCollection(dim, distance)

    def delete_collection(self, name: str) -> None:
        self.collections.pop(name, None)

    def get_collections(self) -> list[str]:
        return list(self.collections)

    def _get(self, name: str) -> OracleCollection:
        if name not in self.collections:
            raise ValueError(f"Collection {name} not found")
        return self.collections[name]

    def upsert(self, name, ids, vectors, payloads):
        self._get(name).upsert(ids, vectors, payloads)

    def query_points(self, name, query, limit=10, query_filter=None, with_payload=True):
        return self._get(name).search(query, limit, query_filter, with_payload)

    def delete(self, name, flt):
        return self._get(name).delete(flt)

    def scroll(self, name, scroll_filter=None, limit=10, with_payload=True):
        return self._get(name).scroll(scroll_filter, limit, with_payload)

    def count(self, name, count_filter=None):
        return self._get(name).count(count_filter)

    def points_count(self, name) -> int:
        return self._get(name).count(None)


class OracleManager:
    """``QdrantManager`` (client.py:18-228) over :class:`OracleClient`: same call surface, synchronous."""

    CODE_CHUNKS = "code_chunks"
    SUMMARIES = "summaries"

    def __init__(self, dim: int):
        self.dim = dim
        self.client = OracleClient()

    def create_collections(self) -> None:  # client.py:72-91
        for name in (self.CODE_CHUNKS, self.SUMMARIES):
            if name not in self.client.get_collections():
                self.client.create_collection(name, self.dim, "cosine")

    def upsert(self, collection, ids, vectors, payloads) -> None:  # client.py:115-130
        self.client.upsert(collection, ids, vectors, payloads)

    def search(self, collection, query_vector, limit=10, filters=None):  # client.py:132-157
        pts = self.client.query_points(collection, query_vector, limit, build_filter(filters))
        return [{"id": str(p["id"]), "score": p["score"], "payload": p["payload"]} for p in pts]

    def delete(self, collection, filters) -> None:  # client.py:159-169
        self.client.delete(collection, build_filter(filters))

    def file_needs_update(self, collection, file_path, content_hash) -> bool:  # client.py:178-202
        try:
            pts = self.client.scroll(collection, build_filter({"file_path": file_path}), limit=1)
            if not pts:
                return True
            return pts[0]["payload"].get("content_hash") != content_hash
        except Exception:
            return True

    def points_count(self, collection) -> int:  # client.py:204-210 + query/engine.py:302
        return self.client.points_count(collection)
