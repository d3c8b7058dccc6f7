"""``B200VectorStore``: drop-in for lattice's ``QdrantManager`` backed by a GPU-resident collection.

Mirrors, method for method, ``lattice.embeddings.client.QdrantManager`` (reference
``src/lattice/embeddings/client.py:18-228``; protocol ``VectorStore``, ``src/lattice/core/protocols.py:34-52``):
same names, keyword arguments, result dictionaries and error behaviour, so it can be injected wherever the
reference passes a ``QdrantManager`` (``QueryEngine(qdrant=...)`` ``query/engine.py:37``,
``VectorSearcher(qdrant, embedder)`` ``query/vector_search.py:46-58``, ``VectorIndexer(qdrant, ...)``
``embeddings/indexer.py:36-44``, ``ContextBuilder(memgraph, qdrant)`` ``query/context/builder.py:24-30``).

Division of labour: ids (uuid strings), payload dicts and the value dictionaries of the keyword columns live in
host RAM, keyed by row; vectors, tombstones and the dictionary codes live in HBM.  All arithmetic (normalisation,
scoring, selection, filter evaluation) runs in liblattice_b200.so.
"""
from __future__ import annotations

import asyncio
import logging
import os
import threading
import uuid
from enum import Enum
from types import SimpleNamespace
from typing import Any, Sequence

import numpy as np

import struct

from . import _native as N
from .collection import DeviceCollection
from .errors import VectorStoreError

_PACKERS: dict[int, struct.Struct] = {}


def _pack_query(query_vector) -> bytes | None:
    """list[float] -> the bytes of its float64 values, or None when ``struct`` refuses the list (the numpy way reports the error)."""
    if type(query_vector) is not list:
        return None
    n = len(query_vector)
    pk = _PACKERS.get(n)
    if pk is None:
        pk = _PACKERS[n] = struct.Struct(f"{n}d")
    try:
        return pk.pack(*query_vector)
    except (struct.error, TypeError):
        return None


def _as_query(query_vector) -> np.ndarray:
    """list[float] (what an embedding provider hands to lattice) -> float64 [1, dim].  ``struct.pack`` converts a 768-element list
    in a quarter of the time ``np.asarray`` takes; anything it refuses (nested sequences, strings) goes the numpy way and fails there
    the way it always did."""
    packed = _pack_query(query_vector)
    if packed is not None:
        return np.frombuffer(packed, dtype=np.float64)[None, :]
    return np.asarray(query_vector, dtype=np.float64)[None, :]


class _CollectionLock:
    """The lock of a collection's host half.  ``with lock:`` is exclusive - writes, blocking searches - and waits until the
    pipelined searches in flight have been collected.  Such a search (``B200VectorStore.search`` on the event loop) holds the mutex
    only while it submits (``try_enter`` ... ``entered``) and while it collects (``try_reenter`` ... ``reader_done``); several may be
    in flight, they execute in submission order on the collection's stream, and no write can slip between a search's submission
    and the moment its rows are turned into payloads.  A waiting writer keeps new readers out."""

    def __init__(self):
        self._m = threading.Lock()
        self._c = threading.Condition(self._m)
        self.readers = 0
        self._writers = 0

    def acquire(self, blocking: bool = True) -> bool:
        if not self._m.acquire(blocking):
            return False
        if self.readers:
            if not blocking:
                self._m.release()
                return False
            self._writers += 1
            try:
                while self.readers:
                    self._c.wait()
            finally:
                self._writers -= 1
        return True

    def release(self) -> None:
        self._m.release()

    def __enter__(self):
        self.acquire()
        return self

    def __exit__(self, *exc):
        self._m.release()

    def try_enter(self, max_readers: int) -> bool:
        """Non-blocking: the mutex for a reader about to submit; refused while a writer waits or `max_readers` are in flight."""
        if not self._m.acquire(False):
            return False
        if self._writers or self.readers >= max_readers:
            self._m.release()
            return False
        return True

    def entered(self) -> None:
        """The reader has submitted: it counts as in flight and gives the mutex back."""
        self.readers += 1
        self._m.release()

    def try_reenter(self) -> bool:
        return self._m.acquire(False)

    def reenter(self) -> None:
        self._m.acquire()

    def reader_done(self) -> None:
        """The reader (holding the mutex) has collected its result."""
        self.readers -= 1
        self._c.notify_all()
        self._m.release()

logger = logging.getLogger(__name__)

DEFAULT_DIMENSIONS = 1536  # reference config/settings.py:53 (AISettings.embedding_dimensions)


class CollectionName(str, Enum):
    """reference embeddings/client.py:13-15."""
    CODE_CHUNKS = "code_chunks"
    SUMMARIES = "summaries"


# keyword payload indexes the reference creates (client.py:76-89)
_INDEX_FIELDS = {
    CollectionName.CODE_CHUNKS.value: ["file_path", "entity_type", "language", "content_hash", "project_name"],
    CollectionName.SUMMARIES.value: ["file_path", "entity_type"],
}


def _canonical_id(point_id: Any) -> Any:
    """Qdrant point ids are unsigned ints or UUID strings (local mode validates the same way)."""
    if isinstance(point_id, bool):
        raise ValueError(f"invalid point id {point_id!r}")
    if isinstance(point_id, int):
        if point_id < 0:
            raise ValueError(f"invalid point id {point_id!r}")
        return point_id
    if isinstance(point_id, uuid.UUID):
        return str(point_id)
    if isinstance(point_id, str):
        return str(uuid.UUID(point_id))
    raise ValueError(f"invalid point id {point_id!r}")


def _tie_key(point_id: Any) -> int:
    """64-bit key whose order follows the id order (ints numerically; uuids by their leading 64 bits)."""
    if isinstance(point_id, int):
        return min(point_id, 0xFFFFFFFFFFFFFFFF)
    return uuid.UUID(point_id).int >> 64


def vector_result_from_payload(p: dict[str, Any] | None, score: float, kind: str | None = "code") -> dict[str, Any]:
    """The dict VectorSearcher hands to the ranker for a hit (query/vector_search.py:221-260): ``kind`` "summary" is
    ``_transform_summary_result``, anything else ``_transform_code_result``."""
    p = p or {}
    if kind == "summary":
        return {"score": score, "file_path": p.get("file_path"), "entity_type": p.get("entity_type"),
                "entity_name": p.get("entity_name"), "summary": p.get("summary"), "graph_node_id": p.get("graph_node_id")}
    return {"score": score, "file_path": p.get("file_path"), "entity_type": p.get("entity_type"),
            "entity_name": p.get("entity_name"), "language": p.get("language"), "content": p.get("content"),
            "start_line": p.get("start_line"), "end_line": p.get("end_line"), "graph_node_id": p.get("graph_node_id")}


def _id_sort_key(point_id: Any):
    return (0, point_id, "") if isinstance(point_id, int) else (1, 0, point_id)


_INLINE_SEARCH_BYTES = 256 << 20      # shards up to this size (a ~50 us scan) are searched without leaving the event loop's thread
_POLLED_SEARCH = os.environ.get("LATTICE_B200_POLLED_SEARCH", "1") != "0"   # 0: every large search takes a worker thread (A/B switch)
_ENTER_SPINS = 2000          # event-loop turns a search waits for a free submit slot / for a writer before it takes a thread
_POLL_SPINS = 20000          # event-loop turns a search polls its completion word before a thread waits for it


class _HostCollection:
    """Host half of a collection: ids, payloads and the per-column value dictionaries."""

    def inline_bytes(self) -> int:
        """Bytes one search has to stream (decides whether ``B200VectorStore.search`` hops to a worker thread)."""
        return len(self.ids) * self.dim * (2 if getattr(self.dev, "storage", "f32") == "bf16" else 4)

    def __init__(self, name: str, dim: int, storage: str, index_fields: Sequence[str], device: int, dev_factory=None,
                 rank_kind: str | None = None, rank_dicts: tuple[dict, dict, dict] | None = None):
        self.name = name
        # fused search -> rank (include/lvs.h lvs_search_rank): per-row ranking attributes are derived from the payload at
        # upsert, the way VectorSearcher._transform_code_result / _transform_summary_result (query/vector_search.py:221-260)
        # and HybridRanker._process_vector_results (ranking/ranker.py:150-169) would read them.  None = feature off.
        self.rank_kind = rank_kind
        # interned keys / file paths / centrality keys: ONE id space per store, so that the hits of `code_chunks` and of
        # `summaries` can be ranked together (per-file cap, merges with graph candidates); names are per collection
        self.rk_keys, self.rk_files, self.rk_cent = rank_dicts if rank_dicts is not None else ({}, {}, {})
        self.rk_names: dict[str, int] = {}
        self.dim = dim
        self.columns: list[str] = list(index_fields)[: N.MAX_FILTER_COLS]
        self.dicts: list[dict[Any, int]] = [dict() for _ in self.columns]
        self.ids: list[Any] = []                 # row -> id
        self.id_to_row: dict[Any, int] = {}
        self.payloads: list[dict[str, Any] | None] = []
        # rows of deleted points, reused by later upserts: VectorIndexer.index_file (embeddings/indexer.py:61-77) deletes a
        # file's chunks and inserts new ones under fresh uuid4 ids on every re-index, so without reuse a shard only grows
        self.free_rows: list[int] = []
        # The device orders equal scores by a 64-bit key derived from the id (_tie_key), then by row.  Ids that share their key
        # (uuids equal in their leading 64 bits: never uuid4, but e.g. uuid.UUID(int=small)) would come back in row order, so
        # the keys carried by more than one live point are tracked and such hits are put in id order on the host (search()).
        self.tie_counts: dict[int, int] = {}
        self.dup_keys: dict[int, int] = {}
        self.lock = _CollectionLock()
        # dev_factory exists so that the host-side bookkeeping can be unit-tested without a GPU (tests only)
        factory = dev_factory or DeviceCollection
        # timing off: no CUDA events around the kernels, so a search completes through the word its kernel stores into the pinned
        # slot (lower latency) and consecutive searches of concurrent awaits overlap on the GPU
        self.dev = factory(name, dim, storage=storage, metric="cosine", n_filter_cols=N.MAX_FILTER_COLS,
                           capacity=0, device=device, timing=False)

    # -- dictionary encoding ---------------------------------------------------------------------------
    def _encode_value(self, col: int, value: Any, create: bool) -> int:
        if value is None:
            return N.NULL_CODE
        if isinstance(value, (list, tuple, dict, set)):
            raise ValueError(f"payload key {self.columns[col]!r}: list/dict values are not supported in keyword columns")
        d = self.dicts[col]
        code = d.get(value)
        if code is None:
            if not create:
                return N.NO_MATCH
            code = len(d) + 1
            d[value] = code
        return code

    def encode_payloads(self, payloads: Sequence[dict[str, Any] | None]) -> np.ndarray:
        codes = np.zeros((len(payloads), N.MAX_FILTER_COLS), dtype=np.uint32)
        for i, p in enumerate(payloads):
            if not p:
                continue
            for c, key in enumerate(self.columns):
                if key in p:
                    codes[i, c] = self._encode_value(c, p[key], create=True)
        return codes

    def ensure_column(self, key: str) -> int:
        """Index a payload key on first use in a filter (Qdrant filters on any key, indexed or not)."""
        if key in self.columns:
            return self.columns.index(key)
        if len(self.columns) >= N.MAX_FILTER_COLS:
            raise ValueError(f"cannot filter on {key!r}: all {N.MAX_FILTER_COLS} keyword columns are in use ({self.columns})")
        col = len(self.columns)
        self.columns.append(key)
        self.dicts.append(dict())
        self.backfill_column(col)
        return col

    def backfill_column(self, col: int) -> None:
        """Codes of a column that was added after rows were written, from the payloads kept on the host."""
        key, n = self.columns[col], len(self.payloads)
        if n:
            codes = np.zeros(n, dtype=np.uint32)
            for r, p in enumerate(self.payloads):
                if p and key in p:
                    codes[r] = self._encode_value(col, p[key], create=True)
            self.dev.set_codes(col, codes, row0=0)

    def want_codes(self, filters: dict[str, Any] | None) -> np.ndarray | None:
        if not filters:
            return None
        want = np.full(N.MAX_FILTER_COLS, N.ANY, dtype=np.uint32)
        for key, value in filters.items():
            if value is None or isinstance(value, (float, list, tuple, dict, set)):
                # models.MatchValue only accepts bool | int | str (pydantic validation error in the reference)
                raise ValueError(f"filter value for {key!r} must be a bool, int or str, got {type(value).__name__}")
            col = self.ensure_column(key)
            code = self._encode_value(col, value, create=False)
            if want[col] != N.ANY and want[col] != code:
                code = N.NO_MATCH
            want[col] = code
        return want

    # -- order among exactly equal scores ----------------------------------------------------------------------
    def _tie_add(self, key: int) -> None:
        c = self.tie_counts.get(key, 0) + 1
        self.tie_counts[key] = c
        if c > 1:
            self.dup_keys[key] = c

    def _tie_drop(self, key: int) -> None:
        c = self.tie_counts.get(key, 0) - 1
        if c <= 0:
            self.tie_counts.pop(key, None)
        else:
            self.tie_counts[key] = c
        if c > 1:
            self.dup_keys[key] = c
        else:
            self.dup_keys.pop(key, None)

    def rebuild_tie_counts(self) -> None:
        self.tie_counts.clear(); self.dup_keys.clear()
        for pid in self.ids:
            if pid is not None:
                self._tie_add(_tie_key(pid))

    def device_limit(self, limit: int) -> int:
        """How many hits to ask the device for so that the best `limit` in (score desc, id asc) order are among them: the run of
        equal (score, key) hits that straddles the cut has at most max(dup_keys) members.  Bounded by MAX_TIE_RUN - 1 extra
        hits (one device call, whatever the ids look like): runs of more than MAX_TIE_RUN bit-identical vectors under ids with
        one key are cut in row order."""
        if not self.dup_keys:
            return limit
        return min(N.MAX_K, limit + min(max(self.dup_keys.values()), self.MAX_TIE_RUN) - 1)

    MAX_TIE_RUN = 16

    def in_id_order(self, rows: np.ndarray, scores: np.ndarray, limit: int, id_of=None) -> tuple[np.ndarray, np.ndarray]:
        """(score desc, id asc) - BASELINE.json's rule - for hits that came back ordered by (score desc, key asc, row asc)."""
        if self.dup_keys and len(rows) > 1:
            id_of = id_of or (lambda r: self.ids[r])
            order = sorted(range(len(rows)), key=lambda j: (-float(scores[j]), _id_sort_key(id_of(int(rows[j])))))
            rows, scores = rows[order], scores[order]
        return rows[:limit], scores[:limit]

    # -- operations (called under self.lock) ---------------------------------------------------------------
    def upsert(self, ids, vectors, payloads) -> None:
        n = len(ids)
        if not (len(vectors) == n and len(payloads) == n):
            n = min(n, len(vectors), len(payloads))  # the reference zips the three lists (client.py:123-126)
        if n == 0:
            return
        vec = np.asarray(vectors[:n], dtype=np.float64)
        if vec.ndim != 2 or vec.shape[1] != self.dim:
            raise ValueError(f"vectors must have dimension {self.dim}, got shape {vec.shape}")
        if not np.isfinite(vec).all():
            raise ValueError("vectors must be finite")
        self._upsert_with(ids[:n], payloads[:n], lambda keep, rows, codes, ties: self.dev.upsert(vec[keep], rows=rows, codes=codes, ties=ties))

    def upsert_tokens(self, ids, token_ids: np.ndarray, payloads, encoder) -> None:
        """Embed and upsert in one device call: `token_ids` [n, L] go through `encoder` (embedding.B200CodeEncoder)
        and the pooled vectors pass from its last kernel to the upsert kernel without leaving HBM - the step that in the reference is
        embed_batch -> list[list[float]] -> QdrantManager.upsert (embeddings/indexer.py:77-86)."""
        tok = np.ascontiguousarray(token_ids, dtype=np.int32)
        n = min(len(ids), tok.shape[0], len(payloads))
        if n == 0:
            return
        if tok.ndim != 2:
            raise ValueError("token_ids must be [n, L]")
        if encoder.hidden != self.dim:
            raise ValueError(f"the encoder produces {encoder.hidden}-dimensional vectors, the collection holds {self.dim}")
        self._upsert_with(ids[:n], payloads[:n],
                          lambda keep, rows, codes, ties: encoder.embed_upsert(self.dev, tok[:n][keep], rows=rows, codes=codes, ties=ties))

    def _upsert_with(self, ids, payloads, write) -> None:
        """Host bookkeeping of an upsert (row assignment, overwrite by id, row reuse, dictionary codes, tie keys) around
        `write(keep, rows, codes, ties)`, which puts the vectors of the points `keep` into `rows` on the device."""
        n = len(ids)
        canon = [_canonical_id(i) for i in ids[:n]]
        # a repeated id inside one batch: the last occurrence wins, as with sequential point upserts
        last = {pid: i for i, pid in enumerate(canon)}
        keep = sorted(last.values())
        rows = np.empty(len(keep), dtype=np.int64)
        next_row = len(self.ids)
        new_ids = []
        reused: list[tuple[Any, int]] = []
        for j, i in enumerate(keep):
            pid = canon[i]
            r = self.id_to_row.get(pid)
            if r is None:
                if self.free_rows:
                    r = self.free_rows.pop()
                    reused.append((pid, r))
                else:
                    r = next_row
                    next_row += 1
                    new_ids.append(pid)
            rows[j] = r
        pl = [dict(payloads[i]) if payloads[i] is not None else None for i in keep]
        codes = self.encode_payloads(pl)
        ties = np.array([_tie_key(canon[i]) for i in keep], dtype=np.uint64)
        try:
            write(keep, rows, codes, ties)
        except Exception:
            self.free_rows.extend(r for _, r in reused)          # nothing was written: the rows stay reusable
            raise
        for pid, r in reused:
            self.id_to_row[pid] = r
            self.ids[r] = pid
            self._tie_add(_tie_key(pid))
        for pid in new_ids:
            self.id_to_row[pid] = len(self.ids)
            self.ids.append(pid)
            self.payloads.append(None)
            self._tie_add(_tie_key(pid))
        for j, r in enumerate(rows):
            self.payloads[int(r)] = pl[j]
        if self.rank_kind is not None:
            self._set_rank_attrs(rows, pl)

    def vector_result(self, row: int, score: float) -> dict[str, Any]:
        """The dict VectorSearcher hands to the ranker for this row (query/vector_search.py:221-260)."""
        return vector_result_from_payload(self.payloads[row], score, self.rank_kind)

    def vector_result_from_hit(self, hit: dict[str, Any]) -> dict[str, Any]:
        return self.vector_result(self.id_to_row[_canonical_id(hit["id"])], hit["score"])

    def _set_rank_attrs(self, rows: np.ndarray, pl: Sequence[dict[str, Any] | None]) -> None:
        n = len(rows)
        key = np.empty(n, dtype=np.uint32); fil = np.empty(n, dtype=np.uint32); cent = np.empty(n, dtype=np.uint32)
        nam = np.empty(n, dtype=np.uint32); clen = np.empty(n, dtype=np.int32); flg = np.empty(n, dtype=np.uint8)
        new_names: list[bytes] = []
        first_new = len(self.rk_names)
        for j in range(n):
            vr = self.vector_result(int(rows[j]), 0.0)
            name = vr.get("entity_name")
            content, summary = vr.get("content"), vr.get("summary")
            key[j] = self.rk_keys.setdefault(f"{vr.get('file_path')}:{name}:{vr.get('start_line')}", len(self.rk_keys))
            fil[j] = self.rk_files.setdefault(vr.get("file_path"), len(self.rk_files))
            cent[j] = self.rk_cent.setdefault(vr.get("graph_node_id") or name, len(self.rk_cent))
            low = name.lower() if isinstance(name, str) else ""      # the reference raises on a missing name; here: no match
            nid = self.rk_names.get(low)
            if nid is None:
                nid = len(self.rk_names)
                self.rk_names[low] = nid
                new_names.append(low.encode("utf-8"))
            nam[j] = nid
            clen[j] = len(content) if content else -1
            flg[j] = (1 if summary else 0) | (8 if content else 0)
        if new_names:
            got = self.dev.rank_names_append(new_names)
            if got != first_new:
                raise RuntimeError(f"entity-name pool out of step (device {got}, host {first_new})")
        self.dev.rank_attrs_set(rows, key, fil, cent, nam, clen, flg)

    def _hits(self, rows, scores) -> list[dict[str, Any]]:
        """rows / scores: arrays or plain lists."""
        if not isinstance(rows, list):
            rows, scores = rows.tolist(), scores.tolist()
        ids, payloads = self.ids, self.payloads
        out = []
        for r, s in zip(rows, scores):
            if r < 0:
                continue
            p = payloads[r]
            out.append({"id": str(ids[r]), "score": float(s), "payload": dict(p) if p is not None else None})
        return out

    def search_one(self, query_vector, limit: int, filters: dict[str, Any] | None) -> list[dict[str, Any]]:
        """``search`` for one query given as the caller's list, without numpy in between (the inline path of a small collection)."""
        packed = _pack_query(query_vector) if 0 < limit <= N.MAX_K and hasattr(self.dev, "search_packed") else None
        if packed is None or len(packed) != 8 * self.dim:
            return self.search(_as_query(query_vector), limit, filters)[0]
        n, flags, rows, scores = self.dev.search_packed(packed, self.device_limit(limit), self.want_codes(filters)).single()
        if flags & N.FLAG_UNPROVEN:
            logger.warning("search on %s: exactness bound not met (many near-ties); result is the best of the largest candidate set", self.name)
        rows, scores = rows[:n], scores[:n]
        if self.dup_keys and n > 1:                      # (score desc, id asc): see in_id_order
            order = sorted(range(n), key=lambda j: (-scores[j], _id_sort_key(self.ids[rows[j]])))
            rows, scores = [rows[j] for j in order], [scores[j] for j in order]
        return self._hits(rows[:limit], scores[:limit])

    def search(self, query_vectors: np.ndarray | None, limit: int, filters: dict[str, Any] | None) -> list[list[dict[str, Any]]]:
        want = self.want_codes(filters)
        if limit <= 0:
            return [[] for _ in range(1 if query_vectors is None else len(query_vectors))]
        if query_vectors is None:
            # query=None degenerates to a scroll: matching points in ascending id order, score 0.0
            rows, _ = self.dev.match_rows(want)
            order = sorted(rows.tolist(), key=lambda r: _id_sort_key(self.ids[r]))[:limit]
            return [self._hits(np.asarray(order, dtype=np.int64), np.zeros(len(order)))]
        if limit > N.MAX_K:
            raise ValueError(f"limit {limit} exceeds the largest supported top-k ({N.MAX_K})")
        return self._shape(self.dev.search(query_vectors, self.device_limit(limit), want), limit)

    def _shape(self, res, limit: int) -> list[list[dict[str, Any]]]:
        out = []
        for qi in range(res.rows.shape[0]):
            n = int(res.counts[qi])
            if res.flags[qi] & N.FLAG_UNPROVEN:
                logger.warning("search on %s: exactness bound not met for query %d (many near-ties); "
                               "result is the best of the largest candidate set", self.name, qi)
            out.append(self._hits(*self.in_id_order(res.rows[qi, :n], res.scores[qi, :n], limit)))
        return out

    # The same search in three steps, for callers that must not block (``B200VectorStore.search`` on the event loop): begin
    # submits (``lvs_search_submit``) and returns, ready polls the completion word the kernel stores, end collects and shapes.
    # begin and end run under the collection's mutex, the wait in between does not (``_CollectionLock``).
    MAX_IN_FLIGHT = 4                # the library's submit slots per collection

    def search_begin(self, query_vectors: np.ndarray | None, limit: int, filters: dict[str, Any] | None) -> dict:
        if query_vectors is None or limit <= 0 or len(query_vectors) == 0 or not hasattr(self.dev, "search_poll"):
            return {"result": self.search(query_vectors, limit, filters)}
        if limit > N.MAX_K:
            raise ValueError(f"limit {limit} exceeds the largest supported top-k ({N.MAX_K})")
        return {"ticket": self.dev.search_submit(query_vectors, self.device_limit(limit), self.want_codes(filters)), "limit": limit}

    def search_ready(self, h: dict) -> bool:
        return "ticket" not in h or self.dev.search_poll(h["ticket"])

    def search_block(self, h: dict) -> None:
        """Blocking wait (from a worker thread, when polling has gone on for too long); needs no host-side lock."""
        if "ticket" in h:
            h["res"] = self.dev.search_wait(h.pop("ticket"))

    def search_end(self, h: dict) -> list[list[dict[str, Any]]]:
        if "result" in h:
            return h["result"]
        self.search_block(h)
        return self._shape(h["res"], h["limit"])

    def search_discard(self, h: dict) -> None:
        """The caller gave up (cancelled): the ticket must not stay in flight."""
        try:
            self.search_block(h)
        except Exception:  # noqa: BLE001
            pass

    def release_rows(self, rows) -> None:
        """Forget the ids and payloads of deleted rows and queue the rows for reuse."""
        for r in np.asarray(rows, dtype=np.int64).tolist():
            pid = self.ids[r]
            if pid is None:
                continue
            self.id_to_row.pop(pid, None)
            self._tie_drop(_tie_key(pid))
            self.ids[r] = None
            self.payloads[r] = None
            self.free_rows.append(r)

    COMPACT_MIN_FREE = 4096          # compaction runs when at least this many rows AND a quarter of the shard are free

    def maybe_compact(self, force: bool = False) -> int:
        """Move the live rows of the tail into the holes and truncate (after a mass delete - projects/cleanup.py:38-73 drops a
        whole project - tombstones would otherwise keep costing scan bandwidth).  Returns the number of rows dropped."""
        nfree, n = len(self.free_rows), len(self.ids)
        if nfree == 0 or not (force or (nfree >= self.COMPACT_MIN_FREE and 4 * nfree >= n)):
            return 0
        m = n - nfree                                        # rows after compaction
        holes = sorted(r for r in self.free_rows if r < m)
        tail_live = [r for r in range(m, n) if self.ids[r] is not None]
        assert len(holes) == len(tail_live)
        if tail_live:
            self.dev.move_rows(np.asarray(tail_live, dtype=np.int64) + self.dev.row_base, np.asarray(holes, dtype=np.int64) + self.dev.row_base)
            for src, dst in zip(tail_live, holes):
                pid = self.ids[src]
                self.ids[dst], self.payloads[dst] = pid, self.payloads[src]
                self.id_to_row[pid] = dst
        self.dev.truncate(m)
        del self.ids[m:], self.payloads[m:]
        self.free_rows = []
        return nfree

    def delete(self, filters: dict[str, Any]) -> int:
        want = self.want_codes(filters)
        rows, n = self.dev.delete_where(want)
        self.release_rows(rows)
        self.maybe_compact()
        return n

    def scroll(self, filters: dict[str, Any] | None, limit: int) -> list[dict[str, Any]]:
        return self.search(None, limit, filters)[0]

    def count(self, filters: dict[str, Any] | None = None) -> int:
        if not filters:
            return self.dev.count()
        _, n = self.dev.match_rows(self.want_codes(filters), cap=0)
        return n

    # -- what ``manager.client`` needs (projects/cleanup.py:38-73): equality conditions on the device, MatchText on the host --
    def rows_matching(self, eq: dict[str, Any], text: Sequence[tuple[str, str]]) -> list[int]:
        rows, _ = self.dev.match_rows(self.want_codes(eq))
        out = []
        for r in rows.tolist():
            p = self.payloads[r] or {}
            # MatchText in local mode is a substring test on the string payload value
            if all(isinstance(p.get(k), str) and t in p[k] for k, t in text):
                out.append(r)
        return out

    def point(self, row: int) -> tuple[Any, dict[str, Any] | None]:
        return self.ids[row], self.payloads[row]

    def delete_found(self, rows: Sequence[int]) -> None:
        if len(rows):
            self.dev.delete_rows(np.asarray(rows, dtype=np.int64))
            self.release_rows(rows)
            self.maybe_compact()

    def close(self) -> None:
        self.dev.close()

    # -- snapshots ---------------------------------------------------------------------------------------------
    _HOST_STATE = ("name", "dim", "columns", "dicts", "ids", "id_to_row", "payloads", "free_rows", "rank_kind", "rk_keys", "rk_files",
                   "rk_cent", "rk_names")

    def save(self, directory: str) -> None:
        """Device arrays -> ``<name>.lvs`` (raw, row order), host half (ids, payloads, dictionaries) -> ``<name>.host.pkl``."""
        import pickle
        os.makedirs(directory, exist_ok=True)
        self.dev.save_snapshot(os.path.join(directory, f"{self.name}.lvs"))
        with open(os.path.join(directory, f"{self.name}.host.pkl"), "wb") as f:
            pickle.dump({k: getattr(self, k) for k in self._HOST_STATE}, f, protocol=pickle.HIGHEST_PROTOCOL)

    @classmethod
    def load(cls, directory: str, name: str, device: int) -> "_HostCollection":
        import pickle
        with open(os.path.join(directory, f"{name}.host.pkl"), "rb") as f:
            state = pickle.load(f)
        self = cls.__new__(cls)
        for k in cls._HOST_STATE:
            setattr(self, k, state[k])
        self.lock = _CollectionLock()
        self.tie_counts, self.dup_keys = {}, {}
        self.rebuild_tie_counts()
        self.dev = DeviceCollection.load_snapshot(os.path.join(directory, f"{name}.lvs"), name=name, device=device)
        if self.dev.rows != len(self.ids):
            raise ValueError(f"snapshot of {name}: {self.dev.rows} device rows but {len(self.ids)} host rows")
        return self


class _ClientShim:
    """What ``manager.client`` exposes (reference callers reach through it: projects/cleanup.py:41-61,
    tests/test_database.py:81; ``scroll`` and ``query_points`` are the two calls QdrantManager itself makes, client.py:142,182).  Filters are duck-typed ``models.Filter`` objects (``.must[i].key``,
    ``.match.value`` / ``.match.text``)."""

    def __init__(self, store: "B200VectorStore"):
        self._store = store

    async def get_collections(self):
        return SimpleNamespace(collections=[SimpleNamespace(name=n) for n in self._store._collections])

    async def get_collection(self, collection_name: str):
        return await self._store.get_collection_info(collection_name)

    def _split_filter(self, flt) -> tuple[dict[str, Any], list[tuple[str, str]]]:
        eq: dict[str, Any] = {}
        text: list[tuple[str, str]] = []
        for cond in (getattr(flt, "must", None) or []):
            m = cond.match
            if hasattr(m, "value"):
                eq[cond.key] = m.value
            elif hasattr(m, "text"):
                text.append((cond.key, m.text))
            else:
                raise ValueError(f"unsupported match {type(m).__name__}")
        return eq, text

    def _rows_matching(self, coll: _HostCollection, flt) -> list[int]:
        eq, text = self._split_filter(flt) if flt is not None else ({}, [])
        return coll.rows_matching(eq, text)

    async def count(self, collection_name: str, count_filter=None, exact: bool = True):
        coll = self._store._get(collection_name)

        def work():
            with coll.lock:
                return len(self._rows_matching(coll, count_filter))
        return SimpleNamespace(count=await asyncio.to_thread(work))

    async def delete(self, collection_name: str, points_selector=None):
        coll = self._store._get(collection_name)
        flt = getattr(points_selector, "filter", points_selector)

        def work():
            with coll.lock:
                rows = self._rows_matching(coll, flt)
                coll.delete_found(rows)
                return len(rows)
        await asyncio.to_thread(work)
        return SimpleNamespace(status="completed")

    async def scroll(self, collection_name: str, scroll_filter=None, limit: int = 10, offset=None, with_payload: bool = True,
                     with_vectors: bool = False, **_ignored):
        """``AsyncQdrantClient.scroll`` as ``QdrantManager.file_needs_update`` uses it (client.py:182-190): matching points in
        ascending id order from ``offset`` (a point id) on -> ``(records, next_offset)``.  Vectors are not returned."""
        coll = self._store._get(collection_name)
        start = None if offset is None else _id_sort_key(_canonical_id(offset))

        def work():
            with coll.lock:
                found = sorted((coll.point(r) for r in self._rows_matching(coll, scroll_filter)), key=lambda t: _id_sort_key(t[0]))
            if start is not None:
                found = [t for t in found if _id_sort_key(t[0]) >= start]
            page, rest = found[:max(0, limit)], found[max(0, limit):]
            records = [SimpleNamespace(id=pid if isinstance(pid, int) else str(pid), payload=(dict(p) if p is not None else None)
                                       if with_payload else None, vector=None) for pid, p in page]
            nxt = rest[0][0] if rest else None
            return records, (nxt if nxt is None or isinstance(nxt, int) else str(nxt))
        return await asyncio.to_thread(work)

    async def query_points(self, collection_name: str, query=None, limit: int = 10, query_filter=None, with_payload: bool = True,
                           **_ignored):
        """``AsyncQdrantClient.query_points`` as ``QdrantManager.search`` calls it (client.py:142-148) -> ``.points`` of scored
        points (``id``, ``score``, ``payload``).  ``MatchValue`` conditions only (``MatchText`` is for count / delete / scroll)."""
        eq, text = self._split_filter(query_filter) if query_filter is not None else ({}, [])
        if text:
            raise ValueError("MatchText is not supported in query_points")
        hits = await self._store.search(collection=collection_name, query_vector=None if query is None else list(query),
                                        limit=limit, filters=eq or None)
        return SimpleNamespace(points=[SimpleNamespace(id=h["id"], score=h["score"], payload=h["payload"] if with_payload else None,
                                                       version=0, vector=None) for h in hits])

    async def close(self):
        return None


class B200VectorStore:
    """Same constructor keywords as ``QdrantManager`` (ignored: there is no server), plus backend knobs."""

    def __init__(self, host: str | None = None, port: int | None = None, grpc_port: int | None = None, *,
                 dimensions: int | None = None, storage: str | None = None, device: int | None = None,
                 rank_attrs: bool = False, _device_factory=None):
        self._host, self._port, self._grpc_port = host, port, grpc_port
        if dimensions is None:
            dimensions = int(os.environ.get("EMBEDDING_DIMENSIONS", DEFAULT_DIMENSIONS))
            try:  # honour lattice's own settings object when it is importable
                from lattice.config import get_settings  # type: ignore
                dimensions = int(get_settings().embedding_dimensions)
            except Exception:  # noqa: BLE001
                pass
        self._dimensions = int(dimensions)
        self._storage = storage or os.environ.get("LATTICE_B200_STORAGE", "f32")
        self._device = int(device if device is not None else os.environ.get("LOCAL_RANK", "0"))
        self._device_factory = _device_factory
        self._rank_attrs = bool(rank_attrs)     # keep per-row ranking attributes on the device (enables search_and_rank)
        self._rank_dicts: tuple[dict, dict, dict] = ({}, {}, {})
        self._connected = False
        self._collections: dict[str, _HostCollection] = {}
        self._shim = _ClientShim(self)

    # ---- connection (client.py:32-70) -------------------------------------------------------------------
    async def connect(self) -> None:
        if not self._connected:
            try:
                if self._device_factory is None:
                    await asyncio.to_thread(N.init, self._device)
                self._connected = True
                logger.info("lattice-b200 vector store bound to cuda:%d", self._device)
            except Exception as e:  # noqa: BLE001
                raise VectorStoreError("Failed to connect to Qdrant", cause=e)

    async def close(self) -> None:
        if self._connected:
            try:
                for coll in self._collections.values():
                    coll.close()
            except Exception as e:  # noqa: BLE001
                logger.warning(f"Error closing vector store: {e}")
            finally:
                self._collections.clear()
                self._connected = False

    @property
    def client(self) -> _ClientShim:
        if not self._connected:
            raise VectorStoreError("Client not connected. Call connect() first.")
        return self._shim

    async def health_check(self) -> bool:
        try:
            await self.client.get_collections()
            return True
        except Exception as e:  # noqa: BLE001
            logger.warning(f"Vector store health check failed: {e}")
            return False

    # ---- collections (client.py:72-113, 204-221) ---------------------------------------------------------
    def _get(self, name: str) -> _HostCollection:
        if not self._connected:
            raise VectorStoreError("Client not connected. Call connect() first.")
        coll = self._collections.get(name)
        if coll is None:
            raise ValueError(f"Collection {name} not found")
        return coll

    async def create_collections(self) -> None:
        try:
            _ = self.client
            for name in (CollectionName.CODE_CHUNKS.value, CollectionName.SUMMARIES.value):
                if name not in self._collections:
                    self._collections[name] = await asyncio.to_thread(
                        _HostCollection, name, self._dimensions, self._storage, _INDEX_FIELDS[name], self._device,
                        self._device_factory,
                        ("summary" if name == CollectionName.SUMMARIES.value else "code") if self._rank_attrs else None,
                        self._rank_dicts)
                    logger.info(f"Created collection: {name}")
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError("Failed to create collections", cause=e)

    # ---- snapshots (the Qdrant volume of docker-compose.yml:42-43) ------------------------------
    async def save(self, directory: str) -> None:
        """Additive: persist every collection (device arrays + ids / payloads / dictionaries) under ``directory``."""
        try:
            _ = self.client

            def work():
                for coll in self._collections.values():
                    with coll.lock:
                        coll.save(directory)
            await asyncio.to_thread(work)
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to save collections to {directory}", cause=e)

    async def load(self, directory: str) -> None:
        """Additive: replace the collections by the ones saved under ``directory``; searches continue exactly where the
        saved store left off (same scores: the local-mode replay state is part of the snapshot).  The host half is a pickle:
        load only snapshots this application wrote itself."""
        try:
            _ = self.client

            def work():
                loaded = {}
                for name in (CollectionName.CODE_CHUNKS.value, CollectionName.SUMMARIES.value):
                    if os.path.exists(os.path.join(directory, f"{name}.lvs")):
                        loaded[name] = _HostCollection.load(directory, name, self._device)
                return loaded
            loaded = await asyncio.to_thread(work)
            for name, coll in loaded.items():
                old = self._collections.pop(name, None)
                if old is not None:
                    old.close()
                self._collections[name] = coll
            if loaded:       # one id space per store again (the pickles of one save() hold equal copies)
                first = next(iter(loaded.values()))
                self._rank_dicts = (first.rk_keys, first.rk_files, first.rk_cent)
                for coll in self._collections.values():
                    coll.rk_keys, coll.rk_files, coll.rk_cent = self._rank_dicts
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to load collections from {directory}", cause=e)

    async def get_collection_info(self, collection: str):
        try:
            coll = self._get(collection)

            def work():
                with coll.lock:
                    n = coll.count()
                return SimpleNamespace(points_count=n, vectors_count=n, indexed_vectors_count=n, status="green",
                                       config=SimpleNamespace(params=SimpleNamespace(
                                           vectors=SimpleNamespace(size=coll.dim, distance="Cosine"))))
            return await asyncio.to_thread(work)
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to get collection info for {collection}", cause=e)

    async def clear_collections(self) -> None:
        self._rank_dicts = ({}, {}, {})
        for name in (CollectionName.CODE_CHUNKS.value, CollectionName.SUMMARIES.value):
            coll = self._collections.pop(name, None)
            if coll is not None:
                coll.close()
                logger.info(f"Deleted collection: {name}")
        await self.create_collections()

    # ---- data path (client.py:115-202) -------------------------------------------------------------------
    async def upsert(self, collection: str, ids: list[str], vectors: list[list[float]], payloads: list[dict[str, Any]]) -> None:
        try:
            coll = self._get(collection)

            def work():
                with coll.lock:
                    coll.upsert(ids, vectors, payloads)
            await asyncio.to_thread(work)
            logger.debug(f"Upserted {len(ids)} vectors to {collection}")
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to upsert vectors to {collection}", cause=e)

    async def upsert_tokens(self, collection: str, ids: list[str], token_ids, payloads: list[dict[str, Any]], encoder) -> None:
        """Additive API: index chunks from their TOKEN IDS.  `encoder` (``embedding.B200CodeEncoder``) embeds
        them on this store's GPU and the vectors go straight into the shard - what ``VectorIndexer.index_file`` does with
        ``embedder.embed_batch`` + ``upsert`` (reference embeddings/indexer.py:77-86) without the list[float] round trip."""
        try:
            coll = self._get(collection)

            def work():
                with coll.lock:
                    coll.upsert_tokens(ids, token_ids, payloads, encoder)
            await asyncio.to_thread(work)
            logger.debug(f"Embedded and upserted {len(ids)} vectors to {collection}")
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to upsert vectors to {collection}", cause=e)

    async def search(self, collection: str, query_vector: list[float] | None, limit: int = 10,
                     filters: dict[str, Any] | None = None) -> list[dict[str, Any]]:
        try:
            coll = self._get(collection)
            flt = filters or None
            inline = getattr(coll, "inline_bytes", None)
            if query_vector is not None and inline is not None and inline() <= _INLINE_SEARCH_BYTES and coll.lock.acquire(blocking=False):
                try:                                     # see below
                    return coll.search_one(query_vector, limit, flt)
                finally:
                    coll.lock.release()
            q = None if query_vector is None else _as_query(query_vector)

            def work():
                with coll.lock:
                    return coll.search(q, limit, flt)[0]
            # A small collection (lattice's usual operating point: 10^4 - 10^5 chunks) answers in tens of microseconds: it is searched
            # inline (above) when nobody else holds the collection - list in, dicts out, no numpy in between.  Anything larger is
            # SUBMITTED from the event loop (a non-blocking library call) and polled between yields - a hop to a worker thread and
            # back costs 0.1-0.2 ms, as much as a whole step on eight GPUs - so concurrent awaits (asyncio.gather in
            # query/engine.py:142-146) pipeline on the device instead of queueing behind a lock.  Filter-only lookups and contended
            # collections take a thread (threading).
            if _POLLED_SEARCH and q is not None and hasattr(coll, "search_begin") and hasattr(coll.lock, "try_enter"):
                results = await self._search_polled(coll, q, limit, flt, work)
            else:
                results = await asyncio.to_thread(work)
            logger.debug(f"Found {len(results)} results in {collection}")
            return results
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to search {collection}", cause=e)

    @staticmethod
    async def _search_polled(coll, q, limit, flt, work):
        lock = coll.lock
        spins = 0
        while not lock.try_enter(coll.MAX_IN_FLIGHT):
            spins += 1
            if spins > _ENTER_SPINS:                     # a long write is under way: queue behind it in a thread
                return await asyncio.to_thread(work)
            await asyncio.sleep(0)
        try:
            h = coll.search_begin(q, limit, flt)
        except BaseException:
            lock.release()
            raise
        if h is None:                                    # not this way (sharded store: another command is under way)
            lock.release()
            return await asyncio.to_thread(work)
        lock.entered()
        try:
            spins = 0
            while not coll.search_ready(h):
                spins += 1
                if spins > _POLL_SPINS:                  # a long search: let a thread wait for it
                    await asyncio.to_thread(coll.search_block, h)
                    break
                await asyncio.sleep(0)
            while not lock.try_reenter():
                await asyncio.sleep(0)
        except BaseException:                            # cancelled (or the poll failed): the ticket must not stay in flight
            lock.reenter()
            try:
                coll.search_discard(h)
            finally:
                lock.reader_done()
            raise
        try:
            return coll.search_end(h)[0]
        finally:
            lock.reader_done()

    async def search_batch(self, collection: str, query_vectors: Sequence[Sequence[float]], limit: int = 10,
                           filters: dict[str, Any] | None = None) -> list[list[dict[str, Any]]]:
        """Additive API: Q searches in one device pass (equivalent to Q consecutive ``search`` calls)."""
        try:
            coll = self._get(collection)
            q = np.asarray(query_vectors, dtype=np.float64)
            if q.ndim != 2:
                raise ValueError("query_vectors must be a list of vectors")

            def work():
                with coll.lock:
                    return coll.search(q, limit, filters or None)
            return await asyncio.to_thread(work)
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to search {collection}", cause=e)

    async def search_and_rank(self, collection: str, items: Sequence[tuple], limit: int = 10,
                              filters: dict[str, Any] | None = None, ranker=None, summaries: bool = False,
                              summaries_filters: dict[str, Any] | None = None):
        """Additive API: vector search + hybrid ranking in one device pass.

        ``items``: one ``(plan, graph_context, query_vector, centrality_scores)`` per query.  Equivalent to
        ``ranker.rank_results(plan, graph_context, <VectorSearcher-shaped results of search(query_vector, limit, filters)>,
        centrality_scores)`` for every item (query/engine.py:176-181), but the top-k hits never leave the GPU between the
        search and the blend.  With ``summaries=True`` the queries whose intent is one of the five that
        ``QueryEngine._execute_vector_search`` (query/engine.py:331-344) extends also get the ``limit // 2`` best ``summaries``
        hits appended behind their code hits (``summaries_filters``: the engine passes the project only).  Needs
        ``rank_attrs=True``.  Returns ``list[list[RankedResult]]``."""
        try:
            coll = self._get(collection)
            if coll.rank_kind is None:
                raise ValueError("search_and_rank needs a store created with rank_attrs=True")
            coll2 = self._get(CollectionName.SUMMARIES.value) if summaries and collection != CollectionName.SUMMARIES.value else None
            from .ranking import HybridRanker
            rk = ranker or HybridRanker()

            def work():
                if coll2 is None:
                    with coll.lock:
                        return rk.rank_batch_fused(coll, items, limit, filters or None)
                with coll.lock, coll2.lock:
                    return rk.rank_batch_fused(coll, items, limit, filters or None,
                                               summaries=(coll2, limit // 2, summaries_filters or None))
            return await asyncio.to_thread(work)
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to search and rank in {collection}", cause=e)

    async def delete(self, collection: str, filters: dict[str, Any]) -> None:
        try:
            coll = self._get(collection)

            def work():
                with coll.lock:
                    return coll.delete(filters)
            await asyncio.to_thread(work)
            logger.debug(f"Deleted vectors from {collection} with filters: {filters}")
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to delete from {collection}", cause=e)

    async def file_needs_update(self, collection: str, file_path: str, content_hash: str) -> bool:
        try:
            coll = self._get(collection)

            def work():
                with coll.lock:
                    return coll.scroll({"file_path": file_path}, 1)
            points = await asyncio.to_thread(work)
            if not points:
                return True
            return (points[0]["payload"] or {}).get("content_hash") != content_hash
        except Exception as e:  # noqa: BLE001
            logger.warning(f"Error checking file update status: {e}")
            return True

    async def __aenter__(self):
        await self.connect()
        return self

    async def __aexit__(self, exc_type, exc_val, exc_tb):
        await self.close()


# The name reference code imports (``from lattice.embeddings.client import QdrantManager``): alias for injection.
QdrantManager = B200VectorStore
