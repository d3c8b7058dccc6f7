"""``DeviceCollection``: one row-major shard of a collection resident in B200 HBM (numpy-facing, synchronous).

Thin object wrapper over the C ABI (``include/lvs.h``).  It is what the reference would call "the Qdrant
collection": vectors, tombstones and dictionary-encoded keyword columns live on the GPU; ids and payload dicts
stay with the caller (``client.py``).  Everything numeric happens in liblattice_b200.so - there is no numpy
arithmetic on this path and no fallback when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _native as N
from .errors import NativeLibraryError

_STORAGE = {"f32": N.STORAGE_F32, "fp32": N.STORAGE_F32, "float32": N.STORAGE_F32,
            "bf16": N.STORAGE_BF16, "bfloat16": N.STORAGE_BF16}
_METRIC = {"cosine": N.METRIC_COSINE, "dot": N.METRIC_DOT}


class SearchResult:
    """Top-k lists of a batch of queries: ``scores`` float64 [Q, k], ``rows`` int64 [Q, k] (global rows, -1 padded), ``ties`` uint64
    [Q, k], ``counts`` uint32 [Q], ``flags`` int32 [Q] (bit0 = exactness not proven).  Built either from the five arrays or from
    the one block the library fills (scores | rows | ties | counts | flags), which is only taken apart when a field is read -
    a single-query search is tens of microseconds, numpy views are one each."""
    __slots__ = ("_f", "_block", "_Q", "_k")
    _NAMES = ("scores", "rows", "ties", "counts", "flags")

    def __init__(self, scores=None, rows=None, ties=None, counts=None, flags=None):
        self._f = [scores, rows, ties, counts, flags]
        self._block = None
        self._Q = self._k = 0

    @classmethod
    def from_block(cls, block, Q: int, k: int) -> "SearchResult":
        r = cls.__new__(cls)
        r._f = None
        r._block, r._Q, r._k = block, Q, k
        return r

    def _fields(self):
        if self._f is None:
            Q, k = self._Q, self._k
            n = Q * k
            a = np.frombuffer(self._block, dtype=np.int64)
            tail = a[3 * n:].view(np.uint32)
            self._f = [a[:n].view(np.float64).reshape(Q, k), a[n:2 * n].reshape(Q, k), a[2 * n:3 * n].view(np.uint64).reshape(Q, k),
                       tail[:Q], tail[Q:2 * Q].view(np.int32)]
        return self._f

    def single(self) -> tuple[int, int, list, list]:
        """(count, flags, rows, scores) of a ONE-query result as plain Python values - no numpy views: behind a small collection
        the whole search is tens of microseconds and five array views are a third of that."""
        if self._f is None and self._Q == 1:
            k = self._k
            st = _UNPACK1.get(k)
            if st is None:
                st = _UNPACK1[k] = struct.Struct(f"<{k}d{k}q{k}QIi")       # scores | rows | ties | count | flags
            v = st.unpack_from(self._block)
            return v[3 * k], v[3 * k + 1], list(v[k:2 * k]), list(v[:k])
        f = self._fields()
        return int(f[3][0]), int(f[4][0]), f[1][0].tolist(), f[0][0].tolist()

    scores = property(lambda self: self._fields()[0], lambda self, v: self._fields().__setitem__(0, v))
    rows = property(lambda self: self._fields()[1], lambda self, v: self._fields().__setitem__(1, v))
    ties = property(lambda self: self._fields()[2], lambda self, v: self._fields().__setitem__(2, v))
    counts = property(lambda self: self._fields()[3], lambda self, v: self._fields().__setitem__(3, v))
    flags = property(lambda self: self._fields()[4], lambda self, v: self._fields().__setitem__(4, v))

    def __repr__(self) -> str:
        return "SearchResult(" + ", ".join(f"{n}={v!r}" for n, v in zip(self._NAMES, self._fields())) + ")"


_UNPACK1: dict[int, struct.Struct] = {}


def _result_block(Q: int, k: int):
    """(ctypes block, the five addresses lvs_search / lvs_search_wait write to)."""
    n = Q * k
    block = (C.c_int64 * (3 * n + Q))()
    base = C.addressof(block)
    return block, (base, base + 8 * n, base + 16 * n, base + 24 * n, base + 24 * n + 4 * Q)


def _addr(a: np.ndarray) -> int:
    """Address of a contiguous array's data (the cheap way when the array is writable)."""
    try:
        return C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError):
        return a.ctypes.data


def _np_dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return N.DT_F32
    if a.dtype == np.float64:
        return N.DT_F64
    raise TypeError(f"vectors must be float32 or float64, got {a.dtype}")


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceCollection:
    def __init__(self, name: str, dim: int, storage: str = "f32", metric: str = "cosine", n_filter_cols: int = 0,
                 capacity: int = 0, row_base: int = 0, device: int = 0, timing: bool = True):
        """`timing`: record CUDA events around every scan launch (``last_timing`` / ``scan_times``).  It serialises consecutive
        searches on a stream, so throughput paths (the adapter, the sharded searcher, ``bench.py``'s timed legs) switch it off."""
        if storage not in _STORAGE:
            raise ValueError(f"storage must be one of {sorted(_STORAGE)}")
        if metric not in _METRIC:
            raise ValueError(f"metric must be one of {sorted(_METRIC)}")
        N.init(device)
        self._lib = N.load()
        self.name = name
        self.dim = int(dim)
        self.storage = "bf16" if _STORAGE[storage] == N.STORAGE_BF16 else "f32"
        self.metric = metric
        self.n_filter_cols = int(n_filter_cols)
        self.row_base = int(row_base)
        h = C.c_void_p()
        N.check(self._lib.lvs_collection_create(name.encode(), self.dim, _STORAGE[storage], _METRIC[metric],
                                                self.n_filter_cols, int(capacity), self.row_base, C.byref(h)),
                "lvs_collection_create")
        self._h = h
        if timing:
            self.set_option("timing", 1)

    # ---- snapshots (include/lvs.h: lvs_snapshot_save / lvs_snapshot_load) -----------------------------------
    def save_snapshot(self, path: str) -> None:
        N.check(self._lib.lvs_snapshot_save(self._handle(), str(path).encode()), "lvs_snapshot_save")

    @classmethod
    def load_snapshot(cls, path: str, name: str | None = None, capacity: int = 0, device: int = 0) -> "DeviceCollection":
        """A shard restored from ``save_snapshot``: same rows, tombstones, codes, search counter and replay state."""
        import struct
        with open(path, "rb") as f:
            head = f.read(24 + 16)
        if head[:8] != b"LVSSNAP2":
            raise ValueError(f"{path} is not a lattice-b200 snapshot")
        dim, storage, metric, n_cols = struct.unpack_from("<4i", head, 8)
        _, row_base = struct.unpack_from("<2q", head, 24)
        N.init(device)
        self = cls.__new__(cls)
        self._lib = N.load()
        self.name = name or "restored"
        self.dim, self.storage = int(dim), ("bf16" if storage == N.STORAGE_BF16 else "f32")
        self.metric = "dot" if metric == N.METRIC_DOT else "cosine"
        self.n_filter_cols, self.row_base = int(n_cols), int(row_base)
        h = C.c_void_p()
        N.check(self._lib.lvs_snapshot_load(str(path).encode(), self.name.encode(), int(capacity), C.byref(h)), "lvs_snapshot_load")
        self._h = h
        return self

    # ---- lifecycle ------------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.lvs_collection_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _handle(self):
        if not self._h:
            raise NativeLibraryError(f"collection {self.name!r} is closed")
        return self._h

    # ---- introspection --------------------------------------------------------------------------------
    @property
    def rows(self) -> int:
        return int(self._lib.lvs_rows(self._handle()))

    @property
    def capacity(self) -> int:
        return int(self._lib.lvs_capacity(self._handle()))

    @property
    def search_counter(self) -> int:
        return int(self._lib.lvs_search_counter(self._handle()))

    def advance_search_counter(self, n: int) -> None:
        """Account `n` reference searches this shard did not execute itself (include/lvs.h)."""
        N.check(self._lib.lvs_advance_search_counter(self._handle(), int(n)), "lvs_advance_search_counter")

    def count(self) -> int:
        n = int(self._lib.lvs_count(self._handle()))
        if n < 0:
            raise NativeLibraryError(f"lvs_count failed: {N.last_error()}")
        return n

    def reserve(self, capacity: int) -> None:
        N.check(self._lib.lvs_collection_reserve(self._handle(), int(capacity)), "lvs_collection_reserve")

    def set_option(self, name: str, value: int) -> None:
        N.check(self._lib.lvs_set_option(self._handle(), name.encode(), int(value)), "lvs_set_option")

    def last_timing(self) -> dict:
        ms = (C.c_float * 4)()
        nl, kind = C.c_int(), C.c_int()
        N.check(self._lib.lvs_last_search_timing(self._handle(), ms, C.byref(nl), C.byref(kind)), "lvs_last_search_timing")
        return {"prep_ms": ms[0], "scan_ms": ms[1], "finalize_ms": ms[2], "total_ms": ms[3],
                "launches": nl.value, "kernel": {0: "none", 1: "scan", 2: "gemm"}.get(kind.value, "?")}

    # ---- writes ---------------------------------------------------------------------------------------
    def upsert(self, vectors: np.ndarray, rows: np.ndarray | None = None, codes: np.ndarray | None = None,
               ties: np.ndarray | None = None) -> None:
        v = np.ascontiguousarray(vectors)
        if v.ndim != 2 or v.shape[1] != self.dim:
            raise ValueError(f"vectors must be [n, {self.dim}], got {v.shape}")
        n = v.shape[0]
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        c = None if codes is None or self.n_filter_cols == 0 else np.ascontiguousarray(codes, dtype=np.uint32)
        t = None if ties is None else np.ascontiguousarray(ties, dtype=np.uint64)
        if r is not None and r.shape != (n,):
            raise ValueError("rows must be [n]")
        if c is not None and c.shape != (n, self.n_filter_cols):
            raise ValueError(f"codes must be [n, {self.n_filter_cols}]")
        if t is not None and t.shape != (n,):
            raise ValueError("ties must be [n]")
        N.check(self._lib.lvs_upsert(self._handle(), _ptr(v), _np_dtype_code(v), n, _ptr(r), _ptr(c), _ptr(t)), "lvs_upsert")

    def upsert_device(self, data_ptr: int, dtype: str, n: int, row0: int, codes_ptr: int = 0, ties_ptr: int = 0,
                      stream: int = 0) -> None:
        """Vectors already on this GPU (e.g. a torch tensor's ``data_ptr()``): ``n`` rows of ``dim`` elements."""
        code = {"f32": N.DT_F32, "f64": N.DT_F64, "bf16": N.DT_BF16}[dtype]
        N.check(self._lib.lvs_upsert_device(self._handle(), C.c_void_p(data_ptr), code, int(n), int(row0),
                                            C.c_void_p(codes_ptr or None), C.c_void_p(ties_ptr or None),
                                            C.c_void_p(stream or None)), "lvs_upsert_device")

    def set_codes(self, col: int, codes: np.ndarray, rows: np.ndarray | None = None, row0: int = 0) -> None:
        c = np.ascontiguousarray(codes, dtype=np.uint32)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        N.check(self._lib.lvs_set_codes(self._handle(), int(col), _ptr(r), int(row0), c.shape[0], _ptr(c)), "lvs_set_codes")

    def delete_rows(self, rows: np.ndarray) -> int:
        r = np.ascontiguousarray(rows, dtype=np.int64)
        nd = C.c_int64()
        N.check(self._lib.lvs_delete_rows(self._handle(), _ptr(r), r.shape[0], C.byref(nd)), "lvs_delete_rows")
        return nd.value

    def move_rows(self, src: np.ndarray, dst: np.ndarray) -> None:
        """Compaction step: row src[i] replaces row dst[i] (all of its device state), src[i] becomes a tombstone."""
        s_ = np.ascontiguousarray(src, dtype=np.int64)
        d_ = np.ascontiguousarray(dst, dtype=np.int64)
        if s_.shape != d_.shape:
            raise ValueError("src and dst must have the same length")
        N.check(self._lib.lvs_move_rows(self._handle(), _ptr(s_), _ptr(d_), s_.shape[0]), "lvs_move_rows")

    def truncate(self, n_rows: int) -> None:
        N.check(self._lib.lvs_truncate(self._handle(), int(n_rows)), "lvs_truncate")

    def _want(self, want) -> np.ndarray | None:
        if want is None or self.n_filter_cols == 0:
            return None
        w = np.ascontiguousarray(want, dtype=np.uint32)
        if w.shape != (self.n_filter_cols,):
            raise ValueError(f"want must have {self.n_filter_cols} codes")
        return w

    def delete_where(self, want, cap: int | None = None) -> tuple[np.ndarray, int]:
        return self._match(want, cap, self._lib.lvs_delete_where, "lvs_delete_where")

    def match_rows(self, want, cap: int | None = None) -> tuple[np.ndarray, int]:
        return self._match(want, cap, self._lib.lvs_match_rows, "lvs_match_rows")

    def _match(self, want, cap, fn, what):
        w = self._want(want)
        cap = self.rows if cap is None else int(cap)
        out = np.empty(max(cap, 1), dtype=np.int64)
        nm = C.c_int64()
        N.check(fn(self._handle(), _ptr(w), _ptr(out), cap, C.byref(nm)), what)
        return out[:min(cap, nm.value)].copy(), int(nm.value)

    # ---- search ---------------------------------------------------------------------------------------
    def _queries(self, queries) -> np.ndarray:
        q = np.ascontiguousarray(queries)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {q.shape}")
        if q.dtype != np.float64 and q.dtype != np.float32:
            q = q.astype(np.float64)
        return q

    def search(self, queries: np.ndarray, k: int, want=None) -> SearchResult:
        """NaN in a query raises ValueError (the library checks while it copies the queries into its pinned slot)."""
        q = self._queries(queries)
        Q = q.shape[0]
        k = int(k)
        block, (ps, pr, pt, pc, pf) = _result_block(Q, k)
        N.check(self._lib.lvs_search(self._handle(), _addr(q), N.DT_F64 if q.dtype == np.float64 else N.DT_F32, Q, k, _ptr(self._want(want)),
                                     ps, pr, pt, pc, pf), "lvs_search")
        return SearchResult.from_block(block, Q, k)

    # ---- fused search -> rank (include/lvs.h: lvs_rank_names_append / lvs_rank_attrs_set / lvs_search_rank) ----------
    def rank_names_append(self, names: list[bytes]) -> int:
        """Append lower-cased UTF-8 entity names to the device-side pool; returns the id of the first one."""
        lens = np.asarray([len(b) for b in names], dtype=np.uint32)
        blob = np.frombuffer(b"".join(names), dtype=np.uint8) if lens.sum() else np.zeros(0, dtype=np.uint8)
        first = C.c_uint32()
        N.check(self._lib.lvs_rank_names_append(self._handle(), _ptr(blob) if len(blob) else None, _ptr(lens), len(names),
                                                C.byref(first)), "lvs_rank_names_append")
        return int(first.value)

    def rank_attrs_set(self, rows, key_id, file_id, cent_id, name_id, content_len, flags) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        a = [np.ascontiguousarray(x, dtype=t) for x, t in ((key_id, np.uint32), (file_id, np.uint32), (cent_id, np.uint32),
                                                             (name_id, np.uint32), (content_len, np.int32), (flags, np.uint8))]
        N.check(self._lib.lvs_rank_attrs_set(self._handle(), _ptr(rows), len(rows), *[_ptr(x) for x in a]), "lvs_rank_attrs_set")

    def search_rank(self, queries: np.ndarray, k: int, want, graph: "N.RankBatch", ctx: "N.RankQueryCtx", n_graph_total: int,
                    max_per_file: int, max_total: int, entity_bonus: float, rel_bonus: float,
                    second: "DeviceCollection | None" = None, k2: int = 0, want2=None, sel2=None) -> dict:
        """lvs_search_rank2: search (this collection, and `second` for the queries sel2) + candidate build + K3 in one call."""
        q = np.ascontiguousarray(queries)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {q.shape}")
        if q.dtype not in (np.float32, np.float64):
            q = q.astype(np.float64)
        if np.isnan(q).any():
            raise ValueError("Query vector must not contain NaN")
        Q, k = q.shape[0], int(k)
        sel = np.ascontiguousarray(sel2 if sel2 is not None else [], dtype=np.int32)
        Q2 = len(sel) if second is not None and k2 > 0 else 0
        k2 = int(k2) if Q2 else 0
        out = {
            "hit_scores": np.zeros((Q, k), dtype=np.float64), "hit_rows": np.full((Q, k), -1, dtype=np.int64),
            "hit_counts": np.zeros(Q, dtype=np.uint32), "flags": np.zeros(Q, dtype=np.int32),
            "hit_scores2": np.zeros((max(Q2, 1), max(k2, 1)), dtype=np.float64), "hit_rows2": np.full((max(Q2, 1), max(k2, 1)), -1, dtype=np.int64),
            "hit_counts2": np.zeros(max(Q2, 1), dtype=np.uint32), "flags2": np.zeros(max(Q2, 1), dtype=np.int32),
            "count": np.zeros(Q, dtype=np.int32), "index": np.zeros((Q, max_total), dtype=np.int32),
            "score": np.zeros((Q, max_total), dtype=np.float64), "signals": np.zeros((Q, max_total, 7), dtype=np.float64),
            "mask": np.zeros((Q, max_total), dtype=np.uint8), "source": np.zeros((Q, max_total), dtype=np.uint8),
            "leader": np.zeros(max(n_graph_total + Q * (k + k2), 1), dtype=np.int32),
        }
        h1, h2 = N.RankHits(), N.RankHits()
        for h, sfx in ((h1, ""), (h2, "2")):
            h.scores, h.rows = out["hit_scores" + sfx].ctypes.data_as(C.c_void_p), out["hit_rows" + sfx].ctypes.data_as(C.c_void_p)
            h.counts, h.flags = out["hit_counts" + sfx].ctypes.data_as(C.c_void_p), out["flags" + sfx].ctypes.data_as(C.c_void_p)
        ms = (C.c_float * 2)()
        w = self._want(want)
        w2 = second._want(want2) if Q2 else None
        N.check(self._lib.lvs_search_rank2(self._handle(), second._handle() if Q2 else None, _ptr(q), _np_dtype_code(q), Q, k, k2,
                                           _ptr(w), _ptr(w2), _ptr(sel) if Q2 else None, Q2, C.byref(graph), C.byref(ctx),
                                           int(max_per_file), int(max_total), float(entity_bonus), float(rel_bonus),
                                           C.byref(h1), C.byref(h2) if Q2 else None,
                                           _ptr(out["count"]), _ptr(out["index"]), _ptr(out["score"]), _ptr(out["signals"]),
                                           _ptr(out["mask"]), _ptr(out["source"]), _ptr(out["leader"]), ms), "lvs_search_rank2")
        out["search_ms"], out["rank_ms"] = float(ms[0]), float(ms[1])
        out["k2"], out["Q2"] = k2, Q2
        return out

    def search_submit(self, queries: np.ndarray, k: int, want=None) -> tuple[int, int, int]:
        """Pipelined search: returns a ticket for :meth:`search_wait`; up to 4 searches may be in flight."""
        q = self._queries(queries)
        t = C.c_int()
        N.check(self._lib.lvs_search_submit(self._handle(), _addr(q), N.DT_F64 if q.dtype == np.float64 else N.DT_F32, q.shape[0], int(k),
                                            _ptr(self._want(want)), C.byref(t)), "lvs_search_submit")
        return (t.value, q.shape[0], int(k))

    def search_submit_sharded(self, ex, queries: np.ndarray, k: int, want=None) -> tuple[int, int, int]:
        """Pipelined SHARDED search (``lvs_search_submit_sharded``): every rank submits the same queries; ``search_wait`` returns the
        merged lists and merged flags.  `ex` is the rank's ``lvs_exchange`` handle."""
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {q.shape}")
        t = C.c_int()
        N.check(self._lib.lvs_search_submit_sharded(self._handle(), ex, _addr(q), N.DT_F64, q.shape[0], int(k), _ptr(self._want(want)),
                                                    C.byref(t)), "lvs_search_submit_sharded")
        return (t.value, q.shape[0], int(k))

    def search_packed(self, packed: bytes, k: int, want=None) -> SearchResult:
        """``search`` for ONE float64 query that is already a bytes object of ``dim`` doubles (``struct.pack``): nothing is converted
        or checked twice on the way to ``lvs_search`` (which refuses NaN itself)."""
        if len(packed) != 8 * self.dim:
            raise ValueError(f"query must have {self.dim} components")
        k = int(k)
        block, (ps, pr, pt, pc, pf) = _result_block(1, k)
        N.check(self._lib.lvs_search(self._handle(), packed, N.DT_F64, 1, k, _ptr(self._want(want)), ps, pr, pt, pc, pf), "lvs_search")
        return SearchResult.from_block(block, 1, k)

    def search_poll(self, ticket: tuple[int, int, int]) -> bool:
        """True once :meth:`search_wait` will not block on the device (``lvs_search_poll``)."""
        d = C.c_int()
        N.check(self._lib.lvs_search_poll(self._handle(), ticket[0], C.byref(d)), "lvs_search_poll")
        return bool(d.value)

    def search_wait(self, ticket: tuple[int, int, int]) -> SearchResult:
        t, Q, k = ticket
        block, (ps, pr, pt, pc, pf) = _result_block(Q, k)
        N.check(self._lib.lvs_search_wait(self._handle(), t, ps, pr, pt, pc, pf), "lvs_search_wait")
        return SearchResult.from_block(block, Q, k)

    def search_device(self, q_ptr: int, q_dtype: str, Q: int, k: int, want, scores_ptr: int, rows_ptr: int, ties_ptr: int,
                      counts_ptr: int, stream: int = 0) -> np.ndarray:
        """Device-pointer form (sharded path): outputs stay on the GPU; returns the host flags."""
        flags = np.zeros(Q, dtype=np.int32)
        w = self._want(want)
        code = {"f32": N.DT_F32, "f64": N.DT_F64}[q_dtype]
        N.check(self._lib.lvs_search_device(self._handle(), C.c_void_p(q_ptr), code, int(Q), int(k), _ptr(w),
                                            C.c_void_p(scores_ptr), C.c_void_p(rows_ptr), C.c_void_p(ties_ptr),
                                            C.c_void_p(counts_ptr), _ptr(flags), C.c_void_p(stream or None)),
                "lvs_search_device")
        return flags

    def search_device_at(self, search_no: int, q_ptr: int, q_dtype: str, Q: int, k: int, want, scores_ptr: int, rows_ptr: int,
                         ties_ptr: int, counts_ptr: int, flags: np.ndarray, stream: int = 0) -> None:
        """Repeat of flagged queries on the exact scan, numbered as the reference searches they repeat (``lvs_search_device_at``)."""
        w = self._want(want)
        code = {"f32": N.DT_F32, "f64": N.DT_F64}[q_dtype]
        N.check(self._lib.lvs_search_device_at(self._handle(), int(search_no), C.c_void_p(q_ptr), code, int(Q), int(k), _ptr(w),
                                               C.c_void_p(scores_ptr), C.c_void_p(rows_ptr), C.c_void_p(ties_ptr),
                                               C.c_void_p(counts_ptr), _ptr(flags), C.c_void_p(stream or None)),
                "lvs_search_device_at")

    def search_device_async(self, q_ptr: int, q_dtype: str, Q: int, k: int, want, scores_ptr: int, rows_ptr: int,
                            ties_ptr: int, counts_ptr: int, flags_ptr: int, stream: int = 0) -> None:
        """Enqueue-only search (pipelined callers): nothing is synchronised; flags land in a device buffer."""
        w = self._want(want)
        code = {"f32": N.DT_F32, "f64": N.DT_F64}[q_dtype]
        N.check(self._lib.lvs_search_device_async(self._handle(), C.c_void_p(q_ptr), code, int(Q), int(k), _ptr(w),
                                                  C.c_void_p(scores_ptr), C.c_void_p(rows_ptr), C.c_void_p(ties_ptr),
                                                  C.c_void_p(counts_ptr), C.c_void_p(flags_ptr), C.c_void_p(stream or None)),
                "lvs_search_device_async")

    def search_sharded_device_async(self, ex, q_ptr: int, q_dtype: str, Q: int, k: int, want, out_ptr: int, counts_ptr: int,
                                    flags_ptr: int, stream: int = 0) -> None:
        """This rank's part of a sharded search, enqueue only: local search + exchange over peer memory + merge
        (``lvs_search_sharded_device_async``); `ex` is the rank's ``lvs_exchange`` handle."""
        w = self._want(want)
        code = {"f32": N.DT_F32, "f64": N.DT_F64}[q_dtype]
        N.check(self._lib.lvs_search_sharded_device_async(self._handle(), ex, C.c_void_p(q_ptr), code, int(Q), int(k), _ptr(w),
                                                          C.c_void_p(out_ptr), C.c_void_p(counts_ptr), C.c_void_p(flags_ptr),
                                                          C.c_void_p(stream or None)), "lvs_search_sharded_device_async")

    def last_kernel_phases(self) -> np.ndarray:
        """Profiling aid (``set_option("dbg_times", 1)`` first): microseconds from the first CTA's start to each phase of the last
        scan-kernel launch, 15 values in the order of ``include/lvs.h`` (queries ready, shard scanned, list written, lists visible,
        candidates selected, rescoring done, result stored; then the finer stamps inside the selection and the ordering)."""
        ns = np.zeros(16, dtype=np.uint64)
        N.check(self._lib.lvs_last_kernel_phases(self._handle(), ns.ctypes.data_as(C.POINTER(C.c_uint64))), "lvs_last_kernel_phases")
        return (ns[1:].astype(np.int64) - np.int64(ns[0])) / 1e3

    def scan_times(self, max_n: int = 256) -> tuple[np.ndarray, np.ndarray]:
        """(ms, algorithmic bytes) of the last scan-kernel launches (CUDA events on the launching stream)."""
        ms = np.zeros(max_n, dtype=np.float32)
        by = np.zeros(max_n, dtype=np.float64)
        n = C.c_int()
        N.check(self._lib.lvs_scan_times(self._handle(), int(max_n), ms.ctypes.data_as(C.POINTER(C.c_float)),
                                         by.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)), "lvs_scan_times")
        return ms[:n.value], by[:n.value]

    def fetch_rows(self, rows: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((r.shape[0], self.dim), dtype=np.float32)
        N.check(self._lib.lvs_fetch_rows_f32(self._handle(), _ptr(r), r.shape[0], _ptr(out)), "lvs_fetch_rows_f32")
        return out


def merge_topk_device(scores_ptr: int, rows_ptr: int, ties_ptr: int, G: int, Q: int, k: int, out_scores_ptr: int,
                      out_rows_ptr: int, out_ties_ptr: int, out_counts_ptr: int, stream: int = 0,
                      shard_stride: int = 0) -> None:
    """K5: merge G gathered per-shard lists ([G, Q, k] device buffers) into the global top-k on this GPU."""
    lib = N.load()
    N.check(lib.lvs_merge_topk_device(C.c_void_p(scores_ptr), C.c_void_p(rows_ptr), C.c_void_p(ties_ptr), int(shard_stride), int(G), int(Q), int(k),
                                      C.c_void_p(out_scores_ptr), C.c_void_p(out_rows_ptr), C.c_void_p(out_ties_ptr),
                                      C.c_void_p(out_counts_ptr), C.c_void_p(stream or None)), "lvs_merge_topk_device")
