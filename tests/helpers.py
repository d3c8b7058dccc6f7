"""Test-only helpers: a CPU stand-in for DeviceCollection (built on the oracle) so that the adapter's HOST logic
(ids, payloads, dictionary codes, filters) can be exercised without a GPU, and the K5 merge rule in numpy."""
from __future__ import annotations

import numpy as np

from code_rag_b200.collection import SearchResult
from oracle.qdrant_local import OracleCollection

ANY = 0xFFFFFFFF


class FakeDevice:
    """Same methods as code_rag_b200.collection.DeviceCollection; arithmetic by the oracle.  TESTS ONLY."""

    def __init__(self, name, dim, storage="f32", metric="cosine", n_filter_cols=0, capacity=0, row_base=0, device=0, timing=True):
        self.name, self.dim, self.n_filter_cols = name, dim, n_filter_cols
        self.ora = OracleCollection(dim)
        self.codes = np.zeros((0, n_filter_cols), dtype=np.uint32)
        self.ties = np.zeros(0, dtype=np.uint64)
        self.closed = False

    @property
    def rows(self):
        return len(self.ora.payload)

    def count(self):
        return self.ora.count(None)

    def _ensure(self, n):
        if n > len(self.ora.payload):
            add = n - len(self.ora.payload)
            base = len(self.ora.payload)
            self.ora._grow(n)
            for i in range(add):
                self.ora.ids[base + i] = base + i
                self.ora.ids_inv.append(base + i)
                self.ora.payload.append(None)
                self.ora.deleted[base + i] = True
            self.codes = np.vstack([self.codes, np.zeros((add, self.n_filter_cols), dtype=np.uint32)])
            self.ties = np.concatenate([self.ties, np.zeros(add, dtype=np.uint64)])

    def upsert(self, vectors, rows=None, codes=None, ties=None):
        v = np.asarray(vectors, dtype=np.float64)
        n = v.shape[0]
        rows = np.arange(self.rows, self.rows + n) if rows is None else np.asarray(rows)
        self._ensure(int(rows.max()) + 1)
        for i, r in enumerate(rows.tolist()):
            nrm = np.linalg.norm(v[i])
            self.ora.vectors[r] = v[i] / nrm if nrm > 0 else v[i]
            self.ora.deleted[r] = False
            if codes is not None:
                self.codes[r] = codes[i]
            self.ties[r] = ties[i] if ties is not None else r

    def set_codes(self, col, codes, rows=None, row0=0):
        rows = np.arange(row0, row0 + len(codes)) if rows is None else np.asarray(rows)
        self.codes[rows, col] = codes

    def _mask(self, want):
        m = ~self.ora.deleted[:self.rows]
        if want is not None:
            for c, w in enumerate(np.asarray(want).tolist()):
                if w != ANY:
                    m &= self.codes[:self.rows, c] == w
        return m

    def match_rows(self, want, cap=None):
        rows = np.nonzero(self._mask(want))[0].astype(np.int64)
        n = len(rows)
        cap = n if cap is None else cap
        return rows[:cap], n

    def delete_where(self, want, cap=None):
        rows, n = self.match_rows(want, None)
        self.ora.deleted[rows] = True
        return rows[: (n if cap is None else cap)], n

    def move_rows(self, src, dst):
        for s_, d_ in zip(np.asarray(src).tolist(), np.asarray(dst).tolist()):
            self.ora.vectors[d_] = self.ora.vectors[s_]
            self.ora.deleted[d_] = self.ora.deleted[s_]
            self.ora.deleted[s_] = True
            self.codes[d_] = self.codes[s_]
            self.ties[d_] = self.ties[s_]

    def truncate(self, n):
        assert self.ora.deleted[n:self.rows].all()
        for i in range(n, self.rows):
            self.ora.ids.pop(i, None)
        del self.ora.payload[n:], self.ora.ids_inv[n:]
        self.codes, self.ties = self.codes[:n], self.ties[:n]

    @property
    def row_base(self):
        return 0

    def delete_rows(self, rows):
        rows = np.asarray(rows)
        live = ~self.ora.deleted[rows]
        self.ora.deleted[rows] = True
        return int(live.sum())

    def search(self, queries, k, want=None):
        q = np.atleast_2d(np.asarray(queries, dtype=np.float64))
        Q = q.shape[0]
        res = SearchResult(np.zeros((Q, k)), np.full((Q, k), -1, dtype=np.int64), np.zeros((Q, k), dtype=np.uint64),
                           np.zeros(Q, dtype=np.uint32), np.zeros(Q, dtype=np.int32))
        mask = self._mask(want)
        for i in range(Q):
            scores = self.ora._scores(q[i])
            cand = np.nonzero(mask)[0]
            order = cand[np.lexsort((cand, self.ties[cand], -scores[cand]))][:k]
            res.rows[i, :len(order)] = order
            res.scores[i, :len(order)] = scores[order]
            res.counts[i] = len(order)
        return res

    def search_packed(self, packed, k, want=None):
        """DeviceCollection.search_packed: one float64 query as bytes."""
        q = np.frombuffer(packed, dtype=np.float64)
        if q.shape[0] != self.dim:
            raise ValueError(f"query must have {self.dim} components")
        if np.isnan(q).any():
            raise ValueError("Query vector must not contain NaN")
        return self.search(q[None, :], k, want)

    # pipelined pair + poll (DeviceCollection.search_submit / search_poll / search_wait): the search runs at submit, the first
    # poll says "not yet" so that callers' polling loops are exercised
    def search_submit(self, queries, k, want=None):
        if len(getattr(self, "_tickets", {})) >= 4:
            raise RuntimeError("4 searches already in flight")
        if not hasattr(self, "_tickets"):
            self._tickets, self._next_ticket = {}, 0
        t = self._next_ticket
        self._next_ticket += 1
        self._tickets[t] = [self.search(queries, k, want), 0]
        return (t, 0, k)

    def search_poll(self, ticket):
        ent = self._tickets[ticket[0]]
        ent[1] += 1
        return ent[1] > 1

    def search_wait(self, ticket):
        return self._tickets.pop(ticket[0])[0]

    def close(self):
        self.closed = True

    def fetch_rows(self, rows):
        return self.ora.vectors[np.asarray(rows)].copy()

    def save_snapshot(self, path):
        import pickle
        with open(path, "wb") as f:
            pickle.dump({"name": self.name, "dim": self.dim, "n_filter_cols": self.n_filter_cols, "ora": self.ora, "codes": self.codes,
                         "ties": self.ties}, f)

    @classmethod
    def load_snapshot(cls, path, name=None, capacity=0, device=0):
        import pickle
        with open(path, "rb") as f:
            st = pickle.load(f)
        self = cls(name or st["name"], st["dim"], n_filter_cols=st["n_filter_cols"])
        self.ora, self.codes, self.ties = st["ora"], st["codes"], st["ties"]
        return self


class ExactTieDevice(FakeDevice):
    """FakeDevice whose scores do not depend on where a row sits in the matrix: identical rows get bit-identical scores, as on the
    GPU (one warp re-scores a candidate with a fixed reduction order).  The oracle's BLAS matrix-vector product does not have that
    property (rows in a tail block are accumulated differently), so exact-tie behaviour is tested against the RULE
    (score desc, id asc), not against the oracle's noise.  TESTS ONLY."""

    def search(self, queries, k, want=None):
        q = np.atleast_2d(np.asarray(queries, dtype=np.float64))
        Q = q.shape[0]
        res = SearchResult(np.zeros((Q, k)), np.full((Q, k), -1, dtype=np.int64), np.zeros((Q, k), dtype=np.uint64),
                           np.zeros(Q, dtype=np.uint32), np.zeros(Q, dtype=np.int32))
        mask = self._mask(want)
        for i in range(Q):
            self.ora._scores(q[i])                                  # the in-place re-normalisation of local mode
            nq = np.linalg.norm(q[i])
            qn = q[i] / (nq if nq != 0.0 else np.finfo(np.float64).eps)      # local mode guards a zero norm the same way
            m = self.ora.vectors[:self.rows].astype(np.float64)
            scores = np.array([np.dot(m[r], qn) for r in range(self.rows)])      # 1-D dot: the same arithmetic for every row
            cand = np.nonzero(mask)[0]
            order = cand[np.lexsort((cand, self.ties[cand], -scores[cand]))][:k]
            res.rows[i, :len(order)] = order
            res.scores[i, :len(order)] = scores[order]
            res.counts[i] = len(order)
        return res


def merge_lists_numpy(scores: np.ndarray, rows: np.ndarray, ties: np.ndarray, k: int):
    """The K5 merge rule (score desc, tie asc, row asc) over [G, Q, k] lists; rows < 0 are padding."""
    G, Q, _ = scores.shape
    o_s = np.zeros((Q, k)); o_r = np.full((Q, k), -1, dtype=np.int64); o_t = np.zeros((Q, k), dtype=np.uint64)
    for q in range(Q):
        s = scores[:, q].reshape(-1); r = rows[:, q].reshape(-1); t = ties[:, q].reshape(-1)
        keep = np.nonzero(r >= 0)[0]
        order = keep[np.lexsort((r[keep], t[keep], -s[keep]))][:k]
        o_s[q, :len(order)] = s[order]; o_r[q, :len(order)] = r[order]; o_t[q, :len(order)] = t[order]
    return o_s, o_r, o_t


class FakeShardSearcher:
    """CPU stand-in for code_rag_b200.sharded.ShardedSearcher in the sharded adapter's tests: the shard's own top-k (FakeDevice),
    the path's real exchange step over gloo (``allgather_packed``) and the K5 merge rule in numpy.  TESTS ONLY."""

    def __init__(self, shard, rank, world, group=None):
        self.shard, self.rank, self.world, self.group = shard, rank, world, group

    def search(self, queries, k, want=None):
        import torch
        from code_rag_b200.sharded import allgather_packed
        res = self.shard.search(queries, k, want)
        Q = res.rows.shape[0]
        local = torch.zeros((3, Q, k), dtype=torch.int64)
        rows = np.where(res.rows >= 0, res.rows + (self.rank << 32), -1)            # global rows, as the real shard returns them
        local[0] = torch.from_numpy(res.scores.view(np.int64).copy())
        local[1] = torch.from_numpy(rows)
        local[2] = torch.from_numpy(res.ties.view(np.int64).copy())
        gathered = torch.zeros((self.world, 3, Q, k), dtype=torch.int64)
        allgather_packed(local, gathered, self.group)
        g = gathered.numpy()
        s, r, t = merge_lists_numpy(g[:, 0].view(np.float64), g[:, 1], g[:, 2].view(np.uint64), k)
        return s, r, t, (r >= 0).sum(axis=1).astype(np.uint32), np.zeros(Q, dtype=np.int32)

    # the pipelined pair the controller's event-loop path uses (ShardedSearcher.submit / poll / wait): the search (and its
    # collective) runs at submit; the first poll says "not yet"
    def submit(self, queries, k, want=None):
        return {"res": self.search(queries, k, want), "polls": 0}

    def poll(self, handle):
        handle["polls"] += 1
        return handle["polls"] > 1

    def wait(self, handle):
        return handle["res"]

    def close(self):
        pass


def fake_embed(token_ids, hidden: int, pad_id: int = 1) -> np.ndarray:
    """A deterministic stand-in for the code encoder on CPU: the mean of fixed random token vectors over the non-pad tokens."""
    tok = np.asarray(token_ids)
    table = np.random.default_rng(77).standard_normal((1000, hidden))
    m = (tok != pad_id)[..., None]
    return (table[tok % 1000] * m).sum(1) / np.maximum(m.sum(1), 1)


class FakeEncoder:
    """``embedding.B200CodeEncoder`` as far as the sharded adapter uses it (``hidden``, ``embed_upsert``, ``close``).  TESTS ONLY."""

    def __init__(self, spec):
        self.hidden = int(spec["random"]["hidden"])
        self.closed = False

    def embed_upsert(self, dev, token_ids, rows=None, codes=None, ties=None):
        dev.upsert(fake_embed(token_ids, self.hidden), rows=rows, codes=codes, ties=ties)

    def close(self):
        self.closed = True
