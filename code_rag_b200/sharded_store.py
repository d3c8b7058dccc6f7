"""``ShardedB200VectorStore``: the ``QdrantManager`` drop-in over the row shards of ALL GPUs of one box.

SURVEY section 8(e): rows are independent, so a collection is split into one row shard per GPU, new points go to the least-full
shard, every search scans all shards at once and the per-shard top-k lists are merged after one exchange step
(``sharded.ShardedSearcher``: NCCL all-gather or the peer-memory kernel, then the K5 merge).  ``B200VectorStore`` (``client.py``)
is that adapter for ONE GPU; this module puts the same call surface (reference ``embeddings/client.py:18-228``) in front of N.

Process model - one process per GPU (``torchrun``), one of them in charge:

* rank 0 is the **controller**: the lattice application runs there and talks to a ``ShardedB200VectorStore`` exactly as it would
  to a ``QdrantManager``.  Ids, payload dicts and the value dictionaries of the keyword columns live in its RAM (one
  ``_HostCollection`` per shard, sharing the dictionaries), as they do in the single-GPU adapter.
* ranks 1..N-1 are **shard workers**: after ``ShardPlane.start()`` they sit in ``plane.serve()`` and execute what the
  controller sends.  Nobody else issues commands, so the order of collectives is the controller's order - concurrent awaits on
  rank 0 (``asyncio.gather`` of graph + vector search, ``query/engine.py:142-146``) are serialised by the plane's lock.
* control plane: pickled commands over a **gloo** group (scatter to the ranks, gather of the replies); data plane: the search's
  one exchange step on the NCCL group.  Vectors travel only to the rank that owns their shard.

Global row = ``shard << 32 | local row`` (a shard is created with ``row_base = rank << 32``), so a merged hit names its owner.

A write is queued on the controller and sent with the next command (or at the end of the adapter call), one message per
rank and call.  After deletes the shards compact themselves and, when a mass delete left them uneven, are rebalanced.  Snapshots: one ``.lvs`` file per rank plus the controller's host half.  ``search_and_rank`` is there too, as one
batched search + one K3 launch (the device-resident form of it is single-GPU).
"""
from __future__ import annotations

import asyncio
import logging
import os
import threading
from typing import Any, Callable, Sequence

import numpy as np

from . import _native as N
from .client import (B200VectorStore, CollectionName, _CollectionLock, _HostCollection, _INDEX_FIELDS, _canonical_id, _id_sort_key,
                     _tie_key, vector_result_from_payload)
from .errors import VectorStoreError

logger = logging.getLogger(__name__)

SHARD_BITS = 32
LOCAL_MASK = (1 << SHARD_BITS) - 1


def split_row(global_row: int) -> tuple[int, int]:
    """(shard, local row) of a merged hit."""
    return int(global_row) >> SHARD_BITS, int(global_row) & LOCAL_MASK


def least_full(live: list[int]) -> int:
    """Shard that takes the next new point: fewest live points, lowest rank on a draw."""
    return min(range(len(live)), key=live.__getitem__)



# ------------------------------------------------------------------------------------------------------------------
# the search fast path of the control plane: a shared-memory mailbox
# ------------------------------------------------------------------------------------------------------------------
class _Mailbox:
    """Fixed-layout command block in POSIX shared memory (all ranks are processes of ONE box).

    A gloo scatter + gather of pickled objects costs hundreds of microseconds per command - more than a whole single-query step
    on 8 GPUs.  Searches therefore do not travel through gloo: the controller writes (collection, Q, k, filter codes, query
    vectors) into this block and bumps a sequence word; the workers poll that word (a short spin, then micro-sleeps) and run the
    search; each worker answers in its own status slot.  Everything else (creates, writes, deletes, snapshots) is announced here
    with op GLOO and then takes the general scatter / gather path, so there is still exactly one command stream and the
    controller's order is every rank's order.  x86 total store order makes "payload first, sequence word last" sufficient.
    """

    OP_GLOO, OP_SEARCH = 1, 2
    HDR = 256                      # seq u64 | op u32 | Q u32 | k u32 | dim u32 | has_want u32 | want[8] u32 | name_len u32 | name[64]
    STATUS = 256                   # per rank: seq u64 | ok u32 | msg_len u32 | msg[240]
    QUERY_BYTES = 4 << 20

    def __init__(self, rank: int, world: int, ctl_group):
        import torch.distributed as dist
        from multiprocessing import resource_tracker, shared_memory
        self.rank, self.world = rank, world
        size = self.HDR + world * self.STATUS + self.QUERY_BYTES
        name = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=size)
            self.shm.buf[:self.HDR + world * self.STATUS] = bytes(self.HDR + world * self.STATUS)
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=ctl_group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
            try:        # the creator owns the segment: keep this process's resource tracker from unlinking it at exit
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        buf = self.shm.buf
        self.seq = np.ndarray(1, dtype=np.uint64, buffer=buf, offset=0)
        self.u32 = np.ndarray(14, dtype=np.uint32, buffer=buf, offset=8)          # op Q k dim has_want want[8] name_len
        self.name = np.ndarray(64, dtype=np.uint8, buffer=buf, offset=64)
        self.status_seq = [np.ndarray(1, dtype=np.uint64, buffer=buf, offset=self.HDR + r * self.STATUS) for r in range(world)]
        self.status_u32 = [np.ndarray(2, dtype=np.uint32, buffer=buf, offset=self.HDR + r * self.STATUS + 8) for r in range(world)]
        self.status_msg = [np.ndarray(240, dtype=np.uint8, buffer=buf, offset=self.HDR + r * self.STATUS + 16) for r in range(world)]
        self.qoff = self.HDR + world * self.STATUS
        self.last = 0

    def fits(self, Q: int, dim: int, name: str) -> bool:
        return Q * dim * 8 <= self.QUERY_BYTES and len(name.encode()) <= 64

    # ---- controller ----------------------------------------------------------------------------------------------
    def post(self, op: int, name: str = "", queries: np.ndarray | None = None, k: int = 0, want=None) -> int:
        u = self.u32
        u[0] = op
        if op == self.OP_SEARCH:
            Q, dim = queries.shape
            u[1], u[2], u[3] = Q, k, dim
            u[4] = 0 if want is None else 1
            if want is not None:
                u[5:13] = np.asarray(want, dtype=np.uint32)
            nb = name.encode()
            u[13] = len(nb)
            self.name[:len(nb)] = np.frombuffer(nb, dtype=np.uint8)
            np.ndarray((Q, dim), dtype=np.float64, buffer=self.shm.buf, offset=self.qoff)[...] = queries
        self.last += 1
        self.seq[0] = self.last                       # published last
        return self.last

    def collect(self, seq: int, timeout_s: float = 60.0) -> list[tuple[bool, str]]:
        """Replies of ranks 1..N-1 to command `seq`."""
        import time
        out = []
        t0 = time.perf_counter()
        for r in range(1, self.world):
            spins = 0
            while int(self.status_seq[r][0]) != seq:
                spins += 1
                if spins > 2000:
                    if time.perf_counter() - t0 > timeout_s:
                        out.append((False, f"no reply within {timeout_s:.0f} s"))
                        break
                    time.sleep(20e-6)
            else:
                ok, n = int(self.status_u32[r][0]) != 0, int(self.status_u32[r][1])
                out.append((ok, bytes(self.status_msg[r][:n]).decode(errors="replace")))
        return out

    # ---- workers -------------------------------------------------------------------------------------------------
    def wait_command(self, is_closed) -> tuple[int, int]:
        """Blocks until the controller posts the next command; returns (seq, op).  Spins while commands keep coming (a search
        step on 8 GPUs is ~0.3 ms), backs off to micro-sleeps when the plane is idle."""
        import time
        spins = 0
        while True:
            seq = int(self.seq[0])
            if seq != self.last:
                self.last = seq
                return seq, int(self.u32[0])
            spins += 1
            if spins > 20000:
                if is_closed():
                    return -1, 0
                time.sleep(50e-6 if spins < 200000 else 500e-6)

    def read_search(self):
        u = self.u32
        Q, k, dim = int(u[1]), int(u[2]), int(u[3])
        want = np.array(u[5:13], dtype=np.uint32) if int(u[4]) else None
        name = bytes(self.name[:int(u[13])]).decode()
        q = np.ndarray((Q, dim), dtype=np.float64, buffer=self.shm.buf, offset=self.qoff).copy()
        return name, q, k, want

    def reply(self, seq: int, ok: bool, msg: str = "") -> None:
        mb = msg.encode()[:240]
        self.status_u32[self.rank][0] = 1 if ok else 0
        self.status_u32[self.rank][1] = len(mb)
        if mb:
            self.status_msg[self.rank][:len(mb)] = np.frombuffer(mb, dtype=np.uint8)
        self.status_seq[self.rank][0] = seq           # published last

    def close(self) -> None:
        for a in ("seq", "u32", "name", "status_seq", "status_u32", "status_msg"):
            setattr(self, a, None)
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass


# ------------------------------------------------------------------------------------------------------------------
# the plane: process group plumbing + what every rank (controller included) does with a command
# ------------------------------------------------------------------------------------------------------------------
class ShardPlane:
    """One per process.  ``start()`` is collective; afterwards rank 0 drives and the others ``serve()``."""

    def __init__(self, rank: int, world: int, ctl_group, device: int, device_factory: Callable | None,
                 searcher_factory: Callable | None, encoder_factory: Callable | None = None):
        self.rank, self.world, self.ctl, self.device = rank, world, ctl_group, device
        self._device_factory, self._searcher_factory, self._encoder_factory = device_factory, searcher_factory, encoder_factory
        self.encoder = None                       # this rank's code encoder (op "encoder"): chunks are embedded where they will live
        self.shards: dict[str, Any] = {}          # collection name -> this rank's shard (DeviceCollection)
        self.searchers: dict[str, Any] = {}
        self.lock = threading.RLock()             # controller: one command at a time
        self.pending: list[list[tuple]] = [[] for _ in range(world)]     # controller: queued writes per rank
        self.closed = False
        self._deferred: int | None = None
        self.encoder_hidden: int | None = None    # controller: width of the encoders the ranks hold (attach_encoder)
        self._polled: dict | None = None          # controller: the search begun with search_begin and not yet ended
        self.polled_searches = 0                  # ... how many took that route (tests, diagnostics)
        self.mailbox: _Mailbox | None = None
        if world > 1 and os.environ.get("LATTICE_B200_MAILBOX", "1") != "0":
            self.mailbox = _Mailbox(rank, world, ctl_group)              # collective

    @classmethod
    def start(cls, device_factory: Callable | None = None, searcher_factory: Callable | None = None,
              ctl_group=None, encoder_factory: Callable | None = None) -> "ShardPlane":
        """Collective over the default process group (``torchrun`` env; initialised here when the caller has not).  Binds the
        process to ``cuda:LOCAL_RANK`` and opens the gloo control group.  The two factories exist for the CPU tests."""
        import torch.distributed as dist
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        if device_factory is None:
            import torch
            N.init(device)                          # raises without the library or an sm_100 device: no fallback
            torch.cuda.set_device(device)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if device_factory is None:
                dist.init_process_group("nccl", device_id=torch.device("cuda", device))
            else:
                dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        if ctl_group is None:
            ctl_group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
        return cls(rank, world, ctl_group, device, device_factory, searcher_factory, encoder_factory)

    # ---- command transport ---------------------------------------------------------------------------------------
    def _scatter(self, per_rank: list | None):
        import torch.distributed as dist
        out = [None]
        dist.scatter_object_list(out, per_rank if self.rank == 0 else None, src=0, group=self.ctl)
        return out[0]

    def _gather(self, reply):
        import torch.distributed as dist
        got = [None] * self.world if self.rank == 0 else None
        dist.gather_object(reply, got, dst=0, group=self.ctl)
        return got

    def call(self, op: str, name: str | None, common=None) -> list:
        """Controller only.  Sends ``op`` (with the writes queued so far in front of it) to every rank, runs its own share and
        returns the per-rank results; a failure on any rank is raised here."""
        if self.rank != 0:
            raise RuntimeError("only rank 0 drives the shard plane")
        with self.lock:
            if self.closed:
                raise RuntimeError("shard plane is shut down")
            if op == "search" and any(self.pending):
                # the search ends in a data-plane collective that every rank must enter: writes still queued go first, in a
                # command of their own, so that a rank failing one of them is reported instead of missing the exchange
                self.call("noop", None)
            mb = self.mailbox
            if self._deferred is not None:
                # the workers' replies to the previous fast-path search (see below) are due before the mailbox is written again
                seq, self._deferred = self._deferred, None
                bad = [(r + 1, msg) for r, (ok, msg) in enumerate(mb.collect(seq)) if not ok]
                if bad and op == "shutdown":             # the workers must still be released
                    logger.warning("previous search failed on %s", "; ".join(f"rank {r}: {msg}" for r, msg in bad))
                elif bad:
                    raise RuntimeError("previous search: " + "; ".join(f"rank {r}: {msg}" for r, msg in bad))
            if op == "search" and mb is not None and mb.fits(common[0].shape[0], common[0].shape[1], name):
                # fast path: the command travels through shared memory, the replies through the workers' status slots.  The controller
                # holds the MERGED result as soon as its own shard's kernel has finished (the exchange made every rank wait for every
                # other one), so it returns without waiting for the workers' host-side epilogue; their replies are collected before
                # the next command (a worker that failed before its exchange shows up right away as an exchange timeout)
                seq = mb.post(mb.OP_SEARCH, name, common[0], int(common[1]), common[2])
                mine = self._execute(op, name, common, [])
                self._deferred = seq
                replies = [mine]
            else:
                if mb is not None:
                    mb.post(mb.OP_GLOO)                  # the workers leave the mailbox poll and enter the scatter
                writes, self.pending = self.pending, [[] for _ in range(self.world)]
                mine = self._scatter([(op, name, common, writes[r]) for r in range(self.world)])
                replies = self._gather(self._execute(*mine))
        bad = [(r, rep[1]) for r, rep in enumerate(replies) if not rep[0]]
        if bad:
            raise RuntimeError("; ".join(f"rank {r}: {msg}" for r, msg in bad))
        return [rep[1] for rep in replies]

    # ---- the same fast-path search in three steps, for the event loop (client.B200VectorStore._search_polled) ------------------
    def search_begin(self, name: str, queries: np.ndarray, k: int, want) -> dict | None:
        """Controller, non-blocking: posts the search to the workers and submits it on this rank's shard.  Returns a handle for
        ``search_ready`` / ``search_end``, or None when this search has to take ``call`` (another command under way, writes queued,
        no mailbox, a searcher without the pipelined pair).  The plane stays locked until ``search_end``: one search at a time
        across the shards (a flagged query is repeated collectively - every rank must be at the same point of the command stream)."""
        if self.rank != 0:
            raise RuntimeError("only rank 0 drives the shard plane")
        if self._polled is not None or not self.lock.acquire(blocking=False):
            return None
        try:
            mb, searcher = self.mailbox, self.searchers.get(name)
            if (self.closed or mb is None or any(self.pending) or searcher is None or not hasattr(searcher, "poll")
                    or not mb.fits(queries.shape[0], queries.shape[1], name)):
                self.lock.release()
                return None
            if self._deferred is not None:
                seq, self._deferred = self._deferred, None
                bad = [(r + 1, msg) for r, (ok, msg) in enumerate(mb.collect(seq)) if not ok]
                if bad:
                    raise RuntimeError("previous search: " + "; ".join(f"rank {r}: {msg}" for r, msg in bad))
            if self._device_factory is None:
                import torch
                torch.cuda.set_device(self.device)
            seq = mb.post(mb.OP_SEARCH, name, queries, int(k), want)
            self._deferred = seq                         # the workers' replies are due before the next command, whatever happens below
            self._polled = {"seq": seq, "searcher": searcher, "handle": searcher.submit(queries, int(k), want)}
            self.polled_searches += 1
            return self._polled
        except BaseException:
            self._polled = None
            self.lock.release()
            raise

    def search_ready(self, h: dict) -> bool:
        return h["searcher"].poll(h["handle"])

    def search_end(self, h: dict):
        """(scores, rows, counts, flags) of the merged result; blocks if the search has not finished.  Unlocks the plane."""
        try:
            hd = h["handle"]
            if isinstance(hd, dict) and "ticket" in hd and hd["ticket"][1] == 1 and hasattr(h["searcher"], "shard"):
                # one query, nothing flagged (the usual case): plain Python values, no numpy views (see SearchResult.single)
                res = h["searcher"].shard.search_wait(hd["ticket"])
                n, flag, rows, scores = res.single()
                if flag == 0:
                    return [scores], [rows], [n], [0]
                hd = dict(hd, done=res)                  # flagged: the general path repeats / raises
            scores, rows, _ties, counts, flags = h["searcher"].wait(hd)
            return scores, rows, counts, flags
        finally:
            self._polled = None
            self.lock.release()

    def queue(self, shard: int, name: str, method: str, *args) -> None:
        with self.lock:
            self.pending[shard].append((name, method, args))

    def flush(self) -> None:
        with self.lock:
            if any(self.pending):
                self.call("noop", None)

    def serve(self) -> None:
        """Ranks 1..N-1: execute the controller's commands until it shuts the plane down."""
        if self.rank == 0:
            raise RuntimeError("rank 0 is the controller")
        mb = self.mailbox
        while not self.closed:
            if mb is not None:
                seq, op = mb.wait_command(lambda: self.closed)
                if seq < 0:
                    break
                if op == mb.OP_SEARCH:
                    name, q, k, want = mb.read_search()
                    ok, res = self._execute("search", name, (q, k, want), [])
                    mb.reply(seq, ok, "" if ok else str(res))
                    continue
            cmd = self._scatter(None)
            self._gather(self._execute(*cmd))
        if mb is not None:
            mb.close()

    def shutdown(self) -> None:
        """Controller: release the workers (collective with their ``serve()`` loops)."""
        if self.rank == 0 and not self.closed:
            self.call("shutdown", None)
            if self.mailbox is not None:
                self.mailbox.close()
                self.mailbox = None

    # ---- what a rank does with a command -------------------------------------------------------------------------
    def _execute(self, op: str, name: str | None, common, writes: list) -> tuple[bool, Any]:
        """Returns (ok, result or error text).  Data-plane collectives (the search's exchange) are entered by every rank or by
        none: arguments are validated on the controller before anything is sent."""
        try:
            if self._device_factory is None:          # commands may arrive on another thread (asyncio.to_thread on rank 0)
                import torch
                torch.cuda.set_device(self.device)
            for w_name, method, args in writes:
                self._write(self.shards[w_name], method, args)
            return True, getattr(self, "_op_" + op)(name, common)
        except Exception as e:  # noqa: BLE001
            logger.exception("shard rank %d: %s(%s) failed", self.rank, op, name)
            return False, f"{type(e).__name__}: {e}"

    def _write(self, dev, method: str, args: tuple) -> None:
        base = int(getattr(dev, "row_base", 0))          # the wire carries local rows
        if method == "upsert_tokens":
            # SURVEY section 8f row 4 over N GPUs: the token ids came over the control plane, the vectors never exist outside this GPU
            tok, rows, codes, ties = args
            if self.encoder is None:
                raise RuntimeError("no encoder attached to the shard plane (ShardedB200VectorStore.attach_encoder)")
            self.encoder.embed_upsert(dev, tok, rows=rows + base, codes=codes, ties=ties)
        elif method == "upsert":
            vec, rows, codes, ties = args
            dev.upsert(vec, rows=rows + base, codes=codes, ties=ties)
        elif method == "set_codes":
            col, codes, rows, row0 = args
            dev.set_codes(col, codes, rows=None if rows is None else rows + base, row0=row0 + base)
        elif method == "delete_rows":
            dev.delete_rows(args[0] + base)
        elif method == "move_rows":
            dev.move_rows(args[0] + base, args[1] + base)
        elif method == "truncate":
            dev.truncate(args[0])
        else:
            raise ValueError(f"unknown shard write {method!r}")

    def _op_noop(self, name, common):
        return None

    def _op_encoder(self, name, spec):
        """Every rank builds the same code encoder on its own GPU.  spec: {"pretrained": path-or-name} (transformers'
        ``RobertaModel`` checkpoint, as unixcoder_provider.py:65-83 loads it), or {"random": {vocab, hidden, layers, intermediate,
        max_pos, seed}} (benchmarks, tests), plus n_heads / pad_id where they are not the checkpoint's.  None detaches."""
        if self.encoder is not None:
            close = getattr(self.encoder, "close", None)
            if close:
                close()
            self.encoder = None
        if spec is None:
            return None
        if self._encoder_factory is not None:
            self.encoder = self._encoder_factory(spec)
        else:
            from .embedding import B200CodeEncoder, random_state_dict
            if "pretrained" in spec:
                self.encoder = B200CodeEncoder.from_pretrained(spec["pretrained"], device=self.device)
            else:
                r = spec["random"]
                sd = random_state_dict(r["vocab"], r["hidden"], r["layers"], r["intermediate"], r["max_pos"], seed=r.get("seed", 0))
                self.encoder = B200CodeEncoder(sd, n_layers=r["layers"], n_heads=spec.get("n_heads", r["hidden"] // 64), pad_id=spec.get("pad_id", 1),
                                               device=self.device)
        return int(self.encoder.hidden)

    def _op_shutdown(self, name, common):
        self._op_encoder(None, None)
        for nm in list(self.shards):
            self._op_destroy(nm, None)
        self.closed = True
        return None

    def _op_create(self, name, common):
        dim, storage, n_cols = common
        if name in self.shards:
            self._op_destroy(name, None)
        if self._device_factory is None:
            from .collection import DeviceCollection as factory
        else:
            factory = self._device_factory
        shard = factory(name, dim, storage=storage, metric="cosine", n_filter_cols=n_cols, capacity=0,
                        row_base=self.rank << SHARD_BITS, device=self.device, timing=False)
        self.shards[name] = shard
        self._attach_searcher(name, shard)
        return None

    def _attach_searcher(self, name, shard) -> None:
        if self._searcher_factory is None:
            from .sharded import ShardedSearcher
            self.searchers[name] = ShardedSearcher(shard)
        else:
            self.searchers[name] = self._searcher_factory(shard, self.rank, self.world)

    def _op_destroy(self, name, common):
        s = self.searchers.pop(name, None)
        if s is not None and hasattr(s, "close"):
            s.close()
        d = self.shards.pop(name, None)
        if d is not None:
            d.close()
        return None

    def _op_search(self, name, common):
        queries, k, want = common
        scores, rows, ties, counts, flags = self.searchers[name].search(queries, k, want)
        # every rank holds the merged answer; only the controller needs it
        return (scores, rows, counts, flags) if self.rank == 0 else None

    def _local_rows(self, dev, found) -> tuple[np.ndarray, int]:
        rows, n = found
        return np.asarray(rows, dtype=np.int64) - int(getattr(dev, "row_base", 0)), int(n)

    def _op_match(self, name, common):
        want, cap = common
        dev = self.shards[name]
        return self._local_rows(dev, dev.match_rows(want, cap))

    def _op_delete_where(self, name, common):
        dev = self.shards[name]
        return self._local_rows(dev, dev.delete_where(common))

    def _op_count(self, name, common):
        return int(self.shards[name].count())

    def _op_fetch(self, name, common):
        """common: {rank: local rows}; the named ranks return their stored vectors (float32, as stored)."""
        rows = common.get(self.rank)
        if rows is None:
            return None
        dev = self.shards[name]
        return dev.fetch_rows(np.asarray(rows, dtype=np.int64) + int(getattr(dev, "row_base", 0)))

    def _shard_file(self, directory: str, name: str) -> str:
        return os.path.join(directory, f"{name}.shard{self.rank}of{self.world}.lvs")

    def _op_snapshot_save(self, name, common):
        os.makedirs(common, exist_ok=True)
        self.shards[name].save_snapshot(self._shard_file(common, name))
        return int(self.shards[name].rows)

    def _op_snapshot_load(self, name, common):
        if self._device_factory is None:
            from .collection import DeviceCollection as factory
        else:
            factory = self._device_factory
        shard = factory.load_snapshot(self._shard_file(common, name), name=name, device=self.device)
        if int(getattr(shard, "row_base", 0)) not in (0, self.rank << SHARD_BITS):
            shard.close()
            raise ValueError(f"{self._shard_file(common, name)} holds rows of another shard")
        self._op_destroy(name, None)
        self.shards[name] = shard
        self._attach_searcher(name, shard)
        return int(shard.rows)


# ------------------------------------------------------------------------------------------------------------------
# controller-side host state
# ------------------------------------------------------------------------------------------------------------------
class _ShardProxy:
    """What a shard's ``_HostCollection`` sees as its device on the controller: local rows, writes queued for the owner."""

    row_base = 0

    def __init__(self, plane: ShardPlane, name: str, shard: int):
        self._plane, self._name, self._shard = plane, name, shard

    def upsert(self, vectors, rows=None, codes=None, ties=None) -> None:
        self._plane.queue(self._shard, self._name, "upsert", np.ascontiguousarray(vectors), np.asarray(rows, dtype=np.int64),
                          None if codes is None else np.ascontiguousarray(codes, dtype=np.uint32),
                          None if ties is None else np.ascontiguousarray(ties, dtype=np.uint64))

    def upsert_tokens(self, token_ids, rows=None, codes=None, ties=None) -> None:
        self._plane.queue(self._shard, self._name, "upsert_tokens", np.ascontiguousarray(token_ids, dtype=np.int32),
                          np.asarray(rows, dtype=np.int64), None if codes is None else np.ascontiguousarray(codes, dtype=np.uint32),
                          None if ties is None else np.ascontiguousarray(ties, dtype=np.uint64))

    def set_codes(self, col, codes, rows=None, row0=0) -> None:
        self._plane.queue(self._shard, self._name, "set_codes", int(col), np.ascontiguousarray(codes, dtype=np.uint32),
                          None if rows is None else np.asarray(rows, dtype=np.int64), int(row0))

    def delete_rows(self, rows) -> None:
        self._plane.queue(self._shard, self._name, "delete_rows", np.asarray(rows, dtype=np.int64))

    def move_rows(self, src, dst) -> None:
        self._plane.queue(self._shard, self._name, "move_rows", np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64))

    def truncate(self, n_rows) -> None:
        self._plane.queue(self._shard, self._name, "truncate", int(n_rows))

    def close(self) -> None:
        pass


class _PlaneEncoder:
    """What ``_HostCollection.upsert_tokens`` takes for an encoder on the controller: the embedding itself happens on the rank that
    owns the shard (``ShardPlane._write``: "upsert_tokens"), so this only forwards the token ids to the shard's proxy."""

    def __init__(self, hidden: int):
        self.hidden = int(hidden)

    def embed_upsert(self, dev, token_ids, rows=None, codes=None, ties=None) -> None:
        dev.upsert_tokens(token_ids, rows=rows, codes=codes, ties=ties)


class _HostShard(_HostCollection):
    """Bookkeeping of one shard.  Keyword columns and their dictionaries belong to the collection, not to the shard."""

    def __init__(self, owner: "_ShardedHostCollection", shard: int):
        super().__init__(owner.name, owner.dim, owner.storage, owner.columns, owner.plane.device,
                         dev_factory=lambda *a, **kw: _ShardProxy(owner.plane, owner.name, shard))
        self.columns, self.dicts = owner.columns, owner.dicts          # shared objects
        self.tie_counts, self.dup_keys = owner.tie_counts, owner.dup_keys     # ids meet in the merged list whatever their shard
        self.owner = owner

    def ensure_column(self, key: str) -> int:
        return self.owner.ensure_column(key)


def _whole_op(fn):
    """An adapter-level operation owns the plane from its first queued write to its last command, so that its writes (and
    their failures) are not carried by another collection's command."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        with self.plane.lock:
            return fn(self, *args, **kwargs)
    return wrapped


class _ShardedHostCollection:
    """Same operations as ``client._HostCollection`` (what ``B200VectorStore`` calls under ``.lock``), over N shards."""

    rank_kind = None          # fused search -> rank is a single-GPU feature

    def __init__(self, plane: ShardPlane, name: str, dim: int, storage: str, index_fields: Sequence[str]):
        self.plane, self.name, self.dim, self.storage = plane, name, dim, storage
        self.columns: list[str] = list(index_fields)[: N.MAX_FILTER_COLS]
        self.dicts: list[dict[Any, int]] = [dict() for _ in self.columns]
        self.tie_counts: dict[int, int] = {}
        self.dup_keys: dict[int, int] = {}
        self.lock = _CollectionLock()
        plane.call("create", name, (dim, storage, N.MAX_FILTER_COLS))
        self.shards = [_HostShard(self, s) for s in range(plane.world)]

    # -- snapshots: one raw ``.lvs`` per rank (written and read by its owner) + the controller's host half ---------------
    _SHARD_STATE = ("ids", "id_to_row", "payloads", "free_rows")

    @_whole_op
    def save(self, directory: str) -> None:
        import pickle
        self.plane.flush()
        rows = self.plane.call("snapshot_save", self.name, str(directory))
        if rows != [len(sh.ids) for sh in self.shards]:
            raise RuntimeError(f"snapshot of {self.name}: device rows {rows} but host rows {[len(sh.ids) for sh in self.shards]}")
        state = {"name": self.name, "dim": self.dim, "storage": self.storage, "world": self.plane.world, "columns": self.columns,
                 "dicts": self.dicts, "shards": [{k: getattr(sh, k) for k in self._SHARD_STATE} for sh in self.shards]}
        with open(os.path.join(directory, f"{self.name}.host{self.plane.world}.pkl"), "wb") as f:
            pickle.dump(state, f, protocol=pickle.HIGHEST_PROTOCOL)

    @classmethod
    def load(cls, plane: ShardPlane, directory: str, name: str) -> "_ShardedHostCollection":
        """The host half is a pickle: load only snapshots this application wrote itself (as with the single-GPU store)."""
        import pickle
        with open(os.path.join(directory, f"{name}.host{plane.world}.pkl"), "rb") as f:
            state = pickle.load(f)
        if state["world"] != plane.world:
            raise ValueError(f"snapshot of {name} was written by {state['world']} shard(s), this job has {plane.world}")
        self = cls.__new__(cls)
        self.plane, self.name, self.dim, self.storage = plane, name, state["dim"], state["storage"]
        self.columns, self.dicts = state["columns"], state["dicts"]
        self.tie_counts, self.dup_keys = {}, {}
        self.lock = _CollectionLock()
        self.shards = [_HostShard(self, s) for s in range(plane.world)]
        for sh, st in zip(self.shards, state["shards"]):
            for k in cls._SHARD_STATE:
                setattr(sh, k, st[k])
            for pid in sh.ids:
                if pid is not None:
                    sh._tie_add(_tie_key(pid))
        with plane.lock:
            rows = plane.call("snapshot_load", name, str(directory))
        if rows != [len(sh.ids) for sh in self.shards]:
            raise ValueError(f"snapshot of {name}: device rows {rows} but host rows {[len(sh.ids) for sh in self.shards]}")
        return self

    # -- columns and filters -------------------------------------------------------------------------------------
    def ensure_column(self, key: str) -> int:
        if key in self.columns:
            return self.columns.index(key)
        if len(self.columns) >= N.MAX_FILTER_COLS:
            raise ValueError(f"cannot filter on {key!r}: all {N.MAX_FILTER_COLS} keyword columns are in use ({self.columns})")
        self.columns.append(key)
        self.dicts.append(dict())
        for sh in self.shards:
            sh.backfill_column(len(self.columns) - 1)
        return len(self.columns) - 1

    def want_codes(self, filters):
        return self.shards[0].want_codes(filters)

    # -- operations ----------------------------------------------------------------------------------------------
    def live_counts(self) -> list[int]:
        return [len(sh.ids) - len(sh.free_rows) for sh in self.shards]

    def shard_of(self, pid) -> int | None:
        for s, sh in enumerate(self.shards):
            if pid in sh.id_to_row:
                return s
        return None

    @_whole_op
    def upsert(self, ids, vectors, payloads) -> None:
        n = min(len(ids), len(vectors), len(payloads))           # the reference zips the three lists (client.py:123-126)
        if n == 0:
            return
        canon = [_canonical_id(i) for i in ids[:n]]
        vec = np.asarray(vectors[:n], dtype=np.float64)
        if vec.ndim != 2 or vec.shape[1] != self.dim:
            raise ValueError(f"vectors must have dimension {self.dim}, got shape {vec.shape}")
        if not np.isfinite(vec).all():
            raise ValueError("vectors must be finite")
        last = {pid: i for i, pid in enumerate(canon)}           # a repeated id: the last occurrence wins
        live = self.live_counts()
        parts: list[list[int]] = [[] for _ in self.shards]
        for i in sorted(last.values()):
            s = self.shard_of(canon[i])
            if s is None:                                        # a new point goes to the least-full shard
                s = least_full(live)
                live[s] += 1
            parts[s].append(i)
        for s, idx in enumerate(parts):
            if idx:
                self.shards[s].upsert([canon[i] for i in idx], vec[idx], [payloads[i] for i in idx])
        self.plane.flush()

    @_whole_op
    def upsert_tokens(self, ids, token_ids, payloads, encoder=None) -> None:
        """``upsert`` from token ids: every point's tokens travel to the rank that owns (or will own) it, which embeds them with its
        own encoder and writes the vectors into its shard in place (``attach_encoder`` first).  `encoder` is ignored: a vector
        computed on one GPU would have to cross the control plane to reach another."""
        hidden = getattr(self.plane, "encoder_hidden", None)
        if hidden is None:
            raise ValueError("no encoder attached: call ShardedB200VectorStore.attach_encoder(...) first")
        if hidden != self.dim:
            raise ValueError(f"the encoder produces {hidden}-dimensional vectors, the collection holds {self.dim}")
        tok = np.ascontiguousarray(token_ids, dtype=np.int32)
        if tok.ndim != 2:
            raise ValueError("token_ids must be [n, L]")
        n = min(len(ids), tok.shape[0], len(payloads))
        if n == 0:
            return
        canon = [_canonical_id(i) for i in ids[:n]]
        last = {pid: i for i, pid in enumerate(canon)}           # a repeated id: the last occurrence wins
        live = self.live_counts()
        parts: list[list[int]] = [[] for _ in self.shards]
        for i in sorted(last.values()):
            s = self.shard_of(canon[i])
            if s is None:
                s = least_full(live)
                live[s] += 1
            parts[s].append(i)
        enc = _PlaneEncoder(hidden)
        for s, idx in enumerate(parts):
            if idx:
                self.shards[s].upsert_tokens([canon[i] for i in idx], tok[idx], [payloads[i] for i in idx], enc)
        self.plane.flush()

    def _id_of(self, global_row: int):
        s, r = split_row(global_row)
        return self.shards[s].ids[r]

    def _hits(self, rows, scores) -> list[dict[str, Any]]:
        if not isinstance(rows, list):
            rows, scores = rows.tolist(), scores.tolist()
        shards = self.shards
        out = []
        for g, sc in zip(rows, scores):
            if g < 0:
                continue
            sh = shards[g >> SHARD_BITS]
            r = g & LOCAL_MASK
            p = sh.payloads[r]
            out.append({"id": str(sh.ids[r]), "score": float(sc), "payload": dict(p) if p is not None else None})
        return out

    def _match(self, want, cap=None) -> tuple[list[list[int]], int]:
        found = self.plane.call("match", self.name, (want, cap))
        return [rows.tolist() for rows, _ in found], sum(n for _, n in found)

    @_whole_op
    def search(self, query_vectors, limit: int, filters) -> list[list[dict[str, Any]]]:
        want = self.want_codes(filters)
        if limit <= 0:
            return [[] for _ in range(1 if query_vectors is None else len(query_vectors))]
        if query_vectors is None:
            # filter-only lookup (query/context/builder.py:111-119): matching points in ascending id order, score 0.0
            per_shard, _ = self._match(want)
            found = [(self.shards[s].ids[r], s, r) for s, rows in enumerate(per_shard) for r in rows]
            found.sort(key=lambda t: _id_sort_key(t[0]))
            return [[h for _, s, r in found[:limit] for h in self.shards[s]._hits(np.asarray([r]), np.zeros(1))]]
        if limit > N.MAX_K:
            raise ValueError(f"limit {limit} exceeds the largest supported top-k ({N.MAX_K})")
        q = np.ascontiguousarray(query_vectors, dtype=np.float64)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {q.shape}")
        if np.isnan(q).any():
            raise ValueError("Query vector must not contain NaN")
        if q.shape[0] == 0:
            return []
        k_dev = self.shards[0].device_limit(int(limit))
        scores, rows, counts, flags = self.plane.call("search", self.name, (q, k_dev, want))[0]
        return self._shape(scores, rows, counts, flags, limit)

    def _shape(self, scores, rows, counts, flags, limit: int) -> list[list[dict[str, Any]]]:
        out = []
        if isinstance(rows, list):                       # the one-query fast path of ShardPlane.search_end: lists of lists
            for qi in range(len(rows)):
                n = counts[qi]
                r, sc = rows[qi][:n], scores[qi][:n]
                if self.dup_keys and n > 1:              # (score desc, id asc): see client._HostCollection.in_id_order
                    order = sorted(range(n), key=lambda j: (-sc[j], _id_sort_key(self._id_of(r[j]))))
                    r, sc = [r[j] for j in order], [sc[j] for j in order]
                out.append(self._hits(r[:limit], sc[:limit]))
            return out
        for qi in range(rows.shape[0]):
            n = int(counts[qi])
            if int(flags[qi]) & N.FLAG_UNPROVEN:
                logger.warning("search on %s: exactness bound not met for query %d (many near-ties)", self.name, qi)
            out.append(self._hits(*self.shards[0].in_id_order(rows[qi, :n], scores[qi, :n], limit, self._id_of)))
        return out

    # three-step form (see client._HostCollection.search_begin): begin may return None = "take a thread and call search()"
    MAX_IN_FLIGHT = 1

    def search_begin(self, query_vectors, limit: int, filters) -> dict | None:
        if query_vectors is None or limit <= 0:
            return None
        if limit > N.MAX_K:
            raise ValueError(f"limit {limit} exceeds the largest supported top-k ({N.MAX_K})")
        q = np.ascontiguousarray(query_vectors, dtype=np.float64)
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [Q, {self.dim}], got {q.shape}")
        if np.isnan(q).any():
            raise ValueError("Query vector must not contain NaN")
        if q.shape[0] == 0:
            return None
        h = self.plane.search_begin(self.name, q, self.shards[0].device_limit(int(limit)), self.want_codes(filters))
        if h is not None:
            h["limit"] = limit
        return h

    def search_ready(self, h: dict) -> bool:
        return self.plane.search_ready(h)

    def search_block(self, h: dict) -> None:
        if "res" not in h:
            h["res"] = self.plane.search_end(h)

    def search_end(self, h: dict):
        self.search_block(h)
        scores, rows, counts, flags = h["res"]
        return self._shape(scores, rows, counts, flags, h["limit"])

    def search_discard(self, h: dict) -> None:
        try:
            self.search_block(h)
        except Exception:  # noqa: BLE001
            pass

    def scroll(self, filters, limit: int):
        return self.search(None, limit, filters)[0]

    def _after_delete(self, per_shard: list[list[int]]) -> None:
        for sh, rows in zip(self.shards, per_shard):
            if rows:
                sh.release_rows(rows)
                sh.maybe_compact()
        self.plane.flush()
        self.rebalance()

    REBALANCE_MIN_GAP = 4096       # like compaction: only when it pays (the step time of a search is its LARGEST shard's scan)

    @_whole_op
    def rebalance(self, force: bool = False) -> int:
        """Shard rebalancing.  New points already go to the least-full shard, so growth evens a
        collection out by itself; after a mass delete that hit the shards unevenly (a project whose files were indexed in one
        stretch) the fullest shard keeps bounding every search.  When the fullest and the emptiest shard differ by at least
        REBALANCE_MIN_GAP points AND a quarter of the mean (`force`: by more than one point), half of the gap moves: the tail
        rows of the fullest shard are read back (`fetch_rows`), written to the emptiest one under the same ids and payloads,
        deleted at the source, and the source is compacted.  Returns the number of points moved.  A moved point is a fresh
        write at its new home (bf16 shards: the same bits; fp32 shards: re-normalised, i.e. within one float32 ulp)."""
        moved = 0
        for _ in range(4 * len(self.shards)):
            live = self.live_counts()
            hi = max(range(len(live)), key=live.__getitem__)
            lo = least_full(live)
            gap = live[hi] - live[lo]
            mean = sum(live) / len(live)
            if gap <= 1 or not (force or (gap >= self.REBALANCE_MIN_GAP and 4 * gap >= mean)):
                break
            src, dst, n = self.shards[hi], self.shards[lo], gap // 2
            rows = [r for r in range(len(src.ids) - 1, -1, -1) if src.ids[r] is not None][:n]
            self.plane.flush()
            vec = self.plane.call("fetch", self.name, {hi: np.asarray(rows, dtype=np.int64)})[hi]
            ids, pls = [src.ids[r] for r in rows], [src.payloads[r] for r in rows]
            src.dev.delete_rows(rows)
            src.release_rows(rows)
            dst.upsert(ids, np.asarray(vec, dtype=np.float64), pls)
            src.maybe_compact(force=True)
            self.plane.flush()
            moved += len(rows)
        return moved

    @_whole_op
    def delete(self, filters) -> int:
        want = self.want_codes(filters)
        found = self.plane.call("delete_where", self.name, want)
        self._after_delete([rows.tolist() for rows, _ in found])
        return sum(n for _, n in found)

    @_whole_op
    def count(self, filters=None) -> int:
        if not filters:
            return sum(self.plane.call("count", self.name))
        return self._match(self.want_codes(filters), cap=0)[1]

    @_whole_op
    def rows_matching(self, eq, text) -> list[tuple[int, int]]:
        per_shard, _ = self._match(self.want_codes(eq))
        out = []
        for s, rows in enumerate(per_shard):
            for r in rows:
                p = self.shards[s].payloads[r] or {}
                if all(isinstance(p.get(k), str) and t in p[k] for k, t in text):
                    out.append((s, r))
        return out

    @_whole_op
    def point(self, where: tuple[int, int]):
        return self.shards[where[0]].point(where[1])

    def delete_found(self, found: Sequence[tuple[int, int]]) -> None:
        per_shard: list[list[int]] = [[] for _ in self.shards]
        for s, r in found:
            per_shard[s].append(r)
        for sh, rows in zip(self.shards, per_shard):
            if rows:
                sh.dev.delete_rows(rows)
        self._after_delete(per_shard)

    def close(self) -> None:
        if not self.plane.closed:
            self.plane.call("destroy", self.name)


# ------------------------------------------------------------------------------------------------------------------
# the adapter
# ------------------------------------------------------------------------------------------------------------------
class ShardedB200VectorStore(B200VectorStore):
    """``QdrantManager``'s surface (see ``client.B200VectorStore``) over every GPU of the ``torchrun`` job.

    Every rank::

        plane = ShardPlane.start()                       # collective
        if plane.rank != 0:
            plane.serve()                                # until rank 0 calls plane.shutdown()
        else:
            store = ShardedB200VectorStore(dimensions=768, storage="bf16", plane=plane)
            await store.connect(); await store.create_collections()
            ...                                          # QueryEngine(qdrant=store), VectorIndexer(store, ...), ...
            await store.close(); plane.shutdown()
    """

    def __init__(self, host: str | None = None, port: int | None = None, grpc_port: int | None = None, *,
                 dimensions: int | None = None, storage: str | None = None, plane: ShardPlane | None = None):
        super().__init__(host, port, grpc_port, dimensions=dimensions, storage=storage, device=plane.device if plane else None)
        self._plane = plane

    @property
    def plane(self) -> ShardPlane:
        if self._plane is None:
            raise VectorStoreError("Client not connected. Call connect() first.")
        return self._plane

    async def connect(self) -> None:
        if not self._connected:
            try:
                if self._plane is None:
                    self._plane = await asyncio.to_thread(ShardPlane.start)
                if self._plane.rank != 0:
                    raise RuntimeError("ShardedB200VectorStore lives on rank 0; the other ranks run plane.serve()")
                if self._plane.closed:
                    raise RuntimeError("shard plane is shut down")
                self._connected = True
                logger.info("lattice-b200 sharded vector store: %d shard(s)", self._plane.world)
            except Exception as e:  # noqa: BLE001
                raise VectorStoreError("Failed to connect to Qdrant", cause=e)

    async def create_collections(self) -> None:
        try:
            _ = self.client
            for name in (CollectionName.CODE_CHUNKS.value, CollectionName.SUMMARIES.value):
                if name not in self._collections:
                    self._collections[name] = await asyncio.to_thread(
                        _ShardedHostCollection, self.plane, name, self._dimensions, self._storage, _INDEX_FIELDS[name])
                    logger.info(f"Created collection: {name}")
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError("Failed to create collections", cause=e)

    async def get_collection_info(self, collection: str):
        info = await super().get_collection_info(collection)
        info.shards = self.plane.world
        info.shard_points = self._get(collection).live_counts()
        return info

    async def rebalance(self, collection: str | None = None, force: bool = True) -> int:
        """Additive: even out the shards of one collection (or of both) now; deletes do it by themselves once the gap is large
        (``_ShardedHostCollection.rebalance``).  Returns the number of points moved."""
        try:
            colls = [self._get(collection)] if collection else [self._get(n) for n in self._collections]

            def work():
                total = 0
                for coll in colls:
                    with coll.lock:
                        total += coll.rebalance(force=force)
                return total
            return await asyncio.to_thread(work)
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError("Failed to rebalance the shards", cause=e)

    async def load(self, directory: str) -> None:
        """Additive, like ``B200VectorStore.load``: every rank reads its own shard file, rank 0 the host half.  The job must have as
        many ranks as the one that saved."""
        try:
            _ = self.client

            def work():
                loaded = {}
                for name in (CollectionName.CODE_CHUNKS.value, CollectionName.SUMMARIES.value):
                    if os.path.exists(os.path.join(directory, f"{name}.host{self.plane.world}.pkl")):
                        loaded[name] = _ShardedHostCollection.load(self.plane, directory, name)
                return loaded
            self._collections.update(await asyncio.to_thread(work))      # the workers replaced their shards in place
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to load collections from {directory}", cause=e)

    async def attach_encoder(self, spec: dict | None) -> int | None:
        """Additive API (SURVEY section 8f row 4 over N GPUs): every rank builds the code encoder `spec` describes on its own GPU
        (``ShardPlane._op_encoder``: {"pretrained": path} or {"random": {...}}); afterwards ``upsert_tokens`` embeds every chunk on
        the GPU whose shard will hold it - N encoders work in parallel and no vector crosses the control plane.  Returns the
        embedding width; None detaches."""
        try:
            widths = await asyncio.to_thread(self.plane.call, "encoder", None, spec)
            if spec is None:
                self.plane.encoder_hidden = None
                return None
            if len(set(widths)) != 1:
                raise RuntimeError(f"the ranks built encoders of different widths: {widths}")
            self.plane.encoder_hidden = int(widths[0])
            return self.plane.encoder_hidden
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError("Failed to attach the encoder", cause=e)

    async def upsert_tokens(self, collection: str, ids: list[str], token_ids, payloads: list[dict[str, Any]], encoder=None) -> None:
        """``B200VectorStore.upsert_tokens`` over the shards: see ``attach_encoder``."""
        return await super().upsert_tokens(collection, ids, token_ids, payloads, encoder)

    async def search_and_rank(self, collection: str, items: Sequence[tuple], limit: int = 10,
                              filters: dict[str, Any] | None = None, ranker=None, summaries: bool = False,
                              summaries_filters: dict[str, Any] | None = None):
        """Same call and same results as ``B200VectorStore.search_and_rank``, by the two-step route: ONE batched search over all
        shards (plus one over ``summaries`` for the queries whose intent ``QueryEngine._execute_vector_search`` extends,
        query/engine.py:331-344), the hits shaped as ``VectorSearcher`` shapes them (query/vector_search.py:221-260), then ONE
        K3 launch on rank 0's GPU for the whole batch (``HybridRanker.rank_batch``).  The single-GPU store keeps the hits on
        the device between the two; over shards the ranking attributes would have to be gathered per shard before the exchange."""
        try:
            coll = self._get(collection)
            summ = CollectionName.SUMMARIES.value
            coll2 = self._get(summ) if summaries and collection != summ and limit // 2 > 0 else None
            from .ranking import SUMMARY_INTENTS, HybridRanker, _intent_value
            rk = ranker or HybridRanker()
            items = list(items)
            if not items:
                return []
            q = np.asarray([it[2] for it in items], dtype=np.float64)
            sel2 = [i for i, it in enumerate(items) if _intent_value(it[0].primary_intent) in SUMMARY_INTENTS] if coll2 else []
            kind = "summary" if collection == summ else "code"

            def work():
                with coll.lock:
                    hits = coll.search(q, limit, filters or None)
                vector_results = [[vector_result_from_payload(h["payload"], h["score"], kind) for h in hs] for hs in hits]
                if sel2:
                    with coll2.lock:
                        hits2 = coll2.search(q[sel2], limit // 2, summaries_filters or None)
                    for i, hs in zip(sel2, hits2):
                        vector_results[i].extend(vector_result_from_payload(h["payload"], h["score"], "summary") for h in hs)
                return rk.rank_batch([(it[0], it[1], vr, it[3]) for it, vr in zip(items, vector_results)])
            return await asyncio.to_thread(work)
        except Exception as e:  # noqa: BLE001
            raise VectorStoreError(f"Failed to search and rank in {collection}", cause=e)


def run(main: Callable, **plane_kwargs) -> Any:
    """Entry point of a ``torchrun`` job that uses the sharded store: every rank calls ``run(main)``; rank 0 executes
    ``await main(plane)`` (the application: build a ``ShardedB200VectorStore(plane=plane, ...)`` and use it), the other ranks serve
    their shard until it returns or raises.  Returns ``main``'s result on rank 0, None elsewhere."""
    import torch.distributed as dist
    plane = ShardPlane.start(**plane_kwargs)
    result = None
    try:
        if plane.rank != 0:
            plane.serve()
        else:
            try:
                result = asyncio.run(main(plane))
            finally:
                plane.shutdown()
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
    return result


__all__ = ["ShardPlane", "ShardedB200VectorStore", "run", "split_row", "least_full", "SHARD_BITS"]
