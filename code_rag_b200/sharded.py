"""Row-sharded search across the GPUs of one box: one process per GPU (torchrun), NCCL over NVLink.

Exact top-k decomposes over row shards (global top-k is a subset of the union of the shard top-k lists), so the
path has exactly one exchange step: every rank scans its shard (K1/K2 + exact rescoring), the Q x k
(score, row, tie) lists are all-gathered, and K5 merges G*k -> k on every rank in the reference's order
(score desc, id asc).  Message size is 24*Q*k bytes per rank (240 B for Q=1, k=10): latency-bound, so a single
``all_gather_into_tensor`` of one packed buffer is used.  torch.distributed is the plumbing; the scan and the merge
are this repo's kernels.
"""
from __future__ import annotations

import numpy as np

import ctypes as C
import os

from . import _native as N
from .collection import DeviceCollection, merge_topk_device
from .errors import NativeLibraryError


def shard_bounds(n_total: int, world: int, align: int = 1) -> list[tuple[int, int]]:
    """Contiguous row blocks [lo, hi) per rank, sizes differing by at most `align` rows."""
    if world < 1 or n_total < 0:
        raise ValueError("bad world / n_total")
    units = (n_total + align - 1) // align
    out, lo = [], 0
    for r in range(world):
        u = units // world + (1 if r < units % world else 0)
        hi = min(n_total, lo + u * align)
        out.append((lo, hi))
        lo = hi
    return out


def allgather_packed(local, gathered, group=None):
    """The path's one exchange step: every rank contributes its packed [3, Q, k] int64 block (float64 score bits,
    global rows, tie keys) and receives all of them as [G, 3, Q, k].  NCCL on GPUs; gloo in the CPU tests."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=group)
    return gathered


class ShardedSearcher:
    """Wraps this rank's shard; every search returns the same global result on every rank.

    All device work is ordered on one side stream (``self.stream``).  Host entry points are zero-copy: queries are read
    by the prep kernel straight from pinned host memory and the final kernel (finalize, or the merge when world > 1)
    stores the result straight into pinned host memory, so a step is kernels + one collective, no copy-engine hops.
    """

    def __init__(self, shard: DeviceCollection, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._bufs: dict[tuple, dict] = {}
        self.merge_launches = 0
        self.stream = torch.cuda.Stream(device=self.device)
        shard.set_option("timing", 0)      # consecutive searches overlap (programmatic dependent launch); bench.py's roofline leg turns it on
        self.n_slots = 4
        self._next_slot = 0
        # exchange step: "p2p" = one kernel per rank that stores its block into every peer's gather buffer over NVLink
        # (CUDA-IPC mapped), flags, waits and merges; "nccl" = all_gather_into_tensor + merge kernel
        self.exchange_mode = os.environ.get("LATTICE_B200_EXCHANGE", "p2p") if self.world > 1 else "none"
        self._ex = None
        self._ex_cap = (0, 0)

    def _exchange(self, Q: int, k: int):
        """Create (or grow) the peer-memory exchange; collective: every rank calls it with the same arguments."""
        if self._ex is not None and Q <= self._ex_cap[0] and k <= self._ex_cap[1]:
            return self._ex
        lib = N.load()
        if self._ex is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            lib.lvs_exchange_destroy(self._ex)
            self._ex = None
        cap_q, cap_k = max(Q, self._ex_cap[0], 1), max(k, self._ex_cap[1], 16)
        ex = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        N.check(lib.lvs_exchange_create(self.world, self.rank, cap_q, cap_k, C.byref(ex), handle), "lvs_exchange_create")
        handles = [None] * self.world
        self.dist.all_gather_object(handles, bytes(handle), group=self.group)
        blob = b"".join(handles)
        N.check(lib.lvs_exchange_connect(ex, C.c_char_p(blob)), "lvs_exchange_connect")
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)       # every rank has mapped every buffer before the first store
        self._ex, self._ex_cap = ex, (cap_q, cap_k)
        return ex

    def close(self) -> None:
        if self._ex is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            N.load().lvs_exchange_destroy(self._ex)
            self._ex = None

    def _slot(self, Q: int, k: int, slot: int, host: bool) -> dict:
        key = (Q, k, slot, host)
        b = self._bufs.get(key)
        if b is None:
            t = self.torch
            mk = (lambda *shape, dtype: t.zeros(shape, dtype=dtype).pin_memory()) if host else \
                 (lambda *shape, dtype: t.zeros(shape, dtype=dtype, device=self.device))
            b = {"out": mk(3, Q, k, dtype=t.int64), "counts": mk(Q, dtype=t.int32), "flags": mk(Q, dtype=t.int32),
                 "event": t.cuda.Event()}
            if self.world > 1:
                b["local"] = t.zeros((3, Q, k), dtype=t.int64, device=self.device)
                b["local_counts"] = t.zeros(Q, dtype=t.int32, device=self.device)
                b["local_flags"] = t.zeros(Q, dtype=t.int32, device=self.device)
                b["gathered"] = t.zeros((self.world, 3, Q, k), dtype=t.int64, device=self.device)
            self._bufs[key] = b
        return b

    def _enqueue(self, q_ptr: int, q_dtype: str, Q: int, k: int, want, b: dict) -> None:
        """This rank's search + the exchange + the merge, all on self.stream.  "p2p" (default): ONE call into the library - for
        up to 4 queries on the scan path that is one kernel per rank (scan + exact rescoring + peer-memory exchange + merge),
        batches add the exchange kernel behind the tensor-core path's finalize.  "nccl": local search, all_gather_into_tensor,
        merge kernel (the form BASELINE.json's north_star names).  Either way every rank ends up with the same merged lists
        and the same merged flags (OR over the shards)."""
        stream = self.stream.cuda_stream
        out = b["out"]
        b["base"] = self.shard.search_counter + 1          # reference search number of query 0 (a repeat re-uses it)
        b["want"], b["q"] = want, (q_ptr, q_dtype, Q, k)
        if self.world == 1:
            self.shard.search_device_async(q_ptr, q_dtype, Q, k, want, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                           b["counts"].data_ptr(), b["flags"].data_ptr(), stream)
            return
        if self.exchange_mode == "p2p":
            ex = self._exchange(Q, k)
            self.shard.search_sharded_device_async(ex, q_ptr, q_dtype, Q, k, want, out.data_ptr(), b["counts"].data_ptr(),
                                                   b["flags"].data_ptr(), stream)
            return
        local = b["local"]
        self.shard.search_device_async(q_ptr, q_dtype, Q, k, want, local[0].data_ptr(), local[1].data_ptr(), local[2].data_ptr(),
                                       b["local_counts"].data_ptr(), b["local_flags"].data_ptr(), stream)
        with self.torch.cuda.stream(self.stream):
            allgather_packed(local, b["gathered"], self.group)
            self.dist.all_reduce(b["local_flags"], op=self.dist.ReduceOp.MAX, group=self.group)
            b["flags"].copy_(b["local_flags"], non_blocking=True)
        n = Q * k
        base = b["gathered"].data_ptr()
        merge_topk_device(base, base + 8 * n, base + 16 * n, self.world, Q, k, out[0].data_ptr(), out[1].data_ptr(),
                          out[2].data_ptr(), b["counts"].data_ptr(), stream, shard_stride=3 * n)
        self.merge_launches += 1

    def _repeat_query(self, qvec: np.ndarray, search_no: int, k: int, want):
        """Collective (every rank sees the same merged flags, so every rank gets here for the same queries): redo ONE query through
        the synchronous path of the library - larger candidate sets, the exact scan instead of the tensor-core path, numbered as the
        reference search it repeats - then exchange and merge the lists again.  Returns ([3, k] int64 merged block, count, flag)."""
        t = self.torch
        lib = N.load()
        rb = self._repeat_bufs(k, qvec.shape[-1])
        rb["hq"].numpy()[0] = qvec
        st = self.stream.cuda_stream
        self.shard.search_device_at(int(search_no), rb["hq"].data_ptr(), "f64", 1, k, want, rb["local"][0].data_ptr(), rb["local"][1].data_ptr(),
                                    rb["local"][2].data_ptr(), rb["counts"].data_ptr(), rb["hflags"], st)
        if self.world == 1:
            self.stream.synchronize()
            return rb["local"][:, 0, :].cpu().numpy(), int(rb["counts"].item()), int(rb["hflags"][0])
        rb["dflags"].copy_(t.from_numpy(rb["hflags"]), non_blocking=False)
        if self.exchange_mode == "p2p":
            N.check(lib.lvs_exchange_merge_device(self._exchange(1, k), C.c_void_p(rb["local"].data_ptr()), C.c_void_p(rb["dflags"].data_ptr()),
                                                  1, k, C.c_void_p(rb["out"].data_ptr()), C.c_void_p(rb["counts"].data_ptr()),
                                                  C.c_void_p(rb["dflags2"].data_ptr()), C.c_void_p(st)), "lvs_exchange_merge_device")
        else:
            with t.cuda.stream(self.stream):
                allgather_packed(rb["local"], rb["gathered"], self.group)
                self.dist.all_reduce(rb["dflags"], op=self.dist.ReduceOp.MAX, group=self.group)
                rb["dflags2"].copy_(rb["dflags"])
            base = rb["gathered"].data_ptr()
            merge_topk_device(base, base + 8 * k, base + 16 * k, self.world, 1, k, rb["out"][0].data_ptr(), rb["out"][1].data_ptr(),
                              rb["out"][2].data_ptr(), rb["counts"].data_ptr(), st, shard_stride=3 * k)
        self.stream.synchronize()
        return rb["out"][:, 0, :].cpu().numpy(), int(rb["counts"].item()), int(rb["dflags2"].item())

    def _settle(self, queries: np.ndarray, base: int, k: int, want, flags: np.ndarray, store) -> np.ndarray:
        """After a search has completed: raise if the exchange failed, and REPEAT the queries that some shard could not prove
        exact.  `store(qi, block, count)` puts a repeated query's merged lists where the caller keeps its results."""
        if (flags & N.FLAG_EXCHANGE).any():
            raise NativeLibraryError("sharded search: a peer's result lists did not arrive (exchange timeout); a rank is down or stalled")
        for qi in np.nonzero(flags & N.FLAG_UNPROVEN)[0].tolist():
            block, count, f = self._repeat_query(queries[qi], base + qi, k, want)
            store(qi, block, count)
            flags[qi] = f
        return flags

    def _repeat_bufs(self, k: int, dim: int) -> dict:
        rb = self._bufs.get(("repeat", k))
        if rb is None:
            t = self.torch
            dev = lambda *shape, dtype: t.zeros(shape, dtype=dtype, device=self.device)  # noqa: E731
            rb = {"local": dev(3, 1, k, dtype=t.int64), "out": dev(3, 1, k, dtype=t.int64), "counts": dev(1, dtype=t.int32),
                  "dflags": dev(1, dtype=t.int32), "dflags2": dev(1, dtype=t.int32), "hflags": np.zeros(1, dtype=np.int32),
                  "gathered": dev(self.world, 3, 1, k, dtype=t.int64), "hq": t.zeros((1, dim), dtype=t.float64).pin_memory()}
            self._bufs[("repeat", k)] = rb
        return rb

    # ---- device entry points ---------------------------------------------------------------------------
    def search_device_async(self, dq, k: int, want=None, slot: int = 0):
        """dq: CUDA tensor [Q, dim] float32/float64 (same on every rank), already valid on ``self.stream``.  Enqueue only;
        returns CUDA tensors (scores f64 [Q,k], rows i64, ties i64, counts i32 [Q], flags i32 [Q]) that are valid once
        ``self.stream`` has been synchronised.  `slot` rotates result buffers between consecutive searches.  Flagged queries
        (``flags != 0``: some shard could not prove its list exact) are NOT repeated here; :meth:`search_device` does that."""
        t = self.torch
        Q = int(dq.shape[0])
        b = self._slot(Q, k, slot, host=False)
        self._enqueue(dq.data_ptr(), "f64" if dq.dtype == t.float64 else "f32", Q, k, want, b)
        out = b["out"]
        return out[0].view(t.float64), out[1], out[2], b["counts"], b["flags"]

    def search_device(self, dq, k: int, want=None):
        """Synchronous device entry; flagged queries are repeated (collectively); returns the tensors plus the host flags."""
        t = self.torch
        s, r, ti, c, f = self.search_device_async(dq, k, want, slot=0)
        self.stream.synchronize()
        b = self._slot(int(dq.shape[0]), k, 0, host=False)
        flags = f.cpu().numpy().copy()
        if flags.any():
            def store(qi, block, count):
                b["out"][:, qi, :].copy_(t.from_numpy(block))
                b["counts"][qi] = count
            flags = self._settle(dq.double().cpu().numpy(), int(b["base"]), k, want, flags, store)
            t.cuda.synchronize()
        return s, r, ti, c, flags

    # ---- host entry points -----------------------------------------------------------------------------
    def submit(self, queries: np.ndarray, k: int, want=None):
        """Pipelined host entry; returns a handle for :meth:`wait`.  Up to ``n_slots`` searches may be in flight.  One GPU or the
        peer-memory exchange: the library's own pipelined pair (``lvs_search_submit[_sharded]`` / ``lvs_search_wait``: pinned slots,
        the query staged by one CTA, the result and a completion word stored by the kernel, no event between consecutive
        searches).  "nccl" exchange: pinned torch buffers + all_gather_into_tensor + merge kernel."""
        t = self.torch
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None, :]
        Q, dim = q.shape
        if Q == 0:                                     # nothing to search (every rank sees the same empty batch: no collective)
            return {"empty": int(k)}
        if self.world == 1 or self.exchange_mode == "p2p":
            base = self.shard.search_counter + 1
            ticket = self.shard.search_submit(q, k, want) if self.world == 1 else \
                self.shard.search_submit_sharded(self._exchange(Q, k), q, k, want)
            return {"ticket": ticket, "queries": q, "base": base, "want": want, "k": k}
        slot = self._next_slot
        self._next_slot = (slot + 1) % self.n_slots
        b = self._slot(Q, k, slot, host=True)
        if "hq" not in b or b["hq"].shape != (Q, dim):
            b["hq"] = t.zeros((Q, dim), dtype=t.float64).pin_memory()
        b["hq"].numpy()[...] = q
        # pinned memory is mapped into the device address space (UVA): the kernels read/write it directly
        self._enqueue(b["hq"].data_ptr(), "f64", Q, k, want, b)
        b["event"].record(self.stream)
        b["queries"], b["k"] = q, k
        return b

    def poll(self, handle) -> bool:
        """True once :meth:`wait` will not block on the device."""
        if "empty" in handle:
            return True
        if "ticket" in handle:
            return self.shard.search_poll(handle["ticket"])
        return bool(handle["event"].query())

    def wait(self, handle):
        """Blocks until the search has finished; flagged queries are repeated (every rank takes the same decision, so this stays
        a collective); raises when the exchange reported a missing peer."""
        if "empty" in handle:
            k = handle["empty"]
            return (np.zeros((0, k)), np.zeros((0, k), dtype=np.int64), np.zeros((0, k), dtype=np.uint64), np.zeros(0, dtype=np.uint32),
                    np.zeros(0, dtype=np.int32))
        if "ticket" in handle:
            res = handle["done"] if "done" in handle else self.shard.search_wait(handle["ticket"])
            scores, rows, ties, counts, flags = res.scores, res.rows, res.ties, res.counts, res.flags
        else:
            handle["event"].synchronize()
            o = handle["out"].numpy()
            scores, rows, ties = o[0].view(np.float64).copy(), o[1].copy(), o[2].view(np.uint64).copy()
            counts, flags = handle["counts"].numpy().copy(), handle["flags"].numpy().copy()
        if flags.any() and (self.world > 1 or "ticket" not in handle):      # one GPU through the library: already repeated there
            def store(qi, block, count):
                scores[qi], rows[qi], ties[qi], counts[qi] = block[0].view(np.float64), block[1], block[2].view(np.uint64), count
            flags = self._settle(handle["queries"], int(handle["base"]), handle["k"], handle["want"], flags, store)
        return scores, rows, ties, counts, flags

    def search(self, queries: np.ndarray, k: int, want=None):
        """Synchronous host entry (one search at a time)."""
        return self.wait(self.submit(queries, k, want))
