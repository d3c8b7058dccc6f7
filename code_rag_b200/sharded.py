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

from .collection import DeviceCollection, merge_topk_device


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


class ShardedSearcher:
    """Wraps this rank's shard; `search` returns the same global result on every rank."""

    def __init__(self, shard: DeviceCollection, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._bufs: dict[tuple, tuple] = {}
        self._flag_bufs: dict[tuple, object] = {}
        self._io_bufs: dict[tuple, tuple] = {}
        self.merge_launches = 0
        # all device work of this searcher is ordered on one side stream (a real handle, never the legacy stream 0)
        self.stream = torch.cuda.Stream(device=self.device)
        self.n_slots = 4
        self._next_slot = 0

    def _buffers(self, Q: int, k: int, slot: int = 0):
        key = (Q, k, slot)
        if key not in self._bufs:
            t = self.torch
            local = t.zeros((3, Q, k), dtype=t.int64, device=self.device)
            counts = t.zeros(Q, dtype=t.int32, device=self.device)
            gathered = t.zeros((self.world, 3, Q, k), dtype=t.int64, device=self.device) if self.world > 1 else None
            out = t.zeros((3, Q, k), dtype=t.int64, device=self.device)
            out_counts = t.zeros(Q, dtype=t.int32, device=self.device)
            self._bufs[key] = (local, counts, gathered, out, out_counts)
        return self._bufs[key]

    def search_device_async(self, dq, k: int, want=None, slot: int = 0):
        """Enqueue-only variant for pipelined callers: no host synchronisation, flags stay on the device.  `slot`
        selects one of several result buffers so that consecutive searches do not overwrite each other.  Returns
        (scores, rows, ties, counts, flags, packed) where packed is the [3, Q, k] int64 block holding the first three.
        Call it under ``torch.cuda.stream(searcher.stream)`` (or any non-default stream)."""
        t = self.torch
        Q = int(dq.shape[0])
        local, counts, gathered, out, out_counts = self._buffers(Q, k, slot)
        flags = self._flag_bufs.setdefault((Q, slot), t.zeros(Q, dtype=t.int32, device=self.device))
        stream = t.cuda.current_stream().cuda_stream
        qd = "f64" if dq.dtype == t.float64 else "f32"
        self.shard.search_device_async(dq.data_ptr(), qd, Q, k, want, local[0].data_ptr(), local[1].data_ptr(),
                                       local[2].data_ptr(), counts.data_ptr(), flags.data_ptr(), stream)
        if self.world == 1:
            return local[0].view(t.float64), local[1], local[2], counts, flags, local
        self.dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        n = Q * k
        base = gathered.data_ptr()
        merge_topk_device(base, base + 8 * n, base + 16 * n, self.world, Q, k, out[0].data_ptr(), out[1].data_ptr(),
                          out[2].data_ptr(), out_counts.data_ptr(), stream, shard_stride=3 * n)
        self.merge_launches += 1
        return out[0].view(t.float64), out[1], out[2], out_counts, flags, out

    def search_device(self, dq, k: int, want=None):
        """dq: CUDA tensor [Q, dim] float32/float64 (same on every rank).  Returns CUDA tensors
        (scores f64 [Q,k], rows i64 [Q,k], ties i64 [Q,k], counts i32 [Q]) and the host flags."""
        t = self.torch
        Q = int(dq.shape[0])
        local, counts, gathered, out, out_counts = self._buffers(Q, k)
        stream = t.cuda.current_stream().cuda_stream
        qd = "f64" if dq.dtype == t.float64 else "f32"
        flags = self.shard.search_device(dq.data_ptr(), qd, Q, k, want, local[0].data_ptr(), local[1].data_ptr(),
                                         local[2].data_ptr(), counts.data_ptr(), stream)
        if self.world == 1:
            return local[0].view(t.float64), local[1], local[2], counts, flags
        self.dist.all_gather_into_tensor(gathered.view(-1), local.view(-1), group=self.group)
        n = Q * k
        base = gathered.data_ptr()
        merge_topk_device(base, base + 8 * n, base + 16 * n, self.world, Q, k, out[0].data_ptr(), out[1].data_ptr(),
                          out[2].data_ptr(), out_counts.data_ptr(), stream, shard_stride=3 * n)
        self.merge_launches += 1
        return out[0].view(t.float64), out[1], out[2], out_counts, flags

    # ---- host entry points ----------------------------------------------------------------------------
    def submit(self, queries: np.ndarray, k: int, want=None):
        """Pipelined host entry: pinned H2D of the queries, sharded search, D2H of the merged result, all enqueued
        on the searcher's stream.  Returns a handle for :meth:`wait`; up to ``n_slots`` may be in flight."""
        t = self.torch
        q = np.ascontiguousarray(queries, dtype=np.float64)
        if q.ndim == 1:
            q = q[None, :]
        Q, dim = q.shape
        slot = self._next_slot
        self._next_slot = (slot + 1) % self.n_slots
        key = (Q, k, dim, slot)
        if key not in self._io_bufs:
            self._io_bufs[key] = (
                t.empty((Q, dim), dtype=t.float64).pin_memory(), t.empty((Q, dim), dtype=t.float64, device=self.device),
                t.empty((3, Q, k), dtype=t.int64).pin_memory(), t.empty(Q, dtype=t.int32).pin_memory(),
                t.empty(Q, dtype=t.int32).pin_memory(), t.cuda.Event())
        hq, dq, h_out, h_counts, h_flags, ev = self._io_bufs[key]
        hq.numpy()[...] = q
        with t.cuda.stream(self.stream):
            dq.copy_(hq, non_blocking=True)
            s, r, ti, c, flags, packed = self.search_device_async(dq, k, want, slot)
            h_out.copy_(packed, non_blocking=True)
            h_counts.copy_(c, non_blocking=True)
            h_flags.copy_(flags, non_blocking=True)
            ev.record(self.stream)
        return key

    def wait(self, handle):
        hq, dq, h_out, h_counts, h_flags, ev = self._io_bufs[handle]
        ev.synchronize()
        o = h_out.numpy()
        return o[0].view(np.float64).copy(), o[1].copy(), o[2].view(np.uint64).copy(), h_counts.numpy().copy(), h_flags.numpy().copy()

    def search(self, queries: np.ndarray, k: int, want=None):
        """Synchronous host entry (one search at a time)."""
        return self.wait(self.submit(queries, k, want))
