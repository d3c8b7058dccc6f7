#!/usr/bin/env python
"""Headline benchmark: QPS of exact top-10 cosine search over a 10M x 768 bf16 corpus (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A step = one search (Q queries, default 1) over the whole corpus.  With N > 1 (torchrun, one rank per GPU) the
10M rows are split into N contiguous shards (strong scaling), every rank scans its shard and the per-shard
top-k lists are merged after one NCCL all-gather.  Timed region: W warm-up steps, barrier + synchronize,
exactly K steps, barrier + synchronize; device time by CUDA events on the launching stream, MAX over ranks.
`value` has the queries resident in HBM; `e2e` goes through the host-buffer API (pinned H2D of the query,
D2H of ids+scores inside the timed region).  Each step scans >= 1.9 GB per GPU, far more than the 126 MB L2,
so no explicit L2 flush is needed between steps (config.l2: "inputs_exceed_l2").
`roofline` is the HBM one for the scan (and for the tensor-core kernel up to 128 queries per step) and the tensor one
(TFLOP/s against the measured cuBLAS bf16 peak) above that.  At N=1 with the default single-query step the line also carries
`batched`: BASELINE.json's configs[2] (256 queries x top-100 on the tcgen05 path) measured after the headline legs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

# torchrun exports OMP_NUM_THREADS=1 to every rank.  The reference arm is a CPU measurement on rank 0 alone (the other ranks
# exit at once), so it gets every host thread back - before numpy loads its BLAS, which reads the variable once.
if any(a in ("reference", "--impl=reference") for a in sys.argv) and os.environ.get("RANK", "0") == "0" and "LATTICE_B200_KEEP_OMP" not in os.environ:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "qps_exact_top10_cosine_10Mx768_bf16"
CPU_KEYS = ("value", "unit", "cores", "kind", "sample", "extrapolated", "best_effort_cpu")
UNIT = "queries/s"
CHUNK_ROWS = 250_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=1, help="queries per step (batch)")
    ap.add_argument("--storage", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--cpu-sample-rows", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the C3 (256 queries x top-100) measurement beside the headline")
    ap.add_argument("--no-adapter", action="store_true", help="skip the e2e_adapter leg (await store.search(...) through the QdrantManager drop-in)")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the 100M-row configuration (configs[4])")
    ap.add_argument("--stage-kb", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    return ap.parse_args()


def bf16_round_np(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32)).view(np.float32).reshape(x.shape)


def make_queries(n: int, dim: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((n, dim))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def blas_threads() -> int:
    """Threads numpy's BLAS actually runs with in this process (what `cpu_baseline.cores` reports)."""
    try:
        from threadpoolctl import threadpool_info
        n = [int(i["num_threads"]) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:  # noqa: BLE001
        pass
    return int(os.environ.get("OMP_NUM_THREADS") or os.cpu_count() or 1)


def host_chunk(chunk_id: int, rows: int, dim: int, seed: int = 3456) -> np.ndarray:
    """CPU twin of the corpus generator for the CPU arm (same distribution, numpy RNG; one call per chunk of rows)."""
    rng = np.random.default_rng([seed, chunk_id])
    x = rng.standard_normal((rows, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return bf16_round_np(x)


# --------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (qdrant-client local mode, restated in oracle/qdrant_local.py)
# --------------------------------------------------------------------------------------------------------
def cpu_reference_qps(args, steps: int, warmup: int) -> dict:
    """The reference's CPU path on this box's host cores.  Engine: the real qdrant-client in local mode when it is importable
    (oracle/real_qdrant.py probes site-packages, baseline/_ref and oracle/_ref), else the numpy restatement of it
    (oracle/qdrant_local.py, kind "port").  By default a bounded row sample is timed and scaled linearly to the full corpus (the
    work is a dense O(N*D) pass + a sort; "extrapolated": true); --cpu-sample-rows 0 times the FULL corpus (fp32 in RAM: 30.7 GB
    for 10M x 768, a minute or two per search) so that one real point validates the scaling."""
    from oracle import real_qdrant
    from oracle.qdrant_local import OracleCollection
    full = args.cpu_sample_rows <= 0
    n = args.rows if full else min(args.cpu_sample_rows, args.rows)
    if full:
        steps, warmup = min(steps, 2), 1
    qs = make_queries(steps + warmup, args.dim, seed=11)
    real = real_qdrant.find() is not None and n <= 2_000_000       # PointStruct upserts of python lists: bounded samples only
    if real:
        eng = real_qdrant.RealManager(args.dim)
        eng.create_collections()
        for c0 in range(0, n, 20_000):
            x = host_chunk(c0 // 20_000, min(20_000, n - c0), args.dim)
            eng.upsert("code_chunks", list(range(c0, c0 + len(x))), x.astype(np.float64).tolist(), [None] * len(x))
        search = lambda q: eng.search("code_chunks", q.tolist(), limit=args.k)
        kind = f"qdrant-client {real_qdrant.version()} (local mode)"
    else:
        ora = OracleCollection(args.dim, capacity=n)               # sized up front: no doubling copy of a 30 GB matrix
        for c0 in range(0, n, CHUNK_ROWS):
            x = host_chunk(c0 // CHUNK_ROWS, min(CHUNK_ROWS, n - c0), args.dim)
            ora.upsert_rows_f32(c0, x, [None] * len(x))
        search = lambda q: ora.search_topk_rows(q, args.k)
        kind = "port"
    for i in range(warmup):
        search(qs[i])
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        for _ in range(args.queries):
            search(qs[i])
    dt = (time.perf_counter() - t0) / max(1, steps)
    scale = args.rows / n
    out = {"value": args.queries / (dt * scale), "unit": UNIT, "cores": blas_threads(), "kind": kind, "extrapolated": not full,
           "sample": (f"all {n} rows" if full else f"{n} of {args.rows} rows") + f" x {args.dim} (bf16-rounded, fp32 in RAM), {steps} searches after "
                     f"{warmup} warm-up" + ("" if full else f"; time scaled x{scale:.0f} (scan + sort is linear in rows)") +
                     f"; numpy {np.__version__}, {blas_threads()} BLAS threads",
           "ms_per_step_sample": dt * 1e3, "ms_per_step_scaled": dt * 1e3 * scale}
    if not real and not full:
        # a CPU path that is NOT the reference's arithmetic, only the fastest plain-numpy way to the same ids: rows normalised once,
        # float32 matrix-vector product (BLAS, all threads), argpartition + sort of the k winners.  Reported so that the GPU/CPU ratio
        # is not inflated by local mode's per-search re-normalisation, float64 product and full argsort.
        x = ora.vectors[:n]
        xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-30)
        q32 = qs.astype(np.float32)
        for i in range(warmup):
            sc = xn @ q32[i]
        t0 = time.perf_counter()
        for i in range(warmup, warmup + steps):
            for _ in range(args.queries):
                sc = xn @ q32[i]
                top = np.argpartition(-sc, args.k)[:args.k]
                top = top[np.argsort(-sc[top], kind="stable")]
        dt_fast = (time.perf_counter() - t0) / max(1, steps)
        out["best_effort_cpu"] = {"value": args.queries / (dt_fast * scale), "unit": UNIT,
                                  "what": "pre-normalised float32 X @ q (BLAS, all threads) + argpartition, same sample and scaling; not the reference's arithmetic"}
    return out


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warmup = max(1, min(args.warmup, 3))
    cb = cpu_reference_qps(args, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": cb["ms_per_step_scaled"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"exact top-{args.k} cosine, {args.rows}x{args.dim} bf16 corpus, {args.queries} query/step",
                   "rows": args.rows, "dim": args.dim, "k": args.k, "queries_per_step": args.queries,
                   "engine": cb["kind"] if cb["kind"] != "port" else
                             "qdrant-client local-mode restatement (oracle/qdrant_local.py); qdrant-client itself is not installable here"},
        "cpu_baseline": {k: cb[k] for k in CPU_KEYS if k in cb},
        "extrapolated": cb["extrapolated"],
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, throttle reasons and board power through NVML every few ms while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        self.index = index
        self.period_s = period_s
        self.samples: list[tuple[float, int]] = []
        self.power_w: list[float] = []
        self.max_mhz = None
        self.power_limit_w = None
        self._stop = threading.Event()
        self._thread = None
        self._err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except (ValueError, IndexError):
                    idx = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self._err = repr(e)
            return
        try:
            self.power_limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
        except Exception:  # noqa: BLE001
            pass

        def loop():
            while not self._stop.is_set():
                try:
                    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    try:
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:  # noqa: BLE001
                        rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((float(mhz), int(rs)))
                except Exception as e:  # noqa: BLE001
                    self._err = repr(e)
                    return
                try:
                    self.power_w.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(self.period_s)
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [f"no samples: {self._err}"], "samples": 0}
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        seen = 0
        for _, r in self.samples:
            seen |= r
        out = {"sm_mhz": statistics.median(m for m, _ in self.samples), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(n for b_, n in bits.items() if seen & b_), "samples": len(self.samples)}
        if self.power_w:      # what the power cap is about: board power next to the enforced limit (the K2 leg runs into it)
            out["power_w"] = statistics.median(self.power_w)
            out["power_limit_w"] = self.power_limit_w
        return out


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------
class _LazyIds:
    """Synthetic host metadata for the adapter leg: the point in row r has id r (qdrant ids are unsigned ints or uuids) and the
    payload {"row": r}.  A real 10M-point store would hold 10M Python strings and dicts; the leg measures the adapter's call path
    (list -> ndarray, asyncio.to_thread, the C ABI, payload dict copies), not Python's memory allocator."""

    def __init__(self, n): self.n = n
    def __len__(self): return self.n
    def __getitem__(self, r): return int(r)
    def get(self, pid, default=None): return int(pid) if isinstance(pid, int) and 0 <= pid < self.n else default
    def __contains__(self, pid): return self.get(pid) is not None


class _LazyPayloads(_LazyIds):
    def __getitem__(self, r): return {"row": int(r), "file_path": f"f{int(r) % 50000}.py", "entity_name": f"e{int(r)}"}


def build_shard(args, torch, dev, rank, world, rows_total, name):
    """This rank's row shard of the synthetic corpus: unit-norm gaussian rows rounded to bf16, generated on the GPU chunk by
    chunk (seeded per global chunk, so the corpus does not depend on the number of shards)."""
    from code_rag_b200.collection import DeviceCollection
    from code_rag_b200.sharded import shard_bounds
    lo, hi = shard_bounds(rows_total, world, align=CHUNK_ROWS if rows_total % (CHUNK_ROWS * world) == 0 else 1)[rank]
    # global row = shard << 32 | local row when the collection is sharded (what the multi-GPU adapter expects)
    shard = DeviceCollection(name, args.dim, storage=args.storage, metric="cosine", n_filter_cols=0, capacity=hi - lo,
                             row_base=(rank << 32) if world > 1 else 0, device=dev.index, timing=False)
    if args.stage_kb:
        shard.set_option("stage_kb", args.stage_kb)
    if args.stages:
        shard.set_option("stages", args.stages)
    t0 = time.perf_counter()
    row = lo
    while row < hi:
        n = min(CHUNK_ROWS - (row % CHUNK_ROWS), hi - row)
        g = torch.Generator(device=dev)
        g.manual_seed(3456 * 1_000_003 + row // CHUNK_ROWS)
        full = torch.randn((CHUNK_ROWS, args.dim), generator=g, device=dev, dtype=torch.float32)
        x = full[row % CHUNK_ROWS: row % CHUNK_ROWS + n]
        x = x / x.norm(dim=1, keepdim=True)
        xb = x.to(torch.bfloat16).contiguous() if args.storage == "bf16" else x.contiguous()
        torch.cuda.synchronize()
        shard.upsert_device(xb.data_ptr(), "bf16" if args.storage == "bf16" else "f32", n, shard.row_base + (row - lo))
        row += n
        del full, x, xb
    torch.cuda.synchronize()
    return shard, hi - lo, time.perf_counter() - t0


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from code_rag_b200 import build as lvs_build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (one process per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        lvs_build.build()
    if world > 1:
        dist.barrier()
    from code_rag_b200.sharded import ShardedSearcher

    dev = torch.device("cuda", local_rank)
    peaks = {}
    pk_file = ROOT / "MEASURED_PEAKS.json"
    if pk_file.exists():
        peaks = json.loads(pk_file.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tc_peak = float(peaks.get("bf16_tflops", 1590.0)) * (1.0 if args.storage == "bf16" else 0.5)
    tc_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    row_bytes = args.dim * (2 if args.storage == "bf16" else 4)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def legs(shard, searcher, n_local, Q, k, K, W, seed, with_sync=True):
        """The measurements of one workload (Q queries per step, top-k) on the resident shard.  Returns a dict of max-over-ranks
        numbers; `sampler` clocks are taken during the value leg."""
        tensor_path = Q >= ((2 if n_local >= 2_000_000 else 3) if args.storage == "bf16" else 5)   # lvs_api.cu gemm_eligible
        qs = make_queries((K + W) * Q, args.dim, seed=seed).reshape(K + W, Q, args.dim)
        dq_all = torch.from_numpy(qs).to(dev)          # resident queries for the `value` leg
        NSLOT, stream = searcher.n_slots, searcher.stream
        # ---- roofline leg: per-launch kernel time by CUDA events around every scan / tensor-core launch (timing on serialises
        #      consecutive searches: this is the kernel ALONE - on the scan path one launch that prepares the query, scans, rescores
        #      and, at N > 1, exchanges and merges)
        KR = min(K, 30)
        shard.set_option("timing", 1)
        torch.cuda.synchronize()
        for i in range(max(W, NSLOT)):                  # every result slot exists before anything is timed
            searcher.search_device_async(dq_all[i % (K + W)], k, slot=i % NSLOT)
        barrier()
        for i in range(W, W + KR):
            searcher.search_device_async(dq_all[i], k, slot=i % NSLOT)
        barrier()
        kernel_ms = float(np.mean(shard.scan_times(KR)[0]))
        shard.set_option("timing", 0)
        # ---- value: device-resident queries, K searches enqueued back to back (each one is a full pass over the corpus; results
        #      stay in HBM; consecutive searches overlap by programmatic dependent launch)
        for i in range(W):
            searcher.search_device_async(dq_all[i], k, slot=i % NSLOT)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        flag_bufs = {}
        e0.record(stream)
        for i in range(W, W + K):
            out = searcher.search_device_async(dq_all[i], k, slot=i % NSLOT)
            flag_bufs[id(out[4])] = out[4]
            launches += shard.last_timing()["launches"]       # kernels the library launched for this search
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
        n_flagged = int(sum(int((f != 0).sum().item()) for f in flag_bufs.values()))
        launches += searcher.merge_launches if world > 1 else 0
        clocks = sampler.stop() if rank == 0 else None
        # ---- sync: one search at a time, the host waits for each result (latency-bound)
        sync_ms = 0.0
        if with_sync:
            barrier()
            t0 = time.perf_counter()
            for i in range(W, W + K):
                searcher.search_device(dq_all[i], k)
            barrier()
            sync_ms = (time.perf_counter() - t0) * 1e3
        # ---- e2e: host buffers through the public API, every step H2D(query) + D2H(result).  N=1: the C ABI's pipelined pair
        #      lvs_search_submit / lvs_search_wait (host pointers in, host pointers out); N>1: ShardedSearcher.submit / wait (the same
        #      plus the exchange).  Two searches in flight, so the host work of step i+1 overlaps the scan of step i.
        DEPTH = 2
        submit = (lambda qh: searcher.submit(qh, k)) if world > 1 else (lambda qh: shard.search_submit(qh, k))
        wait = searcher.wait if world > 1 else shard.search_wait
        for i in range(W):
            wait(submit(qs[i % (K + W)]))
        # every result slot (pinned + device staging buffers are allocated at a slot's first use) exists before the clock starts:
        # the library hands out the first free slot, so all of them have to be in flight at once
        for h in [submit(qs[i % (K + W)]) for i in range(NSLOT)]:
            wait(h)
        barrier()
        t0 = time.perf_counter()
        inflight, done_at = [], []
        for i in range(W, W + K):
            inflight.append(submit(qs[i]))
            if len(inflight) >= DEPTH:
                wait(inflight.pop(0))
                done_at.append(time.perf_counter())
        while inflight:
            wait(inflight.pop(0))
            done_at.append(time.perf_counter())
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        gaps = np.diff(np.array(done_at)) * 1e3         # time between consecutive completions: the median ignores a host hiccup
        barrier()
        t0 = time.perf_counter()
        each = []
        for i in range(W, W + K):                       # depth 1 (strict request / response)
            t1 = time.perf_counter()
            wait(submit(qs[i]))
            each.append((time.perf_counter() - t1) * 1e3)
        barrier()
        e2e1_ms = (time.perf_counter() - t0) * 1e3
        dev_ms, e2e_ms, e2e1_ms, kernel_ms, sync_ms, n_flagged = allmax(dev_ms, e2e_ms, e2e1_ms, kernel_ms, sync_ms, float(n_flagged))
        algo_bytes = n_local * row_bytes
        gbs = algo_bytes / (kernel_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "gemm_topk_kernel" if tensor_path else "scan_topk_kernel (query prep + scan + exact rescoring" +
                                                                                 (" + exchange + merge)" if world > 1 else ")"),
                "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
        tfile = ROOT / "profiles" / ("gemm_traffic.json" if tensor_path else "scan_traffic.json")
        if tfile.exists():
            try:   # ncu --set full capture of this kernel on a 15.36 GB shard; other shard sizes scale with the rows read
                tj = json.loads(tfile.read_text())
                per = tj["dram_bytes_per_launch"]
                if tensor_path:
                    per = per[("pair_q256_bf16" if Q > 128 else "single_q128_bf16") if args.storage == "bf16" else "single_q128_tf32"]
                roof["traffic"] = int(per * (algo_bytes / tj["algorithmic_bytes_per_launch"]))
                roof["traffic_source"] = f"ncu capture in profiles/{tfile.name}, scaled by rows"
            except Exception:  # noqa: BLE001
                pass
        if Q > 128:     # above the ridge (2 Q flop per corpus byte): the tensor pipe bounds the kernel
            flops = 2.0 * min(Q, 256) * n_local * args.dim
            tfl = flops / (kernel_ms * 1e-3) / 1e12
            roof.update({"bound": "tensor", "achieved": tfl, "peak": tc_peak, "unit": "TFLOP/s", "frac": tfl / tc_peak,
                         "frac_of_sustained_peak": tfl / tc_sustained, "algorithmic_flops_per_launch": flops,
                         "hbm_gbs_over_algorithmic_bytes": gbs,
                         "peak_source": ("MEASURED_PEAKS.json bf16_tflops burst" if peaks else "fallback 1590 TFLOP/s") +
                                        ("" if args.storage == "bf16" else " x 0.5 (tf32)")})
        return {"value": K * Q / (dev_ms * 1e-3), "ms_per_step": dev_ms / K, "kernel_ms": kernel_ms, "roofline": roof,
                "sync_qps": K * Q / (sync_ms * 1e-3) if sync_ms else None, "sync_ms_per_step": sync_ms / K if sync_ms else None,
                "e2e": {"value": K * Q / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": Q * args.dim * 8,
                        "d2h_bytes_per_step": Q * k * 24 + Q * 8, "ms_per_step": e2e_ms / K, "in_flight": DEPTH,
                        "api": "lvs_search_submit/lvs_search_wait (C ABI, host buffers)" if world == 1 else "ShardedSearcher.submit/wait (host buffers)",
                        "depth1_qps": K * Q / (e2e1_ms * 1e-3), "depth1_ms_per_step": e2e1_ms / K,
                        "this_rank": {"p50_ms_between_completions": float(np.median(gaps)) if len(gaps) else None,
                                      "max_ms_between_completions": float(gaps.max()) if len(gaps) else None,
                                      "depth1_p50_ms": float(np.median(each)), "depth1_max_ms": float(max(each))}},
                "unproven_queries": int(n_flagged), "gpu_launches": launches, "clocks": clocks, "steps": K, "warmup": W,
                "queries_per_step": Q, "k": k}

    # ========================= the metric workload: 10M x 768 bf16, Q queries per step, top-k =========================
    shard, n_local, t_gen = build_shard(args, torch, dev, rank, world, args.rows, f"bench_r{rank}")
    searcher = ShardedSearcher(shard)
    Q, K, W, k = args.queries, args.steps, max(args.warmup, 3), args.k
    main_leg = legs(shard, searcher, n_local, Q, k, K, W, seed=11)

    # configs[2] (C3) beside the headline at every N: 256 queries x top-100 on the tensor-core path, same corpus
    batched = None
    if Q == 1 and args.storage == "bf16" and not args.no_batched:
        b = legs(shard, searcher, n_local, 256, 100, 6, 3, seed=12, with_sync=False)
        batched = {"workload": f"C3: 256 queries x top-100 per step, {args.rows}x{args.dim} bf16 over {world} GPU(s)",
                   "qps": b["value"], "ms_per_batch": b["ms_per_step"], "gemm_topk_kernel_ms": b["kernel_ms"], "roofline": b["roofline"],
                   "e2e": b["e2e"], "flagged": b["unproven_queries"], "gpu_launches": b["gpu_launches"], "steps": b["steps"],
                   "clocks": b["clocks"]}

    # ---------------- e2e_adapter: the call lattice makes - await store.search(collection=, query_vector=list, limit=10) ----------------
    # (reference query/vector_search.py:60-116 -> QdrantManager.search).  The store is put in front of the SAME resident shard(s);
    # ids and payloads are synthetic and lazy (see _LazyIds).  N=1: B200VectorStore; N>1: ShardedB200VectorStore on rank 0, the other
    # ranks serve their shard (commands through the shared-memory mailbox).
    adapter = None
    if Q == 1 and not args.no_adapter:
        adapter = adapter_leg(args, torch, dist, shard, searcher, n_local, rank, world, K, W, k)
        shard = searcher = None           # at N > 1 the plane's shutdown has closed them
    else:
        searcher.close()
        shard.close()
        shard = searcher = None

    # ---------------- configs[4] (C5): 100M x 768 bf16 row-sharded over the N GPUs (N >= 2: 153.6 GB do not fit one) ----------------
    c5 = None
    if world > 1 and Q == 1 and args.storage == "bf16" and not args.no_c5 and args.rows == 10_000_000:
        torch.cuda.empty_cache()
        sh5, n5, t5 = build_shard(args, torch, dev, rank, world, 100_000_000, f"c5_r{rank}")
        se5 = ShardedSearcher(sh5)
        q1 = legs(sh5, se5, n5, 1, 10, 12, 3, seed=13, with_sync=False)
        q256 = legs(sh5, se5, n5, 256, 10, 3, 3, seed=14, with_sync=False)
        c5 = {"workload": f"C5: 100000000x{args.dim} bf16 row-sharded over {world} GPUs, top-10", "rows_per_gpu": n5, "corpus_gen_s": round(t5, 1),
              "q1": {kk: q1[kk] for kk in ("value", "ms_per_step", "kernel_ms", "roofline", "e2e", "unproven_queries", "steps", "clocks")},
              "q256": {kk: q256[kk] for kk in ("value", "ms_per_step", "kernel_ms", "roofline", "e2e", "unproven_queries", "steps", "clocks")}}
        se5.close()
        sh5.close()

    if rank == 0:
        m = main_leg
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.storage == "f32" else "bf16->f32 scan, f64 rescoring", "data": "synthetic",
            "config": {"workload": f"exact top-{k} cosine, {args.rows}x{args.dim} {args.storage} corpus, {Q} query/step, "
                                   f"row-sharded over {world} GPU(s)",
                       "rows": args.rows, "dim": args.dim, "k": k, "queries_per_step": Q, "storage": args.storage,
                       "rows_per_gpu": n_local, "l2": "inputs_exceed_l2", "corpus_gen_s": round(t_gen, 1),
                       "parallelism": f"row-shard x{world}; per search ONE kernel per GPU: scan + exact rescoring" +
                                      (" + peer-memory exchange + merge" if world > 1 else ""),
                       "value_mode": "K searches enqueued back to back on one stream (device-resident queries/results), "
                                     "consecutive searches overlapped by programmatic dependent launch",
                       "sync_qps": m["sync_qps"], "sync_ms_per_step": m["sync_ms_per_step"],
                       "unproven_queries": m["unproven_queries"]},
            "e2e": m["e2e"],
            "gpu_launches": m["gpu_launches"],
            "roofline": m["roofline"],
            "clocks": m["clocks"],
        }
        if batched is not None:
            line["batched"] = batched
        if adapter is not None:
            line["e2e_adapter"] = adapter
        if c5 is not None:
            line["c5"] = c5
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_qps(args, steps=8, warmup=2)
            line["cpu_baseline"] = {kk: cb[kk] for kk in CPU_KEYS if kk in cb}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def adapter_leg(args, torch, dist, shard, searcher, n_local, rank, world, K, W, k):
    """QPS of `await store.search(collection="code_chunks", query_vector=<list of floats>, limit=k)` on the resident corpus."""
    import asyncio

    from code_rag_b200 import client as C_
    from code_rag_b200.client import B200VectorStore, CollectionName, _HostCollection
    CODE = CollectionName.CODE_CHUNKS.value
    qs = make_queries(K + W, args.dim, seed=15)
    qlists = [q.tolist() for q in qs]                    # what an embedding provider hands to lattice: list[float]

    async def timed(store):
        for i in range(W):
            await store.search(collection=CODE, query_vector=qlists[i], limit=k)
        t0 = time.perf_counter()
        for i in range(W, W + K):
            hits = await store.search(collection=CODE, query_vector=qlists[i], limit=k)
        dt = time.perf_counter() - t0
        assert len(hits) == k and all("payload" in h and "score" in h and "id" in h for h in hits)
        # concurrent awaits, as QueryEngine does with asyncio.gather (query/engine.py:142-146): calls serialise per collection
        t0 = time.perf_counter()
        await asyncio.gather(*[store.search(collection=CODE, query_vector=qlists[i], limit=k) for i in range(W, W + K)])
        dtc = time.perf_counter() - t0
        return dt, dtc

    def host_half(dev_factory, n_rows):
        hc = _HostCollection(CODE, args.dim, args.storage, ["file_path", "entity_type", "language", "content_hash", "project_name"], 0,
                             dev_factory=dev_factory)
        hc.ids = _LazyIds(n_rows); hc.id_to_row = _LazyIds(n_rows); hc.payloads = _LazyPayloads(n_rows)
        return hc

    if world == 1:
        store = B200VectorStore(dimensions=args.dim, storage=args.storage)
        asyncio.run(store.connect())
        store._collections[CODE] = host_half(lambda *a, **kw: shard, n_local)
        dt, dtc = asyncio.run(timed(store))
        C_._POLLED_SEARCH = False                        # the same calls through a worker thread (asyncio.to_thread), for comparison
        dt_thread, dtc_thread = asyncio.run(timed(store))
        C_._POLLED_SEARCH = True
        searcher.close()
        asyncio.run(store.close())                       # closes the shard
        api = "B200VectorStore.search (asyncio, list[float] in, list[dict] out)"
    else:
        from code_rag_b200 import sharded_store as SS
        plane = SS.ShardPlane.start()                    # collective; re-uses the NCCL group, opens the gloo control group + mailbox
        plane.shards[CODE], plane.searchers[CODE] = shard, searcher
        dt = dtc = dt_thread = dtc_thread = 0.0
        if rank != 0:
            plane.serve()                                # until rank 0 shuts the plane down
        else:
            store = SS.ShardedB200VectorStore(dimensions=args.dim, storage=args.storage, plane=plane)
            asyncio.run(store.connect())
            coll = SS._ShardedHostCollection.__new__(SS._ShardedHostCollection)
            coll.plane, coll.name, coll.dim, coll.storage = plane, CODE, args.dim, args.storage
            coll.columns, coll.dicts = ["file_path", "entity_type", "language", "content_hash", "project_name"], [dict() for _ in range(5)]
            coll.tie_counts, coll.dup_keys = {}, {}
            from code_rag_b200.client import _CollectionLock
            coll.lock = _CollectionLock()
            coll.shards = []
            for s_ in range(world):
                hs = SS._HostShard(coll, s_)
                hs.ids = _LazyIds(n_local); hs.id_to_row = _LazyIds(n_local); hs.payloads = _LazyPayloads(n_local)
                coll.shards.append(hs)
            store._collections[CODE] = coll
            dt, dtc = asyncio.run(timed(store))
            C_._POLLED_SEARCH = False
            dt_thread, dtc_thread = asyncio.run(timed(store))
            C_._POLLED_SEARCH = True
            store._collections.clear()
            plane.shutdown()                             # the workers leave serve(); every rank closes its shard
        api = "ShardedB200VectorStore.search on rank 0 (asyncio; search commands through the shared-memory mailbox)"
    if rank != 0:
        return None
    return {"value": K / dt, "unit": UNIT, "ms_per_call": dt / K * 1e3, "api": api, "calls": K,
            "gathered_qps": K / dtc, "through_a_worker_thread": {"ms_per_call": dt_thread / K * 1e3, "gathered_qps": K / dtc_thread},
            "metadata": "synthetic lazy ids / payloads over the resident shard(s)",
            "h2d_bytes_per_step": args.dim * 8, "d2h_bytes_per_step": k * 24 + 8}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
