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
    """CPU twin of the corpus generator for the bounded CPU sample (same distribution, numpy RNG)."""
    rng = np.random.default_rng([seed, chunk_id])
    x = rng.standard_normal((rows, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return bf16_round_np(x)


# --------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (qdrant-client local mode, restated in oracle/qdrant_local.py)
# --------------------------------------------------------------------------------------------------------
def cpu_reference_qps(args, steps: int, warmup: int) -> dict:
    """Times the oracle on a bounded row sample and scales linearly to the full corpus (the work is a dense
    O(N*D) scan + sort, linear in N).  Uses every host thread numpy's BLAS will take."""
    from oracle.qdrant_local import OracleCollection
    n = min(args.cpu_sample_rows, args.rows)
    x = host_chunk(0, n, args.dim)
    ora = OracleCollection(args.dim)
    ora.upsert_rows_f32(0, x, [None] * n)
    qs = make_queries(steps + warmup, args.dim, seed=11)
    for i in range(warmup):
        ora.search_topk_rows(qs[i], args.k)
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        for _ in range(args.queries):
            ora.search_topk_rows(qs[i], args.k)
    dt = (time.perf_counter() - t0) / max(1, steps)
    scale = args.rows / n
    ms_full = dt * 1e3 * scale
    # a CPU path that is NOT the reference's arithmetic, only the fastest plain-numpy way to the same ids: rows normalised once,
    # float32 matrix-vector product (BLAS, all threads), argpartition + sort of the k winners.  Reported so that the GPU/CPU ratio
    # is not inflated by local mode's per-search re-normalisation, float64 product and full argsort.
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
    return {
        "best_effort_cpu": {"value": args.queries / (dt_fast * scale), "unit": UNIT,
                            "what": "pre-normalised float32 X @ q (BLAS, all threads) + argpartition, same sample and scaling; not the reference's arithmetic"},
        "value": args.queries / (dt * scale), "unit": UNIT, "cores": blas_threads(), "kind": "port",
        "sample": f"{n} of {args.rows} rows x {args.dim} (bf16-rounded, fp32 in RAM), {steps} searches after {warmup} warm-up; "
                  f"time scaled x{scale:.0f} (scan+sort is linear in rows); numpy {np.__version__}, {blas_threads()} BLAS threads",
        "ms_per_step_sample": dt * 1e3, "ms_per_step_scaled": ms_full,
    }


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
                   "engine": "qdrant-client local-mode restatement (oracle/qdrant_local.py); qdrant-client itself is not installable here"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "best_effort_cpu")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every ~10 ms while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[float, int]] = []
        self.max_mhz = None
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
                time.sleep(0.01)
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
        return {"sm_mhz": statistics.median(m for m, _ in self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b_, n in bits.items() if seen & b_), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------
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
    from code_rag_b200.collection import DeviceCollection
    from code_rag_b200.sharded import ShardedSearcher, shard_bounds

    dev = torch.device("cuda", local_rank)
    lo, hi = shard_bounds(args.rows, world, align=CHUNK_ROWS if args.rows % (CHUNK_ROWS * world) == 0 else 1)[rank]
    n_local = hi - lo
    shard = DeviceCollection(f"bench_r{rank}", args.dim, storage=args.storage, metric="cosine", n_filter_cols=0,
                             capacity=n_local, row_base=lo, device=local_rank)
    if args.stage_kb:
        shard.set_option("stage_kb", args.stage_kb)
    if args.stages:
        shard.set_option("stages", args.stages)
    # synthetic corpus: unit-norm gaussian rows rounded to bf16, generated on the GPU chunk by chunk (seeded per
    # global chunk so the corpus does not depend on the number of shards)
    t_gen = time.perf_counter()
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
        shard.upsert_device(xb.data_ptr(), "bf16" if args.storage == "bf16" else "f32", n, row)
        row += n
        del full, x, xb
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    searcher = ShardedSearcher(shard)

    Q, K, W, k = args.queries, args.steps, args.warmup, args.k
    W = max(W, 3)
    tensor_path = Q >= (3 if args.storage == "bf16" else 5)       # K2 (tcgen05) takes batches from this size on (lvs_api.cu)
    qs = make_queries((K + W) * Q, args.dim, seed=11).reshape(K + W, Q, args.dim)
    dq_all = torch.from_numpy(qs).to(dev)          # resident queries for the `value` leg
    row_bytes = args.dim * (2 if args.storage == "bf16" else 4)
    algo_bytes = n_local * row_bytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- roofline leg: per-launch kernel time by CUDA events around every scan launch ----------------
    # (timing on serialises consecutive searches, so this is the kernel ALONE: query prep + scan + exact rescoring, and at
    #  N > 1 the exchange wait + merge, all one launch; the value leg below runs with it off)
    NSLOT = searcher.n_slots
    stream = searcher.stream
    KR = min(K, 30)
    shard.set_option("timing", 1)
    torch.cuda.synchronize()
    for i in range(W):
        searcher.search_device_async(dq_all[i], k, slot=i % NSLOT)
    barrier()
    for i in range(W, W + KR):
        searcher.search_device_async(dq_all[i], k, slot=i % NSLOT)
    barrier()
    ms_ring, bytes_ring = shard.scan_times(KR)
    scan_ms = [float(v) for v in ms_ring]
    shard.set_option("timing", 0)

    # ---------------- value: device-resident queries, searches enqueued back to back ----------------
    # (each search is still one full pass over the corpus; results stay in HBM, flags are checked after the loop; consecutive
    #  searches overlap by programmatic dependent launch: the next one streams while the last CTAs of this one rescore / exchange)
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
        launches += shard.last_timing()["launches"]       # kernels the library launched for this search (1 on the scan path)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    n_flagged = int(sum(int((f != 0).sum().item()) for f in flag_bufs.values()))
    launches += searcher.merge_launches if world > 1 else 0
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- sync: one search at a time, the host waits for each result (latency-bound) ----------------
    barrier()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        searcher.search_device(dq_all[i], k)
    barrier()
    sync_ms = (time.perf_counter() - t0) * 1e3

    # ---------------- e2e: host buffers through the public API, every step H2D(query) + D2H(result) ----------------
    # N=1: the C ABI's pipelined pair lvs_search_submit / lvs_search_wait (host pointers in, host pointers out);
    # N>1: ShardedSearcher.submit / wait (adds the all-gather + merge).  Two searches are kept in flight, so the copies
    # and the host work of step i+1 overlap the scan of step i.
    DEPTH = 2

    def e2e_submit(qh):
        return searcher.submit(qh, k) if world > 1 else shard.search_submit(qh, k)

    def e2e_wait(h):
        return searcher.wait(h) if world > 1 else shard.search_wait(h)

    for i in range(W):
        e2e_wait(e2e_submit(qs[i]))
    barrier()
    t0 = time.perf_counter()
    inflight = []
    last = None
    for i in range(W, W + K):
        inflight.append(e2e_submit(qs[i]))
        if len(inflight) >= DEPTH:
            last = e2e_wait(inflight.pop(0))
    while inflight:
        last = e2e_wait(inflight.pop(0))
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # depth 1 (strict request/response latency)
    barrier()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        last = e2e_wait(e2e_submit(qs[i]))
    barrier()
    e2e1_ms = (time.perf_counter() - t0) * 1e3

    t_dev = torch.tensor([dev_ms, e2e_ms, statistics.mean(scan_ms), sync_ms, float(n_flagged)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, scan_mean, sync_ms, n_flagged = (float(v) for v in t_dev.cpu())

    if rank == 0:
        peaks = {}
        pk_file = ROOT / "MEASURED_PEAKS.json"
        if pk_file.exists():
            peaks = json.loads(pk_file.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = algo_bytes / (scan_mean * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "scan_traffic.json"
        if tf.exists() and Q == 1:
            try:   # ncu capture of this kernel on the 10M-row shard; other shard sizes scale with the rows scanned
                tj = json.loads(tf.read_text())
                traffic = int(tj["dram_bytes_per_launch"] * (algo_bytes / tj["algorithmic_bytes_per_launch"]))
            except Exception:
                traffic = None
        gf = ROOT / "profiles" / "gemm_traffic.json"
        if tensor_path and gf.exists():
            try:   # ncu capture of the tensor-core kernel on a 15.36 GB shard; other shard sizes scale with the rows read
                gj = json.loads(gf.read_text())
                key = ("pair_q256_bf16" if Q > 128 else "single_q128_bf16") if args.storage == "bf16" else "single_q128_tf32"
                traffic = int(gj["dram_bytes_per_launch"][key] * (algo_bytes / gj["algorithmic_bytes_per_launch"]))
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "kernel": "gemm_topk_kernel" if tensor_path else "scan_topk_kernel", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": scan_mean,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"}
        if Q > 128:     # above the ridge (2 Q flop per corpus byte): the tensor pipe bounds the kernel
            tpeak = float(peaks.get("bf16_tflops", 1649.5)) * (1.0 if args.storage == "bf16" else 0.5)
            tfl = 2.0 * min(Q, 256) * n_local * args.dim / (scan_mean * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_topk_kernel", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak,
                    "traffic": traffic, "algorithmic_flops_per_launch": 2.0 * min(Q, 256) * n_local * args.dim, "kernel_ms": scan_mean,
                    "hbm_gbs_over_algorithmic_bytes": achieved,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops burst" if peaks else "fallback 1649.5 TFLOP/s") +
                                   ("" if args.storage == "bf16" else " x 0.5 (tf32)")}
        batched = None
        if world == 1 and Q == 1 and args.storage == "bf16" and not args.no_batched:
            # configs[2] (C3) beside the headline: 256 queries, top-100, tcgen05 path, host buffers through lvs_search
            qb = make_queries(3 * 256, args.dim, seed=12).reshape(3, 256, args.dim)
            shard.set_option("timing", 1)
            walls, gms = [], []
            for r_ in range(3):
                t0 = time.perf_counter()
                rb_ = shard.search(qb[r_], 100)
                walls.append((time.perf_counter() - t0) * 1e3)
                gms.append(shard.last_timing()["scan_ms"])
            w_, g_ = min(walls[1:]), min(gms[1:])
            tfl = 2.0 * 256 * n_local * args.dim / (g_ * 1e-3) / 1e12
            batched = {"workload": f"C3: 256 queries x top-100, {args.rows}x{args.dim} bf16, one lvs_search call (host buffers)",
                       "ms_per_batch": w_, "qps": 256e3 / w_, "gemm_topk_kernel_ms": g_, "tflops": tfl,
                       "frac_of_bf16_burst_peak": tfl / float(peaks.get("bf16_tflops", 1649.5)),
                       "frac_of_bf16_sustained_peak": tfl / float(peaks.get("bf16_tflops_sustained", 1361.6)),
                       "flagged": int(rb_.flags.sum())}
        line = {
            "metric": METRIC, "value": K * Q / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.storage == "f32" else "bf16->f32 scan, f64 rescoring", "data": "synthetic",
            "config": {"workload": f"exact top-{k} cosine, {args.rows}x{args.dim} {args.storage} corpus, {Q} query/step, "
                                   f"row-sharded over {world} GPU(s)",
                       "rows": args.rows, "dim": args.dim, "k": k, "queries_per_step": Q, "storage": args.storage,
                       "rows_per_gpu": n_local, "l2": "inputs_exceed_l2", "corpus_gen_s": round(t_gen, 1),
                       "parallelism": f"row-shard x{world} + top-k exchange ({searcher.exchange_mode}) + merge",
                       "value_mode": "K searches enqueued back to back on one stream (device-resident queries/results)",
                       "sync_qps": K * Q / (sync_ms * 1e-3), "sync_ms_per_step": sync_ms / K,
                       "unproven_queries": int(n_flagged)},
            "e2e": {"value": K * Q / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": Q * args.dim * 8,
                    "d2h_bytes_per_step": Q * k * 24 + Q * 8, "ms_per_step": e2e_ms / K, "in_flight": DEPTH,
                    "api": "lvs_search_submit/lvs_search_wait (C ABI, host buffers)" if world == 1 else "ShardedSearcher.submit/wait",
                    "depth1_qps": K * Q / (e2e1_ms * 1e-3), "depth1_ms_per_step": e2e1_ms / K},
            "gpu_launches": launches,
            "roofline": roof,
            "clocks": clocks,
        }
        if batched is not None:
            line["batched"] = batched
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_qps(args, steps=8, warmup=2)
            line["cpu_baseline"] = {kk: cb[kk] for kk in ("value", "unit", "cores", "kind", "sample", "best_effort_cpu")}
        print(json.dumps(line), flush=True)
    searcher.close()
    shard.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
