#!/usr/bin/env python
"""Timings of BASELINE.json's non-headline configs on one B200 (parity for them lives in tests/).

    python benchmarks/configs.py [c1] [c2] [c3] ...

Each line is JSON: the config, device time per search from CUDA events (scan kernel and whole call), achieved GB/s over
the algorithmic bytes (rows x row_bytes per pass) and QPS.  Corpora are generated on the GPU with torch (seeded).
"""
from __future__ import annotations

import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from code_rag_b200.collection import DeviceCollection  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
ANY = 0xFFFFFFFF


def fill(dev: DeviceCollection, n: int, dim: int, storage: str, seed: int, unit: bool = True, codes=None, chunk=250_000):
    d = torch.device("cuda")
    row = 0
    while row < n:
        m = min(chunk, n - row)
        g = torch.Generator(device=d)
        g.manual_seed(seed * 1_000_003 + row // chunk)
        x = torch.randn((m, dim), generator=g, device=d, dtype=torch.float32)
        if unit:
            x = x / x.norm(dim=1, keepdim=True)
        xs = x.to(torch.bfloat16).contiguous() if storage == "bf16" else x.contiguous()
        cp = 0
        if codes is not None:
            cc = codes[row:row + m].contiguous()
            cp = cc.data_ptr()
        torch.cuda.synchronize()
        dev.upsert_device(xs.data_ptr(), "bf16" if storage == "bf16" else "f32", m, row, codes_ptr=cp)
        row += m


def timed(dev: DeviceCollection, queries: np.ndarray, k: int, want, reps: int, warm: int = 3):
    wall, scan, total = [], [], []
    for i in range(warm + reps):
        q = queries[i % len(queries)]
        t0 = time.perf_counter()
        res = dev.search(q, k, want)
        dt = time.perf_counter() - t0
        if i >= warm:
            t = dev.last_timing()
            wall.append(dt * 1e3); scan.append(t["scan_ms"]); total.append(t["total_ms"])
    return res, statistics.median(wall), statistics.median(scan), statistics.median(total)


def emit(name, **kw):
    print(json.dumps({"config": name, **kw}), flush=True)


def c1():
    n, dim = 10_000, 768
    dev = DeviceCollection("c1", dim, storage="f32")
    rng = np.random.default_rng(1234)
    mu = 0.5 * rng.standard_normal(dim)
    x = (mu + rng.standard_normal((n, dim))).astype(np.float32)
    dev.upsert(x)
    qs = (mu + rng.standard_normal((32, 1, dim)))
    _, wall, scan, total = timed(dev, qs, 10, None, 200)
    emit("C1 10k x 768 fp32, Q=1, top-10 (L2-resident: latency config)", wall_ms=wall, scan_ms=scan, device_ms=total,
         qps=1e3 / wall, bytes_per_pass=n * dim * 4)
    dev.close()


def c2():
    n, dim = 1_000_000, 1536
    d = torch.device("cuda")
    g = torch.Generator(device=d); g.manual_seed(2345)
    proj = torch.multinomial(torch.tensor([.40, .20, .15, .10, .05, .04, .03, .03], device=d), n, True, generator=g).to(torch.int32) + 1
    lang = torch.multinomial(torch.tensor([.6, .25, .15], device=d), n, True, generator=g).to(torch.int32) + 1
    fpath = torch.randint(1, 50_001, (n,), device=d, generator=g, dtype=torch.int32)
    codes = torch.stack([proj, lang, fpath], dim=1).contiguous()      # columns: project_name, language, file_path
    dev = DeviceCollection("c2", dim, storage="f32", n_filter_cols=3, capacity=n)
    fill(dev, n, dim, "f32", seed=2345, codes=codes)
    rng = np.random.default_rng(5)
    qs = rng.standard_normal((16, 1, dim))
    cases = {"none": None, "project=p0 (40%)": [1, ANY, ANY], "project=p7 & language=python (~1.8%)": [8, 1, ANY],
             "file_path=f (~20 rows)": [ANY, ANY, 777]}
    for name, want in cases.items():
        res, wall, scan, total = timed(dev, qs, 10, None if want is None else np.array(want, dtype=np.uint32), 50)
        emit(f"C2 1M x 1536 fp32, Q=1, top-10, filter {name}", wall_ms=wall, scan_ms=scan, device_ms=total, qps=1e3 / wall,
             full_scan_gbs=n * dim * 4 / (scan * 1e-3) / 1e9, frac_of_measured_peak=n * dim * 4 / (scan * 1e-3) / 1e9 / PEAK,
             hits=int(res.counts[0]))
    dev.close()


def c3(Q=256, k=100, n=10_000_000):
    dim = 768
    dev = DeviceCollection("c3", dim, storage="bf16", capacity=n)
    fill(dev, n, dim, "bf16", seed=3456)
    rng = np.random.default_rng(6)
    qs = rng.standard_normal((2, Q, dim))
    res, wall, scan, total = timed(dev, qs, k, None, 3, warm=1)
    t = dev.last_timing()
    emit(f"C3 10M x 768 bf16, Q={Q}, top-{k}", wall_ms=wall, device_ms=total, batch_qps=Q * 1e3 / wall, kernel=t["kernel"],
         launches=t["launches"], flagged=int(res.flags.sum()),
         tflops=2.0 * Q * n * dim / (wall * 1e-3) / 1e12, gbs_equiv=n * dim * 2 / (wall * 1e-3) / 1e9)
    dev.close()


def c4(Q=64, k=100, n=10_000_000, G=64):
    """Hybrid ranking: vector top-100 (K2) fused with synthetic graph-relevance candidates (K3) at batch 64."""
    import random
    from types import SimpleNamespace as NS

    from code_rag_b200.ranking import HybridRanker
    dim = 768
    dev = DeviceCollection("c4", dim, storage="bf16", capacity=n)
    fill(dev, n, dim, "bf16", seed=3456)
    rng = np.random.default_rng(4567)
    prng = random.Random(4567)
    intents = ["find_callers", "find_callees", "find_call_chain", "find_hierarchy", "find_implementations", "find_usages",
               "find_dependencies", "find_dependents", "locate_entity", "locate_file", "explain_implementation",
               "explain_relationship", "explain_data_flow", "explain_architecture", "find_similar", "search_functionality",
               "search_pattern"]
    ranker = HybridRanker()
    for rep in range(3):
        qs = rng.standard_normal((Q, dim))
        t0 = time.perf_counter()
        res = dev.search(qs, k)
        t_search = time.perf_counter() - t0
        tm = dev.last_timing()
        t1 = time.perf_counter()
        items = []
        for qi in range(Q):
            rows, scores = res.rows[qi, :res.counts[qi]], res.scores[qi, :res.counts[qi]]
            vec = [{"score": float(s), "file_path": f"src/f{int(r) // 20}.py", "entity_type": "function", "entity_name": f"e{int(r)}",
                    "content": "c" * (10 + int(r) % 300), "start_line": int(r) % 500, "end_line": int(r) % 500 + 9,
                    "graph_node_id": f"m.e{int(r)}"} for r, s in zip(rows, scores)]
            nodes = []
            for g in range(G):
                if prng.random() < 0.25 and vec:
                    v = prng.choice(vec)
                    nm, fp, sl = v["entity_name"], v["file_path"], v["start_line"]
                else:
                    nm, fp, sl = f"g{qi}_{g}", f"src/g{prng.randrange(40)}.py", 1000 + g
                nodes.append(NS(node_type="Function", name=nm, qualified_name=f"m.{nm}", file_path=fp, signature=prng.choice([None, "s"]),
                                docstring=prng.choice([None, "d"]), summary=prng.choice([None, "x"]), start_line=sl, end_line=sl + 1,
                                metadata={"depth": prng.choice([1, 2, 3])}))
            ctx = NS(primary_entities=nodes[:4], callers=nodes[4:24], callees=nodes[24:44], methods=nodes[44:54],
                     parent_classes=nodes[54:59], child_classes=nodes[59:])
            cent = {f"m.e{int(r)}": {"total_degree": int(rng.poisson(12))} for r in rows[:10]}
            plan = NS(primary_intent=NS(value=intents[qi % len(intents)]), entities=[NS(name=f"e{int(rows[0])}" if len(rows) else "x")])
            items.append((plan, ctx, vec, cent))
        t_prep = time.perf_counter() - t1
        t2 = time.perf_counter()
        ranked = ranker.rank_batch(items)
        t_rank = time.perf_counter() - t2
    emit(f"C4 10M x 768 bf16, Q={Q}, vector top-{k} + {G} graph candidates, hybrid fusion", search_wall_ms=t_search * 1e3,
         search_kernel_ms=tm["scan_ms"], search_finalize_ms=tm["finalize_ms"], kernel=tm["kernel"],
         synthetic_candidate_build_ms=t_prep * 1e3, rank_wall_ms=t_rank * 1e3, rank_kernel_us=ranker.last_device_ms * 1e3,
         fused_qps=Q / (t_search + t_rank), results_per_query=len(ranked[0]))
    dev.close()


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2"]
    torch.cuda.init()
    for w in which:
        {"c1": c1, "c2": c2, "c3": c3, "c4": c4}[w]()
