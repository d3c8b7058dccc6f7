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
    # the same calls without CUDA events around the kernel: the search completes through the word its kernel stores into the pinned slot
    dev.set_option("timing", 0)
    names = ("queries_ready", "scanned", "list_written", "lists_visible", "selected", "rescored", "stored", "sel_loaded", "sel_threshold",
             "sel_gathered", "scores_collected", "ordered", "host_told", "rows_staged", "replayed")
    variants = {}
    for label, inline in (("query_in_kernel_params", 1), ("query_staged_by_cta0", 0)):
        dev.set_option("inline_query", inline)
        w = []
        for i in range(230):
            t0 = time.perf_counter()
            dev.search(qs[i % len(qs)], 10)
            if i >= 30:
                w.append((time.perf_counter() - t0) * 1e3)
        dev.set_option("dbg_times", 1)
        ph = []
        for i in range(20):
            dev.search(qs[i % len(qs)], 10)
            ph.append(dev.last_kernel_phases())
        dev.set_option("dbg_times", 0)
        variants[label] = {"wall_ms": statistics.median(w), "wall_ms_p10": sorted(w)[len(w) // 10],
                           "kernel_phases_us": dict(zip(names, [round(float(v), 2) for v in np.median(np.array(ph), axis=0)]))}
    dev.set_option("inline_query", 1)
    walls = [variants["query_in_kernel_params"]["wall_ms"]]
    # the C-ABI call alone (include/lvs.h: lvs_search), arguments prepared once: what a compiled host pays
    from code_rag_b200 import _native as N
    from code_rag_b200.collection import _result_block
    q0 = np.ascontiguousarray(qs[0], dtype=np.float64)
    block, (ps, pr, pt, pc, pf) = _result_block(1, 10)
    lib, h, qa = dev._lib, dev._handle(), q0.ctypes.data
    cw = []
    for i in range(330):
        t0 = time.perf_counter()
        rc = lib.lvs_search(h, qa, N.DT_F64, 1, 10, None, ps, pr, pt, pc, pf)
        if i >= 30:
            cw.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    variants["c_abi_call_only"] = {"wall_ms": statistics.median(cw), "wall_ms_p10": sorted(cw)[len(cw) // 10]}
    phases = variants["query_in_kernel_params"]["kernel_phases_us"]
    # the same collection behind the QdrantManager drop-in: `await store.search(collection=, query_vector=list, limit=10)` with ids and
    # payload dicts, filtered and not (what lattice's VectorSearcher issues, query/vector_search.py:97-128)
    import asyncio
    import uuid

    from code_rag_b200.client import B200VectorStore

    async def adapter():
        st = B200VectorStore(dimensions=dim, storage="f32")
        await st.connect(); await st.create_collections()
        ids = [str(uuid.UUID(int=i + 1)) for i in range(n)]
        pl = [{"file_path": f"src/f{i % 400}.py", "entity_type": "function", "entity_name": f"fn{i}", "language": "python",
               "content_hash": "h", "project_name": f"p{i % 4}", "content": "def f(): pass"} for i in range(n)]
        xv = x.astype(np.float64)
        for s0 in range(0, n, 2000):
            await st.upsert(collection="code_chunks", ids=ids[s0:s0 + 2000], vectors=xv[s0:s0 + 2000].tolist(), payloads=pl[s0:s0 + 2000])
        ql = [q[0].tolist() for q in qs]
        res = {}
        for label, flt in (("no_filter", None), ("project_filter", {"project_name": "p1"})):
            w = []
            for i in range(260):
                t0 = time.perf_counter()
                hits = await st.search(collection="code_chunks", query_vector=ql[i % len(ql)], limit=10, filters=flt)
                if i >= 60:
                    w.append((time.perf_counter() - t0) * 1e3)
            assert len(hits) == 10
            res[label] = {"ms_per_await": statistics.median(w), "p10": sorted(w)[len(w) // 10]}
        await st.close()
        return res
    variants["adapter_await_search"] = asyncio.run(adapter())
    emit("C1 10k x 768 fp32, Q=1, top-10 (L2-resident: latency config)", wall_ms=wall, scan_ms=scan, device_ms=total, kernel_phases_us=phases,
         wall_ms_no_events=walls[0], wall_ms_no_events_p10=variants["query_in_kernel_params"]["wall_ms_p10"], no_events=variants,
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


def c4f(Q=64, k=100, n=10_000_000, G=64):
    """C4 through the fused C-ABI call (lvs_search_rank): search + on-device candidate build + K3, host buffers in and out.
    Ranking attributes are synthetic per-row columns; graph candidates arrive as packed arrays (what the host packs from a
    GraphContext); the ranked lists come back as arrays (index, score, signals)."""
    import ctypes as C

    from code_rag_b200 import _native as N
    dim = 768
    dev = DeviceCollection("c4f", dim, storage="bf16", capacity=n)
    fill(dev, n, dim, "bf16", seed=3456)
    n_names = 100_000
    first = dev.rank_names_append([f"e{i}".encode() for i in range(n_names)])
    assert first == 0
    t0 = time.perf_counter()
    for lo in range(0, n, 1_000_000):
        r = np.arange(lo, min(n, lo + 1_000_000), dtype=np.int64)
        dev.rank_attrs_set(r, r.astype(np.uint32), (r // 20).astype(np.uint32), r.astype(np.uint32), (r % n_names).astype(np.uint32),
                           (10 + r % 300).astype(np.int32), np.full(len(r), 8, dtype=np.uint8))
    t_attr = time.perf_counter() - t0
    rng = np.random.default_rng(4567)
    walls, sms, rms = [], [], []
    for rep in range(5):
        qs = rng.standard_normal((Q, dim))
        ng = Q * G
        arr = {
            "offsets": (np.arange(Q + 1) * G).astype(np.int32), "kind": rng.choice([0, 1, 2, 3], size=ng).astype(np.uint8),
            "key_id": rng.integers(0, n, size=ng).astype(np.uint32), "file_id": rng.integers(0, n // 20, size=ng).astype(np.uint32),
            "depth": rng.integers(1, 4, size=ng).astype(np.int32), "entity_match": rng.choice([0.0, 0.5, 1.0], size=ng),
            "degree": rng.poisson(12, size=ng).astype(np.int32), "flags": rng.integers(0, 8, size=ng).astype(np.uint8),
            "weights": np.tile(np.array([0.5, 0.5, 0.2, 0.1]), (Q, 1)),
        }
        rb = N.RankBatch(); rb.n_queries = Q
        for name, a in arr.items():
            setattr(rb, name, a.ctypes.data_as(C.c_void_p))
        ents = [f"e{int(x)}".encode() for x in rng.integers(0, n_names, size=Q)]
        cx_arr = {"ent_off": np.arange(Q + 1, dtype=np.int32), "ent_str_off": np.concatenate([[0], np.cumsum([len(e) for e in ents])]).astype(np.uint32),
                  "ent_bytes": np.frombuffer(b"".join(ents), dtype=np.uint8), "cen_off": (np.arange(Q + 1) * 10).astype(np.int32),
                  "cen_id": rng.integers(0, n, size=Q * 10).astype(np.uint32), "cen_deg": rng.poisson(12, size=Q * 10).astype(np.int32)}
        cx = N.RankQueryCtx()
        for name, a in cx_arr.items():
            setattr(cx, name, a.ctypes.data_as(C.c_void_p))
        t1 = time.perf_counter()
        out = dev.search_rank(qs, k, None, rb, cx, ng, 5, 50, 0.3, 0.15)
        walls.append((time.perf_counter() - t1) * 1e3); sms.append(out["search_ms"]); rms.append(out["rank_ms"])
    w = statistics.median(walls[1:])
    emit(f"C4 fused (lvs_search_rank) 10M x 768 bf16, Q={Q}, vector top-{k} + {G} graph candidates", wall_ms=w,
         search_device_ms=statistics.median(sms[1:]), gather_rank_device_us=statistics.median(rms[1:]) * 1e3, fused_qps=Q * 1e3 / w,
         results_per_query=int(out["count"][0]), flagged=int((out["flags"] & 1).sum()), attr_upload_s=t_attr)
    dev.close()


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2"]
    torch.cuda.init()
    for w in which:
        {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c4f": c4f}[w]()
