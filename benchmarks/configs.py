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


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2"]
    torch.cuda.init()
    for w in which:
        {"c1": c1, "c2": c2, "c3": c3}[w]()
