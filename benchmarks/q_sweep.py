#!/usr/bin/env python
"""Batch-size sweep on one B200, 10M x 768 bf16 (or K2_STORAGE=f32 with half the rows): which kernel a batch of Q queries takes
and what it costs.  One JSON line per Q: device time of the whole search (prep + scan/GEMM + finalize) and queries/s."""
import json
import os
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))
from configs import fill  # noqa: E402

from code_rag_b200.collection import DeviceCollection  # noqa: E402

STORAGE = os.environ.get("K2_STORAGE", "bf16")
n = int(sys.argv[1]) if len(sys.argv) > 1 else (10_000_000 if STORAGE == "bf16" else 5_000_000)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = DeviceCollection("sweep", 768, storage=STORAGE, capacity=n)
fill(dev, n, 768, STORAGE, seed=3456)
rng = np.random.default_rng(11)
for Q in (1, 2, 3, 4, 5, 8, 16, 32, 64, 128, 192, 256):
    qs = rng.standard_normal((Q, 768))
    tot, scan, wall = [], [], []
    for rep in range(5):
        t0 = time.perf_counter()
        res = dev.search(qs, k)
        w = (time.perf_counter() - t0) * 1e3
        t = dev.last_timing()
        if rep >= 2:
            tot.append(t["total_ms"]); scan.append(t["scan_ms"]); wall.append(w)
    ms = statistics.median(tot)
    print(json.dumps({"Q": Q, "k": k, "storage": STORAGE, "rows": n, "kernel": t["kernel"], "launches": t["launches"],
                      "first_kernel_ms": statistics.median(scan), "device_ms": ms, "qps": Q * 1e3 / ms,
                      "wall_ms_host_buffers": statistics.median(wall),
                      "flagged": int(res.flags.sum())}), flush=True)
dev.close()
