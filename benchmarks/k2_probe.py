#!/usr/bin/env python
"""K2 probe: 10M x 768 bf16, batches of queries through the tensor-core path; prints kernel / finalize / total times."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "benchmarks"))
from configs import fill  # noqa: E402

from code_rag_b200.collection import DeviceCollection  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
import os
STORAGE = os.environ.get("K2_STORAGE", "bf16")           # "f32": the kind::tf32 form over an fp32 shard
DIM = int(os.environ.get("K2_DIM", "768"))
dev = DeviceCollection("k2", DIM, storage=STORAGE, capacity=n)
fill(dev, n, DIM, STORAGE, seed=3456)
for opt in ("gemm_no_unit", "gemm_stages", "gemm_stages_b", "gemm_dbg", "gemm_keep", "gemm_no_pair"):
    if os.environ.get(opt.upper()):
        dev.set_option(opt, int(os.environ[opt.upper()]))
rng = np.random.default_rng(6)
cfgs = [(256, 100), (256, 10), (128, 10), (64, 100), (64, 10), (16, 10)]
if len(sys.argv) > 3:
    cfgs = [(int(sys.argv[2]), int(sys.argv[3]))]
# K2_SWEEP="4:0,3:6,4:5": (query-chunk buffers : corpus buffers) of the pair form's rings, every configuration on the same resident corpus
# (an optional third field is a gemm_dbg value for that entry)
sweep = [tuple(int(v) for v in item.split(":")) for item in os.environ.get("K2_SWEEP", "").split(",") if item]
sweep = [(e + (0,))[:3] for e in sweep]
for Q, k, sa, sb, dbg in [(Q, k, sa, sb, dbg) for (sa, sb, dbg) in (sweep or [(None, None, 0)]) for (Q, k) in cfgs]:
    if sa is not None:
        dev.set_option("gemm_stages", sa)
        dev.set_option("gemm_stages_b", sb)
        dev.set_option("gemm_dbg", dbg)
    qs = np.random.default_rng(6 + Q).standard_normal((Q, DIM))
    for rep in range(3):
        res = dev.search(qs, k)
    t = dev.last_timing()
    print(json.dumps({"Q": Q, "k": k, "stages": [sa, sb], "dbg": dbg, **t, "flagged": int(res.flags.sum()), "rows_sum": int(res.rows.sum()),
                      "gemm_gbs": n * DIM * (2 if STORAGE == "bf16" else 4) / (t["scan_ms"] * 1e-3) / 1e9, "gemm_tflops": 2.0 * Q * n * DIM / (t["scan_ms"] * 1e-3) / 1e12, "storage": STORAGE}), flush=True)
dev.close()
