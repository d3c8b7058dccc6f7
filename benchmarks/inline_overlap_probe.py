#!/usr/bin/env python
"""Does a launch with a large parameter block (the query riding in the kernel's parameters, scan_kernel.cuh InlineQueries) still
start early behind its predecessor (programmatic dependent launch)?  Pipelined host-buffer searches (two in flight) on a shard that
takes ~1 ms to scan, with the query in the parameter block (option inline_max_mb = huge) and staged by CTA 0 (inline_max_mb = 0),
plus strictly one at a time.  One JSON line.

    python benchmarks/inline_overlap_probe.py [rows]
"""
from __future__ import annotations

import json
import statistics
import sys
import time

import numpy as np

from configs import DeviceCollection, fill


def run(dev, qs, depth, n=120, warm=20):
    t_in = []
    times = []
    for i in range(n + warm):
        if i == warm:
            t0 = time.perf_counter()
        t_in.append(dev.search_submit(qs[i % len(qs)], 10))
        if len(t_in) == depth:
            dev.search_wait(t_in.pop(0))
    while t_in:
        dev.search_wait(t_in.pop(0))
    return (time.perf_counter() - t0) / n * 1e3


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    dim = 768
    dev = DeviceCollection("probe", dim, storage="bf16", capacity=rows, timing=False)
    fill(dev, rows, dim, "bf16", seed=7)
    qs = np.random.default_rng(3).standard_normal((16, 1, dim))
    out = {"what": "pipelined host-buffer searches: query in the parameter block vs staged", "rows": rows, "dim": dim, "storage": "bf16"}
    for label, mb in (("staged", 0), ("in_params", 1 << 20)):
        dev.set_option("inline_max_mb", mb)
        res = {}
        for depth in (1, 2, 3):
            res[f"ms_per_search_depth{depth}"] = statistics.median(run(dev, qs, depth) for _ in range(3))
        out[label] = res
    # where does a synchronous host-buffer search spend its time?  phase stamps of the kernel (option dbg_times) next to the wall time
    dev.set_option("inline_max_mb", 0)
    names = ("queries_ready", "scanned", "list_written", "lists_visible", "selected", "rescored", "stored")
    walls = []
    for i in range(60):
        t0 = time.perf_counter()
        dev.search(qs[i % len(qs)], 10)
        if i >= 10:
            walls.append((time.perf_counter() - t0) * 1e3)
    dev.set_option("dbg_times", 1)
    ph, wd = [], []
    for i in range(30):
        t0 = time.perf_counter()
        dev.search(qs[i % len(qs)], 10)
        wd.append((time.perf_counter() - t0) * 1e3)
        ph.append(dev.last_kernel_phases()[:7])
    dev.set_option("dbg_times", 0)
    out["sync_host_search"] = {"wall_ms": statistics.median(walls), "wall_ms_with_stamps": statistics.median(wd),
                               "kernel_phases_us": dict(zip(names, [round(float(v), 1) for v in np.median(np.array(ph), axis=0)]))}
    # the same search with the query already on the device (lvs_search_device through torch), synchronised per step
    import torch
    dq = torch.from_numpy(np.ascontiguousarray(qs[:, 0, :])).cuda()
    outs = [torch.zeros((1, 10), dtype=torch.float64, device="cuda"), torch.zeros((1, 10), dtype=torch.int64, device="cuda"),
            torch.zeros((1, 10), dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")]
    torch.cuda.synchronize()
    wdv = []
    for i in range(60):
        t0 = time.perf_counter()
        dev.search_device(dq[i % 16].data_ptr(), "f64", 1, 10, None, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr())
        if i >= 10:
            wdv.append((time.perf_counter() - t0) * 1e3)
    out["sync_device_search_wall_ms"] = statistics.median(wdv)
    print(json.dumps(out), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
