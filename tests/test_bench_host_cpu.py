"""Host pieces of bench.py that need no GPU: the NVML sampler's aggregation (against a stand-in pynvml), the bf16 rounding
helper and the reference arm's JSON line (the keys the driver's contract names)."""
import json
import subprocess
import sys
import time
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _fake_pynvml(clocks, reasons, power_mw, limit_mw=1_000_000, max_mhz=1965):
    m = types.ModuleType("pynvml")
    state = {"i": 0}
    m.NVML_CLOCK_SM = 1
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda idx: ("h", idx)
    m.nvmlDeviceGetMaxClockInfo = lambda h, kind: max_mhz
    m.nvmlDeviceGetEnforcedPowerLimit = lambda h: limit_mw

    def clock(h, kind):
        i = min(state["i"], len(clocks) - 1)
        return clocks[i]

    def reason(h):
        i = min(state["i"], len(reasons) - 1)
        return reasons[i]

    def power(h):
        i = min(state["i"], len(power_mw) - 1)
        state["i"] += 1                      # one full sample taken
        return power_mw[i]
    m.nvmlDeviceGetClockInfo = clock
    m.nvmlDeviceGetCurrentClocksEventReasons = reason
    m.nvmlDeviceGetPowerUsage = power
    return m


def test_clock_sampler_aggregates_clock_reasons_and_power(monkeypatch):
    import bench
    monkeypatch.setitem(sys.modules, "pynvml", _fake_pynvml([1965, 1400, 1335, 1335, 1335], [0, 0x4, 0x4, 0x4, 0x4],
                                                            [400_000, 990_000, 1_000_000, 1_000_000, 1_000_000]))
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES", raising=False)
    s = bench.ClockSampler(0, period_s=0.001)
    s.start()
    t0 = time.time()
    while len(s.power_w) < 5 and time.time() - t0 < 5:
        time.sleep(0.002)
    out = s.stop()
    assert out["samples"] >= 5 and out["sm_max_mhz"] == 1965.0
    assert out["sm_mhz"] == 1335.0                         # the median, not the first (idle) sample
    assert out["reasons"] == ["sw_power_cap"]
    assert out["power_w"] == 1000.0 and out["power_limit_w"] == 1000.0


def test_clock_sampler_without_nvml_says_so(monkeypatch):
    import bench
    bad = types.ModuleType("pynvml")

    def boom():
        raise RuntimeError("no driver")
    bad.nvmlInit = boom
    monkeypatch.setitem(sys.modules, "pynvml", bad)
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert out["samples"] == 0 and out["sm_mhz"] is None and "no samples" in out["reasons"][0]


def test_bf16_rounding_is_round_to_nearest_even():
    import bench
    import torch
    x = np.random.default_rng(3).standard_normal(4096).astype(np.float32)
    x[:4] = [1.0, 1.00390625, 1.01171875, -3.0e-39]         # exact, tie to even (down), tie to even (up), subnormal
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(bench.bf16_round_np(x).view(np.uint32), want.view(np.uint32))


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` on a small sample: one JSON line, the keys the driver reads, no GPU touched."""
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--rows", "40000",
                          "--cpu-sample-rows", "20000", "--dim", "64"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches", "extrapolated"):
        assert key in line, key
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" or cb["kind"].startswith("qdrant-client")
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and "20000 of 40000 rows" in cb["sample"] and line["extrapolated"] is True


def test_reference_arm_other_ranks_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
