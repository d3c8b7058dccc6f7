#!/usr/bin/env python
"""One-GPU quick check of the adapter (plain python, no torch; not collected by pytest): exact ties under ids that share their
tie key, random operation sequences, the .client shim and mass-delete compaction on the real device.  The same scenarios run
under pytest (-m gpu); this is the few-second form for a short GPU slot."""
import asyncio
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

import adapter_scenarios as S  # noqa: E402


def main():
    t0 = time.time()
    for name, coro in (("exact ties follow the id", lambda: S.scenario_exact_ties_follow_the_id(None)),
                       ("random ops f32", lambda: S.scenario_random_ops(None, 1, storage="f32", steps=100)),
                       ("random ops bf16", lambda: S.scenario_random_ops(None, 3, storage="bf16", steps=100)),
                       ("edge cases", lambda: S.scenario_edge_cases(None)),
                       ("client shim", lambda: S.scenario_client_shim(None)),
                       ("mass delete compacts", lambda: S.scenario_mass_delete_compacts(None))):
        asyncio.run(coro())
        print(f"{name}: OK ({time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
