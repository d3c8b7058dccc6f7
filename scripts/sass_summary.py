#!/usr/bin/env python
"""Blackwell evidence from the SHIPPED library: per kernel, the counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA
(profiling guide: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP,
mbarrier -> SYNCS, tcgen05.commit -> UTCBAR) plus peer-memory / system-scope accesses of the exchange code.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt        (no GPU needed: cuobjdump -sass on the .so)
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "code_rag_b200" / "lib" / "liblattice_b200.so"
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "ACQBULK", "HMMA", "HGMMA",
       "LDG", "STG", "LDS", "STS", "FFMA", "DFMA", "SHFL", "BAR", "MEMBAR", "ATOM", "RED", "NANOSLEEP"]


def main() -> int:
    sys.path.insert(0, str(ROOT))
    from code_rag_b200 import build
    build.build()
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels: dict[str, collections.Counter] = {}
    extra: dict[str, collections.Counter] = {}
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            extra[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for p in PAT:
            if op == p or op.startswith(p + "."):
                kernels[cur][p] += 1
        for tag in (".2CTA", ".MULTICAST", ".SYS", ".STRONG.SYS", ".128"):
            if tag in op:
                extra[cur][op] += 1
    print(f"# SASS mnemonic counts per kernel of {LIB.relative_to(ROOT)} (cuobjdump -sass; sm_100a)\n")
    order = sorted(kernels, key=lambda k: (-kernels[k]["UTCHMMA"], -kernels[k]["UBLKCP"], k))
    for k in order:
        c = kernels[k]
        if c["_total"] == 0:
            continue
        shown = ", ".join(f"{p} {c[p]}" for p in PAT if c[p])
        print(f"{k}\n    instructions {c['_total']}: {shown}")
        ex = ", ".join(f"{o} x{n}" for o, n in sorted(extra[k].items()) if any(t in o for t in (".2CTA", ".MULTICAST", ".SYS")))
        if ex:
            print(f"    cluster / system-scope forms: {ex}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
