#!/usr/bin/env python
"""One-GPU check (plain python, no torch; not collected by pytest): a shard whose row_base is far above 2^32 - what
sharded_store.py gives rank r (r << 32) - answers with global rows in every entry point and with the oracle's hits."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

import lvs_synth as synth  # noqa: E402
from code_rag_b200.collection import DeviceCollection  # noqa: E402
from oracle.qdrant_local import OracleCollection  # noqa: E402

ANY = 0xFFFFFFFF


def main():
    base, n, dim, k = 5 << 32, 5000, 96, 10
    x, q = synth.unit_rows(n, dim, seed=91, n_queries=40)
    for storage in ("f32", "bf16"):
        xs = synth.bf16_round(x) if storage == "bf16" else x
        ora = OracleCollection(dim)
        ora.upsert_rows_f32(0, xs, [None] * n)
        dev = DeviceCollection(f"rb_{storage}", dim, storage=storage, n_filter_cols=8, row_base=base)
        codes = np.zeros((n, 8), dtype=np.uint32)
        codes[:, 0] = 1 + np.arange(n) % 3
        dev.upsert(xs, rows=base + np.arange(n), codes=codes, ties=np.arange(n, dtype=np.uint64))
        assert dev.rows == n and dev.count() == n
        for Q in (1, 40):                                  # K1 and K2
            res = dev.search(q[:Q].astype(np.float64), k)
            for i in range(Q):
                rows_o, scores_o = ora.search_topk_rows(q[i].astype(np.float64), k)
                assert np.array_equal(res.rows[i] - base, rows_o), (storage, Q, i, res.rows[i], rows_o)
                assert np.allclose(res.scores[i], scores_o, rtol=1e-9, atol=1e-12)
        pick = np.array([0, 17, n - 1, 4242])
        got = dev.fetch_rows(base + pick)                      # what the sharded adapter's rebalancing reads back
        exp = xs[pick] if storage == "bf16" else (xs[pick].astype(np.float64) / np.linalg.norm(xs[pick].astype(np.float64), axis=1, keepdims=True)).astype(np.float32)
        assert got.shape == exp.shape and np.array_equal(got, exp), (storage, np.abs(got - exp).max())
        want = np.full(8, ANY, dtype=np.uint32); want[0] = 2
        rows, m = dev.match_rows(want)
        assert m == len(rows) == (n + 1) // 3 and np.array_equal(np.sort(rows) - base, np.nonzero(codes[:, 0] == 2)[0]), (m, len(rows))   # unordered
        res = dev.search(q[0].astype(np.float64), k, want)
        assert ((res.rows[0] - base) % 3 == 1).all()
        dev.set_codes(1, np.full(10, 7, dtype=np.uint32), row0=base + 20)
        want2 = np.full(8, ANY, dtype=np.uint32); want2[1] = 7
        rows, m = dev.match_rows(want2)
        assert np.array_equal(np.sort(rows) - base, np.arange(20, 30)), rows
        rows, m = dev.delete_where(want2)
        assert m == 10 and np.array_equal(np.sort(rows) - base, np.arange(20, 30)) and dev.count() == n - 10, (m, rows)
        assert dev.delete_rows(base + np.arange(n - 10, n - 5)) == 5
        dev.move_rows(base + np.arange(n - 5, n), base + np.arange(20, 25))      # tail rows into the first holes
        dev.truncate(n - 10)
        assert dev.rows == n - 10 and dev.count() == n - 10 - 5
        res = dev.search(q[1].astype(np.float64), k)
        assert (res.rows[0] >= base).all() and (res.rows[0] < base + n - 10).all()
        dev.close()
        print(f"row_base {base:#x} [{storage}]: OK", flush=True)


if __name__ == "__main__":
    main()
