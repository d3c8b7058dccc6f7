"""GPU parity tests: CUDA path through the C ABI vs the CPU oracle (oracle/qdrant_local.py).

Bars (BASELINE.json north_star): top-k id lists identical (ties by id); scores within 1e-5 relative for fp32
storage, 2e-3 for bf16 storage.  The device replays local mode's arithmetic exactly, so the observed error is
~1e-15; the tests assert the stated bars and additionally a much tighter one to catch regressions of the replay.
"""
import numpy as np
import pytest

from oracle.qdrant_local import OracleCollection
import lvs_synth as synth

pytestmark = pytest.mark.gpu

REL_F32 = 1e-5
REL_BF16 = 2e-3
TIGHT = 1e-12   # absolute, float64 dot of unit vectors computed in a different order


@pytest.fixture(scope="module")
def lib(native_lib):
    from code_rag_b200 import _native
    _native.init(0)
    return native_lib


def _dev(name, dim, storage="f32", ncols=0, metric="cosine", **kw):
    from code_rag_b200.collection import DeviceCollection
    return DeviceCollection(name, dim, storage=storage, metric=metric, n_filter_cols=ncols, **kw)


def _assert_same(res, qi, rows_o, scores_o, rel, tight=TIGHT):
    n = int(res.counts[qi])
    assert n == len(rows_o), f"query {qi}: device returned {n} hits, oracle {len(rows_o)}"
    got = res.rows[qi, :n]
    assert np.array_equal(got, rows_o), (
        f"query {qi}: id lists differ\n device {got.tolist()}\n oracle {rows_o.tolist()}\n"
        f" device scores {res.scores[qi, :n].tolist()}\n oracle scores {scores_o.tolist()}")
    if n == 0:
        assert res.flags[qi] == 0
        return
    err = np.abs(res.scores[qi, :n] - scores_o)
    assert np.all(err <= rel * np.maximum(np.abs(scores_o), 1e-30) + 1e-300), f"query {qi}: score error {err.max()}"
    assert err.max() <= tight, f"query {qi}: replay drift {err.max()} (> {tight})"
    assert res.flags[qi] == 0, f"query {qi}: exactness flag set"


def test_c1_fp32_single_queries(lib):
    """configs[0]: 10k x 768 fp32 UniXcoder-shaped, 1 query at a time, top-10."""
    x, q = synth.unixcoder_like(10_000, 768, seed=1234, n_queries=40)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, x, [None] * len(x))
    dev = _dev("c1", 768)
    dev.upsert(x.astype(np.float64))
    assert dev.rows == 10_000 and dev.count() == 10_000
    for i in range(len(q)):
        res = dev.search(q[i].astype(np.float64), 10)
        rows_o, scores_o = ora.search_topk_rows(q[i].astype(np.float64), 10)
        _assert_same(res, 0, rows_o, scores_o, REL_F32)
    dev.close()


def test_replay_tracks_search_count(lib):
    """Rows written at different times see a different number of in-place re-normalisations."""
    x, q = synth.unit_rows(6_000, 768, seed=77, n_queries=12)
    ora = OracleCollection(768)
    dev = _dev("replay", 768)
    ora.upsert_rows_f32(0, x[:3000], [None] * 3000)
    dev.upsert(x[:3000].astype(np.float64))
    for i in range(6):
        res = dev.search(q[i].astype(np.float64), 10)
        _assert_same(res, 0, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32)
    ora.upsert_rows_f32(3000, x[3000:], [None] * 3000)
    dev.upsert(x[3000:].astype(np.float64))
    for i in range(6, 12):
        res = dev.search(q[i].astype(np.float64), 10)
        _assert_same(res, 0, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32)
    dev.close()


@pytest.mark.parametrize("k", [1, 10, 20, 100, 224])
def test_bf16_storage_parity(lib, k):
    """bf16 shard: the oracle is fed the bf16-rounded rows."""
    x, q = synth.unit_rows(30_000, 768, seed=3456, n_queries=6)
    xb = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xb, [None] * len(xb))
    dev = _dev("bf16", 768, storage="bf16")
    dev.upsert(xb)
    for i in range(len(q)):
        res = dev.search(q[i].astype(np.float64), k)
        _assert_same(res, 0, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_BF16)
    dev.close()


def test_bf16_storage_of_unrounded_inputs_meets_the_stated_bar(lib):
    """north_star: scores within 2e-3 relative for bf16 storage.  Here the inputs are NOT bf16-representable, so storing them
    costs precision: every returned score must be within 2e-3 of what the fp32 reference gives for that same row, and the
    returned rows must be the reference's top-k up to near-ties (rows whose reference score is within that margin of the k-th)."""
    n, k = 30_000, 10
    x, q = synth.unit_rows(n, 768, seed=777, n_queries=6)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, x, [None] * n)                 # the reference keeps float32
    dev = _dev("bf16raw", 768, storage="bf16")
    dev.upsert(x.astype(np.float64))                      # the shard rounds to bf16
    for i in range(len(q)):
        res = dev.search(q[i].astype(np.float64), k)
        ref = np.asarray(ora._scores(q[i].astype(np.float64)), dtype=np.float64)[:n]
        got_rows, got_scores = res.rows[0], res.scores[0]
        assert np.all(np.abs(got_scores - ref[got_rows]) <= 2e-3 * np.abs(ref[got_rows]))
        kth = np.sort(ref)[-k]
        assert np.all(ref[got_rows] >= kth - 2 * 2e-3 * abs(kth)), "a returned row is not among the reference's near-top-k"
        assert len(set(got_rows.tolist()) & set(np.argsort(ref)[-k:].tolist())) >= k - 2
    dev.close()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
@pytest.mark.parametrize("Q", [2, 3, 4, 7, 16])
def test_batch_equals_sequential(lib, storage, Q):
    x, q = synth.unit_rows(12_000, 768, seed=99, n_queries=Q)
    if storage == "bf16":
        x = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, x, [None] * len(x))
    dev = _dev("batch", 768, storage=storage)
    dev.upsert(x)
    res = dev.search(q.astype(np.float64), 10)
    for i in range(Q):   # the oracle runs them one after the other, mutating its matrix each time
        _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32 if storage == "f32" else REL_BF16)
    dev.close()


def test_c2_filters_fp32_1536(lib):
    """configs[1] shape at a CPU-checkable size: 1536-d fp32 unit rows, project/file/language filters, top-10."""
    n = 40_000
    x, q = synth.unit_rows(n, 1536, seed=2345, n_queries=3)
    pl = synth.payloads(n, seed=2345)
    cols = ["project_name", "language", "file_path", "entity_type"]
    dicts = [dict() for _ in cols]
    codes = np.zeros((n, len(cols)), dtype=np.uint32)
    for i, p in enumerate(pl):
        for c, key in enumerate(cols):
            codes[i, c] = dicts[c].setdefault(p[key], len(dicts[c]) + 1)
    ora = OracleCollection(1536)
    ora.upsert_rows_f32(0, x, pl)
    dev = _dev("c2", 1536, ncols=len(cols))
    dev.upsert(x.astype(np.float64), codes=codes)
    ANY = 0xFFFFFFFF
    some_file = pl[123]["file_path"]
    cases = [
        {},
        {"project_name": "proj0"},
        {"project_name": "proj7", "language": "python"},
        {"file_path": some_file},
        {"project_name": "proj3", "entity_type": "class", "language": "javascript"},
    ]
    for flt in cases:
        want = np.full(len(cols), ANY, dtype=np.uint32)
        mask = np.ones(n, dtype=bool)
        for key, val in flt.items():
            c = cols.index(key)
            want[c] = dicts[c][val]
            mask &= codes[:, c] == want[c]
        for i in range(len(q)):
            res = dev.search(q[i].astype(np.float64), 10, want if flt else None)
            rows_o, scores_o = ora.search_topk_rows(q[i].astype(np.float64), 10, mask)
            _assert_same(res, 0, rows_o, scores_o, REL_F32)
    # a value never seen in the column matches nothing
    want = np.full(len(cols), ANY, dtype=np.uint32)
    want[0] = 0xFFFFFFFE
    res = dev.search(q[0].astype(np.float64), 10, want)
    assert res.counts[0] == 0 and (res.rows[0] == -1).all()
    dev.close()


def test_tombstones_overwrite_and_ties(lib):
    x, q = synth.unit_rows(5_000, 256, seed=5, n_queries=4)
    x[100] = x[7]      # exact duplicates: equal scores, order by (tie, row)
    x[4000] = x[7]
    ora = OracleCollection(256)
    ora.upsert_rows_f32(0, x, [None] * len(x))
    dev = _dev("tomb", 256)
    dev.upsert(x.astype(np.float64))
    qq = x[7].astype(np.float64)
    res = dev.search(qq, 5)
    rows_o, scores_o = ora.search_topk_rows(qq, 5)
    assert res.rows[0, :3].tolist() == [7, 100, 4000] == rows_o[:3].tolist()
    assert res.scores[0, 0] == res.scores[0, 1] == res.scores[0, 2]
    # delete two rows, overwrite one
    dev.delete_rows(np.array([7, 4000]))
    ora.deleted[[7, 4000]] = True
    y = synth.unit_rows(1, 256, seed=6)[0]
    dev.upsert(y.astype(np.float64), rows=np.array([100]))
    ora.vectors[100] = y[0].astype(np.float64) / np.linalg.norm(y[0].astype(np.float64))
    assert dev.count() == 4_998
    for i in range(len(q)):
        res = dev.search(q[i].astype(np.float64), 10)
        # row 100 was rewritten after 1 device search: its replay starts from the new write; mirror that in the oracle
        rows_o, scores_o = ora.search_topk_rows(q[i].astype(np.float64), 10)
        _assert_same(res, 0, rows_o, scores_o, REL_F32)
    dev.close()


def test_edges(lib):
    dev = _dev("edge", 64)
    res = dev.search(np.ones(64), 10)
    assert res.counts[0] == 0 and (res.rows == -1).all()
    x, _ = synth.unit_rows(7, 64, seed=3)
    x[3] = 0.0   # a zero vector is stored as is and scores 0
    dev.upsert(x.astype(np.float64))
    ora = OracleCollection(64)
    ora.upsert_rows_f32(0, x, [None] * 7)
    q = x[2].astype(np.float64)
    res = dev.search(q, 10)
    rows_o, scores_o = ora.search_topk_rows(q, 10)
    assert res.counts[0] == 7
    _assert_same(res, 0, rows_o, scores_o, REL_F32)
    with pytest.raises(ValueError):
        dev.search(np.full(64, np.nan), 3)
    from code_rag_b200.errors import NativeLibraryError
    with pytest.raises(NativeLibraryError):
        dev.search(q, 1000)
    dev.close()


@pytest.mark.parametrize("dim", [8, 100, 384, 1000, 3072])
def test_odd_dimensions(lib, dim):
    x, q = synth.unit_rows(3_000, dim, seed=dim, n_queries=3)
    for storage in ("f32", "bf16"):
        xs = synth.bf16_round(x) if storage == "bf16" else x
        ora = OracleCollection(dim)
        ora.upsert_rows_f32(0, xs, [None] * len(xs))
        dev = _dev("odd", dim, storage=storage)
        dev.upsert(xs)
        for i in range(len(q)):
            res = dev.search(q[i].astype(np.float64), 10)
            _assert_same(res, 0, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32 if storage == "f32" else REL_BF16)
        dev.close()


@pytest.mark.parametrize("Q,k", [(8, 10), (100, 10), (128, 100), (256, 100), (300, 20)])
def test_gemm_path_parity(lib, Q, k):
    """K2 (tcgen05 GEMM + fused top-k) on a bf16 shard: same ids and scores as the oracle's consecutive searches."""
    n = 24_000
    x, q = synth.unit_rows(n, 768, seed=3456 + Q, n_queries=Q)
    xb = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xb, [None] * n)
    dev = _dev("gemm", 768, storage="bf16")
    dev.upsert(xb)
    dev.delete_rows(np.array([5, 77, 12_345]))
    ora.deleted[[5, 77, 12_345]] = True
    res = dev.search(q.astype(np.float64), k)
    assert dev.last_timing()["kernel"] == "gemm"
    for i in range(Q):
        _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_BF16)
    dev.close()


def test_gemm_pair_form_ring_geometries(lib):
    """K2's CTA-pair form with every operand-ring layout: split rings (query-chunk buffers : corpus buffers = 4:5, the default;
    3:6; 2:7) and the single ring of combined stages (gemm_stages_b = 0) return the oracle's consecutive searches alike."""
    n, Q, k = 20_000, 160, 20
    x, q = synth.unit_rows(n, 768, seed=977, n_queries=Q)
    xb = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xb, [None] * n)
    dev = _dev("rings", 768, storage="bf16")
    dev.upsert(xb)
    for stages, stages_b in ((0, -1), (4, 0), (3, 6), (2, 7), (4, 5)):
        dev.set_option("gemm_stages", stages)
        dev.set_option("gemm_stages_b", stages_b)
        res = dev.search(q.astype(np.float64), k)
        assert dev.last_timing()["kernel"] == "gemm"
        for i in range(Q):
            _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_BF16)
    dev.close()


@pytest.mark.parametrize("Q,k,dim", [(9, 10, 768), (130, 50, 1536), (64, 100, 100)])
def test_gemm_path_tf32_on_f32_storage(lib, Q, k, dim):
    """K2 over an fp32 shard (tcgen05 kind::tf32 reads the stored fp32 rows): same ids as the oracle, scores within 1e-5."""
    n = 24_000
    x, q = synth.unit_rows(n, dim, seed=4242 + Q, n_queries=Q)
    ora = OracleCollection(dim)
    ora.upsert_rows_f32(0, x, [None] * n)
    dev = _dev("gemmtf32", dim, storage="f32")
    dev.upsert(x.astype(np.float64))
    dev.delete_rows(np.array([1, 300, 23_999]))
    ora.deleted[[1, 300, 23_999]] = True
    res = dev.search(q.astype(np.float64), k)
    assert dev.last_timing()["kernel"] == "gemm"
    for i in range(Q):
        _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_F32)
    dev.close()


@pytest.mark.parametrize("Q", [12, 140])
def test_gemm_path_with_payload_filter(lib, Q):
    """K2 with a payload filter (the mask rides with the tombstone check): wide, selective and empty filters, top-10."""
    n, dim, k = 24_000, 768, 10
    x, q = synth.unit_rows(n, dim, seed=777 + Q, n_queries=Q)
    xb = synth.bf16_round(x)
    rng = np.random.default_rng(5)
    codes = np.stack([rng.choice([1, 2, 3], size=n, p=[0.5, 0.3, 0.2]), rng.integers(1, 401, size=n)], axis=1).astype(np.uint32)
    ANY = 0xFFFFFFFF
    for want_list in ([2, ANY], [1, 17], [ANY, 399], [3, 0xFFFFFFFE]):
        want = np.array(want_list, dtype=np.uint32)
        mask = np.ones(n, dtype=bool)
        for c in range(2):
            if want[c] != ANY:
                mask &= codes[:, c] == want[c]
        ora = OracleCollection(dim)
        ora.upsert_rows_f32(0, xb, [None] * n)
        dev = _dev("gemmflt", dim, storage="bf16", ncols=2)
        dev.upsert(xb, codes=codes)
        res = dev.search(q.astype(np.float64), k, want)
        assert dev.last_timing()["kernel"] == "gemm"
        for i in range(Q):
            _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k, mask), REL_BF16)
        dev.close()


@pytest.mark.parametrize("dim", [100, 1000, 1536])
def test_gemm_path_other_dimensions(lib, dim):
    """K2 with K not a multiple of the 64-element chunk (TMA zero-fills the tail) and with K > 768."""
    n, Q, k = 20_000, 24, 10
    x, q = synth.unit_rows(n, dim, seed=dim + 1, n_queries=Q)
    xb = synth.bf16_round(x)
    ora = OracleCollection(dim)
    ora.upsert_rows_f32(0, xb, [None] * n)
    dev = _dev("gemmdim", dim, storage="bf16")
    dev.upsert(xb)
    res = dev.search(q.astype(np.float64), k)
    assert dev.last_timing()["kernel"] == "gemm"
    for i in range(Q):
        _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_BF16)
    dev.close()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_dot_metric(lib, storage):
    """metric = dot: rows and queries are used as given (no normalisation), float64 scores as np.dot would give."""
    n, dim = 9_000, 256
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((n, dim)) * rng.uniform(0.5, 2.0, size=(n, 1))).astype(np.float32)
    if storage == "bf16":
        x = synth.bf16_round(x)
    q = rng.standard_normal((5, dim))
    ora = OracleCollection(dim, distance="dot")
    ora.upsert_rows_f32(0, x, [None] * n)
    dev = _dev("dot", dim, storage=storage, metric="dot")
    dev.upsert(x)
    for i in range(len(q)):
        res = dev.search(q[i], 10)
        rows_o, scores_o = ora.search_topk_rows(q[i], 10)
        assert np.array_equal(res.rows[0], rows_o)
        assert np.allclose(res.scores[0], scores_o, rtol=1e-12, atol=1e-12)
    res = dev.search(q, 10)            # batch (K1 passes for f32, K1 or K2 for bf16)
    for i in range(len(q)):
        rows_o, scores_o = ora.search_topk_rows(q[i], 10)
        assert np.array_equal(res.rows[i], rows_o)
    dev.close()


def test_f32_storage_large_batch_uses_scan_passes(lib):
    x, q = synth.unit_rows(20_000, 384, seed=21, n_queries=33)
    ora = OracleCollection(384)
    ora.upsert_rows_f32(0, x, [None] * len(x))
    dev = _dev("f32batch", 384)
    dev.set_option("gemm_no_tf32", 1)          # keep the batch on K1 (ceil(Q/4) scan passes)
    dev.upsert(x.astype(np.float64))
    res = dev.search(q.astype(np.float64), 10)
    assert dev.last_timing()["kernel"] == "scan"
    for i in range(len(q)):
        _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32)
    dev.close()


@pytest.mark.parametrize("k", [100, 110])
def test_gemm_falls_back_when_bound_fails(lib, k):
    """Near-duplicate rows concentrated in one tile defeat the 16-key lists: flagged queries are redone on the K1 path - also from
    k = 104 on, where the candidate set is already the largest one (the pipelined lvs_search / lvs_search_wait route used to return
    those queries UNPROVEN instead of giving them the exact scan)."""
    n = 16_384
    x, q = synth.unit_rows(n, 768, seed=11, n_queries=8)
    rng = np.random.default_rng(3)
    x[4096:4096 + 120] = q[0] + 0.01 * rng.standard_normal((120, 768)).astype(np.float32)   # 120 close neighbours of query 0
    xb = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xb, [None] * n)
    dev = _dev("gemmfb", 768, storage="bf16")
    dev.upsert(xb)
    res = dev.search(q.astype(np.float64), k)
    for i in range(8):
        _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_BF16)
    dev.close()


def test_submit_wait_pipeline(lib):
    """lvs_search_submit / lvs_search_wait: 3 searches in flight are accounted as consecutive reference searches."""
    x, q = synth.unit_rows(9_000, 768, seed=31, n_queries=9)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, x, [None] * len(x))
    dev = _dev("pipe", 768)
    dev.upsert(x.astype(np.float64))
    inflight, got = [], []
    for i in range(len(q)):
        inflight.append(dev.search_submit(q[i].astype(np.float64), 10))
        if len(inflight) == 3:
            got.append(dev.search_wait(inflight.pop(0)))
    while inflight:
        got.append(dev.search_wait(inflight.pop(0)))
    for i in range(len(q)):
        _assert_same(got[i], 0, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32)
    # lvs_search_poll: turns true by itself (kernel-stored completion word for a single query, the slot's event for a staged
    # batch), after which the wait returns the same answer a blocking search gives
    import time
    from code_rag_b200.errors import NativeLibraryError
    for Qn in (1, 40):
        t = dev.search_submit(np.tile(q, (5, 1))[:Qn].astype(np.float64), 10)
        t0 = time.perf_counter()
        while not dev.search_poll(t):
            assert time.perf_counter() - t0 < 5.0, "the search never completed"
        res = dev.search_wait(t)
        assert res.rows.shape == (Qn, 10) and (res.counts == 10).all()
        for i in range(Qn):
            rows_o, _ = ora.search_topk_rows(q[i % len(q)].astype(np.float64), 10)
            assert np.array_equal(res.rows[i], rows_o)
    with pytest.raises(NativeLibraryError):
        dev.search_poll((3, 1, 10))            # not in flight
    dev.close()


def test_sharded_searcher_world1(lib):
    """ShardedSearcher (world size 1): async device path and the pipelined host path give the oracle's answer."""
    import torch
    from code_rag_b200.sharded import ShardedSearcher
    x, q = synth.unit_rows(8_000, 768, seed=32, n_queries=6)
    xb = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xb, [None] * len(xb))
    dev = _dev("ss1", 768, storage="bf16")
    dev.upsert(xb)
    ss = ShardedSearcher(dev)
    dq = torch.from_numpy(q.astype(np.float64)).cuda()
    outs = []
    torch.cuda.synchronize()
    for i in range(3):
        s, r, t, c, f = ss.search_device_async(dq[i:i + 1], 10, slot=i)
        outs.append((s, r, c, f))
    ss.stream.synchronize()
    for i in range(3):
        s, r, c, f = outs[i]
        rows_o, scores_o = ora.search_topk_rows(q[i].astype(np.float64), 10)
        assert np.array_equal(r.cpu().numpy()[0], rows_o)
        assert np.abs(s.cpu().numpy()[0] - scores_o).max() <= TIGHT
        assert int(f.sum().item()) == 0
    hs = [ss.submit(q[i].astype(np.float64), 10) for i in range(3, 6)]
    for i, h in zip(range(3, 6), hs):
        s, r, t, c, f = ss.wait(h)
        rows_o, scores_o = ora.search_topk_rows(q[i].astype(np.float64), 10)
        assert np.array_equal(r[0], rows_o) and np.abs(s[0] - scores_o).max() <= TIGHT and f.sum() == 0
    dev.close()


def test_sharded_merge_equals_single(lib):
    """K5: G shards searched separately + merge == one collection (the multi-GPU path, emulated on one GPU)."""
    import torch
    from code_rag_b200.collection import merge_topk_device
    n, dim, k, G, Q = 24_000, 768, 10, 4, 3
    x, q = synth.unit_rows(n, dim, seed=5678, n_queries=Q)
    xb = synth.bf16_round(x)
    single = _dev("single", dim, storage="bf16")
    single.upsert(xb)
    ref = single.search(q.astype(np.float64), k)
    per = n // G
    dq = torch.from_numpy(q.astype(np.float64)).cuda()
    scores = torch.zeros((G, Q, k), dtype=torch.float64, device="cuda")
    rows = torch.full((G, Q, k), -1, dtype=torch.int64, device="cuda")
    ties = torch.zeros((G, Q, k), dtype=torch.int64, device="cuda")
    counts = torch.zeros((G, Q), dtype=torch.int32, device="cuda")
    shards = []
    for g in range(G):
        sh = _dev(f"shard{g}", dim, storage="bf16", row_base=g * per)
        sh.upsert(xb[g * per:(g + 1) * per])
        shards.append(sh)
        torch.cuda.synchronize()
        flags = sh.search_device(dq.data_ptr(), "f64", Q, k, None, scores[g].data_ptr(), rows[g].data_ptr(),
                                 ties[g].data_ptr(), counts[g].data_ptr())
        assert (flags == 0).all()
    o_s = torch.zeros((Q, k), dtype=torch.float64, device="cuda")
    o_r = torch.zeros((Q, k), dtype=torch.int64, device="cuda")
    o_t = torch.zeros((Q, k), dtype=torch.int64, device="cuda")
    o_c = torch.zeros(Q, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    merge_topk_device(scores.data_ptr(), rows.data_ptr(), ties.data_ptr(), G, Q, k, o_s.data_ptr(), o_r.data_ptr(),
                      o_t.data_ptr(), o_c.data_ptr())
    assert np.array_equal(o_r.cpu().numpy(), ref.rows)
    assert np.array_equal(o_s.cpu().numpy(), ref.scores)
    for sh in shards:
        sh.close()
    single.close()


def _exchange_world1(lib, max_q, max_k):
    import ctypes as C
    from code_rag_b200 import _native as N
    ex, handle = C.c_void_p(), (C.c_ubyte * 64)()
    N.check(lib.lvs_exchange_create(1, 0, max_q, max_k, C.byref(ex), handle), "lvs_exchange_create")
    return ex


@pytest.mark.parametrize("storage,Q,k,n", [("bf16", 1, 10, 9_000), ("f32", 3, 10, 9_000), ("f32", 4, 50, 9_000), ("bf16", 2, 120, 9_000),
                                           ("bf16", 7, 10, 9_000), ("bf16", 8, 10, 20_480), ("bf16", 130, 100, 20_480)])
def test_sharded_entry_on_one_rank(lib, storage, Q, k, n):
    """lvs_search_sharded_device_async with a world of one rank: the publish / flag / wait / merge code of the fused scan kernel
    (and of the exchange kernel behind the tensor-core path) runs against the rank's own gather buffer, so the single-GPU tier
    covers it; results must equal the oracle's (and the local entry's)."""
    import torch
    x, q = synth.unit_rows(n, 768, seed=77, n_queries=Q)
    xs = synth.bf16_round(x) if storage == "bf16" else x
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xs, [None] * n)
    dev = _dev("shard1", 768, storage=storage, row_base=5 << 32, timing=False)
    dev.upsert(xs.astype(np.float64))
    ex = _exchange_world1(lib, max(Q, 4), max(k, 16))
    dq = torch.from_numpy(q.astype(np.float64)).cuda()
    out = torch.zeros((3, Q, k), dtype=torch.int64, device="cuda")
    counts = torch.zeros(Q, dtype=torch.int32, device="cuda")
    flags = torch.full((Q,), 7, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(2):        # twice: both slots of the double-buffered exchange, and the replay count moves on
        dev.search_sharded_device_async(ex, dq.data_ptr(), "f64", Q, k, None, out.data_ptr(), counts.data_ptr(), flags.data_ptr(), st.cuda_stream)
        st.synchronize()
        o = out.cpu().numpy()
        from code_rag_b200.collection import SearchResult
        res = SearchResult(o[0].view(np.float64), o[1] - (5 << 32), o[2].view(np.uint64), counts.cpu().numpy().astype(np.uint32), flags.cpu().numpy())
        for i in range(Q):
            _assert_same(res, i, *ora.search_topk_rows(q[i].astype(np.float64), k), REL_BF16 if storage == "bf16" else REL_F32)
    assert lib.lvs_exchange_error(ex) == 0
    lib.lvs_exchange_destroy(ex)
    dev.close()


def test_back_to_back_searches_overlap_correctly(lib):
    """60 single-query searches enqueued back to back on one stream with the programmatic-launch overlap on (timing off): the next
    search's CTAs start while the previous one's last CTAs are still rescoring; every result must still be the oracle's, with the
    replay count of ITS position in the sequence."""
    import torch
    n = 60_000
    x, q = synth.unit_rows(n, 768, seed=78, n_queries=60)
    xb = synth.bf16_round(x)
    ora = OracleCollection(768)
    ora.upsert_rows_f32(0, xb, [None] * n)
    dev = _dev("b2b", 768, storage="bf16", timing=False)
    dev.upsert(xb)
    dq = torch.from_numpy(q.astype(np.float64)).cuda()
    S = torch.zeros((60, 1, 10), dtype=torch.float64, device="cuda")
    R = torch.zeros((60, 1, 10), dtype=torch.int64, device="cuda")
    T = torch.zeros((60, 1, 10), dtype=torch.int64, device="cuda")
    Cn = torch.zeros((60, 1), dtype=torch.int32, device="cuda")
    F = torch.zeros((60, 1), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for i in range(60):
        dev.search_device_async(dq[i:i + 1].data_ptr(), "f64", 1, 10, None, S[i].data_ptr(), R[i].data_ptr(), T[i].data_ptr(),
                                Cn[i].data_ptr(), F[i].data_ptr(), st.cuda_stream)
    st.synchronize()
    from code_rag_b200.collection import SearchResult
    for i in range(60):
        res = SearchResult(S[i].cpu().numpy(), R[i].cpu().numpy(), T[i].cpu().numpy().view(np.uint64), Cn[i].cpu().numpy().astype(np.uint32), F[i].cpu().numpy())
        _assert_same(res, 0, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_BF16)
    dev.close()


def test_search_counter_crosses_2_pow_32(lib):
    """The replay count is (search number - write epoch): both are 64-bit, so a collection that has served 2^32 searches keeps
    answering exactly.  Rows written long ago have reached the fixed point (or 2-cycle) of local mode's in-place re-normalisation -
    only the parity of their age matters - so the oracle can stand in for 2^32 searches with 66 or 67 of them."""
    x, q = synth.unixcoder_like(3_000, 128, seed=91, n_queries=8)
    dev = _dev("wrap", 128)
    dev.upsert(x[:2000].astype(np.float64))
    ora = OracleCollection(128)
    ora.upsert_rows_f32(0, x[:2000], [None] * 2000)
    jump = (1 << 32) - 3
    dev.advance_search_counter(jump)
    for _ in range(66 + (jump & 1)):                    # same parity as `jump`, past the point where every chain has settled
        ora._scores(q[0].astype(np.float64))
    dev.upsert(x[2000:].astype(np.float64))             # written just before the 2^32 boundary
    ora.upsert_rows_f32(2000, x[2000:], [None] * 1000)
    for i in range(8):                                  # searches number 2^32 - 2 ... 2^32 + 5
        res = dev.search(q[i].astype(np.float64), 10)
        _assert_same(res, 0, *ora.search_topk_rows(q[i].astype(np.float64), 10), REL_F32)
    assert dev.search_counter == jump + 8
    dev.close()


@pytest.mark.parametrize("storage", ["f32", "bf16"])
def test_fetch_rows_returns_what_the_shard_holds(lib, storage):
    """lvs_fetch_rows_f32 (what shard rebalancing reads back): fp32 shards hold float32(x / ||x||_64), bf16 shards
    the bf16 rounding of x; rows are GLOBAL numbers of a shard whose row_base is beyond 2^32 (the multi-GPU adapter's layout); a row
    written back from its fetched values answers searches exactly like the original."""
    n, dim, base = 3_000, 200, 3 << 32
    x, q = synth.unixcoder_like(n, dim, seed=55, n_queries=4)
    dev = _dev("fetch", dim, storage=storage, row_base=base)
    dev.upsert(x.astype(np.float64))
    pick = np.array([0, 1, 17, 1500, n - 1], dtype=np.int64)
    got = dev.fetch_rows(base + pick)
    if storage == "f32":
        x64 = x[pick].astype(np.float64)
        exp = (x64 / np.linalg.norm(x64, axis=1, keepdims=True)).astype(np.float32)
    else:
        exp = synth.bf16_round(x[pick])
    assert got.dtype == np.float32 and np.array_equal(got, exp)
    with pytest.raises(Exception):
        dev.fetch_rows(np.array([base + n], dtype=np.int64))          # past the end
    with pytest.raises(Exception):
        dev.fetch_rows(np.array([5], dtype=np.int64))                 # below the shard's row_base
    # move: fetched values re-written into a second shard give the same hits (bf16: the same bits; fp32: re-normalised, <= 1 ulp)
    other = _dev("fetch2", dim, storage=storage, row_base=base)
    other.upsert(dev.fetch_rows(base + np.arange(n, dtype=np.int64)).astype(np.float64))
    for i in range(len(q)):
        a, b = dev.search(q[i].astype(np.float64), 10), other.search(q[i].astype(np.float64), 10)
        assert np.array_equal(a.rows, b.rows) and np.abs(a.scores - b.scores).max() < 1e-6
    dev.close(); other.close()


def test_pipelined_sharded_submit_on_one_rank_and_repeat_at(lib):
    """lvs_search_submit_sharded / lvs_search_wait with a world of one rank (host buffers in, merged lists out, completion by the word
    the kernel stores into the pinned slot), four searches in flight; and lvs_search_device_at: a repeat numbered as the search it
    repeats returns exactly what the first attempt returned, without advancing the search counter."""
    import ctypes as C

    import torch
    x, q = synth.unixcoder_like(12_000, 256, seed=44, n_queries=9)
    ora = OracleCollection(256)
    ora.upsert_rows_f32(0, x, [None] * len(x))
    dev = _dev("subsh", 256, timing=False)
    dev.upsert(x.astype(np.float64))
    ex = _exchange_world1(lib, 4, 16)
    tickets = [dev.search_submit_sharded(ex, q[i].astype(np.float64), 10) for i in range(4)]       # all four slots in flight
    got = [dev.search_wait(t) for t in tickets]
    tickets = [dev.search_submit_sharded(ex, q[4:7].astype(np.float64), 10), dev.search_submit_sharded(ex, q[7].astype(np.float64), 5)]
    got += [dev.search_wait(t) for t in tickets]
    exp = [ora.search_topk_rows(q[i].astype(np.float64), 10) for i in range(7)] + [ora.search_topk_rows(q[7].astype(np.float64), 5)]
    for i in range(4):
        _assert_same(got[i], 0, *exp[i], REL_F32)
    for j in range(3):
        _assert_same(got[4], j, *exp[4 + j], REL_F32)
    _assert_same(got[5], 0, *exp[7], REL_F32)
    assert lib.lvs_exchange_error(ex) == 0 and dev.search_counter == 8
    # repeat search number 3 (query 2): same scores bit for bit, counter untouched
    dq = torch.from_numpy(q[2:3].astype(np.float64)).cuda()
    s = torch.zeros((1, 10), dtype=torch.float64, device="cuda"); r = torch.zeros((1, 10), dtype=torch.int64, device="cuda")
    t = torch.zeros((1, 10), dtype=torch.int64, device="cuda"); c = torch.zeros(1, dtype=torch.int32, device="cuda")
    flags = np.zeros(1, dtype=np.int32)
    torch.cuda.synchronize()
    dev.search_device_at(3, dq.data_ptr(), "f64", 1, 10, None, s.data_ptr(), r.data_ptr(), t.data_ptr(), c.data_ptr(), flags)
    assert np.array_equal(r.cpu().numpy()[0], got[2].rows[0]) and np.array_equal(s.cpu().numpy()[0], got[2].scores[0])
    assert flags[0] == 0 and dev.search_counter == 8
    with pytest.raises(Exception):
        dev.search_device_at(99, dq.data_ptr(), "f64", 1, 10, None, s.data_ptr(), r.data_ptr(), t.data_ptr(), c.data_ptr(), flags)
    lib.lvs_exchange_destroy(ex)
    dev.close()
