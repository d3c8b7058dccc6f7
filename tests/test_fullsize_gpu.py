"""Parity at BASELINE.json's FULL sizes (configs[1] 1M x 1536 fp32 with payload filters; configs[2] / the metric corpus
10M x 768 bf16, 256 queries, top-100), where the CPU oracle cannot scan the corpus in seconds.

Size-independent construction: for every query, k rows are PLANTED at random positions as normalize(q + eps_j * n_j) with n_j a unit
vector orthogonal to q, so their cosine is 1/sqrt(1 + eps_j^2) - strictly decreasing in j, gaps >= 2.9e-3, all >= 0.39 - while
the other (random) rows stay below ~0.2.  The true top-k of the 10M-row corpus is therefore the planted rows in order; the
oracle (qdrant local mode restatement) is run on the planted rows ALONE and must agree with the device on ids AND scores, after
the same number of earlier searches (the local-mode replay state).  The device result containing only planted rows is what
licenses restricting the oracle to them.  Further properties: batch == sequential, delete the best hits -> the list shifts,
filters select exactly the planted rows that carry the code.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ANY = 0xFFFFFFFF


@pytest.fixture(scope="module")
def lib(native_lib):
    from code_rag_b200 import _native
    _native.init(0)
    return native_lib


def _bf16_round(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16).astype(np.uint32)).view(np.float32).reshape(x.shape)


def _build(name, n, dim, storage, Q, kplant, seed, codes_fn=None, n_cols=0, chunk=250_000):
    """Corpus generated on the GPU in chunks; returns (collection, queries f64 [Q, dim], positions [Q, kplant], planted rows as
    the collection stores them (f32 values) [Q * kplant, dim], codes of the planted rows or None)."""
    import torch

    from code_rag_b200.collection import DeviceCollection
    d = torch.device("cuda")
    g = torch.Generator(device=d); g.manual_seed(seed)
    q = torch.randn((Q, dim), generator=g, device=d, dtype=torch.float32)
    q = q / q.norm(dim=1, keepdim=True)
    noise = torch.randn((Q, kplant, dim), generator=g, device=d, dtype=torch.float32)
    noise = noise - (noise * q[:, None, :]).sum(-1, keepdim=True) * q[:, None, :]
    noise = noise / noise.norm(dim=-1, keepdim=True)
    eps = (0.3 + 0.02 * torch.arange(kplant, device=d, dtype=torch.float32))[None, :, None]
    planted = q[:, None, :] + eps * noise
    planted = (planted / planted.norm(dim=-1, keepdim=True)).reshape(Q * kplant, dim)
    if storage == "bf16":
        planted = planted.to(torch.bfloat16).to(torch.float32)      # what the shard will hold
    pos = torch.randperm(n, generator=g, device=d)[: Q * kplant]
    order = torch.argsort(pos)
    pos_sorted, planted_sorted = pos[order], planted[order]
    codes_all = codes_fn(n, g, d) if codes_fn else None
    dev = DeviceCollection(name, dim, storage=storage, n_filter_cols=n_cols, capacity=n)
    lo_idx = 0
    for row in range(0, n, chunk):
        m = min(chunk, n - row)
        x = torch.randn((m, dim), generator=g, device=d, dtype=torch.float32)
        x = x / x.norm(dim=1, keepdim=True)
        hi_idx = int(torch.searchsorted(pos_sorted, torch.tensor([row + m], device=d)).item())
        if hi_idx > lo_idx:
            x[pos_sorted[lo_idx:hi_idx] - row] = planted_sorted[lo_idx:hi_idx]
        lo_idx = hi_idx
        xs = x.to(torch.bfloat16).contiguous() if storage == "bf16" else x.contiguous()
        cptr = 0
        if codes_all is not None:
            cc = codes_all[row:row + m].contiguous()
            cptr = cc.data_ptr()
        torch.cuda.synchronize()
        dev.upsert_device(xs.data_ptr(), "bf16" if storage == "bf16" else "f32", m, row, codes_ptr=cptr)
    torch.cuda.synchronize()
    pcodes = codes_all[pos].cpu().numpy().astype(np.uint32) if codes_all is not None else None
    return (dev, q.double().cpu().numpy(), pos.reshape(Q, kplant).cpu().numpy().astype(np.int64), planted.cpu().numpy(), pcodes)


def _oracle_on_planted(planted_f32, dim, n_prior_searches):
    from oracle.qdrant_local import OracleCollection
    ora = OracleCollection(dim)
    ora.upsert_rows_f32(0, planted_f32, [None] * len(planted_f32))
    probe = np.ones(dim)
    for _ in range(n_prior_searches):
        ora.search_topk_rows(probe, 1)        # a local-mode search re-normalises the matrix in place, whatever the query
    return ora


def test_metric_corpus_10m_bf16_q256_top100_and_single_queries(lib):
    n, dim, Q, kp = 10_000_000, 768, 256, 100
    dev, q, pos, planted, _ = _build("full10m", n, dim, "bf16", Q, kp, seed=3456)
    flat_pos = pos.reshape(-1)
    try:
        # ---- K2 (CTA-pair form): 256 queries x top-100 over the whole corpus ----
        ora = _oracle_on_planted(planted, dim, dev.search_counter)
        res = dev.search(q, kp)
        assert dev.last_timing()["kernel"] == "gemm"
        assert (res.flags == 0).all() and (res.counts == kp).all()
        for i in range(Q):
            assert np.array_equal(res.rows[i], pos[i]), f"query {i}: ids differ from the planted order"
            rows_o, scores_o = ora.search_topk_rows(q[i], kp)            # search number i+1 on both sides
            assert np.array_equal(flat_pos[rows_o], res.rows[i])
            assert np.allclose(res.scores[i], scores_o, rtol=2e-3, atol=0)                 # the bar north_star states for bf16 storage
            assert np.abs(res.scores[i] - scores_o).max() < 1e-9                           # what is actually achieved
            assert np.all(np.diff(res.scores[i]) < 0)
        # ---- K1: single queries, top-10; and a batch of 2 (K1) / 7 (K2, single-CTA form) equals them ----
        singles = []
        for i in (0, 17, 255):
            r1 = dev.search(q[i], 10)
            assert dev.last_timing()["kernel"] == "scan"
            rows_o, scores_o = ora.search_topk_rows(q[i], 10)
            assert np.array_equal(r1.rows[0], pos[i, :10]) and np.array_equal(flat_pos[rows_o], r1.rows[0])
            assert np.abs(r1.scores[0] - scores_o).max() < 1e-9 and r1.flags[0] == 0
            singles.append(r1)
        for idx in ([3, 4], [5, 6, 7, 8, 9, 10, 11]):
            rb = dev.search(q[idx], 10)
            for j, i in enumerate(idx):
                rows_o, scores_o = ora.search_topk_rows(q[i], 10)
                assert np.array_equal(rb.rows[j], pos[i, :10])
                assert np.abs(rb.scores[j] - scores_o).max() < 1e-9
        # ---- deleting the three best hits of query 0 shifts its list ----
        dev.delete_rows(pos[0, :3])
        ora.deleted[[0, 1, 2]] = True
        r2 = dev.search(q[0], 10)
        rows_o, scores_o = ora.search_topk_rows(q[0], 10)
        assert np.array_equal(r2.rows[0], pos[0, 3:13]) and np.array_equal(flat_pos[rows_o], r2.rows[0])
        assert np.abs(r2.scores[0] - scores_o).max() < 1e-9
        assert dev.count() == n - 3
    finally:
        dev.close()


def test_c2_1m_x_1536_fp32_filters(lib):
    """configs[1]: single-query top-10 over 1M x 1536 fp32 with project / language payload filters (K1 scan with the
    filter fused into the producer), plus a batch through K2 (kind::tf32) with the same filter."""
    import torch
    n, dim, Q, kp = 1_000_000, 1536, 12, 40

    def codes_fn(n_, g, d):
        proj = torch.randint(1, 9, (n_,), device=d, generator=g, dtype=torch.int32)
        lang = torch.randint(1, 4, (n_,), device=d, generator=g, dtype=torch.int32)
        return torch.stack([proj, lang], dim=1).contiguous()
    dev, q, pos, planted, pcodes = _build("full1m", n, dim, "f32", Q, kp, seed=2345, codes_fn=codes_fn, n_cols=2)
    pcodes = pcodes.reshape(Q, kp, 2)
    flat_pos = pos.reshape(-1)
    try:
        ora = _oracle_on_planted(planted, dim, dev.search_counter)
        for i in range(6):
            for want in (None, [3, ANY], [5, 2]):
                mask_all = np.ones(Q * kp, dtype=bool)
                sel = np.ones(kp, dtype=bool)
                if want is not None:
                    for c, w in enumerate(want):
                        if w != ANY:
                            sel &= pcodes[i, :, c] == w
                            mask_all &= pcodes.reshape(-1, 2)[:, c] == w
                expect = pos[i][sel][:10]
                res = dev.search(q[i], 10, None if want is None else np.array(want, dtype=np.uint32))
                assert dev.last_timing()["kernel"] == "scan"
                rows_o, scores_o = ora.search_topk_rows(q[i], 10, mask_all)
                got = res.rows[0, :res.counts[0]]
                # planted rows that pass the filter come first, in planted order (random rows may fill the tail of a sparse filter)
                assert np.array_equal(got[:len(expect)], expect), (i, want)
                m = min(len(expect), 10)
                if m == 0:
                    continue
                assert np.array_equal(flat_pos[rows_o[:m]], got[:m])
                assert np.allclose(res.scores[0, :m], scores_o[:m], rtol=1e-5, atol=0)     # the bar for fp32 storage
                assert np.abs(res.scores[0, :m] - scores_o[:m]).max() < 1e-9
        # batch of 12 through the tensor-core path on the fp32 shard, filter project = 3
        want = np.array([3, ANY], dtype=np.uint32)
        mask_all = pcodes.reshape(-1, 2)[:, 0] == 3
        rb = dev.search(q, 10, want)
        assert dev.last_timing()["kernel"] == "gemm"
        for i in range(Q):
            expect = pos[i][pcodes[i, :, 0] == 3][:10]
            rows_o, scores_o = ora.search_topk_rows(q[i], 10, mask_all)
            m = min(len(expect), 10)
            if m == 0:
                continue
            assert np.array_equal(rb.rows[i, :m], expect[:m]) and np.array_equal(flat_pos[rows_o[:m]], rb.rows[i, :m])
            assert np.abs(rb.scores[i, :m] - scores_o[:m]).max() < 1e-9
    finally:
        dev.close()


# ----------------------------------------------------------------------------------------------------------------------------
# SURVEY section 8c, last row: a full-corpus SECONDARY oracle.  No planting: the corpus is random (or adversarial), so the top-100
# gaps are the natural ones (~8e-5 at 10M rows) and the tensor-core path's bf16-query error (~1e-3) really has to be repaired by
# the exact rescoring.  While the corpus is generated chunk by chunk on the GPU, float64 cosines of every row against every query
# are computed with torch (test infrastructure) and a running top-128 per query is kept.  The CPU oracle (qdrant local mode
# restatement) then runs on the UNION of those candidates only - licensed by an assertion that the 100th and the 128th float64
# cosine of every query are further apart than local mode's float32 storage can move a score (1e-6) - in the same search order as
# the device, so ids must agree exactly and scores to 1e-9.
# ----------------------------------------------------------------------------------------------------------------------------
def _build_with_fp64_oracle(name, n, dim, q64, kk, gen_chunk, chunk=250_000):
    import torch

    from code_rag_b200.collection import DeviceCollection
    d = torch.device("cuda")
    Q = q64.shape[0]
    best_s = torch.full((Q, kk), -float("inf"), dtype=torch.float64, device=d)
    best_r = torch.full((Q, kk), -1, dtype=torch.int64, device=d)
    dev = DeviceCollection(name, dim, storage="bf16", capacity=n)
    for row in range(0, n, chunk):
        m = min(chunk, n - row)
        xb = gen_chunk(row, m).to(torch.bfloat16).contiguous()
        torch.cuda.synchronize()
        dev.upsert_device(xb.data_ptr(), "bf16", m, row)
        x = xb.double()
        x = x / x.norm(dim=1, keepdim=True)
        s = q64 @ x.T                                                    # float64 cosines [Q, m]
        ts, ti = torch.topk(s, kk, dim=1)
        cs, cr = torch.cat([best_s, ts], 1), torch.cat([best_r, ti + row], 1)
        o = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :kk]
        best_s, best_r = cs.gather(1, o), cr.gather(1, o)
        del x, s, xb
    torch.cuda.synchronize()
    return dev, best_s.cpu().numpy(), best_r.cpu().numpy()


class _SubsetOracle:
    """The CPU oracle over the union of the candidate rows, addressed by GLOBAL row numbers; searched in the device's order."""

    def __init__(self, dev, cand_rows, dim):
        from oracle.qdrant_local import OracleCollection
        self.rows = np.unique(cand_rows.reshape(-1))                     # ascending: subset order == global row order (ties by row)
        self.ora = OracleCollection(dim)
        self.ora.upsert_rows_f32(0, dev.fetch_rows(self.rows), [None] * len(self.rows))
        probe = np.ones(dim)
        for _ in range(dev.search_counter):
            self.ora.search_topk_rows(probe, 1)

    def search(self, q, k):
        r, s = self.ora.search_topk_rows(q, k)
        return self.rows[r], s


def _check(res, i, exp_rows, exp_scores, what):
    n = len(exp_rows)
    assert int(res.counts[i]) == n, what
    assert np.array_equal(res.rows[i, :n], exp_rows), f"{what}: ids differ\n device {res.rows[i, :n].tolist()}\n oracle {exp_rows.tolist()}"
    assert np.allclose(res.scores[i, :n], exp_scores, rtol=2e-3, atol=0), what            # the bar north_star states for bf16 storage
    assert np.abs(res.scores[i, :n] - exp_scores).max() < 1e-9, what                       # what is actually achieved


def test_random_corpus_10m_against_the_full_corpus_fp64_oracle(lib):
    """10M x 768 bf16, NO planted neighbours, 256 queries, top-100: K2 pair form (256), K2 single form (32), K1 (single queries),
    the pipelined host API and the enqueue-only device API (whose flagged queries are counted, repeated and must then be exact)."""
    import torch

    from code_rag_b200.collection import SearchResult
    n, dim, Q, k, kk = 10_000_000, 768, 256, 100, 128
    d = torch.device("cuda")
    g = torch.Generator(device=d); g.manual_seed(777)
    q = torch.randn((Q, dim), generator=g, device=d, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)

    def gen_chunk(row, m):
        gg = torch.Generator(device=d); gg.manual_seed(9_000_000 + row)
        x = torch.randn((m, dim), generator=gg, device=d, dtype=torch.float32)
        return x / x.norm(dim=1, keepdim=True)
    dev, cs, cr = _build_with_fp64_oracle("rand10m", n, dim, q, kk, gen_chunk)
    qh = q.cpu().numpy()
    try:
        gap = cs[:, k - 1] - cs[:, kk - 1]
        assert gap.min() > 1e-6, f"candidate margin too small ({gap.min()}): raise kk"
        adj = np.diff(-cs[:, :k], axis=1)
        print(f"natural top-100 gaps: median {np.median(adj):.2e}, smallest {adj.min():.2e}; margin to the 128th {gap.min():.2e}")
        ora = _SubsetOracle(dev, cr, dim)
        # ---- K2, CTA-pair form: all 256 queries in one call (host API: flagged queries are repeated on the exact scan) ----
        res = dev.search(qh, k)
        assert dev.last_timing()["kernel"] == "gemm" and (res.flags == 0).all()
        for i in range(Q):
            _check(res, i, *ora.search(qh[i], k), f"K2 pair, query {i}")
        # ---- K2, single-CTA form: 32 queries ----
        res = dev.search(qh[:32], k)
        assert dev.last_timing()["kernel"] == "gemm" and (res.flags == 0).all()
        for i in range(32):
            _check(res, i, *ora.search(qh[i], k), f"K2 single, query {i}")
        # ---- K1: single queries (one fused kernel each) ----
        for i in range(32, 48):
            r1 = dev.search(qh[i], k)
            assert dev.last_timing()["kernel"] == "scan" and r1.flags[0] == 0
            _check(r1, 0, *ora.search(qh[i], k), f"K1, query {i}")
        # ---- pipelined host API: three searches in flight (a batch of 64 on K2, two single queries on K1) ----
        tk = [dev.search_submit(qh[48:112], k), dev.search_submit(qh[112], k), dev.search_submit(qh[113], 10)]
        got = [dev.search_wait(t) for t in tk]
        for j in range(64):
            _check(got[0], j, *ora.search(qh[48 + j], k), f"submit/wait batch, query {48 + j}")
        _check(got[1], 0, *ora.search(qh[112], k), "submit/wait single")
        _check(got[2], 0, *ora.search(qh[113], 10), "submit/wait single, top-10")
        # ---- enqueue-only device API (what bench.py's value leg and every sharded search use): flags stay on the device and
        #      flagged queries are NOT repeated by the library; the rate is recorded, unflagged results must already be exact and
        #      ShardedSearcher.search_device repeats the flagged ones ----
        from code_rag_b200.sharded import ShardedSearcher
        ss = ShardedSearcher(dev)
        dq = q[114:242].contiguous()
        s_, r_, t_, c_, f_ = ss.search_device_async(dq, k)
        ss.stream.synchronize()
        raw = SearchResult(s_.cpu().numpy(), r_.cpu().numpy(), t_.cpu().numpy().view(np.uint64), c_.cpu().numpy().astype(np.uint32), f_.cpu().numpy())
        exp = [ora.search(qh[114 + j], k) for j in range(128)]
        n_flagged = int((raw.flags != 0).sum())
        print(f"enqueue-only K2 pass over a random 10M corpus, 128 queries x top-100: {n_flagged} flagged")
        assert n_flagged <= 6, "the tensor-core path should prove almost every query of a random corpus with its default candidate set"
        for j in range(128):
            if raw.flags[j] == 0:
                _check(raw, j, *exp[j], f"enqueue-only, unflagged query {114 + j}")
        dq2 = q[242:256].contiguous()
        s_, r_, t_, c_, flags = ss.search_device(dq2, k)                # synchronous: flagged queries are repeated
        fin = SearchResult(s_.cpu().numpy(), r_.cpu().numpy(), t_.cpu().numpy().view(np.uint64), c_.cpu().numpy().astype(np.uint32), flags)
        assert (flags == 0).all()
        for j in range(14):
            _check(fin, j, *ora.search(qh[242 + j], k), f"ShardedSearcher.search_device, query {242 + j}")
    finally:
        dev.close()


def test_adversarial_corpus_10m_drifting_scores_and_near_duplicates(lib):
    """10M x 768 bf16 built to hurt a streaming top-k: (a) the score against query 1 DRIFTS UPWARDS with the row number (sorted by
    score up to the bf16 rounding noise), so the running threshold keeps being beaten until the last tile and the tensor-core
    path's 16-key lists overflow; (b) the last 10,000 rows are near-duplicates of query 0 (cosines 0.995 .. 0.99995, ~5e-7 apart),
    far denser than the tensor-core path's error bound, so its queries must be repaired by the exact scan.  Random queries see a
    random corpus.  Everything must still equal the oracle, with no flag left."""
    import torch
    n, dim, Q, k, kk = 10_000_000, 768, 8, 100, 224
    n_dup = 10_000
    d = torch.device("cuda")
    g = torch.Generator(device=d); g.manual_seed(4242)
    q = torch.randn((Q, dim), generator=g, device=d, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    q0, q1 = q[0].float(), q[1].float()

    def gen_chunk(row, m):
        gg = torch.Generator(device=d); gg.manual_seed(5_000_000 + row)
        z = torch.randn((m, dim), generator=gg, device=d, dtype=torch.float32)
        z = z - (z @ q1)[:, None] * q1[None, :]
        z = z / z.norm(dim=1, keepdim=True)
        idx = torch.arange(row, row + m, device=d, dtype=torch.float32)
        a = (-0.25 + 0.5 * idx / n)[:, None]                           # cosine against q1 rises with the row number
        x = a * q1[None, :] + torch.sqrt(1 - a * a) * z
        lo = max(row, n - n_dup)
        if lo < row + m:                                                # the tail: near-duplicates of q0
            t = lo - row
            sig = torch.linspace(0.1, 0.01, n_dup, device=d)[lo - (n - n_dup): lo - (n - n_dup) + (m - t), None]
            w = torch.randn((m - t, dim), generator=gg, device=d, dtype=torch.float32) / dim ** 0.5
            y = q0[None, :] + sig * w
            x[t:] = y / y.norm(dim=1, keepdim=True)
        return x
    dev, cs, cr = _build_with_fp64_oracle("adv10m", n, dim, q, kk, gen_chunk)
    qh = q.cpu().numpy()
    try:
        gap = cs[:, k - 1] - cs[:, kk - 1]
        assert gap.min() > 1e-6, f"candidate margin too small ({gap.min()})"
        assert (cr[0, :k] >= n - n_dup).all() and (cr[1, :k] >= n - 200_000).all()       # the corpus is what the docstring says
        ora = _SubsetOracle(dev, cr, dim)
        res = dev.search(qh, k)                                         # 8 queries: tensor-core pass, then the exact scan for whoever it flags
        assert (res.flags == 0).all(), res.flags
        for i in range(Q):
            _check(res, i, *ora.search(qh[i], k), f"batch, query {i}")
        for i in range(Q):                                              # single queries on the scan path, top-100 and top-10
            r1 = dev.search(qh[i], k)
            assert r1.flags[0] == 0
            _check(r1, 0, *ora.search(qh[i], k), f"single, query {i}")
        for i in (0, 1):
            r1 = dev.search(qh[i], 10)
            assert r1.flags[0] == 0
            _check(r1, 0, *ora.search(qh[i], 10), f"single top-10, query {i}")
    finally:
        dev.close()
