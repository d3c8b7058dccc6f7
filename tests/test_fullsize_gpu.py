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
