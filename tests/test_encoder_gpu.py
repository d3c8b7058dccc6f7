"""SURVEY section 8f row 4 on the device: the code encoder (tcgen05 GEMMs + attention / LayerNorm / pooling kernels behind
lvs_encoder_*) against the float32 oracle (oracle/roberta_encoder.py) and against the fixture made by transformers' own
RobertaModel, and the embed -> upsert hand-over against the ordinary upsert of the same vectors.

Tolerance: the device keeps activations in bf16 (8 mantissa bits) between layers and accumulates in float32; the oracle is float32
throughout.  Stated bar: cosine similarity of every sentence embedding >= 0.999 and max |difference| <= 3e-2 of the embedding's
largest component; observed values are printed."""
import asyncio
from pathlib import Path

import numpy as np
import pytest

from oracle import roberta_encoder as R

pytestmark = pytest.mark.gpu

G = np.load(Path(__file__).parent / "golden" / "roberta_encoder_golden.npz")


@pytest.fixture(scope="module")
def lib(native_lib):
    from code_rag_b200 import _native
    _native.init(0)
    return native_lib


def _close(a, b, what):
    cos = (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
    err = np.abs(a - b).max(1) / np.abs(b).max(1)
    print(f"{what}: min cosine {cos.min():.6f}, max relative component error {err.max():.4f}")
    assert cos.min() >= 0.999, what
    assert err.max() <= 3e-2, what


def test_encoder_matches_transformers_fixture(lib):
    from code_rag_b200.embedding import B200CodeEncoder
    vocab, hidden, layers, heads, inter, max_pos, pad = (int(v) for v in G["config"])
    sd = {k[2:]: G[k] for k in G.files if k.startswith("w:")}
    enc = B200CodeEncoder(sd, n_layers=layers, n_heads=heads, pad_id=pad)
    try:
        got = enc.embed_ids(G["ids"])
        _close(got, G["sentence_embeddings"], "device vs transformers RobertaModel (hidden 128, 3 layers, ragged batch)")
        # rows are independent and padding is invisible: one sequence alone, and the batch re-padded to another length
        one = enc.embed_ids(G["ids"][2:3, :17])
        assert np.abs(one - got[2:3]).max() <= 2e-2 * np.abs(got[2]).max()
        longer = np.full((6, 64), pad, dtype=np.int32); longer[:, :48] = G["ids"]
        assert np.abs(enc.embed_ids(longer) - got).max() <= 2e-2 * np.abs(got).max()
        with pytest.raises(Exception):
            enc.embed_ids(np.full((1, 8), vocab + 5, dtype=np.int32))             # id outside the vocabulary
    finally:
        enc.close()


@pytest.mark.parametrize("B,L", [(3, 40), (5, 200), (2, 512)])
def test_encoder_unixcoder_shape_vs_oracle(lib, B, L):
    """RoBERTa-base dimensions (UniXcoder: 12 layers, 768 wide, 12 heads, 3072 intermediate), random weights, ragged lengths."""
    from code_rag_b200.embedding import B200CodeEncoder
    vocab, hidden, layers, heads, inter, max_pos = 2000, 768, 12, 12, 3072, 1026
    sd = R.random_state_dict(vocab, hidden, layers, inter, max_pos, seed=5)
    rng = np.random.default_rng(B * 1000 + L)
    ids = rng.integers(3, vocab, size=(B, L)).astype(np.int32)
    for b in range(B):
        n = int(rng.integers(max(2, L // 4), L + 1)) if b else L
        ids[b, n:] = 1
    enc = B200CodeEncoder(sd, n_layers=layers, n_heads=heads, pad_id=1)
    try:
        got = enc.embed_ids(ids)
        _, exp = R.encode(sd, ids, n_layers=layers, n_heads=heads, pad_id=1)
        _close(got, exp, f"device vs float32 oracle, B={B} L={L}, forward {enc.last_ms:.3f} ms")
    finally:
        enc.close()


def test_embed_upsert_equals_embed_then_upsert(lib):
    """The vectors that never leave the GPU (lvs_encoder_embed_upsert) index exactly like the same vectors upserted from the host,
    through the adapter: ids, payload filters and search results agree; the reference's flow is embed_batch -> upsert
    (embeddings/indexer.py:77-86)."""
    from code_rag_b200.client import B200VectorStore
    from code_rag_b200.embedding import B200CodeEncoder
    vocab, hidden, layers, heads, inter, max_pos, pad = (int(v) for v in G["config"])
    sd = {k[2:]: G[k] for k in G.files if k.startswith("w:")}
    enc = B200CodeEncoder(sd, n_layers=layers, n_heads=heads, pad_id=pad)
    rng = np.random.default_rng(3)
    n, L = 300, 32
    tok = rng.integers(3, vocab, size=(n, L)).astype(np.int32)
    for i in range(n):
        tok[i, int(rng.integers(4, L + 1)):] = pad
    ids = [str(__import__("uuid").UUID(int=int(v))) for v in rng.integers(1, 2**62, size=n)]
    pl = [{"file_path": f"f{i % 7}.py", "entity_name": f"e{i}", "language": ("python", "go")[i % 2]} for i in range(n)]

    async def run():
        a, b = B200VectorStore(dimensions=hidden), B200VectorStore(dimensions=hidden)
        await a.connect(); await b.connect(); await a.create_collections(); await b.create_collections()
        await a.upsert_tokens("code_chunks", ids, tok, pl, enc)
        vec = enc.embed_ids(tok)
        await b.upsert("code_chunks", ids, vec.astype(np.float64).tolist(), pl)
        assert (await a.get_collection_info("code_chunks")).points_count == n
        q = enc.embed_ids(tok[:5])
        for i in range(5):
            for flt in (None, {"language": "go"}, {"file_path": "f3.py"}):
                ra = await a.search(collection="code_chunks", query_vector=q[i].tolist(), limit=10, filters=flt)
                rb = await b.search(collection="code_chunks", query_vector=q[i].tolist(), limit=10, filters=flt)
                assert [h["id"] for h in ra] == [h["id"] for h in rb]
                assert max(abs(x["score"] - y["score"]) for x, y in zip(ra, rb)) < 1e-9
                assert [h["payload"] for h in ra] == [h["payload"] for h in rb]
            if i == 0:
                assert ra and (await a.search(collection="code_chunks", query_vector=q[0].tolist(), limit=1))[0]["id"] == ids[0]
        # overwrite by id through the token path keeps one point per id
        await a.upsert_tokens("code_chunks", ids[:3], tok[10:13], pl[:3], enc)
        assert (await a.get_collection_info("code_chunks")).points_count == n
        await a.close(); await b.close()
    asyncio.run(run())
    enc.close()


def test_pair_form_of_the_dense_layers_equals_the_single_cta_form(lib, monkeypatch):
    """Large batches run the dense layers as clusters of two CTAs (tcgen05 cta_group::2, 256-row tiles, each CTA loads half of the
    weight tile); small ones as single CTAs.  Same K order, same accumulators: the sentence embeddings must agree bit for bit - on a
    batch whose token count is not a multiple of the 256-row tile - and the large batch must still match the float32 oracle."""
    from code_rag_b200.embedding import B200CodeEncoder
    vocab, hidden, layers, heads, inter, max_pos = 2000, 768, 4, 12, 3072, 1026
    sd = R.random_state_dict(vocab, hidden, layers, inter, max_pos, seed=9)
    rng = np.random.default_rng(12)
    B, L = 47, 384                                     # 18,048 tokens = 70.5 tiles of 256 rows
    ids = rng.integers(3, vocab, size=(B, L)).astype(np.int32)
    for b in range(1, B):
        ids[b, int(rng.integers(L // 3, L + 1)):] = 1
    monkeypatch.setenv("LATTICE_B200_ENCODER_PAIR", "0")
    single = B200CodeEncoder(sd, n_layers=layers, n_heads=heads, pad_id=1)
    monkeypatch.setenv("LATTICE_B200_ENCODER_PAIR", "1")
    pair = B200CodeEncoder(sd, n_layers=layers, n_heads=heads, pad_id=1)
    try:
        a, t_single = single.embed_ids(ids), single.last_ms
        b_, t_pair = pair.embed_ids(ids), pair.last_ms
        print(f"4 layers, {B} x {L}: single-CTA form {t_single:.3f} ms, CTA-pair form {t_pair:.3f} ms")
        assert np.array_equal(a, b_)
        _, exp = R.encode(sd, ids[:6], n_layers=layers, n_heads=heads, pad_id=1)
        _close(b_[:6], exp, "pair form vs float32 oracle (first 6 sequences)")
    finally:
        single.close(); pair.close()
