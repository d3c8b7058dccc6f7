"""CPU oracle for the embedding step in front of the hot path.  TEST INFRASTRUCTURE ONLY.

The reference embeds code chunks with UniXcoder (reference ``src/lattice/providers/unixcoder_provider.py:137-155``):
``RobertaModel.from_pretrained("microsoft/unixcoder-base")`` run over the token ids with an attention mask that lets every
non-pad token attend to every non-pad token (``mask.unsqueeze(1) * mask.unsqueeze(2)``), then a masked mean over the token
embeddings (``(token_embeddings * mask).sum(1) / mask.sum(-1)``); ``embed_batch_sync`` (``:194-215``) turns the result into
python lists, which ``VectorIndexer.index_file`` (``embeddings/indexer.py:46-94``) hands to ``QdrantManager.upsert``.

This module restates that forward pass in float32 numpy, from a Hugging Face ``RobertaModel`` state dict (same parameter names,
so a real ``microsoft/unixcoder-base`` checkpoint can be used when one is on disk; there is none in this image and no network,
tests use random-initialised weights):

* embeddings: ``word[id] + position[pad + cumsum(mask)] + token_type[0]`` -> LayerNorm (eps from the config);
* 12 (n) post-LN encoder layers: multi-head self-attention (``softmax(q k^T / sqrt(d) + key mask) v``), output dense, residual +
  LayerNorm; intermediate dense + exact (erf) GELU, output dense, residual + LayerNorm;
* masked mean pooling over the non-pad tokens.

Pinned: ``tests/golden/roberta_encoder_golden.npz`` was produced by ``transformers``' own ``RobertaModel`` (the class the
reference instantiates) in this container (``tests/golden/make_roberta_golden.py``; a 2-D key mask, which gives the same values
as the reference's 3-D mask on every non-pad row and the pad rows are dropped by the pooling - transformers 5.x no longer accepts
the 3-D form); ``tests/test_encoder_oracle.py`` checks this restatement against it on CPU, ``tests/test_encoder_gpu.py`` checks
the CUDA path against this restatement and against the fixture.
"""
from __future__ import annotations

import math

import numpy as np


def _ln(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


try:
    from scipy.special import erf as _erf          # vectorised; math.erf element by element is the fallback
except Exception:  # noqa: BLE001
    _erf = np.vectorize(math.erf, otypes=[np.float64])


def _gelu(x: np.ndarray) -> np.ndarray:
    return (0.5 * x * (1.0 + _erf(x.astype(np.float64) / math.sqrt(2.0)))).astype(x.dtype)


def encode(sd: dict[str, np.ndarray], ids: np.ndarray, *, n_layers: int, n_heads: int, pad_id: int = 1, eps: float = 1e-5,
           dtype=np.float32) -> tuple[np.ndarray, np.ndarray]:
    """``sd``: RobertaModel state dict as numpy arrays; ``ids`` [B, L] token ids, ``pad_id``-padded.
    Returns (token_embeddings [B, L, H], sentence_embeddings [B, H]) - the pair ``UniXcoder.forward`` returns."""
    ids = np.asarray(ids, dtype=np.int64)
    B, L = ids.shape
    mask = ids != pad_id
    g = lambda k: np.asarray(sd[k], dtype=dtype)  # noqa: E731
    pos = np.cumsum(mask, axis=1) * mask + pad_id                       # transformers create_position_ids_from_input_ids
    x = g("embeddings.word_embeddings.weight")[ids] + g("embeddings.position_embeddings.weight")[pos] + \
        g("embeddings.token_type_embeddings.weight")[0]
    x = _ln(x, g("embeddings.LayerNorm.weight"), g("embeddings.LayerNorm.bias"), eps)
    H = x.shape[-1]
    d = H // n_heads
    neg = np.where(mask[:, None, None, :], 0.0, -np.inf).astype(dtype)   # keys that are padding are never attended to
    for l in range(n_layers):
        p = f"encoder.layer.{l}."
        lin = lambda t, name: t @ g(p + name + ".weight").T + g(p + name + ".bias")  # noqa: E731
        q = lin(x, "attention.self.query").reshape(B, L, n_heads, d).transpose(0, 2, 1, 3)
        k = lin(x, "attention.self.key").reshape(B, L, n_heads, d).transpose(0, 2, 1, 3)
        v = lin(x, "attention.self.value").reshape(B, L, n_heads, d).transpose(0, 2, 1, 3)
        s = q @ k.transpose(0, 1, 3, 2) / math.sqrt(d) + neg
        s = s - s.max(-1, keepdims=True)
        e = np.exp(s)
        a = e / e.sum(-1, keepdims=True)
        ctx = (a @ v).transpose(0, 2, 1, 3).reshape(B, L, H)
        x = _ln(lin(ctx, "attention.output.dense") + x, g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias"), eps)
        h = _gelu(lin(x, "intermediate.dense"))
        x = _ln(lin(h, "output.dense") + x, g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"), eps)
    m = mask[..., None].astype(dtype)
    sent = (x * m).sum(1) / m.sum(1)
    return x, sent


def random_state_dict(vocab: int, hidden: int, n_layers: int, inter: int, max_pos: int, seed: int, pad_id: int = 1) -> dict[str, np.ndarray]:
    """Random weights with the initialisation scale of a RoBERTa checkpoint (N(0, 0.02); LayerNorm weights near 1, biases small
    but non-zero so that a kernel that forgets one fails).  Stands in for the checkpoint in tests and in the benchmark."""
    rng = np.random.default_rng(seed)
    n = lambda *s: (0.02 * rng.standard_normal(s)).astype(np.float32)  # noqa: E731
    sd = {"embeddings.word_embeddings.weight": n(vocab, hidden), "embeddings.position_embeddings.weight": n(max_pos, hidden),
          "embeddings.token_type_embeddings.weight": n(1, hidden),
          "embeddings.LayerNorm.weight": (1 + 0.1 * rng.standard_normal(hidden)).astype(np.float32), "embeddings.LayerNorm.bias": n(hidden)}
    sd["embeddings.word_embeddings.weight"][pad_id] = 0
    for l in range(n_layers):
        p = f"encoder.layer.{l}."
        for name, (o, i) in {"attention.self.query": (hidden, hidden), "attention.self.key": (hidden, hidden), "attention.self.value": (hidden, hidden),
                             "attention.output.dense": (hidden, hidden), "intermediate.dense": (inter, hidden), "output.dense": (hidden, inter)}.items():
            sd[p + name + ".weight"] = (rng.standard_normal((o, i)) / math.sqrt(i)).astype(np.float32)   # keeps activations O(1) through 12 layers
            sd[p + name + ".bias"] = n(o)
        for name in ("attention.output.LayerNorm", "output.LayerNorm"):
            sd[p + name + ".weight"] = (1 + 0.1 * rng.standard_normal(hidden)).astype(np.float32)
            sd[p + name + ".bias"] = n(hidden)
    return sd
