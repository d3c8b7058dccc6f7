"""``B200CodeEncoder``: the embedding step in front of the vector store, on the same GPU.

The reference embeds code chunks with UniXcoder (``src/lattice/providers/unixcoder_provider.py``): ``UniXcoder.tokenize`` builds
``[CLS] <encoder-only> [SEP] tokens [SEP]`` id lists (:85-135), ``UniXcoder.forward`` runs transformers' ``RobertaModel`` with
bidirectional attention among the non-pad tokens and returns the masked mean of the token embeddings (:137-155), and
``embed_batch_sync`` turns the result into python lists (:194-215) for ``QdrantManager.upsert``.  This class is the device half of
that: the forward pass as hand-written CUDA kernels behind the C ABI (``lvs_encoder_*`` in ``include/lvs.h``), plus
``embed_upsert``, which hands the pooled vectors to the shard's upsert kernel without a host round trip.  Tokenisation is a
host-side dictionary lookup and stays with the caller (pass a Hugging Face tokenizer to get ``embed_batch_sync`` with the
reference's signature).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Mapping, Sequence

import numpy as np

from . import _native as N
from .errors import NativeLibraryError

# parameters of a RobertaModel state dict that the forward pass does not use
_SKIP = ("pooler.", "embeddings.position_ids", "embeddings.token_type_ids", "lm_head.")


class B200CodeEncoder:
    EMBEDDING_DIM = 768          # unixcoder_provider.py:229

    def __init__(self, state_dict: Mapping[str, Any], *, n_layers: int, n_heads: int, pad_id: int = 1, ln_eps: float = 1e-5,
                 device: int = 0, tokenizer=None):
        N.init(device)
        self._lib = N.load()
        sd = {k: np.ascontiguousarray(_to_numpy(v), dtype=np.float32) for k, v in state_dict.items()
              if not any(s in k for s in _SKIP)}
        strip = lambda k: k.split("roberta.", 1)[-1] if k.startswith("roberta.") else k  # noqa: E731
        sd = {strip(k): v for k, v in sd.items()}
        word, pos = sd["embeddings.word_embeddings.weight"], sd["embeddings.position_embeddings.weight"]
        inter = sd["encoder.layer.0.intermediate.dense.weight"].shape[0]
        self.vocab, self.hidden = int(word.shape[0]), int(word.shape[1])
        self.n_layers, self.n_heads, self.pad_id, self.max_pos = int(n_layers), int(n_heads), int(pad_id), int(pos.shape[0])
        cfg = N.EncoderConfig(self.vocab, self.hidden, self.n_layers, self.n_heads, int(inter), self.max_pos, self.pad_id, float(ln_eps))
        h = C.c_void_p()
        N.check(self._lib.lvs_encoder_create(C.byref(cfg), C.byref(h)), "lvs_encoder_create")
        self._h = h
        self.tokenizer = tokenizer
        try:
            for name, arr in sd.items():
                N.check(self._lib.lvs_encoder_load(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), int(arr.size)), f"lvs_encoder_load({name})")
        except Exception:
            self.close()
            raise

    @classmethod
    def from_pretrained(cls, path: str, device: int = 0, tokenizer=None) -> "B200CodeEncoder":
        """A Hugging Face checkpoint directory (``config.json`` + ``pytorch_model.bin`` or ``model.safetensors``), e.g. a local copy of
        ``microsoft/unixcoder-base`` - the model the reference loads (unixcoder_provider.py:70-75)."""
        import json
        import os
        with open(os.path.join(path, "config.json")) as f:
            cfg = json.load(f)
        st = os.path.join(path, "model.safetensors")
        if os.path.exists(st):
            from safetensors.numpy import load_file
            sd = load_file(st)
        else:
            import torch
            sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        return cls(sd, n_layers=cfg["num_hidden_layers"], n_heads=cfg["num_attention_heads"], pad_id=cfg.get("pad_token_id", 1),
                   ln_eps=cfg.get("layer_norm_eps", 1e-5), device=device, tokenizer=tokenizer)

    def _handle(self):
        if not getattr(self, "_h", None):
            raise NativeLibraryError("the encoder is closed")
        return self._h

    # ---- token ids in ---------------------------------------------------------------------------------------------
    def _ids(self, token_ids) -> np.ndarray:
        ids = np.ascontiguousarray(token_ids, dtype=np.int32)
        if ids.ndim == 1:
            ids = ids[None, :]
        if ids.ndim != 2 or ids.shape[1] < 1:
            raise ValueError("token_ids must be [B, L]")
        return ids

    def embed_ids(self, token_ids) -> np.ndarray:
        """[B, L] pad-filled token ids -> [B, hidden] float32 sentence embeddings (``UniXcoder.forward``'s second result)."""
        ids = self._ids(token_ids)
        out = np.empty((ids.shape[0], self.hidden), dtype=np.float32)
        N.check(self._lib.lvs_encoder_embed(self._handle(), ids.ctypes.data_as(C.c_void_p), ids.shape[0], ids.shape[1],
                                            out.ctypes.data_as(C.c_void_p)), "lvs_encoder_embed")
        return out

    def embed_upsert(self, dev, token_ids, rows=None, codes=None, ties=None) -> None:
        """Embed `token_ids` and write the vectors into ``dev`` (a ``DeviceCollection``) at GLOBAL rows `rows` - one device-side hand-over."""
        ids = self._ids(token_ids)
        if not hasattr(dev, "_handle"):
            raise NativeLibraryError("embed_upsert needs a DeviceCollection on this process's GPU (the multi-GPU adapter embeds on rank 0 and upserts vectors)")
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        c = None if codes is None or dev.n_filter_cols == 0 else np.ascontiguousarray(codes, dtype=np.uint32)
        t = None if ties is None else np.ascontiguousarray(ties, dtype=np.uint64)
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        N.check(self._lib.lvs_encoder_embed_upsert(self._handle(), dev._handle(), p(ids), ids.shape[0], ids.shape[1], p(r), p(c), p(t)),
                "lvs_encoder_embed_upsert")

    @property
    def last_ms(self) -> float:
        ms = C.c_float()
        N.check(self._lib.lvs_encoder_last_ms(self._handle(), C.byref(ms)), "lvs_encoder_last_ms")
        return float(ms.value)

    # ---- text in (needs a tokenizer): the reference's call shapes -----------------------------------------------------
    def tokenize(self, inputs: Sequence[str], max_length: int = 512) -> np.ndarray:
        """``UniXcoder.tokenize(mode="<encoder-only>", padding=True)`` (unixcoder_provider.py:85-135), padded to the batch's longest
        sequence instead of ``max_length`` (pad positions are invisible to the forward pass)."""
        if self.tokenizer is None:
            raise NativeLibraryError("no tokenizer: pass tokenizer=RobertaTokenizer.from_pretrained(...) or use embed_ids")
        tk = self.tokenizer
        rows = []
        for x in inputs:
            toks = tk.tokenize(x)[: max_length - 4]
            toks = [tk.cls_token, "<encoder-only>", tk.sep_token] + toks + [tk.sep_token]
            rows.append(tk.convert_tokens_to_ids(toks))
        L = max(len(r) for r in rows)
        out = np.full((len(rows), L), self.pad_id, dtype=np.int32)
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return out

    def embed_batch_sync(self, codes: Sequence[str], max_length: int = 512) -> list[list[float]]:
        """Same call and result shape as the reference's ``embed_batch_sync`` (unixcoder_provider.py:194-215)."""
        if not codes:
            return []
        return self.embed_ids(self.tokenize(codes, max_length)).tolist()

    def embed_code_sync(self, code: str, max_length: int = 512) -> list[float]:
        return self.embed_batch_sync([code], max_length)[0]

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.lvs_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def random_state_dict(vocab: int, hidden: int, n_layers: int, inter: int, max_pos: int, seed: int, pad_id: int = 1) -> dict[str, np.ndarray]:
    """Random RobertaModel-shaped weights (Hugging Face parameter names) for benchmarks and smoke runs - there is no checkpoint in an
    offline image.  Dense weights ~ N(0, 1/in) so that activations stay O(1) through every layer; LayerNorm weights near 1."""
    import math
    rng = np.random.default_rng(seed)
    n = lambda *s: (0.02 * rng.standard_normal(s)).astype(np.float32)  # noqa: E731
    sd = {"embeddings.word_embeddings.weight": n(vocab, hidden), "embeddings.position_embeddings.weight": n(max_pos, hidden),
          "embeddings.token_type_embeddings.weight": n(1, hidden),
          "embeddings.LayerNorm.weight": (1 + 0.1 * rng.standard_normal(hidden)).astype(np.float32), "embeddings.LayerNorm.bias": n(hidden)}
    sd["embeddings.word_embeddings.weight"][pad_id] = 0
    shapes = {"attention.self.query": (hidden, hidden), "attention.self.key": (hidden, hidden), "attention.self.value": (hidden, hidden),
              "attention.output.dense": (hidden, hidden), "intermediate.dense": (inter, hidden), "output.dense": (hidden, inter)}
    for layer in range(n_layers):
        p = f"encoder.layer.{layer}."
        for name, (o, i) in shapes.items():
            sd[p + name + ".weight"] = (rng.standard_normal((o, i)) / math.sqrt(i)).astype(np.float32)
            sd[p + name + ".bias"] = n(o)
        for name in ("attention.output.LayerNorm", "output.LayerNorm"):
            sd[p + name + ".weight"] = (1 + 0.1 * rng.standard_normal(hidden)).astype(np.float32)
            sd[p + name + ".bias"] = n(hidden)
    return sd


def _to_numpy(v):
    if isinstance(v, np.ndarray):
        return v
    if hasattr(v, "detach"):
        return v.detach().cpu().float().numpy()
    return np.asarray(v)
