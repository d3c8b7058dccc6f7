#!/usr/bin/env python
"""Generates tests/golden/roberta_encoder_golden.npz with transformers' own RobertaModel - the class the reference instantiates
(reference src/lattice/providers/unixcoder_provider.py:70-75) - on a small random-initialised configuration: weights, token ids
(ragged, pad-filled), token embeddings and masked-mean sentence embeddings as UniXcoder.forward (:137-155) computes them.

    python tests/golden/make_roberta_golden.py            (CPU; transformers 5.5, torch 2.11 in this image)
"""
from pathlib import Path

import numpy as np
import torch
from transformers import RobertaConfig, RobertaModel

ROOT = Path(__file__).resolve().parents[2]


def main():
    torch.manual_seed(20261018)
    cfg = RobertaConfig(vocab_size=300, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256,
                        max_position_embeddings=70, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5)
    m = RobertaModel(cfg, add_pooling_layer=False).eval()
    with torch.no_grad():           # the default init leaves biases at zero and LayerNorms at identity: perturb them so that they matter
        for name, p in m.named_parameters():
            if name.endswith("bias") or "LayerNorm" in name:
                p.add_(0.05 * torch.randn_like(p))
            elif "dense.weight" in name or "query.weight" in name or "key.weight" in name or "value.weight" in name:
                p.mul_(4.0)         # N(0, 0.02) weights make every layer nearly a no-op at this width
    ids = torch.randint(3, 300, (6, 48))
    for b, n in enumerate((48, 31, 17, 5, 1, 40)):
        ids[b, n:] = 1
    mask = ids.ne(1)
    with torch.no_grad():
        tok = m(ids, attention_mask=mask)[0]
        sent = (tok * mask.unsqueeze(-1)).sum(1) / mask.sum(-1).unsqueeze(-1)
    out = {"ids": ids.numpy().astype(np.int32), "token_embeddings": tok.numpy(), "sentence_embeddings": sent.numpy(),
           "config": np.array([cfg.vocab_size, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads, cfg.intermediate_size,
                               cfg.max_position_embeddings, cfg.pad_token_id], dtype=np.int64)}
    for k, v in m.state_dict().items():
        out["w:" + k] = v.numpy()
    path = ROOT / "tests" / "golden" / "roberta_encoder_golden.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
