#!/usr/bin/env python
"""Embedding throughput on one B200: chunks/s and tokens/s of the code encoder at UniXcoder's shape
(RoBERTa-base: 12 layers, 768 wide, 12 heads, 3072 intermediate; random weights - there is no checkpoint in this image), next to
the reference's own path on the same GPU: transformers' RobertaModel run the way unixcoder_provider.py:137-155 runs it (fp32 eager,
and bf16 for context), masked mean pooling included.  One JSON line per (batch, length).

    python benchmarks/encoder_bench.py [--no-torch]

FLOPs per forward = B L (12 layers x 2 (3 H^2 + H^2 + 2 H I)) for the dense layers + 12 x 4 B L^2 H for attention; the fraction of the
measured bf16 tensor peak (MEASURED_PEAKS.json) is reported for the whole forward pass (device time by CUDA events)."""
from __future__ import annotations

import json
import statistics
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from code_rag_b200.embedding import B200CodeEncoder, random_state_dict  # noqa: E402

VOCAB, H, LAYERS, HEADS, INTER, MAXPOS = 51416, 768, 12, 12, 3072, 1026


def main():
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    tpeak = float(peaks.get("bf16_tflops", 1590.0))
    sd = random_state_dict(VOCAB, H, LAYERS, INTER, MAXPOS, seed=1)
    enc = B200CodeEncoder(sd, n_layers=LAYERS, n_heads=HEADS, pad_id=1)
    hf = None
    if "--no-torch" not in sys.argv:
        import torch
        from transformers import RobertaConfig, RobertaModel
        cfg = RobertaConfig(vocab_size=VOCAB, hidden_size=H, num_hidden_layers=LAYERS, num_attention_heads=HEADS, intermediate_size=INTER,
                            max_position_embeddings=MAXPOS, type_vocab_size=1, pad_token_id=1, layer_norm_eps=1e-5)
        hf = RobertaModel(cfg, add_pooling_layer=False).eval().cuda()
        hf.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    rng = np.random.default_rng(0)
    for B, L, ragged in ((1, 128, False), (16, 128, False), (64, 128, True), (256, 128, True), (64, 512, False), (64, 512, True), (128, 512, True)):
        ids = rng.integers(3, VOCAB, size=(B, L)).astype(np.int32)
        if ragged:
            for b in range(1, B):
                ids[b, int(rng.integers(L // 4, L + 1)):] = 1
        n_tok = int((ids != 1).sum())
        for _ in range(3):
            out = enc.embed_ids(ids)
        dev_ms, wall = [], []
        for _ in range(8):
            t0 = time.perf_counter()
            out = enc.embed_ids(ids)
            wall.append((time.perf_counter() - t0) * 1e3)
            dev_ms.append(enc.last_ms)
        ms = statistics.median(dev_ms)
        flops = B * L * LAYERS * 2 * (4 * H * H + 2 * H * INTER) + LAYERS * 4 * B * L * L * H
        line = {"what": "encoder forward", "B": B, "L": L, "ragged": ragged, "non_pad_tokens": n_tok, "device_ms": ms, "wall_ms_host_ids_in_vectors_out": statistics.median(wall),
                "chunks_per_s": B / (ms * 1e-3), "tokens_per_s": B * L / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12,
                "frac_of_bf16_burst_peak": flops / (ms * 1e-3) / 1e12 / tpeak}
        if hf is not None:
            import torch
            t_ids = torch.from_numpy(ids.astype(np.int64)).cuda()
            mask = t_ids.ne(1)
            for dt_name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
                m = hf.to(dt)
                with torch.no_grad():
                    def fwd():
                        tok = m(t_ids, attention_mask=mask)[0]
                        return (tok * mask.unsqueeze(-1)).sum(1) / mask.sum(-1).unsqueeze(-1)
                    for _ in range(2):
                        ref = fwd()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(4):
                        ref = fwd()
                    e1.record()
                    torch.cuda.synchronize()
                    line[f"transformers_{dt_name}_ms"] = e0.elapsed_time(e1) / 4
                if dt_name == "fp32":
                    r = ref.float().cpu().numpy()
                    cos = (out * r).sum(1) / (np.linalg.norm(out, axis=1) * np.linalg.norm(r, axis=1))
                    line["min_cosine_vs_transformers_fp32"] = float(cos.min())
            hf.to(torch.float32)
        print(json.dumps(line), flush=True)
    enc.close()


if __name__ == "__main__":
    main()
