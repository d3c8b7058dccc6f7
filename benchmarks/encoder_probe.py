#!/usr/bin/env python
"""One encoder shape, a few forward passes (for ncu launch lists / captures): python benchmarks/encoder_probe.py [B] [L] [reps]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from code_rag_b200.embedding import B200CodeEncoder, random_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sd = random_state_dict(8000, 768, 12, 3072, 1026, seed=1)
enc = B200CodeEncoder(sd, n_layers=12, n_heads=12, pad_id=1)
ids = np.random.default_rng(0).integers(3, 8000, size=(B, L)).astype(np.int32)
for _ in range(reps):
    enc.embed_ids(ids)
print(f"B={B} L={L}: {enc.last_ms:.3f} ms per forward")
enc.close()
