"""Seeded synthetic corpora shaped like BASELINE.json's configs."""
from __future__ import annotations

import uuid

import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16 -> float32 (what the device stores for bf16 collections)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def unixcoder_like(n: int, dim: int, seed: int, n_queries: int = 0):
    """C1: un-normalised mean-pooled hidden states: x = mu + z, mu ~ 0.5 N(0, I) fixed per corpus."""
    rng = np.random.default_rng(seed)
    mu = 0.5 * rng.standard_normal(dim)
    x = (mu + rng.standard_normal((n, dim))).astype(np.float32)
    q = (mu + rng.standard_normal((n_queries, dim))).astype(np.float32)
    return x, q


def unit_rows(n: int, dim: int, seed: int, n_queries: int = 0, dtype=np.float32):
    """C2/C3: unit-norm rows (text-embedding-3-small shape)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((n_queries, dim)).astype(np.float32)
    if n_queries:
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    return x.astype(dtype), q.astype(dtype)


PROJECTS = [f"proj{i}" for i in range(8)]
PROJECT_P = [.40, .20, .15, .10, .05, .04, .03, .03]
LANGS = ["python", "typescript", "javascript"]
LANG_P = [.6, .25, .15]
ETYPES = ["function", "method", "class"]
ETYPE_P = [.5, .35, .15]


def payloads(n: int, seed: int, n_files: int | None = None) -> list[dict]:
    """C2 payload columns: the 10 keys of CodeChunk.to_payload (reference embeddings/chunker.py:25-37)."""
    rng = np.random.default_rng(seed)
    n_files = n_files or max(1, n // 20)
    proj = rng.choice(len(PROJECTS), size=n, p=PROJECT_P)
    lang = rng.choice(len(LANGS), size=n, p=LANG_P)
    et = rng.choice(len(ETYPES), size=n, p=ETYPE_P)
    fidx = rng.integers(0, n_files, size=n)
    out = []
    for i in range(n):
        out.append({
            "file_path": f"src/pkg{fidx[i] % 37}/file_{fidx[i]}.py",
            "entity_type": ETYPES[et[i]],
            "entity_name": f"entity_{i}",
            "language": LANGS[lang[i]],
            "start_line": int(i % 500) + 1,
            "end_line": int(i % 500) + 20,
            "content": "x" * int(rng.integers(10, 400)),
            "graph_node_id": f"pkg.entity_{i}" if i % 3 else None,
            "content_hash": f"hash{fidx[i]}",
            "project_name": PROJECTS[proj[i]],
        })
    return out


def uuid_for_row(row: int) -> str:
    """Ids monotone in the row number: str(UUID(int=row)) sorts like the integer."""
    return str(uuid.UUID(int=row))


def random_uuids(n: int, seed: int) -> list[str]:
    rng = np.random.default_rng(seed)
    return [str(uuid.UUID(bytes=rng.bytes(16), version=4)) for _ in range(n)]
