"""The embedding oracle (oracle/roberta_encoder.py) against the fixture made by transformers' own
RobertaModel (tests/golden/make_roberta_golden.py): the class the reference's UniXcoder wraps."""
from pathlib import Path

import numpy as np

from oracle import roberta_encoder as R

G = np.load(Path(__file__).parent / "golden" / "roberta_encoder_golden.npz")


def golden_state_dict():
    return {k[2:]: G[k] for k in G.files if k.startswith("w:")}


def test_oracle_matches_transformers_roberta():
    vocab, hidden, layers, heads, inter, max_pos, pad = (int(v) for v in G["config"])
    tok, sent = R.encode(golden_state_dict(), G["ids"], n_layers=layers, n_heads=heads, pad_id=pad)
    mask = G["ids"] != pad
    assert np.abs(tok[mask] - G["token_embeddings"][mask]).max() < 2e-5          # float32 both sides, different summation order
    assert np.abs(sent - G["sentence_embeddings"]).max() < 1e-5
    # float64 restatement agrees to float32 resolution: the fixture is not sitting on a cancellation
    _, sent64 = R.encode(golden_state_dict(), G["ids"], n_layers=layers, n_heads=heads, pad_id=pad, dtype=np.float64)
    assert np.abs(sent64 - G["sentence_embeddings"]).max() < 1e-5


def test_padding_does_not_leak():
    """Rows are independent and pad positions invisible: re-padding a batch to another length changes nothing."""
    vocab, hidden, layers, heads, inter, max_pos, pad = (int(v) for v in G["config"])
    sd = golden_state_dict()
    ids = G["ids"][1:4]
    longer = np.full((3, 60), pad, dtype=ids.dtype)
    longer[:, :ids.shape[1]] = ids
    _, a = R.encode(sd, ids, n_layers=layers, n_heads=heads, pad_id=pad)
    _, b = R.encode(sd, longer, n_layers=layers, n_heads=heads, pad_id=pad)
    assert np.abs(a - b).max() < 1e-6
    _, c = R.encode(sd, ids[1:2], n_layers=layers, n_heads=heads, pad_id=pad)
    assert np.abs(a[1:2] - c).max() < 1e-6
