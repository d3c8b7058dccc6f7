"""The C-ABI library loads without a GPU and exports every symbol include/lvs.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "lvs.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lvs_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    syms = _declared_symbols()
    for must in ("lvs_init", "lvs_collection_create", "lvs_upsert", "lvs_search", "lvs_delete_where", "lvs_match_rows",
                 "lvs_merge_topk_device", "lvs_search_submit", "lvs_search_wait", "lvs_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(native_lib):
    for name in _declared_symbols():
        assert hasattr(native_lib, name), f"liblattice_b200.so does not export {name}"


def test_ctypes_table_matches_header(native_lib):
    from code_rag_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared_symbols()


def test_constants_match_header():
    from code_rag_b200 import _native
    text = (ROOT / "include" / "lvs.h").read_text()
    defs = dict(re.findall(r"#define\s+(LVS_[A-Z0-9_]+)\s+\(?(-?0x[0-9A-Fa-f]+|-?\d+)u?\)?", text))
    val = lambda k: int(defs[k], 0)
    assert val("LVS_MAX_K") == _native.MAX_K
    assert val("LVS_ANY") == _native.ANY and val("LVS_NO_MATCH") == _native.NO_MATCH
    assert val("LVS_MAX_FILTER_COLS") == _native.MAX_FILTER_COLS
    assert (val("LVS_STORAGE_BF16"), val("LVS_METRIC_DOT"), val("LVS_DT_F64")) == (_native.STORAGE_BF16, _native.METRIC_DOT, _native.DT_F64)
    assert (val("LVS_EINVAL"), val("LVS_ESTATE")) == (_native.EINVAL, _native.ESTATE)


def test_no_gpu_fails_loudly(native_lib):
    """Without a CUDA device the product path must raise, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from code_rag_b200 import _native
    from code_rag_b200.errors import NativeLibraryError
    with pytest.raises(NativeLibraryError):
        _native.check(native_lib.lvs_init(0), "lvs_init")
    h = ctypes.c_void_p()
    rc = native_lib.lvs_collection_create(b"x", 8, 0, 0, 0, 0, 0, ctypes.byref(h))
    assert rc == _native.ESTATE and h.value is None


def test_product_code_never_imports_the_oracle():
    for path in (ROOT / "code_rag_b200").rglob("*.py"):
        src = path.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f"{path} mentions the oracle"


def test_snapshot_load_rejects_foreign_and_corrupt_files(native_lib, tmp_path):
    """Header checks run before any allocation (and before any CUDA call, so this runs without a GPU)."""
    import struct
    from code_rag_b200 import _native
    h = ctypes.c_void_p()
    bad = tmp_path / "bad.lvs"
    bad.write_bytes(b"not a snapshot" * 20)
    assert native_lib.lvs_snapshot_load(str(bad).encode(), b"x", 0, ctypes.byref(h)) == _native.EINVAL and h.value is None
    assert b"not a lattice-b200 snapshot" in native_lib.lvs_last_error()
    # right magic, absurd row count / truncated body
    hdr = b"LVSSNAP2" + struct.pack("<4i2q2I2f3q4q", 768, 1, 0, 2, 1 << 40, 0, 0, 1536, 1.0, 0.0, 0, 0, 0, 0, 0, 0, 0)
    bad.write_bytes(hdr)
    assert native_lib.lvs_snapshot_load(str(bad).encode(), b"x", 0, ctypes.byref(h)) == _native.EINVAL
    hdr = b"LVSSNAP2" + struct.pack("<4i2q2I2f3q4q", 768, 1, 0, 2, 1000, 0, 0, 1536, 1.0, 0.0, 0, 0, 0, 0, 0, 0, 0)
    bad.write_bytes(hdr + b"\0" * 100)
    assert native_lib.lvs_snapshot_load(str(bad).encode(), b"x", 0, ctypes.byref(h)) == _native.EINVAL
    assert b"size does not match" in native_lib.lvs_last_error()
    assert native_lib.lvs_snapshot_load(str(tmp_path / "missing.lvs").encode(), b"x", 0, ctypes.byref(h)) == _native.EINVAL


def test_one_process_binds_one_device(native_lib, monkeypatch):
    """One process per GPU: a second init() with another device must fail loudly instead of orphaning the first device's shards."""
    import pytest
    from code_rag_b200 import _native
    from code_rag_b200.errors import NativeLibraryError
    monkeypatch.setattr(_native, "_initialised_device", 0)
    _native.init(0)                                        # idempotent
    with pytest.raises(NativeLibraryError):
        _native.init(1)


def test_search_result_block_decodes_both_ways():
    """The one block lvs_search fills (scores | rows | ties | counts | flags) read through numpy views and, for one query, through
    SearchResult.single() (one struct.unpack, what the adapter's small-collection path uses)."""
    import ctypes as C

    import numpy as np

    from code_rag_b200.collection import SearchResult, _result_block
    for Q, k in ((1, 10), (3, 4), (1, 1)):
        block, ptrs = _result_block(Q, k)
        n = Q * k
        assert ptrs[0] == C.addressof(block) and [p - ptrs[0] for p in ptrs] == [0, 8 * n, 16 * n, 24 * n, 24 * n + 4 * Q]
        a = np.frombuffer(block, dtype=np.int64)
        rng = np.random.default_rng(Q * 10 + k)
        scores = rng.standard_normal(n)
        rows = rng.integers(-1, 1 << 40, n)
        ties = rng.integers(0, 1 << 63, n, dtype=np.uint64)
        a[:n] = scores.view(np.int64); a[n:2 * n] = rows; a[2 * n:3 * n] = ties.view(np.int64)
        tail = a[3 * n:].view(np.uint32)
        tail[:Q] = np.arange(1, Q + 1); tail[Q:2 * Q] = np.arange(Q) % 2
        res = SearchResult.from_block(block, Q, k)
        if Q == 1:
            cnt, flag, r1, s1 = res.single()
            assert (cnt, flag) == (1, 0) and r1 == rows.tolist() and s1 == scores.tolist()
        assert np.array_equal(res.scores, scores.reshape(Q, k)) and np.array_equal(res.rows, rows.reshape(Q, k))
        assert np.array_equal(res.ties, ties.reshape(Q, k)) and res.counts.tolist() == list(range(1, Q + 1))
        assert res.flags.tolist() == [q % 2 for q in range(Q)] and res.flags.dtype == np.int32 and res.counts.dtype == np.uint32
    # the five-array form still works and single() agrees with it
    res = SearchResult(np.array([[0.5, 0.25]]), np.array([[7, -1]]), np.zeros((1, 2), dtype=np.uint64), np.array([1], dtype=np.uint32),
                       np.array([1], dtype=np.int32))
    assert res.single() == (1, 1, [7, -1], [0.5, 0.25])
