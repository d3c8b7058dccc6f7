"""ctypes binding of the C ABI in ``include/lvs.h`` (liblattice_b200.so).  Fails loudly: no fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

from .errors import NativeLibraryError

LIB_PATH = Path(__file__).resolve().parent / "lib" / "liblattice_b200.so"

OK, EINVAL, ECUDA, ENOMEM, ESTATE, ELIMIT, ENAN = 0, -1, -2, -3, -4, -5, -6
STORAGE_F32, STORAGE_BF16 = 0, 1
METRIC_COSINE, METRIC_DOT = 0, 1
DT_F32, DT_F64, DT_BF16 = 0, 1, 2
MAX_FILTER_COLS = 8
ANY = 0xFFFFFFFF
NULL_CODE = 0
NO_MATCH = 0xFFFFFFFE
MAX_K = 224
FLAG_UNPROVEN = 1
FLAG_EXCHANGE = 2

_vp = C.c_void_p


class EncoderConfig(C.Structure):
    """``lvs_encoder_config`` (include/lvs.h)."""
    _fields_ = [("vocab", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32),
                ("intermediate", C.c_int32), ("max_pos", C.c_int32), ("pad_id", C.c_int32), ("ln_eps", C.c_float)]

_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_f32p = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/lvs.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES: dict[str, tuple] = {
    "lvs_version": (C.c_int, []),
    "lvs_last_error": (C.c_char_p, []),
    "lvs_init": (C.c_int, [C.c_int]),
    "lvs_shutdown": (C.c_int, []),
    "lvs_device_info": (C.c_int, [_ip, _ip, _ip, _i64p, _i64p]),
    "lvs_collection_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.POINTER(_vp)]),
    "lvs_collection_destroy": (C.c_int, [_vp]),
    "lvs_collection_reserve": (C.c_int, [_vp, C.c_int64]),
    "lvs_rows": (C.c_int64, [_vp]),
    "lvs_count": (C.c_int64, [_vp]),
    "lvs_capacity": (C.c_int64, [_vp]),
    "lvs_search_counter": (C.c_uint64, [_vp]),
    "lvs_advance_search_counter": (C.c_int, [_vp, C.c_uint64]),
    "lvs_upsert": (C.c_int, [_vp, _vp, C.c_int, C.c_int64, _vp, _vp, _vp]),
    "lvs_upsert_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp]),
    "lvs_set_codes": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, C.c_int64, _vp]),
    "lvs_delete_rows": (C.c_int, [_vp, _vp, C.c_int64, _i64p]),
    "lvs_delete_where": (C.c_int, [_vp, _vp, _vp, C.c_int64, _i64p]),
    "lvs_search": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lvs_search_submit": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _ip]),
    "lvs_search_submit_sharded": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _ip]),
    "lvs_search_wait": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lvs_search_poll": (C.c_int, [_vp, C.c_int, _ip]),
    "lvs_search_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lvs_search_device_at": (C.c_int, [_vp, C.c_uint64, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lvs_search_device_async": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lvs_last_kernel_phases": (C.c_int, [_vp, _u64p]),
    "lvs_scan_times": (C.c_int, [_vp, C.c_int, _f32p, _f64p, _ip]),
    "lvs_match_rows": (C.c_int, [_vp, _vp, _vp, C.c_int64, _i64p]),
    "lvs_merge_topk_device": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lvs_encoder_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "lvs_encoder_load": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int64]),
    "lvs_encoder_embed": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "lvs_encoder_embed_upsert": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "lvs_encoder_last_ms": (C.c_int, [_vp, _f32p]),
    "lvs_encoder_destroy": (C.c_int, [_vp]),
    "lvs_exchange_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), _vp]),
    "lvs_exchange_connect": (C.c_int, [_vp, _vp]),
    "lvs_exchange_merge_device": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "lvs_search_sharded_device_async": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "lvs_exchange_error": (C.c_int, [_vp]),
    "lvs_exchange_destroy": (C.c_int, [_vp]),
    "lvs_rank_fuse": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32p]),
    "lvs_rank_names_append": (C.c_int, [_vp, _vp, _vp, C.c_int, _u32p]),
    "lvs_rank_attrs_set": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lvs_search_rank": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, C.c_double,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32p]),
    "lvs_move_rows": (C.c_int, [_vp, _vp, _vp, C.c_int64]),
    "lvs_truncate": (C.c_int, [_vp, C.c_int64]),
    "lvs_snapshot_save": (C.c_int, [_vp, C.c_char_p]),
    "lvs_snapshot_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(_vp)]),
    "lvs_search_rank2": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int,
                                   C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32p]),
    "lvs_last_search_timing": (C.c_int, [_vp, _f32p, _ip, _ip]),
    "lvs_set_option": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "lvs_fetch_rows_f32": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
}

class RankBatch(C.Structure):
    """``lvs_rank_batch`` of include/lvs.h."""
    _fields_ = [("n_queries", C.c_int32), ("offsets", _vp), ("kind", _vp), ("key_id", _vp), ("file_id", _vp), ("depth", _vp),
                ("entity_match", _vp), ("degree", _vp), ("flags", _vp), ("content_len", _vp), ("vscore", _vp), ("weights", _vp)]


class RankHits(C.Structure):
    """``lvs_rank_hits`` of include/lvs.h (host output arrays of one collection's hits)."""
    _fields_ = [("scores", _vp), ("rows", _vp), ("counts", _vp), ("flags", _vp)]


class RankQueryCtx(C.Structure):
    """``lvs_rank_query_ctx`` of include/lvs.h."""
    _fields_ = [("ent_off", _vp), ("ent_str_off", _vp), ("ent_bytes", _vp), ("cen_off", _vp), ("cen_id", _vp), ("cen_deg", _vp)]


_lib: C.CDLL | None = None


def load() -> C.CDLL:
    """dlopen the in-tree library and type every entry point.  Raises NativeLibraryError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("LVS_LIBRARY", LIB_PATH))
    if not path.exists():
        raise NativeLibraryError(
            f"{path} not found. Build it with `python -m code_rag_b200.build` (nvcc, sm_100a). "
            "lattice-b200 has no CPU fallback."
        )
    try:
        lib = C.CDLL(str(path))
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f"cannot load {path}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{path} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return (load().lvs_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc == ENAN:
        raise ValueError(last_error())          # "Query vector must not contain NaN", as local mode says it
    if rc != OK:
        raise NativeLibraryError(f"{what} failed (code {rc}): {last_error()}")


_initialised_device: int | None = None


def init(device: int = 0) -> None:
    """Bind this process to one GPU (idempotent for the same device)."""
    global _initialised_device
    if _initialised_device == device:
        return
    if _initialised_device is not None:
        # one process drives one GPU (one process per GPU, torchrun-style): collections created on the first device would
        # be orphaned by a re-bind
        raise NativeLibraryError(f"this process is already bound to cuda:{_initialised_device}; cannot re-bind it to cuda:{device}")
    check(load().lvs_init(int(device)), "lvs_init")
    _initialised_device = device


def is_initialised() -> bool:
    return _initialised_device is not None
