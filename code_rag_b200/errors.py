"""Error types of the adapter.

When lattice itself is importable the reference's own classes are re-used (reference
``src/lattice/core/errors.py:1-9,37-38``) so that ``except VectorStoreError`` in the reference's callers
(``query/vector_search.py:111-116``, ``embeddings/indexer.py:89-94``) catches what this backend raises.
"""
from __future__ import annotations

try:  # pragma: no cover - only when the reference package is installed next to us
    from lattice.core.errors import CodeRAGError, VectorStoreError  # type: ignore
except Exception:  # noqa: BLE001 - any import problem means "not installed here"

    class CodeRAGError(Exception):
        """Same shape as the reference's base error: message + optional ``cause`` (core/errors.py:1-9)."""

        def __init__(self, message: str, cause: Exception | None = None):
            super().__init__(message)
            self.cause = cause

        def __str__(self) -> str:
            if self.cause:
                return f"{self.args[0]} (caused by: {self.cause})"
            return str(self.args[0])

    class VectorStoreError(CodeRAGError):
        pass


class NativeLibraryError(RuntimeError):
    """liblattice_b200.so is missing, does not load, or reports a CUDA problem.  There is no CPU fallback."""
