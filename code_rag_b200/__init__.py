"""lattice-b200: B200-native exact vector search + hybrid ranking behind lattice's QdrantManager surface.

Import is light (no CUDA work, no library load); the first object that needs the GPU loads
``lib/liblattice_b200.so`` and raises ``NativeLibraryError`` if it, or an sm_100 device, is missing.
"""
from .errors import NativeLibraryError, VectorStoreError  # noqa: F401

__all__ = ["B200VectorStore", "ShardedB200VectorStore", "ShardPlane", "CollectionName", "DeviceCollection", "NativeLibraryError",
           "VectorStoreError"]
__version__ = "0.1.0"


def __getattr__(name):  # lazy: keep `import code_rag_b200` free of numpy/ctypes side effects
    if name in ("B200VectorStore", "CollectionName", "QdrantManager"):
        from . import client
        return getattr(client, name)
    if name in ("ShardedB200VectorStore", "ShardPlane"):
        from . import sharded_store
        return getattr(sharded_store, name)
    if name in ("DeviceCollection", "SearchResult"):
        from . import collection
        return getattr(collection, name)
    raise AttributeError(name)
