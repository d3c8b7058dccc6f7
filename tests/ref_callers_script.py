"""Runs the REFERENCE's own callers of the vector store (query/vector_search.py, embeddings/indexer.py, query/context/builder.py,
projects/cleanup.py, QueryEngine._execute_vector_search, and query/ranking fed with the adapter's hits) against
B200VectorStore.  Executed in a fresh interpreter by tests/test_reference_callers_cpu.py (build container only: it
needs /root/reference).  `qdrant_client` is not installed, so a stub module satisfies the import of
lattice/embeddings/client.py; the reference classes under test are the unmodified files."""
import asyncio
import hashlib
import sys
import types
from pathlib import Path
from types import SimpleNamespace as NS

ROOT = Path(__file__).resolve().parent.parent
SRC = "/root/reference/src"
sys.path[:0] = [SRC, str(ROOT), str(ROOT / "tests")]


def ns(name, path=None):
    m = types.ModuleType(name)
    if path:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


# namespace packages: skip the heavy __init__ chains (tree_sitter, neo4j, qdrant_client ...)
for pkg in ("lattice", "lattice.embeddings", "lattice.query", "lattice.parsing", "lattice.graph"):
    ns(pkg, SRC + "/" + pkg.replace(".", "/"))
qc = ns("qdrant_client")
qc.AsyncQdrantClient = object
qc.models = ns("qdrant_client.models")
class _Model:
    """Keyword-constructed record, as the pydantic models of qdrant_client.models are used by projects/cleanup.py:47-73."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


for _n in ("Filter", "CollectionInfo", "FieldCondition", "MatchValue", "MatchText", "PointStruct", "VectorParams", "Distance",
           "PayloadSchemaType", "FilterSelector"):
    setattr(qc.models, _n, type(_n, (_Model,), {}))
ns("lattice.projects", SRC + "/lattice/projects")
_neo = ns("neo4j")                                     # graph/client.py imports it; nothing on this path talks to a graph
_neo.AsyncGraphDatabase = _neo.AsyncDriver = _neo.AsyncSession = object
_ne = ns("neo4j.exceptions")
_ne.ServiceUnavailable = _ne.Neo4jError = _ne.AuthError = Exception

from lattice.core.errors import IndexingError, QueryError, VectorStoreError  # noqa: E402  (reference classes)
from lattice.embeddings.chunker import CodeChunk  # noqa: E402
from lattice.embeddings.indexer import VectorIndexer, VectorSearcher as LegacySearcher  # noqa: E402
from lattice.query.vector_search import VectorSearcher  # noqa: E402

import numpy as np  # noqa: E402

from code_rag_b200 import errors as our_errors  # noqa: E402
from code_rag_b200.client import B200VectorStore  # noqa: E402
from helpers import FakeDevice  # noqa: E402

assert our_errors.VectorStoreError is VectorStoreError, "the adapter must raise the reference's own exception class"
DIM = 48


def vec(text: str) -> list[float]:
    seed = int.from_bytes(hashlib.sha256(text.encode()).digest()[:8], "little")
    return np.random.default_rng(seed).standard_normal(DIM).tolist()


class Embedder:
    async def embed(self, text):
        return vec(text)

    async def embed_with_progress(self, texts, progress_callback=None):
        return [vec(t) for t in texts]


class Chunker:
    def chunk_file(self, parsed_file, project_name=None):
        fp = str(parsed_file.file_info.path)
        return [CodeChunk(content=f"{fp} chunk {i}", file_path=fp, entity_type="function", entity_name=f"fn{i}", language="python",
                          start_line=10 * i + 1, end_line=10 * i + 9, graph_node_id=f"m.fn{i}",
                          content_hash=parsed_file.file_info.content_hash, project_name=project_name) for i in range(parsed_file.n)]


async def main(use_gpu: bool):
    store = B200VectorStore(dimensions=DIM, _device_factory=None if use_gpu else FakeDevice)
    await store.connect()
    await store.create_collections()
    indexer = VectorIndexer(qdrant=store, embedder=Embedder(), chunker=Chunker())
    files = [NS(file_info=NS(path=Path(f"/repo/pkg/f{j}.py"), content_hash=f"h{j}"), n=4 + j) for j in range(5)]
    total = await indexer.index_files(files, project_name="demo")
    assert total == sum(f.n for f in files) == 30
    assert (await store.get_collection_info("code_chunks")).points_count == 30
    assert await indexer.index_file(files[2]) == 0                        # unchanged hash -> skipped (indexer.py:57-59)
    files[2].file_info.content_hash = "h2-new"
    files[2].n = 3
    assert await indexer.index_file(files[2], project_name="demo") == 3  # delete by file_path, then upsert
    assert (await store.get_collection_info("code_chunks")).points_count == 30 - 6 + 3
    await indexer.index_summary("/repo/pkg/f0.py", "function", "fn0", "summary of fn0", "m.fn0")

    searcher = VectorSearcher(store, Embedder())
    hits = await searcher.search_code("/repo/pkg/f1.py chunk 2", limit=5, language="python", project_name="demo")
    assert hits and set(hits[0]) == {"score", "file_path", "entity_type", "entity_name", "language", "content", "start_line", "end_line", "graph_node_id"}
    assert hits[0]["content"] == "/repo/pkg/f1.py chunk 2" and abs(hits[0]["score"] - 1.0) < 1e-6
    assert [h["score"] for h in hits] == sorted((h["score"] for h in hits), reverse=True)
    assert await searcher.search_code("x", limit=5, language="rust") == []
    sims = await searcher.find_similar_code("/repo/pkg/f1.py chunk 2", limit=3, exclude_file="/repo/pkg/f1.py")
    assert len(sims) == 3 and all(s["file_path"] != "/repo/pkg/f1.py" for s in sims)
    summ = await searcher.search_summaries("summary of fn0", limit=2)
    assert summ and summ[0]["summary"] == "summary of fn0"
    try:
        await searcher.search_code("   ")
        raise AssertionError("empty query must raise QueryError")
    except QueryError:
        pass

    legacy = LegacySearcher(store, Embedder())
    res = await legacy.search_code("/repo/pkg/f3.py chunk 0", limit=5, entity_type="function")
    assert res[0].content == "/repo/pkg/f3.py chunk 0" and res[0].start_line == 1

    # ContextBuilder._build_entity_context (query/context/builder.py:103-133): the filter-only lookup, query_vector=None
    from lattice.query.context.builder import ContextBuilder
    from lattice.query.graph_reasoning import GraphContext, GraphNode
    empty = GraphContext([], [], [], [], [], [], [], [], [], [], [], [])
    node = GraphNode(node_type="Function", name="fn1", qualified_name="m.fn1", file_path="/repo/pkg/f3.py", start_line=11, end_line=19)
    ctx = await ContextBuilder(memgraph=None, qdrant=store)._build_entity_context(node, empty)
    assert ctx.code_snippet is not None and ctx.code_snippet.content == "/repo/pkg/f3.py chunk 1" and ctx.code_snippet.language == "python"
    ghost = GraphNode(node_type="Function", name="nope", qualified_name="m.nope", file_path="/repo/pkg/f3.py")
    assert (await ContextBuilder(memgraph=None, qdrant=store)._build_entity_context(ghost, empty)).code_snippet is None

    # the reference's own HybridRanker over the adapter's hits (query/engine.py:176-181) == this repo's ranking mirror's host half
    # is covered elsewhere; here: the hit dicts are what ranker.py:150-169 reads
    from lattice.query.ranking import HybridRanker as RefRanker
    from lattice.query.query_planner import ExtractedEntity, QueryIntent, QueryPlan
    plan = QueryPlan(original_query="fn2", primary_intent=QueryIntent.FIND_SIMILAR, sub_queries=[],
                     entities=[ExtractedEntity(name="fn2", entity_type="function")], relationships=[])
    ranked = RefRanker().rank_results(plan, empty, hits, {})
    assert ranked and ranked[0].source == "vector" and {r.file_path for r in ranked} <= {h["file_path"] for h in hits}
    assert any(r.entity_name == "fn2" and r.signal_scores["query_entity_match"] == 1.0 for r in ranked)

    # QueryEngine._execute_vector_search (query/engine.py:315-346): code hits, then limit // 2 summaries hits for the
    # five intents it extends - the reference's own engine object over the adapter (no initialize(): only the vector leg runs)
    from lattice.query.engine import QueryEngine
    engine = QueryEngine(qdrant=store, vector_searcher=searcher)
    plan_sf = QueryPlan(original_query="summary of fn0", primary_intent=QueryIntent.SEARCH_FUNCTIONALITY, sub_queries=[], entities=[],
                        relationships=[])
    both = await engine._execute_vector_search("summary of fn0", plan_sf, limit=6, language="python")
    assert len(both) == 6 + 1 and all("content" in h for h in both[:6]) and both[6]["summary"] == "summary of fn0"
    only_code = await engine._execute_vector_search("summary of fn0", plan, limit=6, language="python")
    assert only_code == both[:6]
    ranked2 = engine._ranker.rank_results(plan_sf, empty, both, {})
    assert len(ranked2) == 7 and any(r.summary == "summary of fn0" for r in ranked2)

    # ProjectCleanupService (projects/cleanup.py:12-73): MatchText count + delete through manager.client, both collections
    from lattice.projects.cleanup import ProjectCleanupService
    cleanup = ProjectCleanupService(store)
    assert await cleanup.get_chunk_count("/repo/pkg/f4.py") == 8 and await cleanup.get_chunk_count("/repo/pkg/") == 27
    assert await cleanup.delete_from_qdrant("/repo/pkg/f4.py") == 0          # it returns what is LEFT (cleanup.py:58-63)
    assert (await store.get_collection_info("code_chunks")).points_count == 27 - 8
    assert await cleanup.delete_from_qdrant("/repo/pkg/") == 0
    assert (await store.get_collection_info("code_chunks")).points_count == 0
    assert (await store.get_collection_info("summaries")).points_count == 0
    assert await cleanup.get_chunk_count("/repo/pkg/") == 0

    await store.close()
    try:                                                                   # vector store failures surface as QueryError
        await searcher.search_code("anything")
        raise AssertionError("expected QueryError after close()")
    except QueryError as e:
        assert isinstance(e.cause, VectorStoreError)
    try:
        await indexer.index_file(files[0], force=True)
        raise AssertionError("expected IndexingError after close()")
    except IndexingError:
        pass
    await run_reference_database_tests(use_gpu)
    print("reference callers OK")


async def run_reference_database_tests(use_gpu: bool) -> None:
    """The reference's OWN tests of this seam, unmodified: every test of ``TestQdrantConnection`` in
    /root/reference/tests/test_database.py (connect + health check, create_collections seen through ``.client``, upsert -> search
    -> delete), with ``QdrantManager`` in the module under test bound to the adapter.  pytest-asyncio is not installed, so the
    coroutines are driven here."""
    import importlib.util
    import inspect

    class Manager(B200VectorStore):
        def __init__(self, *a, **kw):
            super().__init__(*a, _device_factory=None if use_gpu else FakeDevice, **kw)

    spec = importlib.util.spec_from_file_location("ref_test_database", "/root/reference/tests/test_database.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.QdrantManager = Manager                     # the name the tests construct; CollectionName stays the reference's enum
    suite = mod.TestQdrantConnection()
    names = [n for n, f in inspect.getmembers(suite, inspect.iscoroutinefunction) if n.startswith("test_")]
    assert set(names) >= {"test_connect_to_qdrant", "test_create_collections", "test_upsert_and_search_vectors"}, names
    for n in names:
        await getattr(suite, n)()
        print(f"reference tests/test_database.py::TestQdrantConnection::{n} passed on the adapter")


if __name__ == "__main__":
    asyncio.run(main(use_gpu="--gpu" in sys.argv))
