"""Drop-in check at the seam: the reference's OWN VectorIndexer / VectorSearcher / ContextBuilder / ProjectCleanupService /
QueryEngine (vector leg) / HybridRanker classes drive B200VectorStore, and the reference's own tests/test_database.py tests of
QdrantManager run unmodified against it.
Needs /root/reference (present in the build container, absent on the GPU box -> skipped there)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not Path("/root/reference/src/lattice").exists(), reason="reference sources not present")
def test_reference_indexer_and_searchers_run_on_the_adapter():
    out = subprocess.run([sys.executable, str(ROOT / "tests" / "ref_callers_script.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "reference callers OK" in out.stdout
    assert out.stdout.count("TestQdrantConnection::test_") >= 3, out.stdout[-2000:]
