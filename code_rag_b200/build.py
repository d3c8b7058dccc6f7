"""Builds liblattice_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with the repo).

Every ``csrc/*.cu`` is one translation unit; they are compiled in parallel (the 36 instantiations of the fused scan kernel
dominate the build and are spread over three units) and linked into one shared library."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
OBJ_DIR = LIB_DIR / "obj"
LIB = LIB_DIR / "liblattice_b200.so"
STAMP = LIB_DIR / "liblattice_b200.stamp"
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _digest() -> str:
    h = hashlib.sha256(_headers_digest().encode())
    for p in _sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def find_nvcc() -> str | None:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source and link one shared library; units whose inputs are unchanged are not recompiled."""
    LIB_DIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build liblattice_b200.so")
    OBJ_DIR.mkdir(exist_ok=True)
    hd = _headers_digest()

    def compile_one(src: Path):
        obj, tag = OBJ_DIR / (src.stem + ".o"), OBJ_DIR / (src.stem + ".stamp")
        want = hashlib.sha256((hd + src.name).encode() + src.read_bytes()).hexdigest()
        if not force and obj.exists() and tag.exists() and tag.read_text().strip() == want:
            return src, 0, ""
        cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", str(src), "-o", str(obj)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode == 0:
            tag.write_text(want)
        return src, proc.returncode, proc.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _sources()))
    for src, rc, err in results:
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {src.name} ({rc}):\n{err[-4000:]}")
        if verbose and err:
            sys.stderr.write(err)
    objs = [str(OBJ_DIR / (s.stem + ".o")) for s in _sources()]
    proc = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, "-o", str(LIB)], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed ({proc.returncode}):\n{proc.stderr[-4000:]}")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
