"""Builds ``libhfnet_b200.so`` (the C-ABI library of include/hfnet_b200.h) in-tree with nvcc for sm_100a.

The built library is git-ignored but travels to the GPU box with the repository snapshot.  ``nvcc`` cross-compiles
without a GPU, so this runs in the CPU-only development container too.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "libhfnet_b200.so"
OBJ_DIR = ROOT / "build"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu"))


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link libhfnet_b200.so.  Incremental: objects are keyed by a content hash of
    the source plus every header."""
    srcs = sources()
    headers = sorted(CSRC.glob("*.cuh")) + sorted((ROOT.parent / "include").glob("*.h"))
    OBJ_DIR.mkdir(exist_ok=True)
    hdr_digest = _digest(headers)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in srcs:
        key = hashlib.sha256((hdr_digest + _digest([s])).encode()).hexdigest()[:16]
        obj = OBJ_DIR / f"{s.stem}.{key}.o"
        objs.append(obj)
        if force or not obj.exists():
            for old in OBJ_DIR.glob(f"{s.stem}.*.o"):
                old.unlink()
            jobs.append((s, obj))

    def compile_one(job):
        s, obj = job
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", str(s), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    stamp = OBJ_DIR / "link.stamp"
    link_key = hashlib.sha256(" ".join(o.name for o in objs).encode()).hexdigest()
    if force or jobs or not LIB.exists() or not stamp.exists() or stamp.read_text() != link_key:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
               "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(link_key)
    return LIB


if __name__ == "__main__":
    p = build_native(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
