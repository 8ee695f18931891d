"""Build libflair_b200.so (sm_100a) in-tree with nvcc.

One shared library, plain C ABI (include/flair_b200.h).  Objects are cached by
source mtime under flair_b200/csrc/_obj so a rebuild after a one-file change
takes seconds.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = CSRC / "_obj"
LIB = ROOT / "libflair_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    *(["-DFLAIR_CONV_TRACE_BUILD"] if os.environ.get("FLAIR_BUILD_TRACE") == "1" else []),
    "-Xptxas", "-v",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "flair_b200.h"]
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src, *headers]):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[flair_b200.build] compiled {name}", file=sys.stderr)
    objs = [OBJ / (s.stem + ".o") for s in sources]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[flair_b200.build] linked {LIB}", file=sys.stderr)
    return LIB


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
