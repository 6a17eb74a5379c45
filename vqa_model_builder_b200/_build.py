"""Build recipe for libb200vqa.so (nvcc, sm_100a only, in-tree).

The shared library is a plain C-ABI object (include/b200vqa.h): no torch headers, cudart linked
statically, the driver API reached through cudaGetDriverEntryPoint — so it loads (and its symbols can be
checked) on a machine without a GPU or libcuda.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libb200vqa.so"
OBJ_DIR = PKG_DIR / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(INCLUDE),
]
if os.environ.get("B200VQA_GEMM_TRACE") == "1":   # instrumented GEMM (scripts/gemm_trace.py): separate library + objects
    NVCC_FLAGS.append("-DB200_GEMM_TRACE")
    LIB_PATH = PKG_DIR / "libb200vqa_trace.so"
    OBJ_DIR = PKG_DIR / "build_trace"
if os.environ.get("B200VQA_EPI_WARPS"):     # experiment knob: epilogue warps of the tcgen05 GEMM (8 | 12 | 16)
    NVCC_FLAGS.append("-DB200_EPI_WARPS=" + os.environ["B200VQA_EPI_WARPS"])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "b200vqa.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    # flags without the include path: the digest must not depend on where the tree is checked out (the GPU box
    # mounts the repo elsewhere; a path-dependent digest made every fresh box rebuild the prebuilt library)
    h.update(" ".join(f for f in NVCC_FLAGS if f != str(INCLUDE)).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    stamp = OBJ_DIR / "digest.txt"
    return LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libb200vqa.so next to this file."""
    if not force and is_fresh():
        return LIB_PATH
    OBJ_DIR.mkdir(exist_ok=True)
    # one builder at a time (torchrun ranks import concurrently); late comers find a fresh library
    import fcntl
    with open(OBJ_DIR / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_fresh():
            return LIB_PATH
        return _build_locked(verbose)


def _source_digest(src: Path) -> str:
    """digest of one translation unit: its own text, every header of csrc/ and the public header, the flags"""
    h = hashlib.sha256()
    for p in [src] + sorted(CSRC.glob("*.cuh")) + [INCLUDE / "b200vqa.h"]:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(f for f in NVCC_FLAGS if f != str(INCLUDE)).encode())
    return h.hexdigest()


def _build_locked(verbose: bool) -> Path:
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        stamp = OBJ_DIR / (src.stem + ".digest")
        dig = _source_digest(src)
        if not verbose and obj.exists() and stamp.exists() and stamp.read_text().strip() == dig:
            return obj                 # unchanged translation unit (router.cu alone takes minutes to compile)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        stamp.write_text(dig)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
           *[str(o) for o in objs], "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)       # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    (OBJ_DIR / "digest.txt").write_text(_digest())
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
