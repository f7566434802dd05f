"""Build recipe for libmagnify_b200.so (nvcc, sm_100a only, in-tree).

`python -m magnify_b200.build [--force] [-v]` or `__graft_entry__.build()`.  Every source is
compiled to an object file under magnify_b200/build/ (in parallel, only when it or a header is
newer than its object) and the objects are linked into the shared library.  nvcc cross-compiles
without a GPU.  The CUDA runtime is linked dynamically (`-cudart shared`): the process already
has torch's libcudart, and the shipped library then carries no runtime of its own.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libmagnify_b200.so")
SOURCES = ("flatfield_stitch.cu", "roi.cu", "roi_tma.cu", "roi_lists.cu", "masks.cu", "circles.cu",
           "circles_sample.cu", "circles_host.cpp", "tiff_pages.cpp")
NVCC_FLAGS = (
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-cudart", "shared",
    "-Xcompiler", "-fPIC",
)


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; magnify_b200 has no CPU fallback and cannot be built without it")


def _headers():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return files


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def _stale(src: str) -> bool:
    obj = _obj(src)
    if not os.path.exists(obj):
        return True
    built = os.path.getmtime(obj)
    return any(os.path.getmtime(f) > built for f in [os.path.join(CSRC, src)] + _headers())


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    inputs = [os.path.join(CSRC, s) for s in SOURCES] + _headers()
    return any(os.path.getmtime(f) > built for f in inputs)


def _run(cmd, verbose: bool) -> None:
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libmagnify_b200.so:\n" + proc.stderr[-4000:])


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [s for s in SOURCES if force or _stale(s)]

    def compile_one(src):
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", _obj(src)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        _run(cmd, verbose)

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as pool:
        list(pool.map(compile_one, todo))
    # zlib: Deflate-compressed TIFF pages
    _run([nvcc, *NVCC_FLAGS, "-shared", *[_obj(s) for s in SOURCES], "-lz", "-o", LIB_PATH], verbose)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
