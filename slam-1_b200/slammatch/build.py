"""In-tree build of libslammatch.so (hand-written CUDA for sm_100a + the C-ABI).

    python -m slammatch.build [--force]        (with slam-1_b200/ on sys.path)

nvcc cross-compiles without a GPU; the resulting .so lives next to the package (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  Objects are rebuilt when a source, a header or the
flag set changes.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # slam-1_b200/
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(PKG_DIR, "build")
LIB = os.path.join(PKG_DIR, "libslammatch.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libslammatch.so cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "slammatch.h"))
    for p in hdrs:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _compile(src: str, obj: str, verbose: bool) -> str:
    cmd = [nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build_rows(force: bool = False) -> str:
    """gcc build of slammatch/_rows.c (CPython extension that creates the DMatch result rows in C)."""
    import sysconfig
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "_rows.c")
    out = os.path.join(here, "_rows" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        cc = shutil.which("gcc") or shutil.which("cc")
        if cc is None:
            raise RuntimeError("gcc not found")
        cmd = [cc, "-O2", "-shared", "-fPIC", "-Wall", "-I", sysconfig.get_paths()["include"], src, "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed on _rows.c:\n{r.stdout}\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    build_rows(force)
    os.makedirs(BUILD, exist_ok=True)
    tag = _headers_digest()
    stamp = os.path.join(BUILD, "flags.stamp")
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if old != tag:
        force = True
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(src):
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(lambda a: _compile(a[0], a[1], verbose), jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(tag)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
