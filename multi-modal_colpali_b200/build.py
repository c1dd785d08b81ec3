"""In-tree build of the C-ABI library (``_lib/liblis.so``) with nvcc for sm_100a.

No torch dependency: the library is plain CUDA runtime + inline PTX.  ``nvcc`` cross-compiles on a
machine without a GPU, so this runs in CI containers as the "does it build" check.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "_lib"
LIB = LIBDIR / "liblis.so"
SOURCES = ["lis_maxsim.cu", "lis_maxsim_pair.cu", "lis_topk.cu", "lis_project.cu", "lis_index.cu", "lis_comm.cu",
           "lis_stream.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-DLIS_BUILD",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA toolkit is required to build liblis.so")
    return exe


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(PKG.parent / "include" / "lis.h")
    return any(d.stat().st_mtime > t for d in deps)


def build_variant(name: str, defines) -> Path:
    """Experiment builds: ``_lib/liblis_<name>.so`` with extra -D flags (select with LIS_LIB=...)."""
    nvcc = _nvcc()
    objdir = LIBDIR / f"obj_{name}"
    objdir.mkdir(parents=True, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = objdir / src.replace(".cu", ".o")
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr}")
        objs.append(obj)
    out = LIBDIR / f"liblis_{name}.so"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *map(str, objs), "-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link ``liblis.so``.  Returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: str) -> Path:
        obj = objdir / (src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIBDIR / "liblis.so.tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *map(str, objs), "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
