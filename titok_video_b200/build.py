"""Builds libtitok_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python titok_video_b200/build.py [--force]      (run as a script: importing the package needs the built library)

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIBDIR = os.path.join(ROOT, "lib")
OBJDIR = os.path.join(ROOT, "build")
LIB = os.path.join(LIBDIR, "libtitok_b200.so")

SOURCES = ["api.cu", "host_util.cu", "fsq.cu", "rowops.cu", "gemm.cu", "attn.cu", "vq.cu", "wgrad.cu", "attn_bwd.cu", "bwd_rows.cu", "sequence.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, tag: str = "", defines: tuple = ()) -> str:
    """tag / defines: experimental variants (lib/libtitok_b200_<tag>.so built with -D<define>, selected at run time
    with TTK_LIB_PATH); the product library is the untagged one."""
    global OBJDIR, LIB
    if tag:
        OBJDIR = os.path.join(ROOT, "build", "variant_" + tag)
        LIB = os.path.join(LIBDIR, f"libtitok_b200_{tag}.so")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        # the CUDA runtime is linked dynamically: PyTorch has already loaded libcudart in-process, and the shipped binary then
        # carries no copy of the runtime
        run([nvcc, "-shared", "-o", LIB, *objs, "-cudart", "shared"])
    return LIB


if __name__ == "__main__":
    _tag = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--tag=")), "")
    _defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    path = build(force="--force" in sys.argv, verbose=True, tag=_tag, defines=_defs)
    print(path)
