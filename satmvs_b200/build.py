"""In-tree build of libsatmvs_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m satmvs_b200.build` (or `__graft_entry__.build()`) runs nvcc directly; the .so lands next
to this file so it travels to the GPU box with the repo snapshot.  No torch headers are involved.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
HEADER = os.path.join(HERE, "..", "include", "satmvs_b200.h")
LIB = os.path.join(HERE, "libsatmvs_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-lineinfo", "-O3", "-std=c++17", "--compiler-options", "-fPIC"]


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER]
    return max(os.path.getmtime(h) for h in hs)


def build_variant(name: str, defines: list[str]) -> str:
    """Experiment helper: a separately named library compiled with extra -D flags (selected at run
    time with SATMVS_B200_LIB=<path>); never used by the default product path."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    out = os.path.join(HERE, f"libsatmvs_b200_{name}.so")
    subprocess.check_call([nvcc, *ARCH, *CFLAGS, *[f"-D{d}" for d in defines], "-shared", "-o", out, *sources()])
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into one shared library; returns its path."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.isfile(LIB):
            return LIB
        raise RuntimeError("nvcc not found and libsatmvs_b200.so is not prebuilt")
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_t = _headers_mtime()
    objs, jobs = [], []
    for src in sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        fresh = os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t)
        if force or not fresh:
            cmd = [nvcc, *ARCH, *CFLAGS, *(["-Xptxas=-v"] if verbose else []), "-c", src, "-o", obj]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in jobs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out, file=sys.stderr)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    if jobs or not os.path.isfile(LIB):
        subprocess.check_call([nvcc, *ARCH, "-shared", "-o", LIB, *objs])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
