"""Builds grafx_b200/lib/libgrafx_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree).

The .so is a plain C-ABI shared library (include/grafx_b200.h); it links only the CUDA runtime
(statically) -- no torch, no Python.  `python -m grafx_b200.build` rebuilds it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, os.environ.get("GFX_LIB_OUT", "libgrafx_b200.so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "-cudart", "static",
]


# coefficient design must round like the separate torch ops it replaces: no FMA contraction there
PER_FILE_FLAGS = {"design.cu": ["-fmad=false"]}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libgrafx_b200.so")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps() -> list[str]:
    deps = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    deps.append(os.path.join(HERE, "..", "include", "grafx_b200.h"))
    return deps


def source_hash() -> str:
    """Digest of everything the library is built from (sources, headers, flags)."""
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + os.environ.get("GFX_NVCC_EXTRA", "").split()).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_rebuild() -> bool:
    """True when the library is absent or was built from other sources.  The comparison is by content (a digest
    recorded next to the .so at build time), not by mtime: a copy of the tree -- e.g. the snapshot sent to a GPU box --
    need not preserve timestamps, and a spurious rebuild there costs a minute per process."""
    if not os.path.exists(LIB):
        return True
    stamp = LIB + ".srchash"
    if os.path.exists(stamp):
        with open(stamp) as f:
            return f.read().strip() != source_hash()
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_rebuild():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build", os.environ.get("GFX_LIB_OUT", "default"))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("GFX_NVCC_EXTRA", "").split(), *PER_FILE_FLAGS.get(os.path.basename(src), ["-fmad=true"]), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-cudart", "static", "-Wno-deprecated-gpu-targets", "-o", LIB, *objs]
    subprocess.check_call(cmd)
    with open(LIB + ".srchash", "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
