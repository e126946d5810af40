"""Builds the CUDA extension in-tree: easyrag_b200/_lib/libeasyrag_b200.so (sm_100a only).

``python -m easyrag_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_lib"
LIB = OUT_DIR / "libeasyrag_b200.so"
STAMP = OUT_DIR / "build.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--cudart", "static",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted((CSRC / "encoder").glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = _sources() + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) \
        + sorted((CSRC / "encoder").glob("*.cuh")) + [PKG.parent / "include" / "easyrag_b200.h", Path(__file__)]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False) -> Path:
    OUT_DIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objs = []
    procs = []
    for src in _sources():
        obj = OUT_DIR / (src.stem + ".o")
        cmd = [nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src.name}\n{out}")
        if pr.returncode != 0:
            failed = True
        objs.append(str(obj))
    (OUT_DIR / "build.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see easyrag_b200/_lib/build.log")
    if verbose:
        print("\n".join(log))
    link = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "--cudart", "static",
            "-Xcompiler", "-fPIC", "-o", str(LIB), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
