"""Builds libdiff3d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m diff3dhpe_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# A/B builds: D3D_LIB_OUT=<path>.so D3D_NVCC_EXTRA="-DFOO=1 ..." python -m diff3dhpe_b200.build   (load it with D3D_LIB=<path>.so)
LIB = Path(os.environ.get("D3D_LIB_OUT", PKG / "libdiff3d_b200.so"))
STAMP = LIB.with_name("." + LIB.stem + ".stamp")
EXTRA = os.environ.get("D3D_NVCC_EXTRA", "").split()

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "diff3d_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + EXTRA).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objs = []
    build_dir = PKG / ("build" if "D3D_LIB_OUT" not in os.environ else "build_" + LIB.stem)
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc_path(), *NVCC_FLAGS, *EXTRA, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src.name}\n{out}", flush=True)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc_path(), "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
