"""Builds libd2pc_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m disparity_to_point_cloud_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but
travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libd2pc_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-I", INCLUDE, "-I", CSRC,
    # no -use_fast_math: gradFilter's division and the exact reprojection need IEEE semantics
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, INCLUDE):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cpp", ".h", ".cuh", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    exe = nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        cmd = [exe] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [exe, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


HARNESS = os.path.join(os.path.dirname(HERE), "tools", "d2pc_offline")


def build_harness(force: bool = False, sanitize: bool = False) -> str:
    """C++ offline harness over the node classes of include/d2pc_b200/nodes.hpp (g++, links libd2pc_b200.so).
    sanitize: the same harness under -fsanitize=address,undefined (the host side of the
    drop-in -- topic bus, launch-file parser, wire (de)serialisation, node classes -- runs under ASan / UBSan)."""
    src = os.path.join(os.path.dirname(HERE), "tools", "d2pc_offline.cpp")
    out = HARNESS + ("_asan" if sanitize else "")
    deps = [src, LIB] + [os.path.join(INCLUDE, "d2pc_b200", f) for f in os.listdir(os.path.join(INCLUDE, "d2pc_b200"))]
    deps.append(os.path.join(INCLUDE, "d2pc_b200.h"))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
             "-fno-omit-frame-pointer"] if sanitize else ["-O2"]
    cmd = ["g++", "-std=c++17", "-Wall"] + flags + ["-I", INCLUDE, src, "-o", out, "-L", HERE, "-ld2pc_b200",
                                                    "-Wl,-rpath,$ORIGIN/../disparity_to_point_cloud_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"harness build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_harness(force="--force" in sys.argv))
    print(build_harness(force="--force" in sys.argv, sanitize=True))
