"""In-tree build of libphylo_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m phylogenetics_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libphylo_b200.so")
SOURCES = ["pb_api.cu", "pb_host_linalg.cpp", "pb_sharded.cpp"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".cpp", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False, defines=(), out=None):
    """`defines` / `out` build a VARIANT of the library beside the shipped one (A/B measurements: load it with
    PHYLO_B200_LIB=<out>, tools/ab_lib.py)."""
    out = out or OUT
    if not force and os.path.exists(out) and os.path.getmtime(out) >= _newest_source_mtime():
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    defs = [argv[i + 1] for i, a in enumerate(argv) if a == "--define"]
    outp = next((argv[i + 1] for i, a in enumerate(argv) if a == "--out"), None)
    print(build(force="--force" in argv, verbose="--verbose" in argv, defines=defs, out=outp))
