"""In-tree build of libaptp_sm100.so (hand-written sm_100a CUDA behind the C ABI in include/aptp_sm100.h).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the repo
snapshot. Re-run `python -m diffusion_pruning_b200.build` after touching csrc/.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libaptp_sm100.so"
STAMP = PKG / ".libaptp_sm100.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "aptp_sm100.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src.name}:\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
        (objdir / (src.stem + ".ptxas.log")).write_text(out)
        objs.append(str(obj))
    if failed:
        raise RuntimeError("libaptp_sm100.so build failed")
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-lcudart"]
    subprocess.check_call(cmd)
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
