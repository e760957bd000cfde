"""Builds libgnnepcsaft_b200.so (sm_100a) in-tree with nvcc — no torch headers, plain C ABI."""

from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgnnepcsaft_b200.so")
SOURCES = ["graph_build.cu", "aggregate.cu", "gemm_simt.cu", "gemm_tc.cu", "norm_pool.cu", "embed.cu", "fold.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas=-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(HERE)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(HERE, name), "rb").read())
    h.update(open(os.path.join(HERE, "..", "..", "include", "gnnepcsaft_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    """profile=True builds libgnnepcsaft_b200_prof.so with -DGNN_B200_PROFILE (per-role cycle counters of the
    tensor-core GEMMs and the gnn_tc_set_profile_buffer hook, tools/tc_probe.py); the product library has neither."""
    lib = LIB.replace(".so", "_prof.so") if profile else LIB
    flags = NVCC_FLAGS + (["-DGNN_B200_PROFILE"] if profile else [])
    stamp = lib + ".stamp"
    digest = _digest() + ("p" if profile else "")
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == digest:
        return lib
    objdir = os.path.join(HERE, "build_prof" if profile else "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(digest)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, profile="--profile" in sys.argv))
