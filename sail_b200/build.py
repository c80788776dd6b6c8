"""Builds sail_b200/_build/libsailgpu.so from sail_b200/csrc/*.cu with nvcc for sm_100a (in-tree, so
the .so travels to the GPU box with the repo snapshot).  Incremental: one object per source."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libsailgpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newer(a: str, b: str) -> bool:
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    sources = sorted(f for f in os.listdir(SRC) if f.endswith(".cu"))
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".hpp", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "sailgpu.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    jobs = []
    for s in sources:
        src, obj = os.path.join(SRC, s), os.path.join(OUT, s[:-3] + ".o")
        if force or _newer(src, obj) or os.path.getmtime(obj) < newest_header:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for warn in ex.map(compile_one, jobs):
                if verbose and warn.strip():
                    print(warn, file=sys.stderr)
    objs = [os.path.join(OUT, s[:-3] + ".o") for s in sources]
    if jobs or not os.path.exists(LIB):
        nccl = []   # filled by ops_more.cu's link needs (libnccl) once the exchange lands
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-lcudart", *nccl]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
