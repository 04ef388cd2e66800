"""Build the C-ABI shared library ``hemelb_b200/libhemelb_b200.so`` for sm_100a with nvcc.

In-tree build (no JIT cache): the .so travels with the repository snapshot to the GPU box.
``-fmad=false`` keeps the kernels' arithmetic bit-identical to the reference's scalar x86-64 path
(the hot path is HBM-bound; the un-fused multiplies are free).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libhemelb_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + os.environ.get("HLB_NVCC_EXTRA", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "hemelb_b200.h"))
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            jobs.append([NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and r.stderr:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([NVCC, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
