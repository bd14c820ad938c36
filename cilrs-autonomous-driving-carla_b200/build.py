"""Build recipe for libcilrs_b200.so: every .cu under csrc/ compiled for sm_100a with nvcc, linked into one
C-ABI shared library that lives in-tree next to this file (so it travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcilrs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_dep():
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
               if f.endswith((".cu", ".cuh", ".h"))) if os.listdir(CSRC) else 0


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr = os.path.join(os.path.dirname(HERE), "include", "cilrs_b200.h")
    newest = max(_newest_dep(), os.path.getmtime(hdr))
    srcs = _sources()
    todo = []
    for s in srcs:
        o = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(o) or os.path.getmtime(o) < newest:
            todo.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        if verbose:
            print("compiled", s)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(cc, todo))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
