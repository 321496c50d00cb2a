"""Build libomg_b200.so in-tree with nvcc for sm_100a.

    python -m openmg_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libomg_b200.so")
SOURCES = ["omg_api.cu", "omg_setup.cu", "omg_cycle.cu", "omg_stencil.cu", "omg_dist.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


def _deps():
    out = [os.path.join(HERE, "..", "include", "omg_b200.h")]
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    dep_time = max(os.path.getmtime(p) for p in _deps())
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        op = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(op)
        if (not force and os.path.exists(op)
                and os.path.getmtime(op) >= max(os.path.getmtime(sp), dep_time)):
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", op]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s ----\n%s\n" % (src, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if (force or procs or not os.path.exists(LIB)):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
