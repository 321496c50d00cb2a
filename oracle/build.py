"""Build the oracle's compiled C restatement (test infrastructure only).

    python -m oracle.build

gcc -O2 (no -ffast-math: keep IEEE evaluation order) -> oracle/_build/liboracle_kernels.so
(git-ignored, travels to the GPU box with the snapshot).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    src = os.path.join(HERE, "csrc", "oracle_kernels.c")
    outdir = os.path.join(HERE, "_build")
    out = os.path.join(outdir, "liboracle_kernels.so")
    os.makedirs(outdir, exist_ok=True)
    if (not force) and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-o", out, src])
    return out


if __name__ == "__main__":
    print(build(force=True))
