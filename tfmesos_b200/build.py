"""Builds tfmesos_b200/lib/libpsx.so (sm_100a only) with nvcc, in-tree.

The .so is git-ignored but travels with the gpurun snapshot; there is no JIT
cache and no fallback: if the library is missing, ``tfmesos_b200.psx`` raises.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "psx.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "psx_kernels.cuh"),
        os.path.join(HERE, "csrc", "psx_nvls.cuh"),
        os.path.join(ROOT, "include", "psx.h")]
OUT = os.path.join(HERE, "lib", "libpsx.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # one IEEE rounding per operation: parity with the CPU oracle is bit-exact
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpsx.so cannot be built")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"),
                                  "-I", os.path.join(HERE, "csrc")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [SRC, "-o", OUT]
    env = dict(os.environ)
    env.pop("CC", None)      # the image's $CC lacks pieces nvcc's host pass needs
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
