"""In-tree build of libbmagwa_b200.so (CUDA kernels + C ABI + host sampler) and the `bmagwa` CLI.

    python -m bmagwa_b200.build            # build if sources are newer than the library
    python -m bmagwa_b200.build --force

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE
(bmagwa_b200/libbmagwa_b200.so) so that it travels to the GPU box with the repo snapshot.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libbmagwa_b200.so")
CLI = os.path.join(PKG, "bmagwa")
OBJ = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3,-march=x86-64-v3,-ffp-contract=off",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _ccbin():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def sources():
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cpp = sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    return cu, cpp


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    cu, cpp = sources()
    headers = (glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "host", "*.hpp")) +
               glob.glob(os.path.join(PKG, "..", "include", "*.h")))
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for src in cu + cpp:
        if os.path.basename(src) == "main.cpp":
            continue
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = [_nvcc(), "-ccbin", _ccbin()] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    if force or _newer(LIB, objs):
        cmd = [_nvcc(), "-ccbin", _ccbin(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lpthread"]
        subprocess.check_call(cmd)
    main = os.path.join(CSRC, "host", "main.cpp")
    if os.path.exists(main) and (force or _newer(CLI, [main, LIB])):
        cmd = [_ccbin(), "-O2", "-std=c++17", "-I", os.path.join(PKG, "..", "include"), main, "-o", CLI,
               "-L", PKG, "-lbmagwa_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
