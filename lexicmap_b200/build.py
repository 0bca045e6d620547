"""Build recipes for the in-tree native artefacts (sm_100a only; no other arch, no JIT cache).

    liblexicmap_gpu.so   CUDA engine + C ABI (include/lexicmap_gpu.h)      <- csrc/engine.cu
    bin/lmi-tools        host tools: minimal .lmi index writer, synthetic genomes / queries   <- csrc/lmi_build.cpp
    bin/lexicmap-gpu     C++ `lexicmap search` look-alike CLI over the C ABI                <- csrc/search_cli.cpp
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblexicmap_gpu.so")
BIN = os.path.join(HERE, "bin")
TOOLS = os.path.join(BIN, "lmi-tools")
CLI = os.path.join(BIN, "lexicmap-gpu")
GXX = "/usr/bin/g++"   # the image exports CXX=/opt/gcc/bin/g++ whose OpenMP spec file is missing
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _newer(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_gpu_lib(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))] + [os.path.join(HERE, "..", "include", "lexicmap_gpu.h")]
    if force or _newer(LIB, srcs):
        _run([NVCC, "-ccbin", GXX, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp,-O3",
              "-shared", os.path.join(CSRC, "engine.cu"), "-o", LIB, "-lz"])
    return LIB


def build_tools(force=False):
    os.makedirs(BIN, exist_ok=True)
    src = os.path.join(CSRC, "lmi_build.cpp")
    if force or _newer(TOOLS, [src, os.path.join(CSRC, "lmi_format.hpp")]):
        _run([GXX, "-O3", "-march=x86-64-v3", "-fopenmp", "-std=c++17", src, "-o", TOOLS, "-lz"])
    cli = os.path.join(CSRC, "search_cli.cpp")
    if os.path.exists(cli) and (force or _newer(CLI, [cli, LIB])):
        _run([GXX, "-O2", "-std=c++17", cli, "-o", CLI, "-I", os.path.join(HERE, "..", "include"), "-L", HERE, "-llexicmap_gpu", "-lz", "-Wl,-rpath,$ORIGIN/.."])
    return TOOLS


def build_all(force=False):
    build_gpu_lib(force)
    build_tools(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB, TOOLS)
