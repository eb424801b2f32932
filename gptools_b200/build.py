"""Build libgptb200.so (hand-written sm_100a CUDA + the C-ABI) in-tree with nvcc.

    python -m gptools_b200.build          # rebuild if sources are newer than the library

The library lives next to its sources (gptools_b200/csrc/libgptb200.so): it is git-ignored but
travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "libgptb200.so")
HOSTCHECK = os.path.join(CSRC, "libgptb200_hostcheck.so")
SOURCES = ["api.cu", "assemble.cu", "gemm.cu", "factor.cu", "predict.cu", "batched4.cu"]
HEADERS = ["common.cuh", "covfn.cuh", "covfn_hyper.cuh", "internal.h", "se_fast.cuh", os.path.join(INCLUDE, "gptb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    objs, jobs = [], []
    for src in srcs:
        obj = src[:-3] + ".o"
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-I", INCLUDE, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            jobs.append(cmd)
    if jobs:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(subprocess.check_call, jobs))
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    hc_src = os.path.join(CSRC, "hostcheck.cpp")
    if force or _stale(HOSTCHECK, [hc_src, os.path.join(CSRC, "covfn.cuh"), os.path.join(CSRC, "covfn_hyper.cuh")]):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", HOSTCHECK, hc_src, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
