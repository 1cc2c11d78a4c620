#!/usr/bin/env python3
"""Build libchromosight_b200.so (sm_100a) in-tree with nvcc.

    python chromosight_b200/csrc/build.py [--force] [--verbose]

The shared library lands next to the package (chromosight_b200/libchromosight_b200.so)
so that it travels with the repository snapshot to the GPU box.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["image.cu", "pearson.cu", "scores.cu", "detrend.cu", "gather.cu", "host_api.cu", "host_expand.cpp"]
CXX = os.environ.get("CXX", "g++")
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "chromosight_b200.h")]
LIB = os.path.join(PKG, "libchromosight_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


VARIANT = os.environ.get("CS_BUILD_VARIANT", "")     # experiments: a tagged library built with
VARIANT_DEFS = os.environ.get("CS_BUILD_DEFS", "").split()  # extra -D flags (scripts/gpu_r2.sh)


def _compile(src, verbose, ablate=False):
    stem, ext = os.path.splitext(src)
    obj = os.path.join(HERE, "build", stem + (".abl.o" if ablate else (f".{VARIANT}.o" if VARIANT else ".o")))
    deps = [os.path.join(HERE, src)] + [os.path.join(HERE, h) for h in HEADERS]
    if not _stale(obj, deps):
        return obj, ""
    if ext == ".cpp":  # plain host code
        cmd = [CXX, "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", os.path.join(HERE, src), "-o", obj]
    else:
        cmd = [NVCC] + FLAGS + (["-DCS_ABLATE"] if ablate else []) + VARIANT_DEFS + (["-Xptxas", "-v"] if verbose else []) \
            + ["-c", os.path.join(HERE, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"compiler failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False, ablate=False):
    """ablate=True builds libchromosight_b200_ablate.so: the same library with the phase
    switches and counters of the timing experiments compiled in (-DCS_ABLATE); the release
    library has none of them."""
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    if force:
        for f in os.listdir(os.path.join(HERE, "build")):
            os.remove(os.path.join(HERE, "build", f))
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose, ablate), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    lib = LIB.replace(".so", "_ablate.so") if ablate else (LIB.replace(".so", f"_{VARIANT}.so") if VARIANT else LIB)
    if force or _stale(lib, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ablate="--ablate" in sys.argv))
