"""Build libparrot_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m parrot_tts_b200.build [--force] [--verbose] [--variant NAME --define MACRO ...]

Objects go to parrot_tts_b200/build/ (git-ignored); the shared library to parrot_tts_b200/lib/.
The library is rebuilt only when a source is newer than it.

``--variant NAME --define MACRO`` builds an EXPERIMENT beside the product: objects in build/NAME/, library
lib/libparrot_b200_NAME.so, loaded instead of the product only when ``PB_LIB_VARIANT=NAME`` is set (``_lib.py``) - so an
A/B run needs no rebuild and the product library is never overwritten by an experiment.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libparrot_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for d in (CSRC, INCLUDE):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def needs_build(lib=LIB):
    return not os.path.isfile(lib) or os.path.getmtime(lib) < _deps_mtime()


def variant_lib(variant):
    return os.path.join(LIBDIR, f"libparrot_b200_{variant}.so") if variant else LIB


def build(force=False, verbose=False, defines=(), variant=""):
    if defines and not variant:
        raise ValueError("a build with extra macros is an experiment: give it a --variant name (the product library keeps its flags)")
    LIB = variant_lib(variant)
    OBJ = os.path.join(HERE, "build", variant) if variant else os.path.join(HERE, "build")
    NVCC_FLAGS = globals()["NVCC_FLAGS"] + [f"-D{d}" for d in defines]
    if not force and not needs_build(LIB):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr_m = max(os.path.getmtime(os.path.join(d, f)) for d in (CSRC, INCLUDE) for f in os.listdir(d)
                if f.endswith((".cuh", ".h")))

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_m):
            return obj, ""
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    objs = [o for o, _ in results]
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    def _values(flag):
        return [sys.argv[i + 1] for i, a in enumerate(sys.argv[:-1]) if a == flag]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=_values("--define"),
                variant=(_values("--variant") or [""])[0]))
