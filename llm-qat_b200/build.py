"""Build libqat_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python llm-qat_b200/build.py [--force] [--verbose]

Objects go to llm-qat_b200/build/, the library to llm-qat_b200/lib/ (both
git-ignored; the .so travels to the GPU box with the gpurun snapshot).
cudart is linked dynamically (libcudart.so.12: the one PyTorch has already loaded in
a Python process; an rpath to the toolkit's copy serves plain C hosts); the driver API
(cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libqat_b200.so")
INCLUDE = os.path.join(ROOT, "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-I" + INCLUDE, "-I" + CSRC, "-Xcompiler", "-fPIC"]
# per-file extra flags.  The fake-quant kernels must not contract a*b+c into an
# FMA; they also use the __f*_rn intrinsics, so the flag
# is belt and braces.
EXTRA = {
    "fakequant.cu": ["-fmad=false"],
    "ste.cu": ["-fmad=false"],
    "lowbit.cu": ["-fmad=false"],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libqat_b200.so")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stamp_inputs():
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    deps += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    deps.append(os.path.abspath(__file__))
    return deps


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link lib/libqat_b200.so.
    Skips the work when sources, header and this script are unchanged."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp.sha256")
    digest = _digest(_stamp_inputs())
    if not force and os.path.exists(LIB) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return LIB
    cc = nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [cc, *ARCH, *COMMON, *EXTRA.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    link = [cc, *ARCH, "-shared", "-cudart", "shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC",
            "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
