"""Build the CUDA C-ABI library in-tree for sm_100a (B200).

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
to the GPU box with the repo snapshot.  Sources are compiled to objects in parallel
(build/, git-ignored) and only when they or a header changed.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhycom_tsadvc_b200.so")
SOURCES = ["tsadvc_kernels.cu", "halo.cu", "tsadvc_abi.cu", "xc_comm.cu", "tsdff.cu", "fct2c.cu", "asselin.cu",
           "synth.cu", "cnuity.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # keep the Fortran operation order: no FMA contraction
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fopenmp",
]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")]
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    deps.append(os.path.abspath(__file__))
    return deps


def _stale(extra: str) -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in _sources()] + _headers()
    tag = os.path.join(OBJ, "flags.txt")
    if not os.path.exists(tag) or open(tag).read() != extra:
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    extra = os.environ.get("HYCOM_TSADVC_NVCC_EXTRA", "")   # experiments (e.g. -DTSADVC_...)
    if not force and not _stale(extra):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libhycom_tsadvc_b200.so")
    os.makedirs(OBJ, exist_ok=True)
    base = [nvcc]
    # the image exports CC/CXX=/opt/gcc/bin/* (no libgomp.spec); use the distro g++
    if os.path.exists("/usr/bin/g++"):
        base += ["-ccbin", "/usr/bin/g++"]
    base += NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + extra.split()
    tag = hashlib.sha1(" ".join(base).encode()).hexdigest()[:10]
    hdr_t = max(os.path.getmtime(d) for d in _headers())

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{tag}.o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t):
            return obj, ""
        r = subprocess.run(base + ["-c", path, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(base + ["-c", path]) + "\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_one, _sources()))
    if verbose:
        for _, err in res:
            sys.stderr.write(err)
    cmd = base + ["-shared", "-o", LIB] + [o for o, _ in res] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    with open(os.path.join(OBJ, "flags.txt"), "w") as f:
        f.write(extra)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
