"""Build libhgs_raster.so (the C-ABI library of include/hgs_raster.h) with nvcc for sm_100a.

In-tree build: objects under horizongs_b200/csrc/build/, library at horizongs_b200/_C/libhgs_raster.so
(git-ignored, but it travels to the GPU box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
Usage:  python -m horizongs_b200.csrc.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT_DIR = os.path.join(PKG, "_C")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(OUT_DIR, "libhgs_raster.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]

# file -> extra flags.  Projection / SH are compiled WITHOUT fma contraction so that every float32
# intermediate rounds exactly like the oracle's explicit torch expressions (integer radii depend on it).
SOURCES = {
    "api.cu": [],
    "project3d.cu": ["-fmad=false"],
    "project2d.cu": ["-fmad=false"],
    "sh.cu": ["-fmad=false"],
    "gauss_bwd.cu": ["-fmad=false"],
    "isect.cu": [],
    "blend3d.cu": [],
    "blend2d.cu": [],
    "blend2d_fast.cu": [],
    "densify.cu": [],
    "normals.cu": [],
    "loss.cu": [],
    "decode.cu": [],
    "exchange.cu": [],
    "exchange_vjp.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, flags) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "hgs_raster.h"))
    sources = {k: v for k, v in SOURCES.items() if os.path.exists(os.path.join(HERE, k))}
    jobs = []
    objs = []
    for src, extra in sources.items():
        path = os.path.join(HERE, src)
        flags = ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else [])
        tag = _digest([path] + headers, flags)
        obj = os.path.join(OBJ_DIR, f"{os.path.splitext(src)[0]}.{tag}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append([nvcc, *flags, "-c", path, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}")

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    stamp = os.path.join(OBJ_DIR, "link.stamp")
    want = " ".join(sorted(objs))
    have = open(stamp).read() if os.path.exists(stamp) else ""
    if force or jobs or want != have or not os.path.exists(LIB):
        run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"])
        with open(stamp, "w") as f:
            f.write(want)
        # drop stale objects
        keep = set(objs)
        for f in os.listdir(OBJ_DIR):
            p = os.path.join(OBJ_DIR, f)
            if f.endswith(".o") and p not in keep:
                os.remove(p)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
