"""Build recipe for libbdf_b200.so (sm_100a only): nvcc on csrc/engine.cu plus one object per padded latent
dimension from csrc/row_inst.cu, compiled in parallel, linked into one shared library next to this file.

    python bayesiandatafusion.jl_b200/build.py [--force] [--ptxas-v]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# variant builds for A/B measurements: BDF_LIB_TAG=<tag> BDF_EXTRA_NVCC="-D..." → libbdf_<tag>.so from its own object directory
TAG = os.environ.get("BDF_LIB_TAG", "")
OBJ = os.path.join(HERE, "build_" + TAG if TAG else "build")
LIB = os.path.join(HERE, f"libbdf_{TAG}.so" if TAG else "libbdf_b200.so")
DPS = [8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "bdf_b200.h"))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(args):
    src, obj, extra = args
    cmd = [NVCC, *FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False, ptxas_v: bool = False) -> str:
    deps = _sources()
    if not force and not os.environ.get("BDF_BUILD_DPS") and not _stale(LIB, deps):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    extra = ["-Xptxas", "-v"] if ptxas_v else []
    if TAG:
        extra = extra + os.environ.get("BDF_EXTRA_NVCC", "").split()
    jobs = []
    for name in ("engine", "features", "predict"):
        if force or _stale(os.path.join(OBJ, f"{name}.o"), deps):
            jobs.append((os.path.join(CSRC, f"{name}.cu"), os.path.join(OBJ, f"{name}.o"), extra))
    only = os.environ.get("BDF_BUILD_DPS")  # dev builds: recompile only these padded dimensions (others keep their objects)
    extra_defs = [] if TAG else os.environ.get("BDF_EXTRA_NVCC", "").split()
    for dp in DPS:
        if only and str(dp) not in only.split(",") and os.path.exists(os.path.join(OBJ, f"row_inst_{dp}.o")):
            continue
        jobs.append((os.path.join(CSRC, "row_inst.cu"), os.path.join(OBJ, f"row_inst_{dp}.o"), [f"-DBDF_DP={dp}", *extra, *extra_defs]))
    jobs = [j for j in jobs if force or (only and "row_inst" in j[1]) or _stale(j[1], deps)]
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        for obj, log in ex.map(_compile, jobs):
            if verbose or ptxas_v:
                print(f"[bdf build] {os.path.basename(obj)}")
                if ptxas_v:
                    print(log)
    objs = [os.path.join(OBJ, "engine.o"), os.path.join(OBJ, "features.o"), os.path.join(OBJ, "predict.o")] + [os.path.join(OBJ, f"row_inst_{dp}.o") for dp in DPS]
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
           "-lcublas", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]  # dense feature products (FᵀF, F·β, UᵀV) are plain library dgemm calls
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_v="--ptxas-v" in sys.argv))
