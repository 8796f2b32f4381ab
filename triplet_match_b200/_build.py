"""In-tree build of the CUDA library (sm_100a).

`build_native()` compiles triplet_match_b200/csrc/*.cu with nvcc into
triplet_match_b200/libtriplet_match_b200.so (git-ignored, travels with gpurun).
Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 and -fmad=false:
the kernels must round every FP32 multiply and add separately to stay
bit-exact with the reference's SSE2 (no-FMA) build.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(ROOT)
CSRC = os.path.join(ROOT, "csrc")
# dev knob: TM_LIB_SUFFIX=_x TM_NVCC_EXTRA="-DTM_SCORE_MIN_BLOCKS=4" builds a tuning variant
_SUFFIX = os.environ.get("TM_LIB_SUFFIX", "")
OBJ = os.path.join(ROOT, "_obj" + _SUFFIX)
LIB = os.path.join(ROOT, "libtriplet_match_b200" + _SUFFIX + ".so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-Xptxas", "-v",
] + os.environ.get("TM_NVCC_EXTRA", "").split()
SOURCES = ["k_util.cu", "k_pairs.cu", "k_score.cu", "k_score2.cu", "k_early2.cu", "k_icp.cu", "k_uvicp.cu", "k_knn.cu", "k_sort.cu", "k_model.cu", "k_query.cu", "capi_core.cu", "capi_icp.cu", "capi_query.cu", "capi_nccl.cu"]
HOST_SOURCES = ["host_model.cpp"]
CXX = os.environ.get("CXX", "g++")
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
             "-Wall", "-Wno-sign-compare"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(REPO, "include", "tm_b200.h"))
    hs.append(os.path.join(REPO, "include", "tm_b200_host.h"))
    for h in ("tm_atan2f.h", "tm_sincosf.h", "tm_voxel_centre.h"):
        hs.append(os.path.join(REPO, "include", "triplet_match", h))
    return hs


def _run(cmd: list[str], log: str | None = None) -> None:
    p = subprocess.run(cmd, capture_output=True, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


def build_native(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    for src in HOST_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cpp", ".o"))
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        if s.endswith(".cpp"):
            _run([CXX] + CXX_FLAGS + ["-c", s, "-o", o], log=o + ".log")
        else:
            _run([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o], log=o + ".log")
        if verbose:
            print("compiled", os.path.basename(s))

    with ThreadPoolExecutor(max_workers=max(1, min(6, os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, jobs))
    if force or jobs or not _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-ldl", "-lgomp", "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True))
