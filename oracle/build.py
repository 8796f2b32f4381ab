"""Builds the test infrastructure under oracle/: liboracle.so (the CPU restatement) and, when
/root/reference is present, oracle/_ref/libtm_ref.so (the reference's own sources compiled against
oracle/shim) and oracle/_ref/libtm_ref_cl.so (its OpenCL kernels).  Called by __graft_entry__.build(); nothing in the product imports this."""
from __future__ import annotations

import os
import subprocess
import sys

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libtm_ref.so")
REF_CL_LIB = os.path.join(ORACLE_DIR, "_ref", "libtm_ref_cl.so")  # the reference's OpenCL kernels as C++


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def _make(target: str) -> None:
    p = subprocess.run(["make", "-C", ORACLE_DIR, "-s", target], capture_output=True, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout + p.stderr)
        raise RuntimeError("oracle build failed: make " + target)


def build_oracle(force: bool = False) -> str:
    deps = [os.path.join(ORACLE_DIR, f) for f in ("oracle.hpp", "oracle_capi.cpp", "Makefile")]
    deps.append(os.path.join(os.path.dirname(ORACLE_DIR), "include", "triplet_match", "tm_sincosf.h"))
    if force or not _newer(ORACLE_LIB, deps):
        _make(ORACLE_DIR + "/liboracle.so")
    if os.path.isdir("/root/reference/include"):
        shim = [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ORACLE_DIR, "shim")) for f in fs]
        ref_deps = deps + [os.path.join(ORACLE_DIR, f) for f in ("ref_shim_capi.cpp", "ref_cl_capi.cpp")] + shim
        if force or not _newer(REF_LIB, ref_deps) or not _newer(REF_CL_LIB, ref_deps):
            _make("ref")
    return ORACLE_LIB


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
