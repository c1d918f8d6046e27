"""Build the native library in-tree: bellpepper_b200/libbp_r1cs.so (nvcc, sm_100a only)."""

from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbp_r1cs.so")
FRONTEND = os.path.join(HERE, "libbp_frontend.so")  # the gadget front-end alone, host recording only (g++, no CUDA)
MICROBENCH = os.path.join(HERE, "bin", "microbench")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3",
]


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources():
    return (glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))
            + glob.glob(os.path.join(CSRC, "host", "*.hpp")) + glob.glob(os.path.join(CSRC, "host", "*.cpp"))
            + glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def build_native(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    if force or _stale(LIB, srcs):
        cmd = ["nvcc", *NVCC_FLAGS, "-shared", "-o", LIB, os.path.join(CSRC, "bp_r1cs.cu"),
               os.path.join(CSRC, "host", "fixtures.cpp"), os.path.join(CSRC, "host", "pack.cpp"), os.path.join(CSRC, "host", "structure_hash.cpp"), "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        if os.environ.get("BP_EXPERIMENTAL_VARIANTS") == "1":
            cmd.insert(1, "-DBP_EXPERIMENTAL_VARIANTS")
        subprocess.run(cmd, check=True, cwd=CSRC)
    host = os.path.join(CSRC, "host")
    if force or _stale(FRONTEND, srcs + [os.path.join(host, "frontend.map")]):
        subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-Wno-subobject-linkage", "-Wl,--version-script=frontend.map",
                        "-o", FRONTEND, "frontend_host.cpp"], check=True, cwd=host)
    for name in ("microbench", "microbench2", "microbench3"):
        mb, out = os.path.join(CSRC, name + ".cu"), os.path.join(os.path.dirname(MICROBENCH), name)
        if os.path.exists(mb) and (force or _stale(out, srcs)):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["nvcc", *NVCC_FLAGS, "-o", out, mb], check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
