"""TEST INFRASTRUCTURE -- the CPU oracle for the R1CS evaluation hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this package.  The product (`bellpepper_b200/`) never does.

* `oracle.r1cs_py`   pure-Python big-int restatement (authoritative, small cases)
* `oracle.bp_oracle` C restatement (`bp_oracle.c`, 4x64 CIOS Montgomery, optional OpenMP) via ctypes
* `oracle.synth`     the counter-based synthetic-instance recipe in Python (checks the C and CUDA ones)

`/root/reference` is Rust and there is no Rust toolchain in this image, so there is no `oracle/_ref`
build of the real reference; see DESIGN.md "Oracle".
"""

from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbp_oracle.so")
_SRC = os.path.join(_HERE, "bp_oracle.c")


def build(force: bool = False) -> str:
    """Compile oracle/bp_oracle.c -> oracle/libbp_oracle.so (gcc -O3 -march=native -fopenmp)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=gnu11", "-o", _SO, _SRC]
    subprocess.run(cmd, check=True)
    with open(_SO + ".cpu", "w") as fh:
        fh.write(_cpu_digest())
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    """Load the C oracle.  On the GPU box the prebuilt .so travels with the snapshot; if its
    -march=native code cannot run there it is rebuilt from source (gcc is in the image)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO) or os.environ.get("BP_ORACLE_REBUILD") == "1" or _foreign_cpu():
        build(force=True)
    L = ctypes.CDLL(_SO)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.bpo_field_params.argtypes = [ctypes.c_int, u64p, u64p, u64p]
    L.bpo_mul.argtypes = [ctypes.c_int, u64p, u64p, u64p]
    L.bpo_add.argtypes = [ctypes.c_int, u64p, u64p, u64p]
    L.bpo_is_canonical.argtypes = [ctypes.c_int, u64p]
    L.bpo_prepare.restype = ctypes.c_void_p
    L.bpo_prepare.argtypes = [ctypes.c_int, ctypes.c_uint64, u32p, u32p, u64p, u64p, ctypes.c_uint64, u64p, ctypes.c_uint64]
    L.bpo_free.argtypes = [ctypes.c_void_p]
    L.bpo_set.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, u64p]
    L.bpo_check.restype = ctypes.c_int64
    L.bpo_check.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, u64p, u64p, u64p]
    L.bpo_max_threads.restype = ctypes.c_int
    L.bpo_synth_len.restype = ctypes.c_uint32
    L.bpo_synth_len.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int]
    L.bpo_synth_lens.restype = ctypes.c_uint64
    L.bpo_synth_lens.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, u32p]
    L.bpo_synth_fill.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                 ctypes.c_uint64, ctypes.c_uint64, u32p, u64p]
    L.bpo_synth_witness.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, u64p]
    L.bpo_synth_witness_at.argtypes = [ctypes.c_int, ctypes.c_uint64, u64p, ctypes.c_uint64, u64p]
    _lib = L
    return L


def _cpu_digest() -> str:
    import hashlib

    try:
        with open("/proc/cpuinfo") as fh:
            flags = next((ln for ln in fh if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha256(flags.encode()).hexdigest()


def _foreign_cpu() -> bool:
    """The .so is built with -march=native; build() records the build host's CPU flags beside it."""
    try:
        with open(_SO + ".cpu") as fh:
            return fh.read().strip() != _cpu_digest()
    except OSError:
        return True
