#!/usr/bin/env python
"""Host-only timing of the witness packing pass (bp_pack_scalars = what bp_cs_recheck_scalars runs before it sends): n 32-byte
scalars -> bits + exception list, on this machine's cores.  No GPU.  One JSON line per (kernel, threads).

    python tools/time_pack.py [n_scalars] > profiles/...jsonl
"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(n):
    import numpy as np

    from bellpepper_b200 import ffi

    L = ffi.load()
    rng = np.random.default_rng(1)
    scal = np.zeros((n, 4), np.uint64)
    scal[:, 0] = rng.integers(0, 2, size=n, dtype=np.uint64)
    scal[:: max(1, n // 1000)] = (7, 0, 1, 0)  # ~1000 exceptions
    bits = np.zeros((n + 7) // 8, np.uint8)
    idx = np.zeros(4096, np.uint64)
    vals = np.zeros((4096, 4), np.uint64)
    k = ctypes.c_uint64()
    best = None
    for _ in range(7):
        t0 = time.perf_counter()
        rc = L.bp_pack_scalars(scal.ctypes.data, n, bits.ctypes.data, idx.ctypes.data, vals.ctypes.data, 4096, ctypes.byref(k))
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    print(json.dumps({"kernel": L.bp_pack_kernel().decode(), "threads": int(os.environ["BP_PACK_THREADS"]), "scalars": n, "bytes": 32 * n,
                      "exceptions": k.value, "best_ms": round(best * 1e3, 2), "GBps": round(32 * n / best / 1e9, 1),
                      "host": "cpus=%d" % len(os.sched_getaffinity(0))}))


if __name__ == "__main__":
    if os.environ.get("BP_TIME_PACK_CHILD"):
        child(int(sys.argv[1]))
    else:
        n = int(sys.argv[1]) if len(sys.argv) > 1 else 109272413  # aux variables of sha256 x4096
        cpus = len(os.sched_getaffinity(0))
        for simd in ("0", "1"):
            for th in sorted({1, max(1, cpus // 2), cpus}):
                env = {**os.environ, "BP_TIME_PACK_CHILD": "1", "BP_PACK_SIMD": simd, "BP_PACK_THREADS": str(th)}
                subprocess.run([sys.executable, os.path.abspath(__file__), str(n)], env=env, check=True)
