"""Time emit mode (bp_cs_eval_async: canonical A.w, B.w, C.w of every row into device buffers) next to the check."""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from bellpepper_b200 import ffi

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sha256_chain_512_pallas")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
L = ffi.load()
h, info = bench.build_workload(L, ffi, a.workload, 0, 1, 0)
stream = torch.cuda.Stream()
assert L.bp_cs_set_stream(h, ctypes.c_void_p(stream.cuda_stream)) == 0
n = info["rows"]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
az, bz, cz = (torch.empty((n, 4), dtype=torch.int64, device="cuda") for _ in range(3))
with torch.cuda.stream(stream):
    for name, fn in (("check", lambda: L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr()))),
                     ("emit", lambda: L.bp_cs_eval_async(h, ctypes.c_void_p(az.data_ptr()), ctypes.c_void_p(bz.data_ptr()),
                                                         ctypes.c_void_p(cz.data_ptr())))):
        for _ in range(3):
            assert fn() == 0, L.bp_cs_last_error(h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"workload": a.workload, "op": name, "ms": round(ms, 4), "rows_per_s": n / ms * 1e3,
                          "out_GBps": (96 * n / ms / 1e6) if name == "emit" else None}), flush=True)
