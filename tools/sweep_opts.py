"""Time one check of a workload under combinations of bp_cs_set_option values (development aid).
   python tools/sweep_opts.py sha256_chain_512_pallas fat_int_ctas_per_sm=1,2,3,6 small_ctas_per_sm=3,4,5"""
import ctypes
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from bellpepper_b200 import ffi

name = sys.argv[1]
axes = [(kv.split("=")[0], [int(x) for x in kv.split("=")[1].split(",")]) for kv in sys.argv[2:]]
ctx = bench.Ctx()
ctx.L, ctx.ffi, ctx.torch, ctx.dist = ffi.load(), ffi, torch, None
ctx.rank, ctx.local_rank, ctx.world = 0, 0, 1
L = ctx.L
h, info = bench.build_workload(ctx, name)
stream = torch.cuda.Stream()
assert L.bp_cs_set_stream(h, ctypes.c_void_p(stream.cuda_stream)) == 0
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for combo in itertools.product(*[v for _, v in axes]):
    for (k, _), v in zip(axes, combo):
        assert L.bp_cs_set_option(h, k.encode(), v) == 0, L.bp_cs_last_error(h)
    with torch.cuda.stream(stream):
        for _ in range(5):
            assert L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr())) == 0, L.bp_cs_last_error(h)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr()))
        e1.record(stream)
        torch.cuda.synchronize()
    print(json.dumps({"workload": name, **{k: v for (k, _), v in zip(axes, combo)}, "ms": round(e0.elapsed_time(e1) / 20, 4),
                      "first_bad": int(out.item()) if int(out.item()) != bench.SAT else None}), flush=True)
