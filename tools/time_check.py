"""Quick device-side timing of the check kernels on a synthetic instance (development aid; bench.py is the contract)."""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bellpepper_b200 import ffi

ap = argparse.ArgumentParser()
ap.add_argument("--field", type=int, default=0)
ap.add_argument("--log-rows", type=int, default=22)
ap.add_argument("--t", type=int, default=6)
ap.add_argument("--log-vars", type=int, default=None)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--kernels", default="0", help="comma list of fat_terms values (0 = the library's default threshold)")
ap.add_argument("--opts", default="")
a = ap.parse_args()

L = ffi.load()
N = 1 << a.log_rows
n_vars = 1 << (a.log_vars if a.log_vars is not None else a.log_rows)
h = ffi.vp()
assert L.bp_cs_new(a.field, 0, N, int(N * 3 * a.t * 1.02) + 1024, n_vars, ctypes.byref(h)) == 0
st = torch.cuda.current_stream().cuda_stream or 1  # 0 = legacy default stream -> cudaStreamLegacy
assert L.bp_cs_set_stream(h, ctypes.c_void_p(st)) == 0
t0 = time.time()
assert L.bp_cs_synth_witness(h, 0x5962BE3D763D318D, n_vars, 16) == 0
CH = 1 << 20
for r0 in range(0, N, CH):
    rc = L.bp_cs_synth_rows(h, 0x5962BE3D763D318D, a.t, n_vars, 16, r0, min(CH, N - r0))
    assert rc == 0, L.bp_cs_last_error(h)
torch.cuda.synchronize()
gen_s = time.time() - t0
c = [ctypes.c_uint64() for _ in range(4)]
L.bp_cs_counts(h, *[ctypes.byref(x) for x in c])
n_in, n_aux, rows, nnz = [x.value for x in c]
alg_bytes = nnz * 36 + rows * 12 + (n_in + n_aux) * 32
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    assert L.bp_cs_set_option(h, k.encode(), int(v)) == 0, L.bp_cs_last_error(h)
for kern in [int(k) for k in a.kernels.split(",")]:
    if kern:
        assert L.bp_cs_set_option(h, b"fat_terms", kern) == 0
    for _ in range(3):
        assert L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr())) == 0, L.bp_cs_last_error(h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        assert L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr())) == 0
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    used = ctypes.c_int64()
    L.bp_cs_get_option(h, b"fat_rows", ctypes.byref(used))
    print(json.dumps({"fat_terms": kern, "fat_rows": used.value, "field": a.field, "rows": rows, "nnz": nnz, "t": a.t, "n_vars": n_in + n_aux,
                      "ms": round(ms, 4), "constraints_per_s": rows / ms * 1e3, "terms_per_s": nnz / ms * 1e3,
                      "alg_GBps": alg_bytes / ms / 1e6, "first_bad": int(out.item()), "gen_s": round(gen_s, 2)}), flush=True)
L.bp_cs_free(h)
