"""Time kernel variants on one workload: bp_cs_set_option("variant") -1 = default, bit 0 no small-operand kernel, bit 1 no
shadows in the fat kernels, bit 2 park, bit 3 no integer pass over the fat rows; 100+k = the experimental template variants of a
library built with BP_EXPERIMENTAL_VARIANTS=1.  --masks: 1 thin-row kernels only, 2 fat-row kernels only, 3 both."""
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
ap.add_argument("--variants", default="-1,1,2,3")
ap.add_argument("--fat-terms", default="96")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--masks", default="3")
ap.add_argument("--opt", default="", help="key=v1,v2,.. extra option to sweep")
a = ap.parse_args()
L = ffi.load()
h, info = bench.build_workload(L, ffi, a.workload, 0, 1, 0)
stream = torch.cuda.Stream()
assert L.bp_cs_set_stream(h, ctypes.c_void_p(stream.cuda_stream)) == 0
out = torch.zeros(1, dtype=torch.int64, device="cuda")
res = []
with torch.cuda.stream(stream):
    okey, ovals = (a.opt.split("=")[0], [int(x) for x in a.opt.split("=")[1].split(",")]) if a.opt else (None, [None])
    for oval in ovals:
      if okey:
        assert L.bp_cs_set_option(h, okey.encode(), oval) == 0, L.bp_cs_last_error(h)
      for ft in [int(x) for x in a.fat_terms.split(",")]:
          assert L.bp_cs_set_option(h, b"fat_terms", ft) == 0
          for mask in [int(x) for x in a.masks.split(",")]:
              assert L.bp_cs_set_option(h, b"kernels_mask", mask) == 0
              for v in [int(x) for x in a.variants.split(",")]:
                  assert L.bp_cs_set_option(h, b"variant", v) == 0
                  for _ in range(3):
                      assert L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr())) == 0, L.bp_cs_last_error(h)
                  torch.cuda.synchronize()
                  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                  e0.record(stream)
                  for _ in range(a.iters):
                      L.bp_cs_check_async(h, ctypes.c_void_p(out.data_ptr()))
                  e1.record(stream)
                  torch.cuda.synchronize()
                  ms = e0.elapsed_time(e1) / a.iters
                  r = {"workload": a.workload, "opt": oval, "fat_terms": ft, "mask": mask, "variant": v, "ms": round(ms, 4),
                       "rows_per_s": info["rows"] / ms * 1e3,
                       "first_bad": int(out.item()) if int(out.item()) != 0x7FFFFFFFFFFFFFFF else None}
                  print(json.dumps(r), flush=True)
