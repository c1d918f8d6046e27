"""Under torchrun: where does a group check's time go?  Per rank: the check alone, the exchange alone, both (development aid)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bellpepper_b200 import ffi, fixtures

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
L = ffi.load()
h, info = fixtures.sha256_chain_into_new_handle(1, lr, blocks, rank, world)
stream = torch.cuda.Stream(device=lr)
assert L.bp_cs_set_stream(h, ctypes.c_void_p(stream.cuda_stream)) == 0
g = ctypes.c_void_p()
idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{lr}")
if world > 1:
    if rank == 0:
        raw = (ctypes.c_uint8 * 128)()
        assert L.bp_group_unique_id(raw) == 0
        idt.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(idt, 0)
raw = (ctypes.c_uint8 * 128)(*idt.cpu().tolist())
assert L.bp_group_init(h, raw if world > 1 else None, rank, world, ctypes.byref(g)) == 0, L.bp_cs_last_error(h)
res = torch.zeros(1, dtype=torch.int64, device=f"cuda:{lr}")
rp = ctypes.c_void_p(res.data_ptr())


def timed(fn, n=200):
    with torch.cuda.stream(stream):
        for _ in range(10):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    return e0.elapsed_time(e1) / n * 1e3


out = {"rank": rank, "rows": info["rows_total"] if world == 1 else None,
       "check_only_us": timed(lambda: L.bp_cs_check_async(h, rp)),
       "exchange_only_us": timed(lambda: L.bp_group_reduce_async(g, rp)),
       "check_plus_exchange_us": timed(lambda: L.bp_group_check_async(g, rp))}
allv = [None] * world
if world > 1:
    dist.all_gather_object(allv, out)
else:
    allv = [out]
if rank == 0:
    print(json.dumps({"world": world, "blocks": blocks, "per_rank": allv}), flush=True)
L.bp_group_free(g)
info["tcs"].close()
if world > 1:
    dist.destroy_process_group()
