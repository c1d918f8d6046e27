#!/usr/bin/env python
"""bench.py -- R1CS constraints/sec of the (A.w) o (B.w) == C.w satisfaction check on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.  For N > 1 it is launched under torchrun (one rank per GPU).

A "step" is one pass of the hot path -- `which_is_unsatisfied` (test_cs.rs:239-253) over every constraint of the workload with
the current witness: one replay of the captured check graph (result init + the check kernels) and, for N > 1, the one-word MIN
exchange over the ranks' peer mailboxes (include/bp_r1cs.h: bp_group_check_async).

The top level of the line is BASELINE configs[1] (sha256 x4096, Pallas: the config the metric is quoted on).  `workloads` holds
the same measurements for the other configs BASELINE names at this N: configs[3] (synthetic 2^24, t = 6, BLS12-381 Fr: the
256-bit Montgomery path and the 1/2/4/8 scaling curve) always, configs[4] (synthetic 2^27, t = 32, Pallas) at N = 8, configs[2]
(blake2s, 64 KiB preimage, Vesta Fr: one GPU by definition) at N = 1.

Per workload:
  value        constraints/s, matrices and witness resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          the same through the C ABI with the witness in pinned HOST memory IN THE REFERENCE'S FORMAT (32-byte canonical
               scalars, what WitnessCS holds): host-side packing (bp_cs_recheck_scalars) or the raw upload, whichever the
               library path is, inside the timed region; `e2e.prepacked_bits` is the same when the caller already holds the
               witness bit-packed
  roofline     achieved bytes / measured HBM peak.  Gadget workloads: the bytes THIS layout reads (frac is a real fraction),
               with the canonical-CSR equivalent beside it; synthetic workloads: the canonical algorithmic bytes of SURVEY 8d
  selfcheck    untimed parity against the CPU oracle AT THIS SIZE (first-unsatisfied row after witness flips; A.w/B.w/C.w of
               sampled rows and of every shard boundary)
  cpu_baseline (N = 1) the CPU restatement of the reference's loop on a bounded sample: all host threads, and one thread
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5962BE3D763D318D
N_INPUTS = 16
SAT = 0x7FFFFFFFFFFFFFFF
HEADLINE = "sha256_chain_4096_pallas"

WORKLOADS = {
    # name: (kind, field, params) -- BASELINE.json configs
    "sha256_chain_4096_pallas": ("sha256", 1, {"blocks": 4096}),                       # configs[1] (metric config)
    "synthetic_2p24_t6_bls12_381": ("synthetic", 0, {"log_rows": 24, "t": 6}),        # configs[3] (scaling curve)
    "synthetic_2p27_t32_pallas": ("synthetic", 1, {"log_rows": 27, "t": 32}),         # configs[4] (8 GPUs)
    "synthetic_2p20_t6_bls12_381": ("synthetic", 0, {"log_rows": 20, "t": 6}),        # small, for quick checks
    "synthetic_2p22_t32_pallas": ("synthetic", 1, {"log_rows": 22, "t": 32}),         # small fat-LC instance
    "blake2s_64KiB_vesta": ("blake2s", 2, {"bytes": 65536}),                          # configs[2]
    "sha256_chain_64_pallas": ("sha256", 1, {"blocks": 64}),
    "sha256_chain_512_pallas": ("sha256", 1, {"blocks": 512}),
}

# dram__bytes_read.sum + dram__bytes_write.sum of ONE check (all its kernels), from the committed ncu pass of this very
# workload at this size on one GPU (profiles/): workload -> (bytes, file)
NCU_TRAFFIC = {
    "sha256_chain_512_pallas": (546_200_000, "profiles/r1_ncu_full_check_small_fat_int_sha256x512_final.csv"),
}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _fh:
        for _k, _v in json.load(_fh).items():
            NCU_TRAFFIC[_k] = (int(_v["bytes"]), _v["file"])
except Exception:
    pass

# What the memory system allows for the synthetic configs (profiles/r2_microbench2_*.jsonl, microbench2.cu `stream_gather`):
# the coefficient stream (36 B/term, coalesced) plus one random 32-byte witness gather per term WITHOUT any arithmetic, over a
# witness of this size.  The gathers miss the 126 MB L2 and HBM serves ~4.9e10 (512 MiB) / ~3.7e10 (4 GiB) random 32-byte
# accesses per second whatever the load width, hints or fetch granularity: that, not the byte count, is the floor.
GATHER_FLOOR_TERMS_PER_S = {512: 3.93e10, 4096: 3.32e10}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # median over the samples taken under load (an idle GPU parks at a few hundred MHz between the timed loops)
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the oracle on a bounded sample (cpu_baseline / --impl reference / the self-checks).  Never on the GPU path.
# ---------------------------------------------------------------------------------------------------------------------
class CpuSample:
    """A bounded sample of a workload held by the C oracle: the FIRST rows of the workload (same row numbering)."""

    def __init__(self, name, frac=0.1, max_rows_log2=22):
        from oracle import c_api

        kind, field, prm = WORKLOADS[name]
        self.name, self.kind, self.field = name, kind, field
        if kind == "synthetic":
            self.n_vars = 1 << prm["log_rows"]
            n = min(1 << max_rows_log2, 1 << prm["log_rows"])
            if self.n_vars > (1 << 25):  # a witness of GiBs: keep only what the sampled rows read (same values, same sums)
                n = min(n, 1 << 18)
                self.inst = c_api.synth_sparse_instance(field, SEED, prm["t"], self.n_vars, N_INPUTS, 0, n)
                self.what = f"first 2^{n.bit_length() - 1} rows of {name} (the witness elements they read)"
            else:
                lens, cols, coeffs = c_api.synth_rows(field, SEED, prm["t"], self.n_vars, N_INPUTS, 0, n)
                w = c_api.synth_witness(field, SEED, 0, self.n_vars)
                self.inst = c_api.Instance(field, lens, cols, coeffs, w[:N_INPUTS], w[N_INPUTS:])
                self.what = f"first 2^{n.bit_length() - 1} rows of {name} (full {self.n_vars}-element witness)"
            self.n_rows = n
            self.n_aux = None
        else:
            from bellpepper_b200 import fixtures

            if kind == "sha256":
                blocks = max(1, min(prm["blocks"], int(round(prm["blocks"] * frac))))
                lens, cols, coeffs, inputs, aux = fixtures.sha256_chain_host_csr(field, blocks)
                self.what = f"sha256 chain of {blocks} blocks: the same per-block circuit as {name}, {frac:.0%} of its blocks"
            else:
                nb = max(64, int(prm["bytes"] * frac) // 64 * 64)
                lens, cols, coeffs, inputs, aux = fixtures.blake2s_host_csr(field, nb)
                self.what = f"blake2s of the first {nb} bytes of {name}"
            self.n_rows = lens.size // 3
            self.n_aux = aux.shape[0]
            self.aux = aux
            self.inst = c_api.Instance(field, lens, cols, coeffs, inputs, aux)
            self.what += f" ({self.n_rows} rows)"

    def time(self, threads, budget_s=12.0, min_steps=1, max_steps=50):
        """Warm, then as many passes as fit the budget: (rows/s, ms per pass, passes)."""
        t0 = time.perf_counter()
        self.inst.check(threads, False)
        one = time.perf_counter() - t0
        steps = int(max(min_steps, min(max_steps, budget_s / max(one, 1e-6))))
        t0 = time.perf_counter()
        for _ in range(steps):
            self.inst.check(threads, False)
        dt = (time.perf_counter() - t0) / steps
        return self.n_rows / dt, dt * 1e3, steps


def host_threads():
    from oracle import lib

    return max(lib().bpo_max_threads(), len(os.sched_getaffinity(0)))


def cpu_baseline_block(sample: CpuSample, budget_s: float):
    threads = host_threads()
    par, par_ms, par_steps = sample.time(threads, budget_s)
    seq, seq_ms, seq_steps = sample.time(1, budget_s)
    return {
        "value": par, "unit": "constraints/s", "cores": threads, "kind": "port", "sample": sample.what, "ms": par_ms, "passes": par_steps,
        "what": "ref_par: C restatement of test_cs.rs:137-155,239-253 (always multiplies, like eval_lc), OpenMP over rows on every "
                "host thread -- stands in for north_star's 'rayon path' (the reference itself has no threads); warm",
        "ref_seq": {"value": seq, "cores": 1, "ms": seq_ms, "passes": seq_steps,
                    "what": "the reference's own shape: ONE thread, sequential over rows, early exit disabled so that it does the same work"},
    }


# ---------------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def build_workload(ctx, name):
    """Create a handle holding this rank's row shard of the workload (contiguous rows, balanced by terms)."""
    L, ffi = ctx.L, ctx.ffi
    kind, field, prm = WORKLOADS[name]
    rank, world, device = ctx.rank, ctx.world, ctx.local_rank
    h = ffi.vp()
    info = {"field": field, "kind": kind}
    t0 = time.time()
    if kind == "synthetic":
        from bellpepper_b200.sharding import split_range

        n_rows_total = 1 << prm["log_rows"]
        n_vars = n_rows_total
        t = prm["t"]
        r0, r1 = split_range(n_rows_total, rank, world)  # (row lengths are i.i.d.: equal rows == equal terms to 0.01 %)
        n = r1 - r0
        rc = L.bp_cs_new(field, device, n, int(n * 3 * t * 1.01) + 4096, n_vars, ctypes.byref(h))
        assert rc == 0, f"bp_cs_new -> {rc} (no CUDA device? there is no CPU path)"
        assert L.bp_cs_synth_witness(h, SEED, n_vars, N_INPUTS) == 0, L.bp_cs_last_error(h)
        CH = 1 << 20
        for s in range(r0, r1, CH):
            rc = L.bp_cs_synth_rows(h, SEED, t, n_vars, N_INPUTS, s, min(CH, r1 - s))
            assert rc == 0, L.bp_cs_last_error(h)
        assert L.bp_cs_set_row_base(h, r0) == 0
        info.update(rows_total=n_rows_total, row0=r0, t=t, n_vars=n_vars)
    elif kind == "sha256":
        from bellpepper_b200 import fixtures

        h, finfo = fixtures.sha256_chain_into_new_handle(field, device, prm["blocks"], rank, world, record_witness_program=True)
        info.update(finfo)
    elif kind == "blake2s":
        from bellpepper_b200 import fixtures

        assert world == 1, "blake2s workload: single shard"
        h, finfo = fixtures.blake2s_into_new_handle(field, device, prm["bytes"])
        info.update(finfo)
    else:
        raise ValueError(kind)
    assert L.bp_cs_sync(h) == 0
    info["ingest_s"] = round(time.time() - t0, 3)
    c = [ctypes.c_uint64() for _ in range(4)]
    assert L.bp_cs_counts(h, *[ctypes.byref(x) for x in c]) == 0
    info["n_inputs"], info["n_aux"], info["rows"], info["nnz"] = [x.value for x in c]
    return h, info


def make_group(ctx, h):
    """Join this rank's handle to the library's group (NCCL id handed round by torch.distributed; world 1: no NCCL)."""
    L, torch = ctx.L, ctx.torch
    g = ctypes.c_void_p()
    if ctx.world == 1:
        assert L.bp_group_init(h, None, 0, 1, ctypes.byref(g)) == 0, L.bp_cs_last_error(h)
        return g
    idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{ctx.local_rank}")
    if ctx.rank == 0:
        raw = (ctypes.c_uint8 * 128)()
        assert L.bp_group_unique_id(raw) == 0, "bp_group_unique_id: NCCL not found"
        idt.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    ctx.dist.broadcast(idt, 0)
    raw = (ctypes.c_uint8 * 128)(*idt.cpu().tolist())
    assert L.bp_group_init(h, raw, ctx.rank, ctx.world, ctypes.byref(g)) == 0, L.bp_cs_last_error(h)
    return g


def opt(L, h, key):
    v = ctypes.c_int64()
    assert L.bp_cs_get_option(h, key.encode(), ctypes.byref(v)) == 0, key
    return v.value


def time_mont_leg(L, h, torch, field, w_in, w_aux, n_rows_total, first_bad, time_e2e):
    """e2e with the witness in the in-memory form of blstrs::Scalar / pasta_curves::{Fp,Fq}: 4 x u64 limbs of x * 2^256 mod p."""
    import numpy as np

    from bellpepper_b200.fields import MODULUS

    p = MODULUS[field]
    r_limbs = torch.from_numpy(np.array([((1 << 256) % p >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], np.uint64).view(np.int64))

    def to_mont(w):
        hi0 = (w[:, 1:] == 0).all(dim=1)
        is_one = hi0 & (w[:, 0] == 1)
        other = ~(is_one | (hi0 & (w[:, 0] == 0)))
        idx = torch.nonzero(other).flatten().tolist()
        if len(idx) > 20000:
            raise RuntimeError("not a bit witness: the Montgomery leg converts only a few values in Python")
        m = torch.zeros_like(w)
        m[is_one] = r_limbs
        wn = w.numpy().view(np.uint64)
        for i in idx:
            v = sum(int(wn[i, j]) << (64 * j) for j in range(4)) * (1 << 256) % p
            m[i] = torch.from_numpy(np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], np.uint64).view(np.int64))
        return m.pin_memory()

    m_in, m_aux = to_mont(w_in), to_mont(w_aux)
    row = ctypes.c_int64()
    pm_in, pm_aux = ctypes.c_void_p(m_in.data_ptr()), ctypes.c_void_p(m_aux.data_ptr())

    def step():
        rc = L.bp_cs_recheck_scalars_mont(h, pm_in, pm_aux, ctypes.byref(row))
        if rc != 0:
            raise RuntimeError(f"bp_cs_recheck_scalars_mont -> {rc}: {L.bp_cs_last_error(h)}")
        return row.value

    s = time_e2e(step)
    same = row.value == (-1 if first_bad == SAT else first_bad)
    up = opt(L, h, "recheck_upload_bytes")
    return {"value": n_rows_total / s, "ms_per_step": s * 1e3, "h2d_bytes_per_step": (up + 7) // 8, "same_verdict": bool(same),
            "what": "32-byte scalars in their in-memory Montgomery form (what Vec<blstrs::Scalar> / Vec<pasta_curves::Fq> holds) in pinned "
                    "host memory -> bp_cs_recheck_scalars_mont: the packing pass matches the limb patterns of 0 and 2^256 mod p, "
                    "exceptions converted on the host; then as bp_cs_recheck_scalars"}


def measure(ctx, name, a, headline):
    """Everything bench.py reports for one workload at this N.  Returns the dict on rank 0 (None elsewhere)."""
    import numpy as np

    L, torch, dist = ctx.L, ctx.torch, ctx.dist
    rank, world, dev = ctx.rank, ctx.world, ctx.local_rank
    kind, field, prm = WORKLOADS[name]
    h, info = build_workload(ctx, name)
    stream = ctx.stream
    assert L.bp_cs_set_stream(h, ctypes.c_void_p(stream.cuda_stream)) == 0
    if a.fat_terms is not None:
        assert L.bp_cs_set_option(h, b"fat_terms", a.fat_terms) == 0
    if a.no_graph:
        assert L.bp_cs_set_option(h, b"graph", 0) == 0
    g = make_group(ctx, h)
    transport = ctypes.c_int()
    assert L.bp_group_info(g, None, None, ctypes.byref(transport)) == 0
    result = torch.zeros(1, dtype=torch.int64, device=f"cuda:{dev}")
    rptr = ctypes.c_void_p(result.data_ptr())
    n_rows_total = info.get("rows_total", info["rows"])
    n_vars = info["n_inputs"] + info["n_aux"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        rc = L.bp_group_check_async(g, rptr)  # check + (N > 1) the MIN exchange: one graph launch
        assert rc == 0, L.bp_cs_last_error(h)

    def group_row():
        """Synchronous which_is_unsatisfied over all shards: global row or -1."""
        step_device()
        v = int(result.item())
        return -1 if v == SAT else v

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {}
    with torch.cuda.stream(stream):
        # ---- value: resident data, device-timed -----------------------------------------------------------------------
        for _ in range(a.warmup):
            step_device()
        barrier()
        sampler = ClockSampler(dev)
        if rank == 0:
            sampler.start()
        l0 = opt(L, h, "launches")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            step_device()
        e1.record(stream)
        barrier()
        n_launch = opt(L, h, "launches") - l0
        ms_step = max_over_ranks(e0.elapsed_time(e1) / a.steps)
        first_bad = int(result.item())

        # ---- e2e: the witness comes from pinned HOST memory every step, the result goes back to the host ----------------
        n_aux = info["n_aux"]
        s0, s1 = rank * n_aux // world, (rank + 1) * n_aux // world  # the slice of the aux witness this rank uploads when sharded
        # synthetic workloads at N > 1 only ever touch this rank's slice on the host (GiBs per rank otherwise)
        a0, a1 = (s0, s1) if (kind == "synthetic" and world > 1) else (0, n_aux)
        w_in = torch.empty((info["n_inputs"], 4), dtype=torch.int64).pin_memory()
        w_aux = torch.empty((a1 - a0, 4), dtype=torch.int64).pin_memory()
        assert L.bp_cs_witness(h, 0, 0, info["n_inputs"], ctypes.c_void_p(w_in.data_ptr())) == 0
        assert L.bp_cs_witness(h, 1, a0, a1 - a0, ctypes.c_void_p(w_aux.data_ptr())) == 0
        row = ctypes.c_int64()
        p_in, p_aux = ctypes.c_void_p(w_in.data_ptr()), ctypes.c_void_p(w_aux.data_ptr())

        def time_e2e(step, n=None):
            n = n or max(3, min(a.steps, 8))
            step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                step()
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / n)

        def finish_group():
            assert L.bp_group_reduce_async(g, rptr) == 0, L.bp_cs_last_error(h)
            return int(result.item())

        # (1) raw 32-byte elements: every PCIe link carries the whole witness (N = 1) or 1/N of it + an NVLink all-gather
        def step_raw():
            if world == 1:
                assert L.bp_cs_set_range(h, 0, 0, info["n_inputs"], p_in) == 0, L.bp_cs_last_error(h)
                assert L.bp_cs_set_range(h, 1, 0, n_aux, p_aux) == 0, L.bp_cs_last_error(h)
                assert L.bp_cs_first_unsatisfied(h, ctypes.byref(row)) == 0, L.bp_cs_last_error(h)
                return row.value
            assert L.bp_cs_set_range(h, 0, 0, info["n_inputs"], p_in) == 0, L.bp_cs_last_error(h)
            assert L.bp_group_set_witness_sharded(g, 1, ctypes.c_void_p(w_aux.data_ptr() + 32 * (s0 - a0))) == 0, L.bp_cs_last_error(h)
            return group_row()

        raw_s = time_e2e(step_raw, 3 if n_vars > (1 << 26) else None)
        raw_h2d = (n_vars * 32) if world == 1 else (info["n_inputs"] + (s1 - s0)) * 32
        e2e_raw = {"value": n_rows_total / raw_s, "ms_per_step": raw_s * 1e3, "h2d_bytes_per_step": raw_h2d,
                   "what": ("32-byte canonical scalars, bp_cs_set_range -> check -> result" if world == 1 else
                            "32-byte canonical scalars: every rank uploads 1/N of the aux witness over its own PCIe link, NVLink "
                            "all-gather (bp_group_set_witness_sharded), validation, check + exchange; bytes are one rank's")}
        e2e_main, e2e_bits, e2e_generated = e2e_raw, None, None
        if kind != "synthetic":
            # (2) same scalars through the library's host-side packer (bits + exception list); row shards pack only what they read
            if world > 1:
                assert L.bp_cs_set_option(h, b"sparse_upload", 1) == 0

            def step_scalars():
                if world == 1:
                    assert L.bp_cs_recheck_scalars(h, p_in, p_aux, ctypes.byref(row)) == 0, L.bp_cs_last_error(h)
                    return row.value
                assert L.bp_cs_recheck_scalars_async(h, p_in, p_aux, rptr) == 0, L.bp_cs_last_error(h)
                return finish_group()

            sc_s = time_e2e(step_scalars)
            up = opt(L, h, "recheck_upload_bytes")
            e2e_scalars = {"value": n_rows_total / sc_s, "ms_per_step": sc_s * 1e3, "h2d_bytes_per_step": (up + 7) // 8,
                           "host_bytes_read_per_step": up * 32,
                           "packer": {"kernel": L.bp_pack_kernel().decode(), "threads": int(os.environ.get("BP_PACK_THREADS", "0")),
                                      "host_GBps": up * 32 / sc_s / 1e9},
                           "what": "32-byte canonical scalars in pinned host memory -> bp_cs_recheck_scalars: packed on the host (one bit "
                                   "per 0/1 value + exception list, all host threads), H2D of the bits, widened into the witness shadows, "
                                   "check" + ("; each rank packs and sends only the chunks its row shard reads, then the exchange" if world > 1 else "")}
            e2e_main = dict(e2e_scalars if sc_s < raw_s else e2e_raw)  # the faster of the two reference-format paths is the headline
            e2e_main["other_reference_format_path"] = e2e_raw if sc_s < raw_s else e2e_scalars
            # (2b) the same witness as it sits IN MEMORY in the reference's Vec<Scalar>: Montgomery limbs (x * 2^256 mod p), no
            # to_repr() pass on the caller's side (bp_cs_recheck_scalars_mont); informational, single rank
            if world == 1:
                try:
                    e2e_main["montgomery_in_memory_form"] = time_mont_leg(L, h, torch, field, w_in, w_aux, n_rows_total, first_bad, time_e2e)
                except Exception as e:  # never lose the line over an informational leg
                    e2e_main["montgomery_in_memory_form"] = {"skipped": str(e)[:200]}
            # (3) the caller already holds the witness bit-packed
            b_in = (w_in[:, 0] & 1).to(torch.uint8)
            b_aux = (w_aux[:, 0] & 1).to(torch.uint8)
            all_bits = bool((w_in[:, 1:] == 0).all() and (w_aux[:, 1:] == 0).all() and (w_in[:, 0] == b_in).all() and (w_aux[:, 0] == b_aux).all())
            if all_bits:
                q_in = torch.from_numpy(np.packbits(b_in.numpy(), bitorder="little")).pin_memory()
                q_aux = torch.from_numpy(np.packbits(b_aux.numpy(), bitorder="little")).pin_memory()

                def step_bits():
                    if world == 1:
                        assert L.bp_cs_recheck_bits(h, ctypes.c_void_p(q_in.data_ptr()), ctypes.c_void_p(q_aux.data_ptr()), ctypes.byref(row)) == 0, \
                            L.bp_cs_last_error(h)
                        return row.value
                    assert L.bp_cs_recheck_bits_async(h, ctypes.c_void_p(q_in.data_ptr()), ctypes.c_void_p(q_aux.data_ptr()), rptr) == 0, \
                        L.bp_cs_last_error(h)
                    return finish_group()

                bits_s = time_e2e(step_bits)
                e2e_bits = {"value": n_rows_total / bits_s, "ms_per_step": bits_s * 1e3, "h2d_bytes_per_step": (up + 7) // 8,
                            "what": "witness ALREADY bit-packed by the caller (1 bit per value, not the reference's format) -> "
                                    "bp_cs_recheck_bits: H2D, widened into the shadows, check"}
            if world > 1:
                assert L.bp_cs_set_option(h, b"sparse_upload", 0) == 0
            # (4) no witness upload at all: the device generates it from the MESSAGE (64 bytes per block) with the program the
            # front-end recorded at synthesis; the host computes the 8-word chaining state per block (plain SHA-256)
            if kind == "sha256" and info.get("witness_program") is not None:
                from bellpepper_b200 import fixtures

                prog = info["witness_program"]
                assert L.bp_cs_set_witness_program(h, ctypes.c_void_p(prog.ctypes.data), prog.size) == 0, L.bp_cs_last_error(h)
                msg = fixtures.chain_message(prm["blocks"])

                def step_generated():
                    st = fixtures.sha256_chain_states(msg)
                    assert L.bp_cs_generate_witness_async(h, msg, len(msg), ctypes.c_void_p(st.ctypes.data), st.size) == 0, L.bp_cs_last_error(h)
                    if world == 1:
                        assert L.bp_cs_first_unsatisfied(h, ctypes.byref(row)) == 0, L.bp_cs_last_error(h)
                        return row.value
                    return group_row()

                gen_s = time_e2e(step_generated)
                assert step_generated() == -1, "the device-generated witness does not satisfy the circuit"
                gen = torch.empty((n_aux, 4), dtype=torch.int64)
                assert L.bp_cs_witness(h, 1, 0, n_aux, ctypes.c_void_p(gen.data_ptr())) == 0
                e2e_generated = {"value": n_rows_total / gen_s, "ms_per_step": gen_s * 1e3, "h2d_bytes_per_step": len(msg) + 32 * prm["blocks"],
                                 "equals_front_end_witness": bool((gen == w_aux).all()), "program_bytes": int(prog.size) * 4,
                                 "what": "witness GENERATED ON THE DEVICE from the message bytes (bp_cs_generate_witness_async: the program "
                                         "the front-end recorded at synthesis, one warp per compression block), chaining states by plain "
                                         "SHA-256 on the host, then the check; no witness crosses PCIe"}
                del gen
            # leave the device witness in its full form again for what follows
            assert L.bp_cs_set_range(h, 0, 0, info["n_inputs"], p_in) == 0
            assert L.bp_cs_set_range(h, 1, 0, n_aux, p_aux) == 0

        # ---- K2, informational: batched LinearCombination::eval (canonical A.w, B.w, C.w of every row, device to device) --
        eval_info, outs = None, None
        try:
            n_loc = info["rows"]
            outs = [torch.empty((n_loc, 4), dtype=torch.int64, device=f"cuda:{dev}") for _ in range(3)]
            ptrs = [ctypes.c_void_p(t.data_ptr()) for t in outs]
            for _ in range(2):
                assert L.bp_cs_eval_async(h, *ptrs) == 0, L.bp_cs_last_error(h)
            barrier()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record(stream)
            for _ in range(3):
                L.bp_cs_eval_async(h, *ptrs)
            v1.record(stream)
            barrier()
            ev_ms = v0.elapsed_time(v1) / 3
            eval_info = {"ms_per_pass": ev_ms, "rows_per_s_this_rank": n_loc / (ev_ms * 1e-3), "out_bytes": 96 * n_loc,
                         "what": "bp_cs_eval_async: canonical A.w, B.w, C.w of every row of this rank's shard, device to device"}
        except Exception as e:  # e.g. not enough memory for the three output vectors
            eval_info = {"skipped": str(e)[:200]}
            outs = None
        clocks = sampler.stop() if rank == 0 else None

        # ---- self-check against the CPU oracle at this size (untimed) -----------------------------------------------------
        selfcheck = None if a.no_selfcheck else oracle_selfcheck(ctx, name, a, h, info, group_row, outs, first_bad)
        del outs

    plan = {}
    for key in ("plain_rows", "generic_rows", "fat_rows", "deferred_rows", "fat_undecided_rows", "plain_row_terms",
                "generic_row_terms", "fat_row_terms", "graph_replays", "graph_captures"):
        plan[key] = opt(L, h, key)
    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        alg_bytes = info["nnz"] * 36 + info["rows"] * 12 + n_vars * 32  # this rank's shard + the whole witness
        canonical = alg_bytes / (ms_step * 1e-3) / 1e9
        small_path = plan["plain_rows"] > 0
        roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_kind": f"of {peak_kind} (MEASURED_PEAKS.json hbm_gbs)"}
        if small_path:
            # bytes the kernels have to read in THIS layout when every operand is small: per plain-row term its 4-byte term
            # word, per plain row 4 bytes of row_meta (+ 8 per 64 rows of range), per fat-row term 4 + 2 bytes (term word,
            # exponents) and per fat row 20 bytes (row list, row_ptr); generic rows the canonical 36 B/term + 12 B/row; every
            # variable's 4-byte shadow once
            laid_out = (plan["plain_row_terms"] * 4 + plan["plain_rows"] * 4 + (info["rows"] // 64 + 1) * 8 + plan["fat_row_terms"] * 6
                        + plan["fat_rows"] * 20 + plan["generic_row_terms"] * 36 + plan["generic_rows"] * 12 + n_vars * 4)
            ach = laid_out / (ms_step * 1e-3) / 1e9
            tr = NCU_TRAFFIC.get(name) if world == 1 else None
            roof.update({
                "achieved": ach, "frac": ach / peak, "bytes_per_launch": laid_out,
                "bytes_kind": "laid out: what this layout makes the kernels read (4-byte term words and witness shadows; coefficient "
                              "bytes of +-1/+-2/+-2^k terms are never read)",
                "traffic": tr[0] if tr else None, "traffic_source": tr[1] if tr else None,
                "canonical_equivalent": {"bytes_per_launch": alg_bytes, "GBps": canonical, "ratio_to_peak": canonical / peak,
                                         "note": "SURVEY 8d's canonical CSR bytes (36 B/term + 12 B/row + 32 B/variable) over the same "
                                                 "time: above 1 means 'faster than streaming the canonical CSR once', not a fraction"},
                "bound_by": "issue slots (integer kernels, ~7 instructions per term), not HBM: profiles/",
            })
            dtype = (f"i64 on 4-byte witness shadows (exact integer shortcut, |X| < p) for {plan['plain_rows']} plain + {plan['fat_rows']} fat "
                     f"rows; u256 (8 x u32 limbs, lazy Montgomery mod p) for {plan['generic_rows']} generic + {plan['deferred_rows']} deferred + "
                     f"{plan['fat_undecided_rows']} undecided rows")
        else:
            wit_mib = n_vars * 32 >> 20
            floor_rate = GATHER_FLOOR_TERMS_PER_S.get(512 if wit_mib <= 512 else 4096)
            floor_ms = info["nnz"] / floor_rate * 1e3
            roof.update({
                "achieved": canonical, "frac": canonical / peak, "bytes_per_launch": alg_bytes,
                "bytes_kind": "canonical algorithmic bytes (SURVEY 8d): 36 B/term + 12 B/row + 32 B/variable, this rank's rows + the whole witness",
                "traffic": (NCU_TRAFFIC.get(name) or (None, None))[0] if world == 1 else None,
                "traffic_source": (NCU_TRAFFIC.get(name) or (None, None))[1] if world == 1 else None,
                "gather_floor": {"ms": floor_ms, "frac_at_floor": alg_bytes / (floor_ms * 1e-3) / 1e9 / peak, "of_floor": floor_ms / ms_step,
                                 "note": "stream + one random 32-byte gather per term with NO arithmetic (microbench2 stream_gather, "
                                         "profiles/r2_microbench2_*.jsonl) over a witness of this size: the memory system's floor for this "
                                         "access pattern; the random gathers miss L2 and HBM serves ~4e10 of them per second"},
            })
            dtype = "u256 (8 x u32 limbs, lazy-reduction Montgomery mod p): every row takes the full-width kernels"
        out = {
            "value": n_rows_total / (ms_step * 1e-3), "ms_per_step": ms_step, "dtype": dtype,
            "e2e": {"value": e2e_main["value"], "unit": "constraints/s", "ms_per_step": e2e_main["ms_per_step"],
                    "h2d_bytes_per_step": e2e_main["h2d_bytes_per_step"], "d2h_bytes_per_step": 12 if world == 1 else 8,
                    "format": "reference (32-byte canonical scalars in pinned host memory); packing, copies and the result read-back are "
                              "inside the timed region",
                    "what": e2e_main["what"],
                    **({"other_reference_format_path": e2e_main["other_reference_format_path"]} if "other_reference_format_path" in e2e_main else {}),
                    **({"montgomery_in_memory_form": e2e_main["montgomery_in_memory_form"]} if "montgomery_in_memory_form" in e2e_main else {}),
                    **({"prepacked_bits": e2e_bits} if e2e_bits else {}),
                    **({"generated_on_device": e2e_generated} if e2e_generated else {})},
            "gpu_launches": n_launch,
            "launch_mechanism": f"{plan['graph_replays']} graph replays / {plan['graph_captures']} captures so far; transport "
                                + {0: "single rank", 1: "NCCL all-reduce", 2: "peer-memory mailboxes (no NCCL launch per step)"}[transport.value],
            "roofline": {**roof, **{k: plan[k] for k in plan if not k.startswith("graph")}},
            "clocks": clocks,
            "first_unsatisfied_row": None if first_bad == SAT else first_bad,
            "selfcheck": selfcheck,
            "eval_all_rows": eval_info,
            "instance": {k: info[k] for k in ("rows", "nnz", "n_inputs", "n_aux", "ingest_s")},
            "config": {"workload": name, "field": ctx.field_name[field], **prm, "l2_policy": "inputs_larger_than_l2",
                       "parallelism": f"row-sharded x{world} (contiguous rows balanced by terms), witness replicated"},
        }
        if world > 1 and kind == "synthetic":
            per_gpu = (info["nnz"] * 36 + info["rows"] * 12) + n_vars * 32
            one_gpu = (info["nnz"] * 36 + info["rows"] * 12) * world + n_vars * 32
            out["scaling_ceiling"] = {"max_speedup_over_1gpu": one_gpu / per_gpu,
                                      "why": "the witness is replicated: per-GPU bytes = rows/N * (36T + 12) + 32n (SURVEY 8d)"}
        if world == 1 and not a.no_cpu_baseline:
            sample = ctx.samples.get(name) or CpuSample(name, a.cpu_sample_frac)
            out["cpu_baseline"] = cpu_baseline_block(sample, a.cpu_budget_s)
    L.bp_group_free(g)
    if "tcs" in info:
        info["tcs"].close()
    else:
        L.bp_cs_free(h)
    ctx.samples.pop(name, None)
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def oracle_selfcheck(ctx, name, a, h, info, group_row, outs, first_bad):
    """Parity against the CPU oracle at the workload's full size (untimed).  Gadget circuits: the first-unsatisfied GLOBAL row
    after a witness flip, early in the chain (oracle on the first blocks) and in the last block (oracle on the last blocks,
    full witness).  Synthetic: A.w, B.w, C.w of this rank's first rows, its last rows (both sides of every shard boundary)
    and a block in the middle against the oracle's, and the first-unsatisfied row."""
    import numpy as np

    from oracle import c_api

    L, torch, dist = ctx.L, ctx.torch, ctx.dist
    rank, world, dev = ctx.rank, ctx.world, ctx.local_rank
    kind, field, prm = WORKLOADS[name]
    res = {"oracle_equal": None}

    def all_ranks_ok(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def set_aux(idx, limbs):
        v = np.asarray(limbs, np.uint64)
        assert L.bp_cs_set(h, 1, idx, ctypes.c_void_p(v.ctypes.data)) == 0, L.bp_cs_last_error(h)

    if kind in ("sha256", "blake2s"):
        if first_bad != SAT:
            return {"oracle_equal": False, "why": f"the honest witness does not satisfy the circuit (row {first_bad})"}
        # ONE host recording of the real circuit (g++ front-end) gives the oracle the rows of a few blocks at the start, in the
        # middle and at the end of the chain under their GLOBAL row numbers, with the whole witness; a witness bit inside each
        # kept range is flipped on every rank's replica and the first unsatisfied global row must be the oracle's.
        checks, t0 = [], time.time()
        n_aux = info["n_aux"]
        if kind == "sha256":
            blocks = prm["blocks"]
            n_bits = 8 * (64 * blocks - 9)
            per_block = (n_aux - n_bits) / blocks
            mid = max(2, blocks // 2)
            ranges = sorted({(0, min(2, blocks)), (mid, min(mid + 2, blocks)), (max(0, blocks - 2), blocks)}) if blocks >= 6 else [(0, blocks)]
            victims = [int(n_bits + per_block * (b0 + 0.5)) for b0, _ in ranges] if blocks >= 6 else [n_aux // 3, (2 * n_aux) // 3]
            victims[-1] = n_aux - 4321 if blocks >= 6 else victims[-1]
        else:
            ranges, victims = None, [n_aux // 3, n_aux - 777]
        oracle = None
        if rank == 0:
            from bellpepper_b200 import fixtures

            with fixtures.Tcs(field, device=-1, named=False) as rec:
                if kind == "sha256":
                    _, where = rec.sha256_ranges(fixtures.chain_message(blocks), ranges)
                else:
                    rec.blake2s(fixtures.xorshift_bytes(prm["bytes"]))
                    where = [(0, 0)]
                lens, cols, coeffs, inputs, aux = rec.host_csr()
            oracle = c_api.Instance(field, lens, cols, coeffs, inputs, aux)
            n_kept = lens.size // 3
            local_starts = [l for _, l in where] + [n_kept]

            def to_global(local_row):
                if local_row < 0:
                    return -1
                for i, (g, l) in enumerate(where):
                    if l <= local_row < local_starts[i + 1]:
                        return g + (local_row - l)
                raise AssertionError(local_row)
        for victim in victims:
            old = np.zeros(4, np.uint64)
            assert L.bp_cs_get(h, 1, victim, ctypes.c_void_p(old.ctypes.data)) == 0
            new = [1 - int(old[0]), 0, 0, 0] if int(old[0]) in (0, 1) and not old[1:].any() else [int(old[0]) ^ 1, int(old[1]), int(old[2]), int(old[3])]
            set_aux(victim, new)
            got = group_row()
            set_aux(victim, old)
            back = group_row()
            want = None
            if rank == 0:
                oracle.set(True, victim, c_api.limbs_to_ints(np.asarray([new], np.uint64))[0])
                want = to_global(oracle.check(host_threads(), False))
                oracle.set(True, victim, c_api.limbs_to_ints(old.reshape(1, 4))[0])
            checks.append({"flipped_aux": victim, "gpu_first_unsatisfied": got, "oracle_first_unsatisfied": want, "holds_after_restore": back == -1})
        ok = True
        if rank == 0:
            ok = all(c["gpu_first_unsatisfied"] == c["oracle_first_unsatisfied"] and c["gpu_first_unsatisfied"] >= 0 and c["holds_after_restore"]
                     for c in checks)
            oracle.close()
        res = {"oracle_equal": all_ranks_ok(ok), "checks": checks,
               "oracle": (f"rows of blocks {ranges} of the REAL {blocks}-block chain ({n_kept if rank == 0 else '?'} rows under their global numbers), "
                          "whole witness" if kind == "sha256" else "the whole circuit") + "; C oracle, first unsatisfied row",
               "oracle_s": round(time.time() - t0, 1)}
    else:
        n_vars, t = info["n_vars"], prm["t"]
        row0, n_loc = info["row0"], info["rows"]
        blk = min(4096, n_loc)
        ranges = sorted({0, max(0, n_loc // 2 - blk // 2), n_loc - blk})
        ok, compared = outs is not None, 0
        if outs is not None:
            for r in ranges:
                inst = c_api.synth_sparse_instance(field, SEED, t, n_vars, N_INPUTS, row0 + r, blk)
                _, az, bz, cz = inst.eval(host_threads())
                for dev_t, ref in zip(outs, (az, bz, cz)):
                    got = dev_t[r:r + blk].cpu().numpy().view(np.uint64)
                    ok = ok and bool((got == ref).all())
                compared += blk
                inst.close()
        big = None
        if rank == 0 and outs is not None:
            # the oracle sample (first rows, the real witness array when it fits): every row of it, and the verdict
            ctx.samples[name] = ctx.samples.get(name) or CpuSample(name, a.cpu_sample_frac)
            s = ctx.samples[name]
            n = min(s.n_rows, n_loc)
            bad, az, bz, cz = s.inst.eval(host_threads())
            eq = True
            for dev_t, ref in zip(outs, (az, bz, cz)):
                eq = eq and bool((dev_t[:n].cpu().numpy().view(np.uint64) == ref[:n]).all())
            big = {"rows": n, "equal": eq, "oracle_first_unsatisfied": bad, "gpu_first_unsatisfied": None if first_bad == SAT else first_bad}
            ok = ok and eq and (bad == (-1 if first_bad == SAT else first_bad) or bad >= n)
        res = {"oracle_equal": all_ranks_ok(ok), "rows_compared_per_rank": compared, "row_blocks": [row0 + r for r in ranges],
               "what": "canonical A.w, B.w, C.w (bp_cs_eval_async) of this rank's first, middle and last 4096 rows -- both sides of every "
                       "shard boundary -- bit for bit against the oracle; rank 0 also its first rows in bulk and the first-unsatisfied row",
               "bulk": big}
        if outs is None:
            res["why"] = "no memory for the emit buffers"
    return res


def reference_arm(a, base):
    """`--impl reference`: the CPU restatement of the reference's loop on the box's host cores (the reference is Rust; no
    toolchain in this image).  No CUDA library is mapped: circuits are recorded by libbp_frontend.so (g++ only)."""
    from bellpepper_b200.fields import NAME as FIELD_NAME

    out = dict(base)
    names = [a.workload] if a.workload != "default" else [HEADLINE, "synthetic_2p24_t6_bls12_381"]
    res = {}
    threads = host_threads()
    for name in names:
        s = CpuSample(name, a.cpu_sample_frac)
        for _ in range(min(a.warmup, 1)):
            s.inst.check(threads, False)
        t0 = time.perf_counter()
        steps = max(1, a.steps if s.kind == "synthetic" or a.steps <= 5 else a.steps)
        # every step = one pass over the bounded sample; cap the run at ~a minute per workload
        one = None
        done = 0
        for _ in range(steps):
            s.inst.check(threads, False)
            done += 1
            one = (time.perf_counter() - t0) / done
            if (time.perf_counter() - t0) > 60:
                break
        rate = s.n_rows / one
        seq, seq_ms, _ = s.time(1, budget_s=8.0)
        kind, field, prm = WORKLOADS[name]
        res[name] = {"value": rate, "ms_per_step": one * 1e3, "steps_run": done,
                     "cpu_baseline": {"value": rate, "unit": "constraints/s", "cores": threads, "kind": "port", "sample": s.what,
                                      "ref_seq": {"value": seq, "cores": 1, "ms": seq_ms}},
                     "config": {"workload": name, "field": FIELD_NAME[field], **prm}}
    head = res[names[0]]
    out.update({
        "impl": "reference", "dtype": "u256 (4 x u64 limbs, CIOS Montgomery, host CPU)", "value": head["value"], "ms_per_step": head["ms_per_step"], "cpu_baseline": head["cpu_baseline"],
        "e2e": {"value": head["value"], "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "workloads": {k: v for k, v in res.items() if k != names[0]},
        "note": "reference is Rust (no toolchain in this image): C restatement of test_cs.rs:137-155,239-253, OpenMP over rows on "
                "every host thread (the reference itself is single-threaded: ref_seq); each step checks the bounded sample named in "
                "cpu_baseline.sample; circuits recorded by the g++-only front-end (no CUDA library in this process)",
    })
    out["config"] = {**base["config"], **head["config"]}
    print(json.dumps(out), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("BP_BENCH_WORKLOAD", "default"),
                    help="default = configs[1] at the top level + the other BASELINE configs under `workloads`; a name = only that one")
    ap.add_argument("--cpu-sample-frac", type=float, default=0.1, help="gadget workloads: fraction of the blocks the CPU sample holds")
    ap.add_argument("--cpu-budget-s", type=float, default=8.0, help="seconds of timed CPU work per baseline leg")
    ap.add_argument("--fat-terms", type=int, default=None, help="rows with more terms than this use the warp-per-row kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the oracle self-check (the gadget one re-synthesizes the chain on the host, ~25 s)")
    ap.add_argument("--no-graph", action="store_true", help="plain kernel launches instead of the captured graph")
    ap.add_argument("--no-extra", action="store_true", help="only the headline workload")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the library's host-side packer (bp_cs_recheck_scalars) uses every hardware thread by default: share them between the ranks
    os.environ.setdefault("BP_PACK_THREADS", str(max(1, len(os.sched_getaffinity(0)) // max(1, world))))
    from bellpepper_b200.fields import NAME as FIELD_NAME

    head_name = HEADLINE if a.workload == "default" else a.workload
    kind, field, prm = WORKLOADS[head_name]
    base = {
        "metric": "R1CS constraints/sec (256-bit Fp SpMV x3 + Hadamard check)",
        "unit": "constraints/s",
        "n_gpus": a.gpus,
        "steps": a.steps,
        "warmup": a.warmup,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "data": "synthetic",
        "config": {"workload": head_name, "field": FIELD_NAME[field], **prm, "l2_policy": "inputs_larger_than_l2",
                   "parallelism": f"row-sharded x{a.gpus}, witness replicated"},
    }
    if a.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(a, base)

    import torch
    import torch.distributed as dist

    from bellpepper_b200 import ffi

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Ctx()
    ctx.L, ctx.ffi, ctx.torch, ctx.dist = ffi.load(), ffi, torch, dist
    ctx.rank, ctx.local_rank, ctx.world = rank, local_rank, world
    ctx.stream = torch.cuda.Stream(device=local_rank)
    ctx.field_name = FIELD_NAME
    ctx.samples = {}

    names = [head_name]
    if a.workload == "default" and not a.no_extra:
        names.append("synthetic_2p24_t6_bls12_381")
        if world == 8:
            names.append("synthetic_2p27_t32_pallas")
    results = {}
    for n in names:
        results[n] = measure(ctx, n, a, n == head_name)
    if a.workload == "default" and not a.no_extra and world == 1:
        # configs[2] (blake2s, 64 KiB preimage, Vesta Fr, one GPU): last, and never at the cost of the line
        try:
            results["blake2s_64KiB_vesta"] = measure(ctx, "blake2s_64KiB_vesta", a, False)
        except Exception as e:
            results["blake2s_64KiB_vesta"] = {"error": str(e)[:300]}
    if rank == 0:
        head = results[head_name]
        out = dict(base)
        out.update(head)
        out["config"] = {**base["config"], **head["config"]}
        extra = {k: v for k, v in results.items() if k != head_name}
        if extra:
            out["workloads"] = extra
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
