"""GPU tests (-m gpu) of the round-2 entry points: batched set (K4), re-check from 32-byte scalars (host packing), the
captured-graph check, and the multi-GPU group (world 1 on any box; world 2 with peer mailboxes when the box has two GPUs).
Every verdict is compared with the CPU oracle."""
import ctypes
import os
import random
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import c_api, synth
from oracle.fields import FIELDS

pytestmark = pytest.mark.gpu

from gpu_util import Handle  # noqa: E402
from test_gpu_parity import _gadget_like_instance, _satisfiable  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAT = 0x7FFFFFFFFFFFFFFF


def test_set_many_matches_oracle_and_rejects_bad_batches():
    """K4 (test_cs.rs:270-282 batched; the set / re-check loop of num.rs:753-762)."""
    fid = 0
    p = FIELDS[fid].p
    lens, cols, coeffs, inputs, aux = _satisfiable(fid, 4, 1500, 2500)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    rng = random.Random(5)
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        assert h.first_unsatisfied() == -1
        for n in (1, 7, 300):
            idx = rng.sample(range(aux.shape[0]), n)
            new = [rng.randrange(p) for _ in idx]
            old = c_api.limbs_to_ints(aux[idx])
            ia = np.asarray(idx, np.uint64)
            va = c_api.ints_to_limbs(new)  # (kept alive across the call: `x.ctypes.data` of a temporary dangles)
            h.ok(h.L.bp_cs_set_many(h.h, 1, n, ia.ctypes.data, va.ctypes.data))
            for i, v in zip(idx, new):
                inst.set(True, i, v)
            assert h.first_unsatisfied() == inst.check(2, False) >= 0
            got = np.zeros((aux.shape[0], 4), np.uint64)
            h.ok(h.L.bp_cs_witness(h.h, 1, 0, aux.shape[0], got.ctypes.data))
            assert c_api.limbs_to_ints(got[idx]) == new
            va = c_api.ints_to_limbs(old)
            h.ok(h.L.bp_cs_set_many(h.h, 1, n, ia.ctypes.data, va.ctypes.data))
            for i, v in zip(idx, old):
                inst.set(True, i, v)
            assert h.first_unsatisfied() == -1
        # a batch with one bad entry changes nothing
        ia = np.asarray([3, aux.shape[0]], np.uint64)
        va = c_api.ints_to_limbs([1, 2])
        assert h.L.bp_cs_set_many(h.h, 1, 2, ia.ctypes.data, va.ctypes.data) == -3
        ia = np.asarray([3, 4], np.uint64)
        va = c_api.ints_to_limbs([1, p])
        assert h.L.bp_cs_set_many(h.h, 1, 2, ia.ctypes.data, va.ctypes.data) == -3
        assert h.first_unsatisfied() == -1
        # inputs too (ONE is an ordinary slot: test_cs.rs:160-169)
        ia = np.asarray([0], np.uint64)
        va = c_api.ints_to_limbs([2])
        h.ok(h.L.bp_cs_set_many(h.h, 0, 1, ia.ctypes.data, va.ctypes.data))
        inst.set(False, 0, 2)
        assert h.first_unsatisfied() == inst.check(2, False)


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_recheck_scalars_packs_bits_and_patches_exceptions(fid):
    """bp_cs_recheck_scalars: the reference's witness format (32-byte scalars) in, host-side packing, same verdicts as the
    oracle -- for an all-bit witness, with a few full-width / byte-sized exceptions, and for a witness that is not bits."""
    p = FIELDS[fid].p
    lens, cols, coeffs, inputs, aux, _ = _gadget_like_instance(fid, 77, 3000, 5000, 9)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    rng = random.Random(fid)
    row = ctypes.c_int64()
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        base = h.first_unsatisfied()
        assert base == inst.check(2, False)
        a = np.zeros_like(aux)
        a[:, 0] = aux[:, 0] & np.uint64(1)  # a bit witness (what a gadget circuit allocates), so that the packer is what runs
        i = inputs.copy()
        i[1:, 0] &= np.uint64(1)
        i[1:, 1:] = 0
        for trial in range(4):
            # new witness: flip some bits, plant some non-bit values (small and full-width)
            for _ in range(5):
                k = rng.randrange(a.shape[0])
                a[k] = c_api.ints_to_limbs([1 - int(a[k][0]) if int(a[k][0]) in (0, 1) and not a[k][1:].any() else 0])[0]
            if trial >= 1:
                for v in (2, 255, (1 << 24) + 1, p - 1, rng.randrange(p)):
                    a[rng.randrange(a.shape[0])] = c_api.ints_to_limbs([v])[0]
            if trial == 3:
                i[rng.randrange(1, i.shape[0])] = c_api.ints_to_limbs([rng.randrange(p)])[0]
            ref = c_api.Instance(fid, lens, cols, coeffs, i, a)
            h.ok(h.L.bp_cs_recheck_scalars(h.h, i.ctypes.data, a.ctypes.data, ctypes.byref(row)))
            assert row.value == ref.check(2, False)
            got = np.zeros_like(a)
            h.ok(h.L.bp_cs_witness(h.h, 1, 0, a.shape[0], got.ctypes.data))
            assert (got == a).all()
            got_i = np.zeros_like(i)
            h.ok(h.L.bp_cs_witness(h.h, 0, 0, i.shape[0], got_i.ctypes.data))
            assert (got_i == i).all()
            # A.w, B.w, C.w of every row too (the shadows and the 32-byte forms agree)
            _, az_r, bz_r, cz_r = ref.eval(2)
            az, bz, cz = h.eval(lens.size // 3)
            assert (az == az_r).all() and (bz == bz_r).all() and (cz == cz_r).all()
        # a non-canonical exception is refused
        a2 = a.copy()
        a2[10] = np.frombuffer(int(p).to_bytes(32, "little"), dtype="<u8")
        assert h.L.bp_cs_recheck_scalars(h.h, i.ctypes.data, a2.ctypes.data, ctypes.byref(row)) == -3
    # a full-width witness goes through as it is
    lens, cols, coeffs, inputs, aux = _satisfiable(fid, 4, 800, 1200)
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        h.ok(h.L.bp_cs_recheck_scalars(h.h, inputs.ctypes.data, aux.ctypes.data, ctypes.byref(row)))
        assert row.value == -1
        a = aux.copy()
        a[700] = c_api.ints_to_limbs([5])[0]
        h.ok(h.L.bp_cs_recheck_scalars(h.h, None, a.ctypes.data, ctypes.byref(row)))
        assert row.value == c_api.Instance(fid, lens, cols, coeffs, inputs, a).check(2, False) >= 0


def test_graph_replay_gives_the_same_verdicts():
    """A check is replayed as one captured CUDA graph while its arguments are unchanged; witness writes, option changes and
    structural changes in between must all be seen (re-capture or plain data dependence)."""
    import torch

    fid = 1
    lens, cols, coeffs, inputs, aux, _ = _gadget_like_instance(fid, 31, 4000, 6000, 11)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    out = torch.zeros(1, dtype=torch.int64, device="cuda:0")
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        assert h.opt("graph") == 1
        want = inst.check(2, False)
        for _ in range(4):
            assert h.first_unsatisfied() == want
        assert h.opt("graph_captures") == 1 and h.opt("graph_replays") == 3
        # a witness write between replays is plain data: no re-capture needed, new verdict
        k = aux.shape[0] // 2
        v = c_api.ints_to_limbs([(int(aux[k][0]) + 1) % 2 if not aux[k][1:].any() and int(aux[k][0]) < 2 else 0])
        h.ok(h.L.bp_cs_set(h.h, 1, k, v.ctypes.data))
        inst.set(True, k, c_api.limbs_to_ints(v)[0])
        assert h.first_unsatisfied() == inst.check(2, False)
        # another result word, another threshold, another variant: re-capture, same answers
        h.ok(h.L.bp_cs_check_async(h.h, ctypes.c_void_p(out.data_ptr())))
        h.ok(h.L.bp_cs_sync(h.h))
        w = inst.check(2, False)
        assert int(out.item()) == (SAT if w < 0 else w)
        for ft, var in ((8, -1), (96, 1), (96, 8)):
            h.opt("fat_terms", ft)
            h.opt("variant", var)
            assert h.first_unsatisfied() == w
            assert h.first_unsatisfied() == w
        # graph off == graph on
        h.opt("graph", 0)
        assert h.first_unsatisfied() == w
        h.opt("graph", 1)
        # more rows appended: the plan changes, the graph follows
        one = np.asarray([[1, 0, 0, 0]], np.uint64)
        l2 = np.asarray([1, 1, 0], np.uint32)
        c2 = np.asarray([0, 0], np.uint32)  # ONE * ONE = 0 : fails
        h.ok(h.L.bp_cs_enforce(h.h, 1, l2.ctypes.data, c2.ctypes.data, np.concatenate([one, one]).ctypes.data))
        n_rows = lens.size // 3
        assert h.first_unsatisfied() == (w if w >= 0 else n_rows)
        assert h.first_unsatisfied() == (w if w >= 0 else n_rows)


def test_split_rows_by_nnz():
    from bellpepper_b200 import ffi

    L = ffi.load()
    lens = np.asarray([1, 1, 1] * 10 + [100, 50, 50] + [0, 0, 0] * 5 + [2, 2, 2] * 20, np.uint32)
    n = lens.size // 3
    for world in (1, 2, 3, 8):
        b = np.zeros(world + 1, np.uint64)
        assert L.bp_split_rows_by_nnz(lens.ctypes.data, n, world, b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) == 0
        assert b[0] == 0 and b[-1] == n and (np.diff(b.astype(np.int64)) >= 0).all()
    b = np.zeros(3, np.uint64)
    L.bp_split_rows_by_nnz(lens.ctypes.data, n, 2, b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    w = lens.reshape(-1, 3).sum(1) + 1
    left = int(w[: int(b[1])].sum())
    assert abs(left - int(w.sum()) / 2) <= 201  # within one (fat) row of the middle


def test_group_of_one_rank():
    """world == 1: no NCCL needed, bp_group_check == bp_cs_first_unsatisfied on GLOBAL rows."""
    fid = 2
    lens, cols, coeffs, inputs, aux, _ = _gadget_like_instance(fid, 3, 1000, 2000, 5)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        h.ok(h.L.bp_cs_set_row_base(h.h, 500))
        g = ctypes.c_void_p()
        h.ok(h.L.bp_group_init(h.h, None, 0, 1, ctypes.byref(g)))
        try:
            row = ctypes.c_int64()
            want = inst.check(2, False)
            for _ in range(3):
                h.ok(h.L.bp_group_check(g, ctypes.byref(row)))
                assert row.value == (want + 500 if want >= 0 else -1)
            r, w, tr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            h.ok(h.L.bp_group_info(g, ctypes.byref(r), ctypes.byref(w), ctypes.byref(tr)))
            assert (r.value, w.value, tr.value) == (0, 1, 0)
            h.ok(h.L.bp_group_broadcast_witness(g, 0))
            h.ok(h.L.bp_group_set_witness_sharded(g, 1, aux.ctypes.data))
            h.ok(h.L.bp_group_check(g, ctypes.byref(row)))
            assert row.value == (want + 500 if want >= 0 else -1)
        finally:
            h.L.bp_group_free(g)


GROUP_WORKER = textwrap.dedent(
    """
    import ctypes, json, os, sys
    sys.path.insert(0, %r)
    sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, torch, torch.distributed as dist
    from bellpepper_b200 import ffi
    from oracle import c_api
    from gpu_util import Handle
    from test_gpu_parity import _satisfiable
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo")
    L = ffi.load()
    fid, n_rows = 0, 3000
    lens, cols, coeffs, inputs, aux = _satisfiable(fid, 4, 2000, n_rows)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    # row shard of this rank, balanced by terms
    b = np.zeros(world + 1, np.uint64)
    assert L.bp_split_rows_by_nnz(lens.ctypes.data, n_rows, world, b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) == 0
    r0, r1 = int(b[rank]), int(b[rank + 1])
    ptr = np.concatenate([[0], np.cumsum(lens.reshape(-1, 3).sum(1).astype(np.int64))])
    h = Handle(fid, device=rank)
    first = ctypes.c_uint64()
    h.ok(L.bp_cs_alloc(h.h, 0, inputs[1:].ctypes.data, inputs.shape[0] - 1, ctypes.byref(first)))
    # every rank but 0 starts from a WRONG witness: the broadcast must fix it
    a0 = aux if rank == 0 else np.zeros_like(aux)
    h.ok(L.bp_cs_alloc(h.h, 1, a0.ctypes.data, aux.shape[0], ctypes.byref(first)))
    sl = lens[3 * r0:3 * r1].copy(); sc = cols[ptr[r0]:ptr[r1]].copy(); sv = coeffs[ptr[r0]:ptr[r1]].copy()
    h.ok(L.bp_cs_enforce(h.h, r1 - r0, sl.ctypes.data, sc.ctypes.data, sv.ctypes.data))
    h.ok(L.bp_cs_set_row_base(h.h, r0))
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (ctypes.c_uint8 * 128)()
        assert L.bp_group_unique_id(raw) == 0
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    dist.broadcast(idbuf, 0)
    raw = (ctypes.c_uint8 * 128)(*idbuf.tolist())
    g = ctypes.c_void_p()
    h.ok(L.bp_group_init(h.h, raw, rank, world, ctypes.byref(g)))
    tr = ctypes.c_int()
    h.ok(L.bp_group_info(g, None, None, ctypes.byref(tr)))
    out = {"transport": tr.value, "checks": []}
    row = ctypes.c_int64()
    h.ok(L.bp_group_broadcast_witness(g, 0))
    h.ok(L.bp_group_check(g, ctypes.byref(row)))
    out["checks"].append([row.value, inst.check(2, False)])
    import random
    rng = random.Random(7)   # same sequence on every rank
    for trial in range(6):
        k = rng.randrange(aux.shape[0])
        v = c_api.ints_to_limbs([rng.randrange(1 << 200)])
        h.ok(L.bp_cs_set(h.h, 1, k, v.ctypes.data))
        inst.set(True, k, c_api.limbs_to_ints(v)[0])
        h.ok(L.bp_group_check(g, ctypes.byref(row)))
        out["checks"].append([row.value, inst.check(2, False)])
    # a new witness, every rank uploading its slice
    a2 = aux.copy(); a2[1234] = c_api.ints_to_limbs([99])[0]
    n = aux.shape[0]; s0, s1 = rank * n // world, (rank + 1) * n // world
    sl2 = np.ascontiguousarray(a2[s0:s1])
    h.ok(L.bp_group_set_witness_sharded(g, 1, sl2.ctypes.data))
    ref = c_api.Instance(fid, lens, cols, coeffs, inputs, a2)
    dev = torch.zeros(1, dtype=torch.int64, device=f"cuda:{rank}")
    for _ in range(3):
        h.ok(L.bp_group_check_async(g, ctypes.c_void_p(dev.data_ptr())))
    h.ok(L.bp_cs_sync(h.h))
    w = ref.check(2, False)
    out["checks"].append([int(dev.item()), w if w >= 0 else 0x7FFFFFFFFFFFFFFF])
    got = np.zeros_like(aux)
    h.ok(L.bp_cs_witness(h.h, 1, 0, n, got.ctypes.data))
    out["witness_equal"] = bool((got == a2).all())
    out["graph_replays"] = h.opt("graph_replays")
    L.bp_group_free(g)
    h.close()
    print("RESULT %%d " %% rank + json.dumps(out), flush=True)
    dist.destroy_process_group()
    """
) % (ROOT, ROOT)


@pytest.mark.parametrize("mailbox", [True, False])
def test_group_of_two_ranks_matches_oracle(tmp_path, mailbox):
    """Two processes, one GPU each: witness broadcast, sharded upload + all-gather, check + exchange (peer mailboxes over NVLink,
    or the NCCL fallback) -- the reduced first-unsatisfied GLOBAL row equals the oracle's on the whole system on both ranks."""
    import json

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(GROUP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if not mailbox:
        env["BP_GROUP_NO_MAILBOX"] = "1"
    import socket

    def launch():
        with socket.socket() as sk:  # a free rendezvous port (a fixed one can still be in TIME_WAIT from an earlier run)
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                               "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600, env=env)

    r = launch()
    if r.returncode != 0:
        tb = [ln for ln in r.stderr.splitlines() if "Error" in ln or "error" in ln or "File" in ln][-12:]
        raise AssertionError("worker failed:\n" + "\n".join(tb) + "\n" + r.stdout[-1500:])
    res = {}
    for ln in r.stdout.splitlines():
        if ln.startswith("RESULT "):
            _, rk, js = ln.split(" ", 2)
            res[int(rk)] = json.loads(js)
    assert set(res) == {0, 1}
    for rk, o in res.items():
        assert o["transport"] == (2 if mailbox else 1), o
        for got, want in o["checks"]:
            assert got == want, (rk, o)
        assert o["witness_equal"]


@pytest.mark.parametrize("fid,log_rows,t", [(0, 22, 6), (1, 20, 32)])
def test_full_width_kernels_at_size_match_oracle(fid, log_rows, t):
    """Parity at a size where 32-bit offsets, tile boundaries and the LC-tile layout matter: a
    device-generated synthetic instance of 2^22 rows (t = 6) / 2^20 rows (t = 32); canonical A.w, B.w, C.w of row blocks at the
    start, at odd offsets in the middle and at the very end bit for bit against the oracle (which generates only the witness
    elements those rows read), with the streaming LC-tile kernel and with the thread-per-row kernel, and the two kernels agree
    on EVERY row."""
    import torch

    n = 1 << log_rows
    n_vars = n
    with Handle(fid, reserve=(n, int(n * 3 * t * 1.02) + 4096, n_vars)) as h:
        h.ok(h.L.bp_cs_synth_witness(h.h, synth.SEED, n_vars, synth.N_INPUTS))
        for s in range(0, n, 1 << 20):
            h.ok(h.L.bp_cs_synth_rows(h.h, synth.SEED, t, n_vars, synth.N_INPUTS, s, min(1 << 20, n - s)))
        outs = {}
        for sk in (1, 0):
            h.opt("stream_kernel", sk)
            assert h.opt("stream_kernel") == sk
            dev = [torch.empty((n, 4), dtype=torch.int64, device="cuda:0") for _ in range(3)]
            h.ok(h.L.bp_cs_eval_async(h.h, *[ctypes.c_void_p(x.data_ptr()) for x in dev]))
            h.ok(h.L.bp_cs_sync(h.h))
            outs[sk] = dev
            assert h.first_unsatisfied() == 0  # random rows: the first one already fails
        for a, b in zip(outs[1], outs[0]):
            assert bool((a == b).all())
        blk = 3000
        for r0 in (0, 255, n // 2 - 1234, n - blk):
            inst = c_api.synth_sparse_instance(fid, synth.SEED, t, n_vars, synth.N_INPUTS, r0, blk)
            bad, az, bz, cz = inst.eval(2)
            for dev_t, ref in zip(outs[1], (az, bz, cz)):
                assert (dev_t[r0:r0 + blk].cpu().numpy().view(np.uint64) == ref).all()
            inst.close()


@pytest.mark.parametrize("fid,blocks", [(1, 3), (0, 7)])
def test_witness_generated_on_device_equals_the_front_end(fid, blocks):
    """the sha256-chain witness generated on the device from the message bytes (the program the front-end
    recorded while it synthesized the circuit) equals, bit for bit, what the gadgets' host closures produce -- for the recorded
    message and for another one -- and the circuit holds with it."""
    from bellpepper_b200 import ffi, fixtures

    L = ffi.load()
    msg = fixtures.chain_message(blocks)
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.record_witness_program()
        t.sha256(msg)
        prog = t.witness_program()
        h = ffi.vp(t.handle)
        n_aux = t.num_aux()
        want = np.zeros((n_aux, 4), np.uint64)
        assert L.bp_cs_witness(h, 1, 0, n_aux, want.ctypes.data) == 0
        assert t.first_unsatisfied_row() == -1
        assert L.bp_cs_set_witness_program(h, prog.ctypes.data, prog.size) == 0, L.bp_cs_last_error(h)
        for m in (msg, bytes((b * 5 + 3) & 0xFF for b in msg)):
            # wreck the witness first: the generator must rewrite every aux value
            junk = np.ones(n_aux, np.uint8)
            assert L.bp_cs_set_range_u8(h, 1, 0, n_aux, junk.ctypes.data) == 0
            assert t.first_unsatisfied_row() >= 0
            st = fixtures.sha256_chain_states(m)
            assert L.bp_cs_generate_witness_async(h, m, len(m), st.ctypes.data, st.size) == 0, L.bp_cs_last_error(h)
            assert t.first_unsatisfied_row() == -1
            got = np.zeros((n_aux, 4), np.uint64)
            assert L.bp_cs_witness(h, 1, 0, n_aux, got.ctypes.data) == 0
            if m is msg:
                assert (got == want).all()
            else:
                with fixtures.Tcs(fid, device=-1, named=False) as rec:
                    rec.sha256(m)
                    assert (got == rec.host_csr()[4]).all()
        # a wrong chaining state is caught by the circuit itself
        st = fixtures.sha256_chain_states(msg).copy()
        st[1][3] ^= 4
        assert L.bp_cs_generate_witness_async(h, msg, len(msg), st.ctypes.data, st.size) == 0
        assert t.first_unsatisfied_row() >= 0
        # malformed programs and mismatched inputs are refused
        bad = prog.copy()
        bad[int(prog[10]) + 1] = n_aux  # unit 0's variables past the end of the aux space
        assert L.bp_cs_set_witness_program(h, bad.ctypes.data, bad.size) == -5
        bad = prog.copy()
        bad[0] ^= 1
        assert L.bp_cs_set_witness_program(h, bad.ctypes.data, bad.size) == -5
        assert L.bp_cs_generate_witness_async(h, msg[:-1], len(msg) - 1, st.ctypes.data, st.size) == -5


def _read_export(path):
    import struct

    raw = open(path, "rb").read()
    assert raw[:8] == b"BPR1CSX\x01"
    version, field, n_rows, n_inputs, n_aux, nnz, row_base, flags, _ = struct.unpack_from("<IIQQQQQII", raw, 8)
    assert version == 1
    off = 8 + struct.calcsize("<IIQQQQQII")
    out = {"field": field, "n_rows": n_rows, "row_base": row_base, "flags": flags}

    def take(count, dtype, shape=None):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a.reshape(shape) if shape else a

    out["lens"] = take(3 * n_rows, "<u4")
    out["cols"] = take(nnz, "<u4")
    out["coeffs"] = take(4 * nnz, "<u8", (-1, 4))
    out["inputs"] = take(4 * n_inputs, "<u8", (-1, 4))
    out["aux"] = take(4 * n_aux, "<u8", (-1, 4))
    if flags & 1:
        for k in ("az", "bz", "cz"):
            out[k] = take(4 * n_rows, "<u8", (-1, 4))
    assert off + 8 == len(raw)
    return out


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_export_round_trips_to_the_oracle_csr(tmp_path, fid):
    """Prover hand-off: bp_cs_export writes A, B, C with CANONICAL coefficients (internal scaling, negated C and
    coefficient classes undone), the witness and A.w / B.w / C.w; the file equals the arrays that were ingested, bit for bit, on
    a gadget-shaped instance (every coefficient class, plain and general LCs) and on a synthetic one; another handle fed with
    the file's arrays gives the same verdict."""
    for make in ("gadget", "synthetic"):
        if make == "gadget":
            lens, cols, coeffs, inputs, aux, _ = _gadget_like_instance(fid, 41, 1200, 2000, 7)
        else:
            lens, cols, coeffs, inputs, aux = c_api.synth_instance(fid, synth.SEED, 5, 3000, synth.N_INPUTS, 900)
        inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
        bad, az, bz, cz = inst.eval(2)
        path = str(tmp_path / f"{make}.bpx").encode()
        with Handle(fid) as h:
            h.load_instance(lens, cols, coeffs, inputs, aux)
            h.ok(h.L.bp_cs_set_row_base(h.h, 77))
            h.ok(h.L.bp_cs_export(h.h, path, 1))
            want_row = h.first_unsatisfied()
        x = _read_export(path.decode())
        assert x["field"] == fid and x["row_base"] == 77 and x["flags"] == 1
        assert (x["lens"] == lens).all() and (x["cols"] == cols).all()
        assert (x["coeffs"] == coeffs).all()
        assert (x["inputs"] == inputs).all() and (x["aux"] == aux).all()
        assert (x["az"] == az).all() and (x["bz"] == bz).all() and (x["cz"] == cz).all()
        with Handle(fid) as h2:  # the file's arrays are exactly what bp_cs_alloc / bp_cs_enforce take
            h2.load_instance(np.array(x["lens"]), np.array(x["cols"]), np.array(x["coeffs"]), np.array(x["inputs"]), np.array(x["aux"]))
            assert h2.first_unsatisfied() == want_row == bad


def test_metric_cs_pretty_print():
    """MetricCS::pretty_print (crates/bellpepper/src/util_cs/metric_cs.rs:130-195) from the host mirror."""
    from bellpepper_b200 import ONE, TestConstraintSystem

    cs = TestConstraintSystem(0)
    p = cs.p
    a = cs.alloc("a", lambda: 3)
    with cs.namespace("ns"):
        b = cs.alloc("b", lambda: 5)
    x = cs.alloc_input("x", lambda: 15)
    cs.enforce("mult", lambda lc: lc + a, lambda lc: lc + b, lambda lc: lc + x)
    cs.enforce("odd", lambda lc: lc + (32, a) - b + (7, ONE), lambda lc: lc + ONE - ONE, lambda lc: lc)
    s = cs.pretty_print_equations()
    d = lambda c: f"Scalar(0x{c:064x})"
    assert s == ("INPUT ONE\nINPUT x\n"
                 "\nmult: (`Aa`) * (`Ans/b`) = (`Ix`)"
                 f"\nodd: ({d(7)} . `IONE` + 2^5 . {d(32)} . `Aa` - `Ans/b`) * (0) = (0)"
                 "\n")
    cs.close()
