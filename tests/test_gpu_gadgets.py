"""Gadget circuits (BASELINE configs 1-3 shapes) through the C++ front-end into the B200 engine vs the oracle."""
import ctypes
import hashlib
import random

import numpy as np
import pytest

from bellpepper_b200 import ffi, fixtures
from oracle import c_api
from oracle.fields import FIELDS

pytestmark = pytest.mark.gpu


def device_eval(t, n_rows):
    L = ffi.load()
    h = ffi.vp(t.handle)
    az, bz, cz = (np.zeros((n_rows, 4), np.uint64) for _ in range(3))
    assert L.bp_cs_eval(h, az.ctypes.data, bz.ctypes.data, cz.ctypes.data) == 0, L.bp_cs_last_error(h)
    return az, bz, cz


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_sha256_two_blocks_bit_exact(fid):
    """configs[0]-shaped: sha256 gadget, values of every LC of every row equal the oracle's, satisfied, digest right."""
    msg = fixtures.xorshift_bytes(64)
    with fixtures.Tcs(fid, device=-1, named=False) as rec:
        rec.sha256(msg)
        lens, cols, coeffs, inputs, aux = rec.host_csr()
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    bad, az_r, bz_r, cz_r = inst.eval(4)
    assert bad == -1
    with fixtures.Tcs(fid, device=0, named=True) as t:
        digest, _ = t.sha256(msg)
        assert digest == hashlib.sha256(msg).digest()
        n = t.num_constraints()
        assert n == 44874 + 512 == lens.size // 3
        assert t.is_satisfied()
        az, bz, cz = device_eval(t, n)
        assert (az == az_r).all() and (bz == bz_r).all() and (cz == cz_r).all()
        L = ffi.load()
        h = ffi.vp(t.handle)
        fat = ctypes.c_int64()
        assert L.bp_cs_get_option(h, b"fat_rows", ctypes.byref(fat)) == 0 and fat.value > 20  # MultiEq rows took the warp path
        # flip-and-recheck by path (uint32.rs:627-633 idiom); first failure must be the reference's
        rng = random.Random(fid)
        paths = ["block 0/w extension 16/computation of w[i]/result bit 0/boolean", "input bit 3 5/boolean",
                 "block 1/compression round 63/maj/maj 31/maj", "block 1/new h7/result bit 32/boolean"]
        for path in paths:
            old = t.get(path)
            assert old in (0, 1)
            t.set(path, 1 - old)
            # oracle: same flip on the flat instance
            idx = None
            got = t.which_is_unsatisfied()
            assert got is not None
            row = t.first_unsatisfied_row()
            # find aux index of the path by probing the device value back (names live in the C++ mirror)
            t.set(path, old)
            assert t.is_satisfied()
            assert t.row_path(row) == got
        # a non-boolean value in a result bit trips its own boolean constraint first or an earlier user of the bit
        t.set("block 0/w extension 16/computation of w[i]/result bit 5/boolean", rng.randrange(2, FIELDS[fid].p))
        assert t.which_is_unsatisfied() == "block 0/w extension 16/computation of w[i]/result bit 5/boolean constraint"


def test_flip_matches_oracle_first_failure():
    fid = 1
    msg = fixtures.xorshift_bytes(130)
    with fixtures.Tcs(fid, device=-1, named=False) as rec:
        rec.sha256(msg)
        lens, cols, coeffs, inputs, aux = rec.host_csr()
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    L = ffi.load()
    rng = random.Random(11)
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.sha256(msg)
        h = ffi.vp(t.handle)
        assert t.is_satisfied()
        for _ in range(12):
            idx = rng.randrange(aux.shape[0])
            old = int(aux[idx][0])
            new = 1 - old if rng.random() < 0.7 else rng.randrange(FIELDS[fid].p)
            v = c_api.ints_to_limbs([new])
            assert L.bp_cs_set(h, 1, idx, v.ctypes.data) == 0
            inst.set(True, idx, new)
            assert t.first_unsatisfied_row() == inst.check(4, False)
            v = c_api.ints_to_limbs([old])
            assert L.bp_cs_set(h, 1, idx, v.ctypes.data) == 0
            inst.set(True, idx, old)
        assert t.is_satisfied()


def test_sha256_chain_64_blocks_pallas_satisfied_and_sharded():
    """configs[1]-shaped (scaled down): 64 chained blocks; whole system satisfied; two row shards agree with the whole."""
    fid, blocks = 1, 64
    msg = fixtures.chain_message(blocks)
    L = ffi.load()
    with fixtures.Tcs(fid, device=0, named=False) as t:
        digest, before = t.sha256(msg)
        assert digest == hashlib.sha256(msg).digest() and before == 0
        n_total = t.num_constraints()
        assert t.is_satisfied()
        h = ffi.vp(t.handle)
        # corrupt one late witness bit: the failing row is reported with its global index by the shard that owns it
        victim = t.num_aux() - 40000
        old = np.zeros(4, np.uint64)
        assert L.bp_cs_get(h, 1, victim, old.ctypes.data) == 0
        new = c_api.ints_to_limbs([1 - int(old[0])])
        assert L.bp_cs_set(h, 1, victim, new.ctypes.data) == 0
        want = t.first_unsatisfied_row()
        assert want > 0
    got = []
    for b0, b1 in ((0, 40), (40, 64)):
        with fixtures.Tcs(fid, device=0, named=False) as s:
            _, before = s.sha256(msg, b0, b1)
            h = ffi.vp(s.handle)
            assert L.bp_cs_set_row_base(h, before) == 0
            assert s.is_satisfied()
            assert L.bp_cs_set(h, 1, victim, new.ctypes.data) == 0
            row = ctypes.c_int64()
            assert L.bp_cs_first_unsatisfied(h, ctypes.byref(row)) == 0
            got.append(row.value + before if row.value >= 0 else None)
    assert min(g for g in got if g is not None) == want


@pytest.mark.parametrize("fid,n_bytes", [(2, 200), (0, 64)])
def test_blake2s_bit_exact_and_flips(fid, n_bytes):
    """configs[2]-shaped: blake2s gadget (Vesta Fr in BASELINE): every LC value equals the oracle's, satisfied, digest right;
    then witness edits (bit flips and non-boolean values) give the oracle's first failing row."""
    msg = fixtures.xorshift_bytes(n_bytes)
    with fixtures.Tcs(fid, device=-1, named=False) as rec:
        rec.blake2s(msg)
        lens, cols, coeffs, inputs, aux = rec.host_csr()
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    bad, az_r, bz_r, cz_r = inst.eval(4)
    assert bad == -1
    L = ffi.load()
    rng = random.Random(5 + fid)
    with fixtures.Tcs(fid, device=0, named=False) as t:
        digest = t.blake2s(msg)
        assert digest == hashlib.blake2s(msg, digest_size=32, person=b"12345678").digest()
        n = t.num_constraints()
        assert n == lens.size // 3
        assert t.is_satisfied()
        az, bz, cz = device_eval(t, n)
        assert (az == az_r).all() and (bz == bz_r).all() and (cz == cz_r).all()
        h = ffi.vp(t.handle)
        stats = {}
        for key in (b"plain_rows", b"fat_rows", b"fat_undecided_rows", b"deferred_rows"):
            v = ctypes.c_int64()
            assert L.bp_cs_get_option(h, key, ctypes.byref(v)) == 0
            stats[key.decode()] = v.value
        assert stats["fat_rows"] > 10 and stats["plain_rows"] > n // 2
        assert stats["fat_undecided_rows"] == 0 and stats["deferred_rows"] == 0  # honest witness: integer kernels decide all
        for _ in range(10):
            idx = rng.randrange(aux.shape[0])
            old = int(aux[idx][0])
            new = 1 - old if rng.random() < 0.7 else rng.randrange(FIELDS[fid].p)
            v = c_api.ints_to_limbs([new])
            assert L.bp_cs_set(h, 1, idx, v.ctypes.data) == 0
            inst.set(True, idx, new)
            assert t.first_unsatisfied_row() == inst.check(4, False)
            v = c_api.ints_to_limbs([old])
            assert L.bp_cs_set(h, 1, idx, v.ctypes.data) == 0
            inst.set(True, idx, old)
        assert t.is_satisfied()


def test_pipelined_recheck_equals_plain_recheck():
    """bp_cs_recheck_u8 with a pinned packed witness (upload pipelined with the check) vs upload-then-check, on a chain long
    enough to take the pipelined path; edits early, late and at chunk boundaries; a non-byte value stays a patch."""
    import torch

    fid, blocks = 1, 170
    L = ffi.load()
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.sha256(fixtures.chain_message(blocks))
        h = ffi.vp(t.handle)
        n_in, n_aux = t.num_inputs(), t.num_aux()
        assert n_aux >= (4 << 20)
        w = np.zeros((n_aux, 4), np.uint64)
        assert L.bp_cs_witness(h, 1, 0, n_aux, w.ctypes.data) == 0
        assert not w[:, 1:].any() and int(w[:, 0].max()) <= 1
        b_aux = torch.from_numpy(w[:, 0].astype(np.uint8)).pin_memory()
        b_in = torch.ones(n_in, dtype=torch.uint8).pin_memory()
        row = ctypes.c_int64()

        def pipelined():
            assert L.bp_cs_recheck_u8(h, ctypes.c_void_p(b_in.data_ptr()), ctypes.c_void_p(b_aux.data_ptr()), ctypes.byref(row)) == 0, \
                L.bp_cs_last_error(h)
            return row.value

        def plain():
            assert L.bp_cs_set_range_u8(h, 1, 0, n_aux, ctypes.c_void_p(b_aux.data_ptr())) == 0
            assert L.bp_cs_first_unsatisfied(h, ctypes.byref(row)) == 0
            return row.value

        assert pipelined() == -1 and plain() == -1
        chunk = (((n_aux + 15) // 16) + 255) & ~255
        rng = random.Random(1)
        for victim in [5, chunk - 1, chunk, 7 * chunk + 3, n_aux - 1, rng.randrange(n_aux), rng.randrange(n_aux)]:
            old = int(b_aux[victim])
            b_aux[victim] = 1 - old
            a, b = pipelined(), plain()
            assert a == b and a >= 0, (victim, a, b)
            b_aux[victim] = old
        assert pipelined() == -1
        # the last rows of the chain: the final blocks' rows only become ready with the last chunk
        b_aux[n_aux - 2000] = 1 - int(b_aux[n_aux - 2000])
        assert pipelined() == plain() > t.num_constraints() - 30000


def test_sparse_upload_on_row_shards():
    """Row shards with "sparse_upload": each shard copies only the witness chunks its rows read, and the minimum over the
    shards still equals the whole system's first failing row."""
    import torch

    fid, blocks = 1, 170
    msg = fixtures.chain_message(blocks)
    L = ffi.load()
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.sha256(msg)
        h = ffi.vp(t.handle)
        n_in, n_aux = t.num_inputs(), t.num_aux()
        w = np.zeros((n_aux, 4), np.uint64)
        assert L.bp_cs_witness(h, 1, 0, n_aux, w.ctypes.data) == 0
        b_aux = torch.from_numpy(w[:, 0].astype(np.uint8)).pin_memory()
        b_in = torch.ones(n_in, dtype=torch.uint8).pin_memory()
        row = ctypes.c_int64()
        victims = [n_aux // 3, n_aux - 5000, 2_000_000 + 7]
        want = []
        for v in victims:
            b_aux[v] = 1 - int(b_aux[v])
            assert L.bp_cs_recheck_u8(h, ctypes.c_void_p(b_in.data_ptr()), ctypes.c_void_p(b_aux.data_ptr()), ctypes.byref(row)) == 0
            want.append(row.value)
            b_aux[v] = 1 - int(b_aux[v])
        assert all(x >= 0 for x in want)
    shards = []
    for b0, b1 in ((0, 60), (60, 120), (120, blocks)):
        s = fixtures.Tcs(fid, device=0, named=False)
        _, before = s.sha256(msg, b0, b1)
        hs = ffi.vp(s.handle)
        assert L.bp_cs_set_row_base(hs, before) == 0
        assert L.bp_cs_set_option(hs, b"sparse_upload", 1) == 0
        up = ctypes.c_int64()
        assert L.bp_cs_get_option(hs, b"recheck_upload_bytes", ctypes.byref(up)) == 0
        assert up.value < 0.6 * n_aux  # a third of the blocks (+ the message bits): far less than the whole witness
        shards.append((s, hs, before))
    try:
        for v, expect in zip(victims, want):
            b_aux[v] = 1 - int(b_aux[v])
            got = []
            for s, hs, before in shards:
                assert L.bp_cs_recheck_u8(hs, ctypes.c_void_p(b_in.data_ptr()), ctypes.c_void_p(b_aux.data_ptr()), ctypes.byref(row)) == 0, \
                    L.bp_cs_last_error(hs)
                if row.value >= 0:
                    got.append(row.value + before)
            assert min(got) == expect, (v, got, expect)
            b_aux[v] = 1 - int(b_aux[v])
        for s, hs, before in shards:
            assert L.bp_cs_recheck_u8(hs, ctypes.c_void_p(b_in.data_ptr()), ctypes.c_void_p(b_aux.data_ptr()), ctypes.byref(row)) == 0
            assert row.value == -1
    finally:
        for s, _, _ in shards:
            s.close()


def test_bit_packed_recheck_equals_byte_packed():
    """bp_cs_recheck_bits / bp_cs_set_range_bits (1 bit per value) vs the byte-packed calls: same verdicts, pipelined and not."""
    import torch

    fid, blocks = 1, 170
    L = ffi.load()
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.sha256(fixtures.chain_message(blocks))
        h = ffi.vp(t.handle)
        n_in, n_aux = t.num_inputs(), t.num_aux()
        w = np.zeros((n_aux, 4), np.uint64)
        assert L.bp_cs_witness(h, 1, 0, n_aux, w.ctypes.data) == 0
        vals = w[:, 0].astype(np.uint8)
        b_in = torch.ones(n_in, dtype=torch.uint8).pin_memory()
        p_in = torch.from_numpy(np.packbits(b_in.numpy(), bitorder="little")).pin_memory()
        row = ctypes.c_int64()
        rng = random.Random(3)
        for victim in [None, 11, n_aux - 1, rng.randrange(n_aux), (n_aux // 16 + 255) // 256 * 256]:
            cur = vals.copy()
            if victim is not None:
                cur[victim] ^= 1
            b_aux = torch.from_numpy(cur).pin_memory()
            p_aux = torch.from_numpy(np.packbits(cur, bitorder="little")).pin_memory()
            assert L.bp_cs_recheck_u8(h, ctypes.c_void_p(b_in.data_ptr()), ctypes.c_void_p(b_aux.data_ptr()), ctypes.byref(row)) == 0
            want = row.value
            assert (want == -1) == (victim is None)
            assert L.bp_cs_recheck_bits(h, ctypes.c_void_p(p_in.data_ptr()), ctypes.c_void_p(p_aux.data_ptr()), ctypes.byref(row)) == 0, \
                L.bp_cs_last_error(h)
            assert row.value == want
            # not pipelined: pageable source, and a sub-range that does not start on a byte boundary of the whole witness
            pageable = np.packbits(cur, bitorder="little")
            assert L.bp_cs_recheck_bits(h, p_in.numpy().ctypes.data, pageable.ctypes.data, ctypes.byref(row)) == 0
            assert row.value == want
            first, n = 1003, 70001
            sub = np.packbits(cur[first: first + n], bitorder="little")
            assert L.bp_cs_set_range_bits(h, 1, first, n, sub.ctypes.data) == 0
            assert L.bp_cs_first_unsatisfied(h, ctypes.byref(row)) == 0 and row.value == want
            back = np.zeros((n, 4), np.uint64)
            assert L.bp_cs_witness(h, 1, first, n, back.ctypes.data) == 0
            assert (back[:, 0] == cur[first: first + n]).all() and not back[:, 1:].any()
