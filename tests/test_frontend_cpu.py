"""C++ host front-end (csrc/host/*.hpp) on the CPU: it must emit exactly the rows, columns, coefficients and
witness the oracle's Python restatement of the reference gadgets emits, and reproduce the reference's structural
known-answers.  No evaluation happens here except by the oracle."""
import hashlib

import numpy as np
import pytest

from bellpepper_b200 import fixtures
from oracle import c_api
from oracle import gadgets_py as G
from oracle.fields import FIELDS
from oracle.r1cs_py import TestConstraintSystem


def oracle_csr(cs):
    lens, cols, coeffs, inputs, aux = cs.to_csr()
    return (np.asarray(lens, np.uint32), np.asarray(cols, np.uint32), c_api.ints_to_limbs(coeffs),
            c_api.ints_to_limbs(inputs), c_api.ints_to_limbs(aux))


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and (x == y).all()


def test_xorshift_streams_agree():
    assert fixtures.xorshift_bytes(100) == G.xorshift_bytes(G.SEED_3D, 100)


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_sha256_block_identical_to_oracle_gadgets(fid):  # sha256.rs:310-336: 25840 (+512)
    F = FIELDS[fid]
    block = fixtures.xorshift_bytes(64)
    cs = TestConstraintSystem(F)
    bits = []
    for i in range(512):
        with cs.namespace(f"input bit {i}") as ns:
            bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, bool((block[i // 8] >> (7 - i % 8)) & 1))))
    out = G.sha256_compression_function(cs, F, bits, G.sha256_iv())
    with fixtures.Tcs(fid, device=-1, named=True) as t:
        out32 = t.sha256_block(block)
        assert t.num_constraints() == cs.num_constraints() == 25840 + 512
        assert t.num_aux() == len(cs.aux) and t.num_inputs() == 1
        same(t.host_csr(), oracle_csr(cs))
        # constraint paths are the reference's (test_cs.rs:363-375)
        for row in (0, 511, 512, 5000, 26351):
            assert t.row_path(row) == cs.constraints[row][3]
    want = b"".join(int(w.value).to_bytes(4, "big") for w in out)
    assert out32 == want


def test_sha256_two_blocks_identical_and_digest():  # sha256.rs:338-363: 44874 (+512); digest vs hashlib
    fid = 0
    F = FIELDS[fid]
    msg = fixtures.xorshift_bytes(64)
    cs = TestConstraintSystem(F)
    bits = []
    for i, byte in enumerate(msg):
        for j in range(7, -1, -1):
            with cs.namespace(f"input bit {i} {j}") as ns:
                bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, bool((byte >> j) & 1))))
    G.sha256(cs, F, bits)
    for named in (True, False):
        with fixtures.Tcs(fid, device=-1, named=named) as t:
            digest, before = t.sha256(msg)
            assert digest == hashlib.sha256(msg).digest() and before == 0
            assert t.num_constraints() - 512 == 44874
            same(t.host_csr(), oracle_csr(cs))


def test_constant_input_gives_no_constraints():  # sha256.rs:283-308
    with fixtures.Tcs(0, device=-1, named=True) as t:
        digest, _ = t.sha256(b"")
        assert digest == hashlib.sha256(b"").digest()
        assert t.num_constraints() == 0


def test_row_sharding_by_block_partitions_the_rows():
    fid = 1
    msg = fixtures.chain_message(4)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.sha256(msg)
        full = t.host_csr()
    n_full = full[0].size // 3
    off = np.concatenate([[0], np.cumsum(full[0].astype(np.int64))])
    got_rows = 0
    for b0, b1 in ((0, 1), (1, 3), (3, 4)):
        with fixtures.Tcs(fid, device=-1, named=False) as t:
            digest, before = t.sha256(msg, b0, b1)
            lens, cols, coeffs, inputs, aux = t.host_csr()
        assert digest == hashlib.sha256(msg).digest()
        assert before == got_rows
        n = lens.size // 3
        assert (lens == full[0][3 * before: 3 * (before + n)]).all()
        k0, k1 = int(off[3 * before]), int(off[3 * (before + n)])
        assert (cols == full[1][k0:k1]).all() and (coeffs == full[2][k0:k1]).all()
        assert (aux == full[4]).all() and (inputs == full[3]).all()  # witness is replicated, rows are sharded
        got_rows += n
    assert got_rows == n_full


def test_front_end_output_satisfies_the_oracle_and_flips():
    fid = 2
    with fixtures.Tcs(fid, device=-1, named=True) as t:
        t.sha256(fixtures.xorshift_bytes(3))
        lens, cols, coeffs, inputs, aux = t.host_csr()
        paths = [t.row_path(r) for r in range(t.num_constraints())]
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    assert inst.check(4, False) == -1
    inst.set(True, 5, 2)  # aux 5 is message bit 5: its own boolean constraint (row 5) is the first to fail
    bad = inst.check(1, True)
    assert bad == 5 and paths[bad] == "input bit 0 2/boolean constraint"


@pytest.mark.parametrize("fid,n_bytes", [(0, 64), (2, 64), (1, 97), (2, 0)])
def test_blake2s_identical_to_oracle_gadgets(fid, n_bytes):  # blake2s.rs:443-457 (21518), :498-555 (digests)
    F = FIELDS[fid]
    msg = fixtures.xorshift_bytes(n_bytes)
    cs = TestConstraintSystem(F)
    bits = []
    for i, byte in enumerate(msg):
        for j in range(8):
            with cs.namespace(f"input bit {i} {j}") as ns:
                bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, bool((byte >> j) & 1))))
    G.blake2s(cs, F, bits, b"12345678")
    with fixtures.Tcs(fid, device=-1, named=True) as t:
        digest = t.blake2s(msg)
        assert digest == hashlib.blake2s(msg, digest_size=32, person=b"12345678").digest()
        assert t.num_constraints() == cs.num_constraints()
        if n_bytes == 64:
            assert t.num_constraints() == 21518
        same(t.host_csr(), oracle_csr(cs))
        for row in (0, cs.num_constraints() // 2, cs.num_constraints() - 1) if cs.num_constraints() else ():
            assert t.row_path(row) == cs.constraints[row][3]


# ---- Boolean xor / and / or / sha256_ch / sha256_maj / enforce_equal over every operand kind (boolean.rs:1109-2003) --------
def _python_boolean_op(F, op, a_kind, b_kind, c_kind):
    import kat_scenarios as S

    cs = TestConstraintSystem(F)
    a, b = S.construct(cs, a_kind, "a"), S.construct(cs, b_kind, "b")
    if op == "xor":
        r = G.Boolean.xor(cs, a, b)
    elif op == "and":
        r = G.Boolean.and_(cs, a, b)
    elif op == "or":
        r = S.boolean_or(cs, a, b)
    elif op == "enforce_equal":
        G.Boolean.enforce_equal(cs, F, a, b)
        r = G.Boolean.constant(False)
    else:
        c = S.construct(cs, c_kind, "c")
        r = (G.Boolean.sha256_ch if op == "sha256_ch" else G.Boolean.sha256_maj)(cs, F, a, b, c)
    return cs, ["Is", "Not", "Constant"][r.kind], bool(r.get_value())


@pytest.mark.parametrize("fid", [0, 1])
def test_boolean_ops_identical_to_oracle_gadgets_over_all_operand_kinds(fid):
    from oracle.r1cs_py import Unsatisfiable

    F = FIELDS[fid]
    kinds = fixtures.Tcs.OPERAND_KINDS
    n = 0
    for op in ("xor", "and", "or", "enforce_equal", "sha256_ch", "sha256_maj"):
        for a in kinds:
            for b in kinds:
                for c in (kinds if op.startswith("sha256") else ["True"]):
                    try:
                        cs, want_kind, want_value = _python_boolean_op(F, op, a, b, c)
                    except Unsatisfiable:
                        with fixtures.Tcs(fid, device=-1, named=True) as t:
                            with pytest.raises(RuntimeError, match="unsatisfiable"):
                                t.boolean_op(op, a, b, c)
                        continue
                    with fixtures.Tcs(fid, device=-1, named=True) as t:
                        kind, value = t.boolean_op(op, a, b, c)
                        assert (kind, value) == (want_kind, want_value), (op, a, b, c)
                        assert t.num_constraints() == cs.num_constraints()
                        same(t.host_csr(), oracle_csr(cs))
                        for row in range(cs.num_constraints()):  # the reference's paths, row for row
                            assert t.row_path(row) == cs.constraints[row][3]
                    n += 1
    assert n == 3 * 36 + (36 - 2) + 2 * 216  # (enforce_equal of True with False, either way round, is the error case)


def test_u64_into_boolean_vec_le():  # boolean.rs:1776-1794
    with fixtures.Tcs(0, device=-1, named=True) as t:
        bits = t.u64_bits(17234652694787248421)
        assert t.num_constraints() == 64 and t.num_aux() == 64
        assert t.row_path(5) == "bit 5/boolean constraint"
        lens, cols, coeffs, inputs, aux = t.host_csr()
    assert [int(bits[63 - i]) for i in (0, 1, 2, 3, 4, 5, 20, 21, 22)] == [1, 1, 1, 0, 1, 1, 1, 0, 0]
    assert sum(int(b) << i for i, b in enumerate(bits)) == 17234652694787248421
    assert c_api.Instance(0, lens, cols, coeffs, inputs, aux).check(1, True) == -1


# ---- UInt32 xor / addmany / sha256_maj / sha256_ch as the reference's tests build them (uint32.rs:492-780) -------------------
def _python_uint32_op(F, op, a, b, c, d):
    cs = TestConstraintSystem(F)
    with cs.namespace("a_bit") as ns:
        a_bit = G.UInt32.alloc(ns, a)
    b_bit = G.UInt32.constant(b)
    if op == "xor":
        with cs.namespace("c_bit") as ns:
            c_bit = G.UInt32.alloc(ns, c)
        with cs.namespace("first xor") as ns:
            r = G.UInt32.xor(a_bit, ns, b_bit)
        with cs.namespace("second xor") as ns:
            r = G.UInt32.xor(r, ns, c_bit)
    elif op == "addmany":
        c_bit = G.UInt32.constant(c)
        with cs.namespace("d_bit") as ns:
            d_bit = G.UInt32.alloc(ns, d)
        with cs.namespace("xor") as ns:
            r = G.UInt32.xor(a_bit, ns, b_bit)
        me = G.MultiEq(cs, F)
        with me.namespace("addition") as ns:
            r = G.UInt32.addmany(ns, F, [r, c_bit, d_bit])
        me.finish()
    else:
        with cs.namespace("c_bit") as ns:
            c_bit = G.UInt32.alloc(ns, c)
        r = (G.UInt32.sha256_maj if op == "sha256_maj" else G.UInt32.sha256_ch)(cs, F, a_bit, b_bit, c_bit)
    return cs, r


@pytest.mark.parametrize("fid", [0, 2])
def test_uint32_ops_identical_to_oracle_gadgets(fid):
    import kat_scenarios as S

    F = FIELDS[fid]
    rng = S.XorShift()
    for op in ("xor", "addmany", "sha256_maj", "sha256_ch"):
        for _ in range(12):
            a, b, c, d = (rng.next_u32() for _ in range(4))
            cs, r = _python_uint32_op(F, op, a, b, c, d)
            want = {"xor": a ^ b ^ c, "addmany": ((a ^ b) + c + d) & 0xFFFFFFFF, "sha256_maj": (a & b) ^ (a & c) ^ (b & c),
                    "sha256_ch": (a & b) ^ (~a & 0xFFFFFFFF & c)}[op]
            assert r.value == want
            with fixtures.Tcs(fid, device=-1, named=True) as t:
                value, n_const = t.uint32_op(op, a, b, c, d)
                assert value == want and n_const == sum(bit.kind == G.CONST for bit in r.bits)
                assert t.num_constraints() == cs.num_constraints()
                same(t.host_csr(), oracle_csr(cs))
                for row in range(cs.num_constraints()):
                    assert t.row_path(row) == cs.constraints[row][3]
            if op == "addmany":
                assert cs.constraints[-1][3] == "multieq 0" and "addition/result bit 0/boolean" in cs.named_objects
