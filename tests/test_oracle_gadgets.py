"""Pin the oracle with the reference's structural and digest known-answers for the gadget circuits
(BLS12-381 Fr = blstrs::Scalar).  Paths relative to /root/reference."""
import hashlib

import numpy as np
import pytest

from oracle import c_api
from oracle import gadgets_py as G
from oracle.fields import FIELDS
from oracle.r1cs_py import TestConstraintSystem

F = FIELDS[0]


def alloc_bits(cs, data: bytes, be: bool = True):
    bits = []
    for i, byte in enumerate(data):
        order = range(7, -1, -1) if be else range(8)
        for j in order:
            with cs.namespace(f"input bit {i} {j}") as ns:
                bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, bool((byte >> j) & 1))))
    return bits


def bits_to_bytes_be(bits):
    vals = [int(b.get_value()) for b in bits]
    return bytes(int("".join(map(str, vals[i: i + 8])), 2) for i in range(0, len(vals), 8))


def test_blank_hash():  # crates/bellpepper/src/gadgets/sha256.rs:283-308: all-constant input -> 0 constraints
    iv = G.sha256_iv()
    cs = TestConstraintSystem(F)
    input_bits = [G.Boolean.constant(False)] * 512
    input_bits[0] = G.Boolean.constant(True)
    out = G.sha256_compression_function(cs, F, input_bits, iv)
    out_bits = [b for w in out for b in w.into_bits_be()]
    assert cs.is_satisfied() and cs.num_constraints() == 0
    expected = bytes.fromhex("e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855")
    assert all(b.is_constant() for b in out_bits)
    assert bits_to_bytes_be(out_bits) == expected


def test_full_block():  # sha256.rs:310-336: 512 allocated bits -> 25840 constraints beyond the 512 input-bit ones
    data = G.xorshift_bytes(G.SEED_3D, 64)
    cs = TestConstraintSystem(F)
    bits = alloc_bits(cs, data)
    G.sha256_compression_function(cs, F, bits, G.sha256_iv())
    assert cs.is_satisfied()
    assert cs.num_constraints() - 512 == 25840
    # and the C oracle agrees with the Python one on the whole circuit, including MultiEq's fat rows
    inst = c_api.from_python_cs(cs)
    assert inst.check(4, False) == -1
    lens = np.asarray(cs.to_csr()[0])
    assert lens.max() > 500  # fat MultiEq rows exist
    # flipping a result bit of an addition breaks the fat row that carries it (uint32.rs:627-633 idiom)
    path = "w extension 16/computation of w[i]/result bit 0/boolean"
    cs.set(path, 1 - cs.get(path))
    bad = cs.first_unsatisfied_row()
    assert bad >= 0
    p = F.p
    failing = [cs.constraints[i][3] for i in range(cs.num_constraints())
               if (lambda abc: (abc[0] * abc[1] - abc[2]) % p != 0)(cs.eval_row(i))]
    assert any("multieq " in n for n in failing)  # the fat row carrying w[16]'s addition fails too
    inst2 = c_api.from_python_cs(cs)
    assert inst2.check(1, True) == bad and inst2.check(4, False) == bad


def test_full_hash_count():  # sha256.rs:338-363: sha256() of 64 bytes (2 blocks) -> 44874 (+512)
    data = G.xorshift_bytes(G.SEED_3D, 64)
    cs = TestConstraintSystem(F)
    bits = alloc_bits(cs, data)
    out = G.sha256(cs, F, bits)
    assert cs.is_satisfied()
    assert cs.num_constraints() - 512 == 44874
    assert bits_to_bytes_be(out) == hashlib.sha256(data).digest()


@pytest.mark.parametrize("n", [0, 1, 31, 55, 56])
def test_against_vectors(n):  # sha256.rs:365-417: digest equals an independent SHA-256 for several lengths
    data = G.xorshift_bytes(G.SEED_3D, n)
    cs = TestConstraintSystem(F)
    bits = alloc_bits(cs, data)
    out = G.sha256(cs, F, bits)
    assert cs.is_satisfied()
    assert bits_to_bytes_be(out) == hashlib.sha256(data).digest()
