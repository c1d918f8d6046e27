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


# ---- blake2s (crates/bellpepper/src/gadgets/blake2s.rs) ----
PERSON = b"12345678"


def bits_to_bytes_le(bits):
    vals = [int(b.get_value()) for b in bits]
    return bytes(sum(vals[i + j] << j for j in range(8)) for i in range(0, len(vals), 8))


def test_blake2s_blank_hash():  # blake2s.rs:420-441: no input -> 0 constraints, pinned digest
    cs = TestConstraintSystem(F)
    out = G.blake2s(cs, F, [], PERSON)
    assert cs.is_satisfied() and cs.num_constraints() == 0
    assert bits_to_bytes_le(out).hex() == "c59f682376d137f3f255e671e207d1f2374ebe504e9314208a52d9f88d69e8c8"


def test_blake2s_constraints():  # blake2s.rs:443-457: 512 allocated true bits -> 21518 constraints (incl. the 512 boolean ones)
    cs = TestConstraintSystem(F)
    bits = []
    for i in range(512):
        with cs.namespace(f"input bit {i}") as ns:
            bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, True)))
    G.blake2s(cs, F, bits, PERSON)
    assert cs.is_satisfied()
    assert cs.num_constraints() == 21518
    inst = c_api.from_python_cs(cs)
    assert inst.check(4, False) == -1


def _rng_bits(rng_bytes):
    return [G.Boolean.constant(b % 2 != 0) for b in rng_bytes]


def test_blake2s_precomp_and_constant_constraints():  # blake2s.rs:459-496
    # the reference draws `rng.next_u32() % 2`; the low bit of next_u32 is the low bit of `next_u32() as u8`
    const_bits = _rng_bits(G.xorshift_bytes(G.SEED_5D, 512))
    cs = TestConstraintSystem(F)
    bits = list(const_bits)
    for i in range(512):
        with cs.namespace(f"input bit {i}") as ns:
            bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, True)))
    G.blake2s(cs, F, bits, PERSON)
    assert cs.is_satisfied()
    assert cs.num_constraints() == 21518  # 512 fixed leading bits cost nothing
    cs2 = TestConstraintSystem(F)
    G.blake2s(cs2, F, const_bits, PERSON)
    assert cs2.num_constraints() == 0


@pytest.mark.parametrize("n", [0, 1, 7, 31, 32, 64, 65, 128, 248])
def test_blake2s_against_hashlib(n):  # blake2s.rs:498-555 (vs blake2s_simd there)
    data = G.xorshift_bytes(G.SEED_5D, n)
    cs = TestConstraintSystem(F)
    bits = alloc_bits(cs, data, be=False)
    out = G.blake2s(cs, F, bits, PERSON)
    assert cs.is_satisfied()
    assert bits_to_bytes_le(out) == hashlib.blake2s(data, digest_size=32, person=PERSON).digest()


def test_blake2s_256_vars():  # blake2s.rs:557-589: pinned input and digest
    data = bytes.fromhex(
        "be9f9c485e670acce8b1516a378176161b20583637b6f1c536fbc1158a0a3296831df2920e57a442d5738f4be4dd6be89dd7913fc8b4d1c0a815646a4d"
        "674b77f7caf313bd880bf759fcac27037c48c2b2a20acd2fd5248e3be426c84a341c0a3c63eaf36e0d537d10b8db5c6e4c801832c41eb1a3ed602177ac"
        "ded8b4b803bd34339d99a18b71df399641cc8dfae2ad193fcd74b5913e704551777160d14c78f2e8d5c32716a8599c1080cb89a40ccd6ba596694a8b4a"
        "065d9f2d0667ef423ed2e418093caff884540858b4f4b62acd47edcea880523e1b1cda8eb225c128c2e9e83f14f6e7448c5733a195cac7d79a53dde508"
        "3172462c45b2f799e42af1c9")
    assert len(data) == 256
    cs = TestConstraintSystem(F)
    out = G.blake2s(cs, F, alloc_bits(cs, data, be=False), PERSON)
    assert cs.is_satisfied()
    assert bits_to_bytes_le(out).hex() == "0af5695115ced92c8a0341e43869209636e9aa6472e4576f0f2b996cf812b30e"
    assert bits_to_bytes_le(out) == hashlib.blake2s(data, digest_size=32, person=PERSON).digest()


def test_blake2s_test_vectors():  # blake2s.rs:623-670: two 1024-byte inputs from the reference's RNG, pinned digests
    stream = G.xorshift_bytes(G.SEED_5D, 2048)
    for k, want in enumerate(["a1309e334376c8f36a736a4ab0e691ef931ee3ebdb9ea96187127136fea622a1",
                              "82fefff60f265cea255252f7c194a7f93965dffee0609ef74eb67f0d76cd41c6"]):
        data = stream[1024 * k: 1024 * k + 1024]
        assert hashlib.blake2s(data, digest_size=32, person=PERSON).hexdigest() == want  # pins the RNG restatement too
        if k == 0:  # the circuit itself once (16 blocks, ~350k constraints in pure Python)
            cs = TestConstraintSystem(F)
            out = G.blake2s(cs, F, alloc_bits(cs, data, be=False), PERSON)
            assert cs.is_satisfied()
            assert bits_to_bytes_le(out).hex() == want
