"""The reference's own known answers (tests/golden/reference_kats.json, extracted from /root/reference by
tests/golden/extract_reference_kats.py) against the oracle's gadget restatements -- the pin the GPU parity tests rest on."""
import hashlib
import json
import os

from oracle import gadgets_py as G
from oracle.fields import FIELDS
from oracle.r1cs_py import TestConstraintSystem

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kats.json")) as fh:
    KATS = json.load(fh)
F = FIELDS[0]  # blstrs::Scalar, the field every reference test runs on


def vals(entries):
    return [e["value"] for e in entries]


def test_golden_file_shape():
    assert vals(KATS["sha256"]["num_constraints_asserts"]) == [0, 25840, 44874]      # sha256.rs:296, 335, 361
    assert vals(KATS["blake2s"]["num_constraints_asserts"]) == [0, 21518, 21518, 0]  # blake2s.rs:427, 456, 479, 493
    assert KATS["sha256"]["rng_seed_bytes"]["value"] == 0x3D and KATS["blake2s"]["rng_seed_bytes"]["value"] == 0x5D
    assert KATS["test_cs"]["which_is_unsatisfied"]["value"] == "mult"


def test_sha256_structure_and_blank_digest_from_golden():
    counts = vals(KATS["sha256"]["num_constraints_asserts"])
    cs = TestConstraintSystem(F)
    bits = [G.Boolean.constant(False)] * 512
    bits[0] = G.Boolean.constant(True)
    out = G.sha256_compression_function(cs, F, bits, G.sha256_iv())
    assert cs.num_constraints() == counts[0]
    got = "".join(str(int(b.get_value())) for w in out for b in w.into_bits_be())
    assert int(got, 2).to_bytes(32, "big").hex() == KATS["sha256"]["blank_block_digest"]["value"]
    data = G.xorshift_bytes(G.SEED_3D, 64)
    cs = TestConstraintSystem(F)
    inp = []
    for i, byte in enumerate(data):
        for j in range(7, -1, -1):
            with cs.namespace(f"input bit {i} {j}") as ns:
                inp.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, bool((byte >> j) & 1))))
    G.sha256_compression_function(cs, F, inp, G.sha256_iv())
    assert cs.num_constraints() - 512 == counts[1] and cs.is_satisfied()


def test_blake2s_digests_from_golden():
    person = KATS["blake2s"]["personalization"]["value"].encode()
    digests = vals(KATS["blake2s"]["digests"])
    cs = TestConstraintSystem(F)
    out = G.blake2s(cs, F, [], person)
    bits = [int(b.get_value()) for b in out]
    assert bytes(sum(bits[i + j] << j for j in range(8)) for i in range(0, 256, 8)).hex() == digests[0]
    # the two 1024-byte RNG inputs (blake2s.rs:623-670): the digests pin the XorShift restatement as well
    stream = G.xorshift_bytes(G.SEED_5D, 2048)
    assert hashlib.blake2s(stream[:1024], digest_size=32, person=person).hexdigest() == digests[3]
    assert hashlib.blake2s(stream[1024:], digest_size=32, person=person).hexdigest() == digests[4]
