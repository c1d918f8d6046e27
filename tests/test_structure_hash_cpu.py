"""bp_structure_hash (host only): TestConstraintSystem::hash (test_cs.rs:64-115, 214-237) computed in C++ from the flat rows
that cross the C ABI equals the oracle's, on all three fields -- including LCs that hold zero coefficients (dropped by proc_lc)
and rows handed over unsorted or with a variable repeated (merged by proc_lc)."""
import ctypes
import random

import numpy as np
import pytest

from bellpepper_b200 import ffi, fixtures
from oracle import c_api
from oracle import gadgets_py as G
from oracle.fields import FIELDS
from oracle.r1cs_py import ONE, TestConstraintSystem


def c_hash(fid, n_inputs, n_aux, lens, cols, coeffs):
    L = ffi.load()
    out = ctypes.create_string_buffer(65)
    rc = L.bp_structure_hash(fid, n_inputs, n_aux, lens.size // 3, lens.ctypes.data if lens.size else None,
                             cols.ctypes.data if cols.size else None, coeffs.ctypes.data if cols.size else None, out)
    assert rc == 0, rc
    return out.value.decode()


def of_cs(fid, cs):
    lens, cols, coeffs, inputs, aux = cs.to_csr()
    return c_hash(fid, len(inputs), len(aux), np.asarray(lens, np.uint32), np.asarray(cols, np.uint32), c_api.ints_to_limbs(coeffs))


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_structure_hash_equals_the_oracle(fid):
    F = FIELDS[fid]
    cs = TestConstraintSystem(F)
    assert of_cs(fid, cs) == cs.hash()  # empty system
    a, b = cs.alloc("a", lambda: 3), cs.alloc("b", lambda: 4)
    c = cs.alloc_input("c", lambda: 12)
    cs.enforce("m", lambda lc: lc + a, lambda lc: lc + b, lambda lc: lc + c)
    cs.enforce("z", lambda lc: lc + a - a + (F.p - 5, b), lambda lc: lc + ONE + (7, c), lambda lc: lc)  # `a - a` stays as a 0 term
    assert of_cs(fid, cs) == cs.hash()
    # a whole sha256 block (fat MultiEq rows, coefficients 2^k): through the oracle's gadgets and through the C++ front-end
    cs = TestConstraintSystem(F)
    block = fixtures.xorshift_bytes(64)
    bits = []
    for i in range(512):
        with cs.namespace(f"input bit {i}") as ns:
            bits.append(G.Boolean.from_bit(G.AllocatedBit.alloc(ns, bool((block[i // 8] >> (7 - i % 8)) & 1))))
    G.sha256_compression_function(cs, F, bits, G.sha256_iv())
    want = cs.hash()
    assert of_cs(fid, cs) == want
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.sha256_block(block)
        lens, cols, coeffs, inputs, aux = t.host_csr()
    assert c_hash(fid, inputs.shape[0], aux.shape[0], lens, cols, coeffs) == want  # two front-ends, one fingerprint


def test_structure_hash_normalises_like_proc_lc():
    fid, F = 0, FIELDS[0]
    rng = random.Random(3)
    # one row whose A is handed over unsorted, with aux 2 three times (coefficients adding up to zero mod p) and input 1 twice
    x, y = rng.randrange(F.p), rng.randrange(F.p)
    terms = [(2 | ffi.COL_AUX, x), (1, 5), (2 | ffi.COL_AUX, y), (0 | ffi.COL_AUX, 9), (1, 6), (2 | ffi.COL_AUX, (-x - y) % F.p)]
    lens = np.asarray([len(terms), 0, 0], np.uint32)
    cols = np.asarray([c for c, _ in terms], np.uint32)
    coeffs = c_api.ints_to_limbs([v for _, v in terms])
    got = c_hash(fid, 2, 3, lens, cols, coeffs)
    # the same LC as the reference's algebra builds it: merged, sorted, the cancelled variable kept as a zero term
    cs = TestConstraintSystem(F)
    cs.alloc_input("i1", lambda: 0)
    v = [cs.alloc(f"v{i}", lambda: 0) for i in range(3)]
    from oracle.r1cs_py import INPUT, Variable

    i1 = Variable(INPUT, 1)
    cs.enforce("r", lambda lc: lc + (x, v[2]) + (5, i1) + (y, v[2]) + (9, v[0]) + (6, i1) + ((-x - y) % F.p, v[2]), lambda lc: lc, lambda lc: lc)
    assert got == cs.hash()
    # errors
    L = ffi.load()
    out = ctypes.create_string_buffer(65)
    assert L.bp_structure_hash(5, 1, 0, 0, None, None, None, out) == ffi.BP_E_ARG
    bad = c_api.ints_to_limbs([F.p])
    one = np.asarray([1, 0, 0], np.uint32)
    col = np.asarray([0], np.uint32)
    assert L.bp_structure_hash(0, 1, 0, 1, one.ctypes.data, col.ctypes.data, bad.ctypes.data, out) == ffi.BP_E_RANGE
