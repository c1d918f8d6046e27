"""Pin the oracle with the reference's own known-answer tests (BLS12-381 Fr = blstrs::Scalar).

Each test cites the reference test it restates (paths relative to /root/reference).  The same
scenarios run against the CUDA path in tests/test_gpu_kat.py.
"""
import numpy as np
import pytest

from oracle import c_api
from oracle.fields import FIELDS, FIELD_BLS12_381_FR
from oracle.r1cs_py import ONE, LinearCombination, TestConstraintSystem, Variable, WitnessCS, compute_path

F = FIELDS[FIELD_BLS12_381_FR]
P = F.p


def both(cs):
    """first-unsatisfied row from the Python oracle and from the C oracle (1 and 4 threads) must agree."""
    r = cs.first_unsatisfied_row()
    inst = c_api.from_python_cs(cs)
    assert inst.check(1, True) == r
    assert inst.check(4, False) == r
    bad, az, bz, cz = inst.eval(1)
    assert bad == r
    for i in range(cs.num_constraints()):
        assert tuple(c_api.limbs_to_ints(np.stack([az[i], bz[i], cz[i]]))) == cs.eval_row(i)
    return r


def test_compute_path():  # crates/bellpepper-core/src/util_cs/test_cs.rs:456-469
    assert compute_path(["hello", "world", "things"], "thing") == "hello/world/things/thing"


def test_cs():  # crates/bellpepper-core/src/util_cs/test_cs.rs:472-510
    cs = TestConstraintSystem(F)
    assert cs.is_satisfied() and cs.num_constraints() == 0
    with cs.namespace("a") as ns:
        a = ns.alloc("var", lambda: 10)
    with cs.namespace("b") as ns:
        b = ns.alloc("var", lambda: 4)
    c = cs.alloc("product", lambda: 40)
    cs.enforce("mult", lambda lc: lc + a, lambda lc: lc + b, lambda lc: lc + c)
    assert cs.is_satisfied() and cs.num_constraints() == 1 and both(cs) == -1
    cs.set("a/var", 4)
    cs.enforce("eq", lambda lc: lc + a, lambda lc: lc + ONE, lambda lc: lc + b)
    assert not cs.is_satisfied()
    assert cs.which_is_unsatisfied() == "mult" and both(cs) == 0
    assert cs.get("product") == 40
    cs.set("product", 16)
    assert cs.is_satisfied() and both(cs) == -1
    with cs.namespace("test1") as n1:
        with n1.namespace("test2") as n2:
            n2.alloc("hehe", lambda: 1)
    assert cs.get("test1/test2/hehe") == 1


def test_duplicate_path_and_slash_panic():  # test_cs.rs:325-333, 363-367
    cs = TestConstraintSystem(F)
    cs.alloc("x", lambda: 1)
    with pytest.raises(AssertionError):
        cs.alloc("x", lambda: 2)
    with pytest.raises(AssertionError):
        cs.alloc("a/b", lambda: 2)
    with pytest.raises(KeyError):
        cs.get("nope")


def test_alloc_closure_error_leaves_no_variable():  # test_cs.rs:386-390
    from oracle.r1cs_py import AssignmentMissing

    cs = TestConstraintSystem(F)

    def boom():
        raise AssignmentMissing()

    with pytest.raises(AssignmentMissing):
        cs.alloc("x", boom)
    assert cs.scalar_aux() == [] and "x" not in cs.named_objects


def test_allocated_bit():  # crates/bellpepper-core/src/gadgets/boolean.rs:777-788 + :86-91 (1-a)*a = 0
    for v, ok in ((0, True), (1, True), (2, False)):
        cs = TestConstraintSystem(F)
        a = cs.alloc("boolean", lambda: v)
        cs.enforce("boolean constraint", lambda lc: lc + ONE - a, lambda lc: lc + a, lambda lc: lc)
        assert cs.is_satisfied() == ok
        assert both(cs) == (-1 if ok else 0)
        if not ok:
            assert cs.which_is_unsatisfied() == "boolean constraint"


def test_xor_shape():  # boolean.rs:101-151: (a+a)*b = a+b-c ; lc + a + a merges to coefficient 2
    for av in (0, 1):
        for bv in (0, 1):
            cs = TestConstraintSystem(F)
            a = cs.alloc("a", lambda: av)
            b = cs.alloc("b", lambda: bv)
            c = cs.alloc("xor result", lambda: av ^ bv)
            cs.enforce("xor constraint", lambda lc: lc + a + a, lambda lc: lc + b, lambda lc: lc + a + b - c)
            A = cs.constraints[0][0]
            assert len(A) == 1 and list(A.iter())[0][1] == 2
            assert both(cs) == -1
            cs.set("xor result", 1 - (av ^ bv))
            assert both(cs) == 0


def test_num_wraps():  # crates/bellpepper-core/src/gadgets/num.rs:591-609: (p-1) + 1 == 0
    cs = TestConstraintSystem(F)
    a = cs.alloc("a", lambda: P - 1)
    s = cs.alloc("sum", lambda: 0)
    # (a + 1) * 1 = sum
    cs.enforce("add", lambda lc: lc + a + ONE, lambda lc: lc + ONE, lambda lc: lc + s)
    assert both(cs) == -1
    cs.set("sum", 1)
    assert both(cs) == 0


def test_num_squaring_and_mul():  # num.rs:612-638: 3^2 = 9, 12*10 = 120; perturb -> unsat
    cs = TestConstraintSystem(F)
    n = cs.alloc("a", lambda: 3)
    n2 = cs.alloc("squared num", lambda: 9)
    cs.enforce("squaring constraint", lambda lc: lc + n, lambda lc: lc + n, lambda lc: lc + n2)
    x = cs.alloc("x", lambda: 12)
    y = cs.alloc("y", lambda: 10)
    xy = cs.alloc("product num", lambda: 120)
    cs.enforce("multiplication constraint", lambda lc: lc + x, lambda lc: lc + y, lambda lc: lc + xy)
    assert both(cs) == -1
    cs.set("product num", 121)
    assert cs.which_is_unsatisfied() == "multiplication constraint" and both(cs) == 1
    cs.set("squared num", 10)
    assert cs.which_is_unsatisfied() == "squaring constraint" and both(cs) == 0


def test_nonzero_assertion():  # num.rs:676-693 + :373-399: a * inv = 1
    cs = TestConstraintSystem(F)
    a = cs.alloc("a", lambda: 3)
    inv = cs.alloc("ephemeral inverse", lambda: pow(3, -1, P))
    cs.enforce("nonzero assertion constraint", lambda lc: lc + a, lambda lc: lc + inv, lambda lc: lc + ONE)
    assert both(cs) == -1
    cs.set("a", 0)
    assert cs.which_is_unsatisfied() == "nonzero assertion constraint" and both(cs) == 0


def test_unpacking_255_bits_flip_each():  # num.rs:717-764 (fat 256-term LC 0*0 = sum 2^i b_i - x), reduced loop
    import random

    rng = random.Random(7)
    x = rng.randrange(P)
    cs = TestConstraintSystem(F)
    xv = cs.alloc("x", lambda: x)
    bits = []
    for i in range(255):
        bv = (x >> i) & 1
        b = cs.alloc(f"bit {i}", lambda bv=bv: bv)
        cs.enforce(f"bit {i} boolean", lambda lc: lc + ONE - b, lambda lc: lc + b, lambda lc: lc)
        bits.append(b)

    def packing(lc):
        for i, b in enumerate(bits):
            lc = lc + (pow(2, i, P), b)
        return lc - xv

    cs.enforce("unpacking constraint", lambda lc: lc, lambda lc: lc, packing)
    assert len(cs.constraints[-1][2]) == 256
    assert both(cs) == -1
    for i in (0, 1, 77, 253, 254):
        old = cs.get(f"bit {i}")
        cs.set(f"bit {i}", 1 - old)
        assert both(cs) == 255  # the packing row is the first (only) failing one: bit stays boolean
        cs.set(f"bit {i}", old)
    assert both(cs) == -1


def test_zero_coefficient_terms_are_retained():  # lc.rs:74-113: x - x stays as a 0-coeff term
    cs = TestConstraintSystem(F)
    a = cs.alloc("a", lambda: 5)
    cs.enforce("z", lambda lc: lc + a - a, lambda lc: lc + ONE, lambda lc: lc)
    A = cs.constraints[0][0]
    assert len(A) == 1 and list(A.iter())[0][1] == 0
    assert both(cs) == -1


def test_lc_order_inputs_before_aux_and_scaled_lc():  # lc.rs:155-160, 339-375
    cs = TestConstraintSystem(F)
    a = cs.alloc("a", lambda: 2)
    i1 = cs.alloc_input("i1", lambda: 3)
    b = cs.alloc("b", lambda: 4)
    lc = LinearCombination.zero(F) + b + a + i1 + ONE
    assert [repr(v) for v, _ in lc.iter()] == ["Input(0)", "Input(1)", "Aux(0)", "Aux(1)"]
    lc2 = LinearCombination.zero(F) + (5, lc) - (2, lc)
    assert [c for _, c in lc2.iter()] == [3, 3, 3, 3]
    assert lc2.eval(cs.scalar_inputs(), cs.scalar_aux()) == 3 * (1 + 3 + 2 + 4)


def test_verify_and_get_input():  # test_cs.rs:284-305
    cs = TestConstraintSystem(F)
    cs.alloc_input("in0", lambda: 7)
    cs.alloc_input("in1", lambda: 9)
    assert cs.verify([7, 9]) and not cs.verify([7, 8])
    assert cs.get_input(1, "in0") == 7 and cs.num_inputs() == 3
    with pytest.raises(AssertionError):
        cs.verify([7])


def test_one_is_mutable():  # test_cs.rs:160-169, 270-275: set("ONE", ..) changes w[0]
    cs = TestConstraintSystem(F)
    a = cs.alloc("a", lambda: 1)
    cs.enforce("c", lambda lc: lc + ONE, lambda lc: lc + ONE, lambda lc: lc + a)
    assert both(cs) == -1
    cs.set("ONE", 2)
    assert both(cs) == 0
    cs.set("a", 4)
    assert both(cs) == -1


def test_witness_cs():  # crates/bellpepper/src/util_cs/witness_cs.rs:94-201
    w = WitnessCS(F)
    assert w.input_assignment == [1] and w.aux_assignment == []
    v = w.alloc("x", lambda: 5)
    assert v == Variable(1, 0)
    vi = w.alloc_input("y", lambda: 6)
    assert vi == Variable(0, 1)
    w.enforce("nop", None, None, None)
    other = WitnessCS(F)
    other.alloc_input("z", lambda: 8)
    other.alloc("q", lambda: 9)
    w.extend(other)
    assert w.input_assignment == [1, 6, 8] and w.aux_assignment == [5, 9]
    a0, i0 = w.allocate_empty(2, 1)
    assert (a0, i0) == (2, 3) and w.aux_assignment[2:] == [0, 0] and w.input_assignment[3:] == [0]
    w.extend_aux([11])
    w.extend_inputs([12])
    assert w.aux_slice()[-1] == 11 and w.inputs_slice()[-1] == 12 and w.is_witness_generator() and WitnessCS.is_extensible()


def test_hash_is_structure_only():  # test_cs.rs:64-115, 214-237
    def build(v):
        cs = TestConstraintSystem(F)
        a = cs.alloc("a", lambda: v)
        cs.enforce("c", lambda lc: lc + a - a + ONE, lambda lc: lc + a, lambda lc: lc + a)
        return cs

    h1, h2 = build(1).hash(), build(2).hash()
    assert h1 == h2 and len(h1) == 64


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_c_oracle_vs_python_on_synthetic(fid):
    """C oracle == Python big-int on a synthetic instance, every row value (all three fields)."""
    from oracle import synth

    f = FIELDS[fid]
    n_vars, n_rows, t = 300, 60, 6
    lens, cols, coeffs, inputs, aux = c_api.synth_instance(fid, synth.SEED, t, n_vars, synth.N_INPUTS, n_rows)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    bad, az, bz, cz = inst.eval(2)
    w = [synth.witness(fid, synth.SEED, i) for i in range(n_vars)]
    first = -1
    for r in range(n_rows):
        vals = []
        for lc in range(3):
            acc = 0
            for tagged, c in synth.lc_terms(fid, synth.SEED, t, n_vars, synth.N_INPUTS, r, lc):
                idx = (tagged & 0x7FFFFFFF) + (synth.N_INPUTS if tagged >> 31 else 0)
                acc = (acc + c * w[idx]) % f.p
            vals.append(acc)
        assert c_api.limbs_to_ints(np.stack([az[r], bz[r], cz[r]])) == vals
        if first < 0 and (vals[0] * vals[1]) % f.p != vals[2]:
            first = r
    assert bad == first == 0
