"""The reference's own known-answer scenarios through the B200-backed TestConstraintSystem mirror.
Same scenarios as tests/test_oracle_kat.py (which pins the oracle); paths relative to /root/reference."""
import random

import pytest

import bellpepper_b200 as bp
from bellpepper_b200 import ONE, LinearCombination, TestConstraintSystem, Variable, WitnessCS
from bellpepper_b200.fields import MODULUS

pytestmark = pytest.mark.gpu
P = MODULUS[0]


def test_cs():  # crates/bellpepper-core/src/util_cs/test_cs.rs:472-510
    cs = TestConstraintSystem.new()
    assert cs.is_satisfied() and cs.num_constraints() == 0
    with cs.namespace("a") as ns:
        a = ns.alloc("var", lambda: 10)
    with cs.namespace("b") as ns:
        b = ns.alloc("var", lambda: 4)
    c = cs.alloc("product", lambda: 40)
    cs.enforce("mult", lambda lc: lc + a, lambda lc: lc + b, lambda lc: lc + c)
    assert cs.is_satisfied() and cs.num_constraints() == 1
    cs.set("a/var", 4)
    cs.enforce("eq", lambda lc: lc + a, lambda lc: lc + ONE, lambda lc: lc + b)
    assert not cs.is_satisfied()
    assert cs.which_is_unsatisfied() == "mult"
    assert cs.get("product") == 40
    cs.set("product", 16)
    assert cs.is_satisfied()
    with cs.namespace("test1") as n1:
        with n1.namespace("test2") as n2:
            n2.alloc("hehe", lambda: 1)
    assert cs.get("test1/test2/hehe") == 1
    assert cs.scalar_aux() == [4, 4, 16, 1] and cs.scalar_inputs() == [1]


def test_panics():  # test_cs.rs:325-333, 363-367, 280, 321
    cs = TestConstraintSystem.new()
    cs.alloc("x", lambda: 1)
    with pytest.raises(AssertionError):
        cs.alloc("x", lambda: 2)
    with pytest.raises(AssertionError):
        cs.alloc("a/b", lambda: 2)
    with pytest.raises(KeyError):
        cs.get("nope")
    with pytest.raises(KeyError):
        cs.set("nope", 1)

    def boom():
        raise bp.AssignmentMissing()

    with pytest.raises(bp.SynthesisError):
        cs.alloc("y", boom)
    # the duplicate-path alloc pushed its value BEFORE the path check panicked (test_cs.rs:386-390), the
    # '/'-name one panicked in compute_path before pushing, the failing closure pushed nothing
    assert cs.scalar_aux() == [1, 2] and "y" not in cs.named_objects


def test_allocated_bit():  # crates/bellpepper-core/src/gadgets/boolean.rs:777-788, :86-91
    for v, ok in ((0, True), (1, True), (2, False)):
        cs = TestConstraintSystem.new()
        a = cs.alloc("boolean", lambda: v)
        cs.enforce("boolean constraint", lambda lc: lc + ONE - a, lambda lc: lc + a, lambda lc: lc)
        assert cs.is_satisfied() == ok
        if not ok:
            assert cs.which_is_unsatisfied() == "boolean constraint"


def test_xor():  # boolean.rs:101-151
    for av in (0, 1):
        for bv in (0, 1):
            cs = TestConstraintSystem.new()
            a = cs.alloc("a", lambda: av)
            b = cs.alloc("b", lambda: bv)
            c = cs.alloc("xor result", lambda: av ^ bv)
            cs.enforce("xor constraint", lambda lc: lc + a + a, lambda lc: lc + b, lambda lc: lc + a + b - c)
            assert cs.is_satisfied()
            cs.set("xor result", 1 - (av ^ bv))
            assert cs.which_is_unsatisfied() == "xor constraint"


def test_num_kats():  # crates/bellpepper-core/src/gadgets/num.rs:591-638, 676-693
    cs = TestConstraintSystem.new()
    a = cs.alloc("a", lambda: P - 1)
    s = cs.alloc("sum", lambda: 0)
    cs.enforce("add", lambda lc: lc + a + ONE, lambda lc: lc + ONE, lambda lc: lc + s)  # (p-1) + 1 wraps to 0
    n = cs.alloc("n", lambda: 3)
    n2 = cs.alloc("squared num", lambda: 9)
    cs.enforce("squaring constraint", lambda lc: lc + n, lambda lc: lc + n, lambda lc: lc + n2)
    x = cs.alloc("x", lambda: 12)
    y = cs.alloc("y", lambda: 10)
    xy = cs.alloc("product num", lambda: 120)
    cs.enforce("multiplication constraint", lambda lc: lc + x, lambda lc: lc + y, lambda lc: lc + xy)
    inv = cs.alloc("ephemeral inverse", lambda: pow(3, -1, P))
    cs.enforce("nonzero assertion constraint", lambda lc: lc + n, lambda lc: lc + inv, lambda lc: lc + ONE)
    assert cs.is_satisfied()
    az, bz, cz = cs.eval_all()
    assert az == [0, 3, 12, 3] and bz == [1, 3, 10, pow(3, -1, P)] and cz == [0, 9, 120, 1]
    cs.set("product num", 121)
    assert cs.which_is_unsatisfied() == "multiplication constraint"
    cs.set("squared num", 10)
    assert cs.which_is_unsatisfied() == "squaring constraint"
    cs.set("sum", 1)
    assert cs.which_is_unsatisfied() == "add"


def test_unpacking_255_bits():  # num.rs:717-764: a 256-term LC; every single-bit flip is caught
    rng = random.Random(7)
    x = rng.randrange(P)
    cs = TestConstraintSystem.new()
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
    assert cs.is_satisfied()
    for i in (0, 1, 100, 254):
        old = cs.get(f"bit {i}")
        cs.set(f"bit {i}", 1 - old)
        assert cs.which_is_unsatisfied() == "unpacking constraint"
        cs.set(f"bit {i}", 2)
        assert cs.which_is_unsatisfied() == f"bit {i} boolean"
        cs.set(f"bit {i}", old)
    assert cs.is_satisfied()


def test_one_is_mutable_and_verify():  # test_cs.rs:160-169, 270-275, 284-305
    cs = TestConstraintSystem.new()
    a = cs.alloc("a", lambda: 1)
    cs.alloc_input("in0", lambda: 7)
    cs.alloc_input("in1", lambda: 9)
    cs.enforce("c", lambda lc: lc + ONE, lambda lc: lc + ONE, lambda lc: lc + a)
    assert cs.is_satisfied()
    cs.set("ONE", 2)
    assert not cs.is_satisfied()
    cs.set("a", 4)
    assert cs.is_satisfied()
    assert cs.verify([7, 9]) and not cs.verify([7, 8]) and cs.get_input(1, "in0") == 7 and cs.num_inputs() == 3
    with pytest.raises(AssertionError):
        cs.verify([7])


def test_lc_eval_and_algebra():  # lc.rs:155-160, 245-267, 339-375
    cs = TestConstraintSystem.new()
    a = cs.alloc("a", lambda: 2)
    i1 = cs.alloc_input("i1", lambda: 3)
    b = cs.alloc("b", lambda: 4)
    lc = LinearCombination.zero(P) + b + a + i1 + ONE
    assert [repr(v) for v, _ in lc.iter()] == ["Variable(Input(0))", "Variable(Input(1))", "Variable(Aux(0))", "Variable(Aux(1))"]
    lc2 = LinearCombination.zero(P) + (5, lc) - (2, lc)
    assert [c for _, c in lc2.iter()] == [3, 3, 3, 3]
    assert lc2.eval(cs) == 30
    z = LinearCombination.zero(P) + a - a
    assert len(z) == 1 and list(z.iter())[0][1] == 0 and z.eval(cs) == 0  # zero coefficient retained (lc.rs:74-113)


def test_witness_cs():  # crates/bellpepper/src/util_cs/witness_cs.rs:94-201
    w = WitnessCS.new()
    assert w.input_assignment() == [1] and w.aux_assignment() == []
    assert w.alloc("x", lambda: 5) == Variable(1, 0)
    assert w.alloc_input("y", lambda: 6) == Variable(0, 1)
    w.enforce("nop", None, None, None)
    other = WitnessCS.new()
    other.alloc_input("z", lambda: 8)
    other.alloc("q", lambda: 9)
    w.extend(other)
    assert w.input_assignment() == [1, 6, 8] and w.aux_assignment() == [5, 9]
    a0, i0 = w.allocate_empty(2, 1)
    assert (a0, i0) == (2, 3) and w.aux_slice()[2:] == [0, 0] and w.inputs_slice()[3:] == [0]
    w.fill_aux(a0, [11, 12])
    w.fill_inputs(i0, [13])
    assert w.aux_slice() == [5, 9, 11, 12] and w.inputs_slice() == [1, 6, 8, 13]
    assert w.is_witness_generator() and WitnessCS.is_extensible()
    w2 = WitnessCS.from_assignments(0, [1, 2, 3], [4, 5])
    assert w2.to_assignments() == ([1, 2, 3], [4, 5])


def test_structure_hash_equals_oracle():  # test_cs.rs:64-115, 214-237: TestConstraintSystem::hash
    """Same little circuit through the product's mirror and through the oracle's: identical Blake2s structure fingerprints;
    a zero coefficient (x - x) is dropped by proc_lc, a changed coefficient changes the hash, witness values do not."""
    from oracle.fields import FIELDS
    from oracle import r1cs_py as O

    def build(mod, cs, one, coeff=3):
        a = cs.alloc("a", lambda: 10)
        b = cs.alloc_input("b", lambda: 4)
        c = cs.alloc("c", lambda: 40)
        cs.enforce("mult", lambda lc: lc + a, lambda lc: lc + b, lambda lc: lc + c)
        cs.enforce("lin", lambda lc: lc + (coeff, a) + one - one, lambda lc: lc + one, lambda lc: lc + (coeff, a) + b - b)
        cs.enforce("empty", lambda lc: lc, lambda lc: lc, lambda lc: lc)
        return cs

    got = build(None, TestConstraintSystem.new(), ONE)
    want = build(None, O.TestConstraintSystem(FIELDS[0]), O.ONE)
    assert got.hash() == want.hash()
    assert build(None, TestConstraintSystem.new(), ONE, coeff=5).hash() != got.hash()
    h0 = got.hash()
    got.set("a", 11)
    assert got.hash() == h0


def test_sized_witness_into_witness_cs():  # witness_cs.rs:7-41, 179-193: SizedWitness::generate_witness_into_cs
    from bellpepper_b200 import SizedWitness

    class Squares(SizedWitness):
        """aux[i] = (i + 2)^2, inputs = [sum of the aux values]; result = that sum."""

        def num_constraints(self):
            return 0

        def num_inputs(self):
            return 1

        def num_aux(self):
            return 300

        def generate_witness_into(self, aux, inputs):
            assert len(aux) == 300 and len(inputs) == 1
            for i in range(300):
                aux[i] = (i + 2) ** 2
            inputs[0] = sum(aux)
            return inputs[0]

    w = WitnessCS.new()
    w.alloc("first", lambda: 7)
    gen = Squares()
    result = gen.generate_witness_into_cs(w)
    aux, inputs = w.aux_slice(), w.inputs_slice()
    assert inputs == [1, result] and aux[0] == 7 and aux[1:] == [(i + 2) ** 2 for i in range(300)]
    assert gen.generate_witness() == (aux[1:], [result], result)
    w.close()


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_cpp_witness_cs_mirror(fid):  # witness_cs.rs:94-201 through the C++ host mirror (csrc/host/cs.hpp)
    from bellpepper_b200 import fixtures

    assert fixtures.witness_cs_selftest(fid, 0) == 0


def test_delta():  # crates/bellpepper-core/src/util_cs/mod.rs:39-76: Comparable::delta
    def build(extra_input=False, coeff=1, second=True, name="second"):
        cs = TestConstraintSystem.new()
        a = cs.alloc("a", lambda: 3)
        b = cs.alloc_input("b", lambda: 3)
        if extra_input:
            cs.alloc_input("c", lambda: 0)
        cs.enforce("first", lambda lc: lc + a, lambda lc: lc + ONE, lambda lc: lc + b)
        if second:
            cs.enforce(name, lambda lc: lc + (coeff, a), lambda lc: lc + ONE, lambda lc: lc + (coeff, b))
        return cs

    base = build()
    assert base.delta(build()) == ("Equal",)
    assert base.delta(build(extra_input=True)) == ("InputCountMismatch", 2, 3)
    assert base.delta(build(second=False)) == ("ConstraintCountMismatch", 2, 1)
    assert base.delta(build(coeff=5)) == ("ConstraintMismatch", 1, "second", "second")
    assert base.delta(build(name="renamed")) == ("ConstraintMismatch", 1, "second", "renamed")
    assert base.delta(build(extra_input=True, coeff=5), ignore_counts=True) == ("ConstraintMismatch", 1, "second", "second")
    assert base.delta(build(extra_input=True), ignore_counts=True) == ("Different",)
