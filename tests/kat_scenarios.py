"""The reference's own gadget tests, restated once and run against two evaluators (TEST INFRASTRUCTURE).

Each scenario builds its circuits with the oracle's gadget restatement (oracle/gadgets_py.py) on the oracle's
TestConstraintSystem and asks `verdict(cs)` for `which_is_unsatisfied()` wherever the reference test asserts
`is_satisfied()` / `which_is_unsatisfied()`:
  * tests/test_oracle_reference_gadget_tests.py (CPU): verdict = the Python oracle, cross-checked with the C oracle;
  * tests/test_gpu_zz_reference_gadget_tests.py (GPU): verdict = the same rows and witness through the C ABI on the device.
So the reference's expectations pin the oracle, and the oracle's verdicts (which row fails first, by path) pin the CUDA path.

Mirrored tests (paths relative to /root/reference):
  crates/bellpepper-core/src/gadgets/boolean.rs:791-933 (AllocatedBit xor / and / and_not / nor), 935-1026 (enforce_equal),
  1028-1070 (negation), 1109-1316 (Boolean::xor x 36 operand kinds), 1318-1548 (and), 1550-1774 (or), 1823-2003 (sha256_ch /
  sha256_maj x 216), 2006-2070 (alloc_conditionally);
  crates/bellpepper/src/gadgets/uint32.rs:492-535 (xor), 537-578 (addmany of constants), 581-635 (addmany + the result-bit
  flip that breaks a MultiEq row), 638-692 (rotr, shr), 694-780 (sha256_maj, sha256_ch).
"""
from oracle import gadgets_py as G
from oracle.r1cs_py import AssignmentMissing, Unsatisfiable

VARIANTS = ["True", "False", "AllocatedTrue", "AllocatedFalse", "NegatedAllocatedTrue", "NegatedAllocatedFalse"]


def is_constant(op):  # boolean.rs:1085-1094
    return op in ("True", "False")


def val(op):  # boolean.rs:1096-1105
    return op in ("True", "AllocatedTrue", "NegatedAllocatedFalse")


def construct(cs, op, name):  # the tests' dyn_construct
    with cs.namespace(name) as ns:
        if is_constant(op):
            return G.Boolean.constant(op == "True")
        b = G.Boolean.from_bit(G.AllocatedBit.alloc(ns, op in ("AllocatedTrue", "NegatedAllocatedTrue")))
        return b.not_() if op.startswith("Negated") else b


def boolean_or(cs, a, b):  # boolean.rs:519-533
    with cs.namespace("not and (not a) (not b)") as ns:
        return G.Boolean.and_(ns, a.not_(), b.not_()).not_()


class XorShift:
    """rand_xorshift 0.3 XorShiftRng (the generator of the reference's tests), seeded from 16 bytes."""

    def __init__(self, seed=G.SEED_5D):
        self.s = [int.from_bytes(seed[4 * i: 4 * i + 4], "little") for i in range(4)]

    def next_u32(self):
        x, y, z, w = self.s
        t = (x ^ (x << 11)) & 0xFFFFFFFF
        w2 = (w ^ (w >> 19) ^ t ^ (t >> 8)) & 0xFFFFFFFF
        self.s = [y, z, w, w2]
        return w2


def flip(cs, verdict, path, expect_path=None):
    """Invert a bit-valued variable: the system must stop being satisfied (at `expect_path` when given); then restore it."""
    old = cs.get(path)
    assert old in (0, 1)
    cs.set(path, 1 - old)
    got = verdict(cs)
    assert got is not None
    if expect_path is not None:
        assert got == expect_path
    cs.set(path, old)
    assert verdict(cs) is None


# ---- boolean.rs ------------------------------------------------------------------------------------------------------------
def allocated_bit_ops(new_cs, verdict):  # boolean.rs:791-933
    ops = {"xor": (G.AllocatedBit.xor, lambda a, b: a ^ b), "and": (G.AllocatedBit.and_, lambda a, b: a & b),
           "and not": (G.AllocatedBit.and_not, lambda a, b: a & (not b)), "nor": (G.AllocatedBit.nor, lambda a, b: (not a) & (not b))}
    for name, (fn, truth) in ops.items():
        for a_val in (False, True):
            for b_val in (False, True):
                cs = new_cs()
                with cs.namespace("a") as ns:
                    a = G.AllocatedBit.alloc(ns, a_val)
                with cs.namespace("b") as ns:
                    b = G.AllocatedBit.alloc(ns, b_val)
                c = fn(cs, a, b)
                assert c.value == bool(truth(a_val, b_val))
                assert verdict(cs) is None
                assert cs.get("a/boolean") == int(a_val) and cs.get("b/boolean") == int(b_val)
                assert cs.get(f"{name} result") == int(truth(a_val, b_val))
                flip(cs, verdict, f"{name} result", f"{name} constraint")


def enforce_equal(new_cs, field, verdict):  # boolean.rs:935-1026
    for a_bool in (False, True):
        for b_bool in (False, True):
            for a_neg in (False, True):
                for b_neg in (False, True):
                    want = (a_bool ^ a_neg) == (b_bool ^ b_neg)
                    for a_const, b_const in ((False, False), (True, False), (False, True), (True, True)):
                        cs = new_cs()
                        if a_const:
                            a = G.Boolean.constant(a_bool)
                        else:
                            with cs.namespace("a") as ns:
                                a = G.Boolean.from_bit(G.AllocatedBit.alloc(ns, a_bool))
                        if b_const:
                            b = G.Boolean.constant(b_bool)
                        else:
                            with cs.namespace("b") as ns:
                                b = G.Boolean.from_bit(G.AllocatedBit.alloc(ns, b_bool))
                        a = a.not_() if a_neg else a
                        b = b.not_() if b_neg else b
                        if a_const and b_const:
                            try:
                                G.Boolean.enforce_equal(cs, field, a, b)
                                ok = True
                            except Unsatisfiable:
                                ok = False
                            assert ok == want
                            if ok:
                                assert verdict(cs) is None
                        else:
                            G.Boolean.enforce_equal(cs, field, a, b)
                            assert (verdict(cs) is None) == want


def boolean_negation(new_cs):  # boolean.rs:1028-1070
    cs = new_cs()
    b = G.Boolean.from_bit(G.AllocatedBit.alloc(cs, True))
    assert b.kind == G.IS
    b = b.not_()
    assert b.kind == G.NOT
    b = b.not_()
    assert b.kind == G.IS
    b = G.Boolean.constant(True)
    assert b.kind == G.CONST and b.c is True
    b = b.not_()
    assert b.kind == G.CONST and b.c is False
    b = b.not_()
    assert b.kind == G.CONST and b.c is True


def binop_result(op, first, second):
    """Both operands allocated: (path of the variable the operation allocates, path of its constraint, the variable's value).
    xor works on the bits under the negations; and picks and / and not / nor by the operands' polarity; or = not(and(not a,
    not b)) inside the namespace "not and (not a) (not b)" (boolean.rs:472-533; the tables at :1154-1771)."""
    na, nb = first.startswith("Negated"), second.startswith("Negated")
    ba, bb = first.endswith("AllocatedTrue"), second.endswith("AllocatedTrue")
    if op == "xor":
        return "xor result", "xor constraint", int(ba ^ bb)
    if op == "and":
        name = "nor" if na and nb else ("and not" if na or nb else "and")
        return f"{name} result", f"{name} constraint", int(val(first) & val(second))
    assert op == "or"
    name = "and" if na and nb else ("and not" if na or nb else "nor")
    ns = "not and (not a) (not b)"
    return f"{ns}/{name} result", f"{ns}/{name} constraint", int(not (val(first) | val(second)))


def boolean_binops(new_cs, verdict):
    """Boolean::xor / and / or over every pair of operand kinds (boolean.rs:1109-1774).  The reference spells out a 36-row table
    per operation; the rows follow one rule each, stated here, and the named result variable is checked where one exists."""
    for first in VARIANTS:
        for second in VARIANTS:
            ca, cb = is_constant(first), is_constant(second)
            na, nb = first.startswith("Negated"), second.startswith("Negated")

            # xor (table 1154-1313): constants fold; a true constant negates the other; Is^Not -> Not(xor); value = bits' xor
            cs = new_cs()
            a, b = construct(cs, first, "a"), construct(cs, second, "b")
            c = G.Boolean.xor(cs, a, b)
            assert verdict(cs) is None
            assert c.get_value() == (val(first) ^ val(second))
            if ca and cb:
                assert c.kind == G.CONST and cs.num_constraints() == 0
            elif ca or cb:
                other_neg, const_true = (nb, first == "True") if ca else (na, second == "True")
                assert c.kind == (G.NOT if other_neg ^ const_true else G.IS)
            else:
                assert c.kind == (G.NOT if na ^ nb else G.IS)
                var, con, v = binop_result("xor", first, second)
                assert cs.get(var) == v and c.bit.value == bool(v)
                flip(cs, verdict, var, con)

            # and (table 1366-1545): false folds to Constant(false); true returns the other; Is&Is -> and, Is&Not -> and not,
            # Not&Not -> nor; the result is always Is
            cs = new_cs()
            a, b = construct(cs, first, "a"), construct(cs, second, "b")
            c = G.Boolean.and_(cs, a, b)
            assert verdict(cs) is None
            assert c.get_value() == (val(first) & val(second))
            if first == "False" or second == "False":
                assert c.kind == G.CONST and c.c is False
            elif ca and cb:
                assert c.kind == G.CONST and c.c is True
            elif ca or cb:
                assert c.kind == (G.NOT if (nb if ca else na) else G.IS)
            else:
                var, con, v = binop_result("and", first, second)
                assert c.kind == G.IS and cs.get(var) == v
                flip(cs, verdict, var, con)

            # or = not(and(not a, not b)) in the namespace "not and (not a) (not b)" (table 1596-1771)
            cs = new_cs()
            a, b = construct(cs, first, "a"), construct(cs, second, "b")
            c = boolean_or(cs, a, b)
            assert verdict(cs) is None
            assert c.get_value() == (val(first) | val(second))
            if first == "True" or second == "True":
                assert c.kind == G.CONST and c.c is True
            elif ca and cb:
                assert c.kind == G.CONST and c.c is False
            elif ca or cb:
                assert c.kind == (G.NOT if (nb if ca else na) else G.IS)
            else:
                # not a / not b are Is where the operand was negated: Is&Is -> and, mixed -> and not, Not&Not -> nor
                var, con, v = binop_result("or", first, second)
                assert c.kind == G.NOT and cs.get(var) == v
                flip(cs, verdict, var, con)


def boolean_sha256_ch_maj(new_cs, field, verdict):  # boolean.rs:1823-2003
    for which in ("ch", "maj"):
        for first in VARIANTS:
            for second in VARIANTS:
                for third in VARIANTS:
                    cs = new_cs()
                    va, vb, vc = val(first), val(second), val(third)
                    expected = ((va & vb) ^ ((not va) & vc)) if which == "ch" else ((va & vb) ^ (va & vc) ^ (vb & vc))
                    a, b, c = construct(cs, first, "a"), construct(cs, second, "b"), construct(cs, third, "c")
                    r = (G.Boolean.sha256_ch if which == "ch" else G.Boolean.sha256_maj)(cs, field, a, b, c)
                    assert verdict(cs) is None
                    assert r.get_value() == bool(expected)
                    consts = [is_constant(first), is_constant(second), is_constant(third)]
                    if any(consts):
                        if all(consts):
                            assert cs.num_constraints() == 0
                    else:
                        assert cs.get(which) == int(expected)
                        cs.set(which, 1 - int(expected))
                        assert verdict(cs) == f"{which} computation"


def alloc_conditionally(new_cs, verdict):  # boolean.rs:2006-2070
    cs = new_cs()
    b = G.AllocatedBit.alloc(cs, False)
    try:
        with cs.namespace("alloc_conditionally") as ns:
            G.AllocatedBit.alloc_conditionally(ns, None, b)
        raise AssertionError("a missing value must be an error")
    except AssignmentMissing:
        pass
    for value, b_val, want in ((True, False, True), (True, True, False), (False, False, True), (False, True, True)):
        cs = new_cs()
        b = G.AllocatedBit.alloc(cs, b_val)
        with cs.namespace("alloc_conditionally") as ns:
            got = G.AllocatedBit.alloc_conditionally(ns, value, b)
        assert got.value == value
        assert (verdict(cs) is None) == want


# ---- uint32.rs -------------------------------------------------------------------------------------------------------------
def check_bits(r, expected):  # the loop every uint32 test ends with
    assert r.value == expected
    for b in r.bits:
        if b.kind == G.IS:
            assert b.bit.value == bool(expected & 1)
        elif b.kind == G.NOT:
            assert b.bit.value != bool(expected & 1)
        else:
            assert b.c == bool(expected & 1)
        expected >>= 1


def uint32_xor(new_cs, verdict, n_iter):  # uint32.rs:492-535
    rng = XorShift()
    for _ in range(n_iter):
        cs = new_cs()
        a, b, c = rng.next_u32(), rng.next_u32(), rng.next_u32()
        with cs.namespace("a_bit") as ns:
            a_bit = G.UInt32.alloc(ns, a)
        b_bit = G.UInt32.constant(b)
        with cs.namespace("c_bit") as ns:
            c_bit = G.UInt32.alloc(ns, c)
        with cs.namespace("first xor") as ns:
            r = G.UInt32.xor(a_bit, ns, b_bit)
        with cs.namespace("second xor") as ns:
            r = G.UInt32.xor(r, ns, c_bit)
        assert verdict(cs) is None
        check_bits(r, a ^ b ^ c)


def uint32_addmany_constants(new_cs, field, n_iter):  # uint32.rs:537-578
    rng = XorShift()
    for _ in range(n_iter):
        cs = new_cs()
        a, b, c = rng.next_u32(), rng.next_u32(), rng.next_u32()
        me = G.MultiEq(cs, field)
        with me.namespace("addition") as ns:
            r = G.UInt32.addmany(ns, field, [G.UInt32.constant(a), G.UInt32.constant(b), G.UInt32.constant(c)])
        me.finish()
        assert all(bit.kind == G.CONST for bit in r.bits) and cs.num_constraints() == 0
        check_bits(r, (a + b + c) & 0xFFFFFFFF)


def uint32_addmany(new_cs, field, verdict, n_iter):  # uint32.rs:581-635
    rng = XorShift()
    for _ in range(n_iter):
        cs = new_cs()
        a, b, c, d = rng.next_u32(), rng.next_u32(), rng.next_u32(), rng.next_u32()
        expected = ((a ^ b) + c + d) & 0xFFFFFFFF
        with cs.namespace("a_bit") as ns:
            a_bit = G.UInt32.alloc(ns, a)
        b_bit, c_bit = G.UInt32.constant(b), G.UInt32.constant(c)
        with cs.namespace("d_bit") as ns:
            d_bit = G.UInt32.alloc(ns, d)
        with cs.namespace("xor") as ns:
            r = G.UInt32.xor(a_bit, ns, b_bit)
        me = G.MultiEq(cs, field)
        with me.namespace("addition") as ns:
            r = G.UInt32.addmany(ns, field, [r, c_bit, d_bit])
        me.finish()  # (Drop of the MultiEq: the wide row `lhs * 1 = rhs` is emitted here)
        assert verdict(cs) is None
        assert all(bit.kind != G.CONST for bit in r.bits)
        check_bits(r, expected)
        # flip a bit and see if the addition constraint still works: the bit stays boolean, so what fails is the MultiEq row
        flip(cs, verdict, "addition/result bit 0/boolean", "multieq 0")


def uint32_rotr_shr():  # uint32.rs:638-692 (constants only: no constraint system involved)
    rng = XorShift()
    num = rng.next_u32()
    a = G.UInt32.constant(num)
    for i in range(32):
        b = a.rotr(i)
        assert len(b.bits) == 32 and all(x.kind == G.CONST for x in b.bits)
        check_bits(b, num)
        num = ((num >> 1) | (num << 31)) & 0xFFFFFFFF
    rng = XorShift()
    for _ in range(50):
        for i in range(60):
            num = rng.next_u32()
            want = num >> (i % 32)  # wrapping_shr
            a = G.UInt32.constant(num).shr(i)
            assert a.value == want and [x.get_value() for x in a.bits] == [x.get_value() for x in G.UInt32.constant(want).bits]


def uint32_sha256_maj_ch(new_cs, field, verdict, n_iter):  # uint32.rs:694-780
    for which in ("maj", "ch"):
        rng = XorShift()
        for _ in range(n_iter):
            cs = new_cs()
            a, b, c = rng.next_u32(), rng.next_u32(), rng.next_u32()
            expected = ((a & b) ^ (a & c) ^ (b & c)) if which == "maj" else ((a & b) ^ (~a & 0xFFFFFFFF & c))
            with cs.namespace("a_bit") as ns:
                a_bit = G.UInt32.alloc(ns, a)
            b_bit = G.UInt32.constant(b)
            with cs.namespace("c_bit") as ns:
                c_bit = G.UInt32.alloc(ns, c)
            r = (G.UInt32.sha256_maj if which == "maj" else G.UInt32.sha256_ch)(cs, field, a_bit, b_bit, c_bit)
            assert verdict(cs) is None
            check_bits(r, expected)
