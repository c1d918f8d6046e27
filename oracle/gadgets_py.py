"""TEST INFRASTRUCTURE (oracle) -- Python restatement of the gadget circuits that produce BASELINE configs 1-3.

These build constraint systems through any object with the oracle's ConstraintSystem surface
(`alloc`, `enforce`, `namespace`, `get_root`); with `oracle.r1cs_py.TestConstraintSystem` they regenerate
the reference's structural and digest known-answers, which is what pins the oracle (tests/test_oracle_gadgets.py).

Restated from (paths relative to /root/reference):
  AllocatedBit, Boolean        crates/bellpepper-core/src/gadgets/boolean.rs:10-272, 369-766
  UInt32                       crates/bellpepper/src/gadgets/uint32.rs:14-406
  MultiEq                      crates/bellpepper/src/gadgets/multieq.rs:6-122
  sha256                       crates/bellpepper/src/gadgets/sha256.rs:16-272
  blake2s                      crates/bellpepper/src/gadgets/blake2s.rs:29-406
Values are Python bools/ints or None (unknown); a needed-but-unknown value raises AssignmentMissing.
"""

from __future__ import annotations

from typing import List, Optional

from .r1cs_py import ONE, AssignmentMissing, LinearCombination, Unsatisfiable, Variable

CAPACITY = 254  # ff::PrimeField::CAPACITY for all three fields


def _bit_value(v: Optional[bool]) -> int:
    if v is None:
        raise AssignmentMissing()
    return 1 if v else 0


class AllocatedBit:
    __slots__ = ("variable", "value")

    def __init__(self, variable: Variable, value: Optional[bool]):
        self.variable, self.value = variable, value

    @staticmethod
    def alloc(cs, value: Optional[bool]) -> "AllocatedBit":  # boolean.rs:68-97
        var = cs.alloc("boolean", lambda: _bit_value(value))
        cs.enforce("boolean constraint", lambda lc: lc + ONE - var, lambda lc: lc + var, lambda lc: lc)
        return AllocatedBit(var, value)

    @staticmethod
    def alloc_conditionally(cs, value, must_be_false: "AllocatedBit") -> "AllocatedBit":  # boolean.rs:27-64
        var = cs.alloc("boolean", lambda: _bit_value(value))
        cs.enforce("boolean constraint", lambda lc: lc + ONE - must_be_false.variable - var, lambda lc: lc + var, lambda lc: lc)
        return AllocatedBit(var, value)

    @staticmethod
    def _binop(cs, a, b, name, fn, ea, eb):
        box = [None]

        def f():
            if a.value is None or b.value is None:
                raise AssignmentMissing()
            box[0] = fn(a.value, b.value)
            return 1 if box[0] else 0

        r = cs.alloc(f"{name} result", f)
        cs.enforce(f"{name} constraint", ea, eb, lambda lc: lc + r)
        return AllocatedBit(r, box[0])

    @staticmethod
    def xor(cs, a, b):  # boolean.rs:101-151: (a + a) * b = a + b - c
        box = [None]

        def f():
            if a.value is None or b.value is None:
                raise AssignmentMissing()
            box[0] = a.value ^ b.value
            return 1 if box[0] else 0

        r = cs.alloc("xor result", f)
        cs.enforce("xor constraint", lambda lc: lc + a.variable + a.variable, lambda lc: lc + b.variable,
                   lambda lc: lc + a.variable + b.variable - r)
        return AllocatedBit(r, box[0])

    @staticmethod
    def and_(cs, a, b):  # boolean.rs:155-191
        return AllocatedBit._binop(cs, a, b, "and", lambda x, y: x and y, lambda lc: lc + a.variable, lambda lc: lc + b.variable)

    @staticmethod
    def and_not(cs, a, b):  # boolean.rs:195-231: a * (1 - b) = c
        return AllocatedBit._binop(cs, a, b, "and not", lambda x, y: x and not y, lambda lc: lc + a.variable,
                                   lambda lc: lc + ONE - b.variable)

    @staticmethod
    def nor(cs, a, b):  # boolean.rs:235-271: (1 - a) * (1 - b) = c
        return AllocatedBit._binop(cs, a, b, "nor", lambda x, y: (not x) and (not y), lambda lc: lc + ONE - a.variable,
                                   lambda lc: lc + ONE - b.variable)


IS, NOT, CONST = 0, 1, 2


class Boolean:
    """boolean.rs:368-376."""

    __slots__ = ("kind", "bit", "c")

    def __init__(self, kind: int, bit: Optional[AllocatedBit] = None, c: bool = False):
        self.kind, self.bit, self.c = kind, bit, c

    @staticmethod
    def constant(b: bool) -> "Boolean":
        return Boolean(CONST, None, bool(b))

    @staticmethod
    def from_bit(bit: AllocatedBit) -> "Boolean":
        return Boolean(IS, bit)

    def is_constant(self) -> bool:
        return self.kind == CONST

    def get_value(self) -> Optional[bool]:  # boolean.rs:429-435
        if self.kind == CONST:
            return self.c
        v = self.bit.value
        if v is None:
            return None
        return v if self.kind == IS else (not v)

    def lc(self, field, one: Variable, coeff: int) -> LinearCombination:  # boolean.rs:437-455
        z = LinearCombination.zero(field)
        if self.kind == CONST:
            return z + (coeff, one) if self.c else z
        if self.kind == IS:
            return z + (coeff, self.bit.variable)
        return z + (coeff, one) - (coeff, self.bit.variable)

    def not_(self) -> "Boolean":  # boolean.rs:463-469
        if self.kind == CONST:
            return Boolean.constant(not self.c)
        return Boolean(NOT if self.kind == IS else IS, self.bit)

    @staticmethod
    def xor(cs, a: "Boolean", b: "Boolean") -> "Boolean":  # boolean.rs:472-491
        if a.kind == CONST and not a.c:
            return b
        if b.kind == CONST and not b.c:
            return a
        if a.kind == CONST and a.c:
            return b.not_()
        if b.kind == CONST and b.c:
            return a.not_()
        if a.kind != b.kind:  # a XOR (NOT b) = NOT(a XOR b)
            is_, not_ = (a, b) if a.kind == IS else (b, a)
            return Boolean.xor(cs, is_, not_.not_()).not_()
        return Boolean.from_bit(AllocatedBit.xor(cs, a.bit, b.bit))

    @staticmethod
    def and_(cs, a: "Boolean", b: "Boolean") -> "Boolean":  # boolean.rs:494-516
        if (a.kind == CONST and not a.c) or (b.kind == CONST and not b.c):
            return Boolean.constant(False)
        if a.kind == CONST:
            return b
        if b.kind == CONST:
            return a
        if a.kind == IS and b.kind == NOT:
            return Boolean.from_bit(AllocatedBit.and_not(cs, a.bit, b.bit))
        if a.kind == NOT and b.kind == IS:
            return Boolean.from_bit(AllocatedBit.and_not(cs, b.bit, a.bit))
        if a.kind == NOT:
            return Boolean.from_bit(AllocatedBit.nor(cs, a.bit, b.bit))
        return Boolean.from_bit(AllocatedBit.and_(cs, a.bit, b.bit))

    @staticmethod
    def enforce_equal(cs, field, a: "Boolean", b: "Boolean"):  # boolean.rs:383-427
        if a.kind == CONST and b.kind == CONST:
            if a.c != b.c:
                raise Unsatisfiable()
            return
        for x, y in ((a, b), (b, a)):
            if x.kind == CONST and x.c:
                cs.enforce("enforce equal to one", lambda lc: lc, lambda lc: lc, lambda lc: lc + ONE - y.lc(field, ONE, 1))
                return
        for x, y in ((a, b), (b, a)):
            if x.kind == CONST and not x.c:
                cs.enforce("enforce equal to zero", lambda lc: lc, lambda lc: lc, lambda _: y.lc(field, ONE, 1))
                return
        cs.enforce("enforce equal", lambda lc: lc, lambda lc: lc, lambda _: a.lc(field, ONE, 1) - b.lc(field, ONE, 1))

    @staticmethod
    def sha256_ch(cs, field, a, b, c) -> "Boolean":  # boolean.rs:536-641
        va, vb, vc = a.get_value(), b.get_value(), c.get_value()
        ch_value = None if None in (va, vb, vc) else ((va and vb) != ((not va) and vc))
        if a.kind == CONST and b.kind == CONST and c.kind == CONST:
            return Boolean.constant(ch_value)
        if a.kind == CONST and not a.c:
            return c
        if b.kind == CONST and not b.c:
            return Boolean.and_(cs, a.not_(), c)
        if c.kind == CONST and not c.c:
            return Boolean.and_(cs, a, b)
        if c.kind == CONST and c.c:
            return Boolean.and_(cs, a, b.not_()).not_()
        if b.kind == CONST and b.c:
            return Boolean.and_(cs, a.not_(), c.not_()).not_()
        ch = cs.alloc("ch", lambda: _bit_value(ch_value))
        cs.enforce("ch computation", lambda _: b.lc(field, ONE, 1) - c.lc(field, ONE, 1), lambda _: a.lc(field, ONE, 1),
                   lambda lc: lc + ch - c.lc(field, ONE, 1))
        return Boolean.from_bit(AllocatedBit(ch, ch_value))

    @staticmethod
    def sha256_maj(cs, field, a, b, c) -> "Boolean":  # boolean.rs:644-759
        va, vb, vc = a.get_value(), b.get_value(), c.get_value()
        maj_value = None if None in (va, vb, vc) else ((va and vb) != (va and vc)) != (vb and vc)
        if a.kind == CONST and b.kind == CONST and c.kind == CONST:
            return Boolean.constant(maj_value)
        if a.kind == CONST and not a.c:
            return Boolean.and_(cs, b, c)
        if b.kind == CONST and not b.c:
            return Boolean.and_(cs, a, c)
        if c.kind == CONST and not c.c:
            return Boolean.and_(cs, a, b)
        if c.kind == CONST and c.c:
            return Boolean.and_(cs, a.not_(), b.not_()).not_()
        if b.kind == CONST and b.c:
            return Boolean.and_(cs, a.not_(), c.not_()).not_()
        if a.kind == CONST and a.c:
            return Boolean.and_(cs, b.not_(), c.not_()).not_()
        maj = cs.alloc("maj", lambda: _bit_value(maj_value))
        with cs.namespace("b and c") as ns:
            bc = Boolean.and_(ns, b, c)
        cs.enforce(
            "maj computation",
            lambda _: bc.lc(field, ONE, 1) + bc.lc(field, ONE, 1) - b.lc(field, ONE, 1) - c.lc(field, ONE, 1),
            lambda _: a.lc(field, ONE, 1),
            lambda _: bc.lc(field, ONE, 1) - maj,
        )
        return Boolean.from_bit(AllocatedBit(maj, maj_value))


class MultiEq:
    """multieq.rs:6-122: packs several narrow equalities into one wide constraint `lhs * 1 = rhs`.
    Rust flushes the tail in Drop; here the owner calls `finish()` when the scope ends."""

    def __init__(self, cs, field):
        self.cs, self.field = cs, field
        self.ops = 0
        self.bits_used = 0
        self.lhs = LinearCombination.zero(field)
        self.rhs = LinearCombination.zero(field)

    def _accumulate(self):
        lhs, rhs = self.lhs, self.rhs
        self.cs.enforce(f"multieq {self.ops}", lambda _: lhs, lambda lc: lc + ONE, lambda _: rhs)
        self.lhs = LinearCombination.zero(self.field)
        self.rhs = LinearCombination.zero(self.field)
        self.bits_used = 0
        self.ops += 1

    def enforce_equal(self, num_bits: int, lhs: LinearCombination, rhs: LinearCombination):
        if CAPACITY <= self.bits_used + num_bits:
            self._accumulate()
        assert CAPACITY > self.bits_used + num_bits
        coeff = pow(2, self.bits_used, self.field.p)
        self.lhs = self.lhs + (coeff, lhs)
        self.rhs = self.rhs + (coeff, rhs)
        self.bits_used += num_bits

    def finish(self):
        if self.bits_used > 0:
            self._accumulate()

    # ConstraintSystem surface, Root = Self (multieq.rs:69-122)
    def alloc(self, annotation, f):
        return self.cs.alloc(annotation, f)

    def alloc_input(self, annotation, f):
        return self.cs.alloc_input(annotation, f)

    def enforce(self, annotation, a, b, c):
        return self.cs.enforce(annotation, a, b, c)

    def push_namespace(self, name):
        self.cs.get_root().push_namespace(name)

    def pop_namespace(self):
        self.cs.get_root().pop_namespace()

    def get_root(self):
        return self

    def namespace(self, name):
        self.push_namespace(name)
        return _Ns(self)


class _Ns:
    """Namespace over an arbitrary root (pops on exit)."""

    def __init__(self, root):
        self._root = root

    def alloc(self, annotation, f):
        return self._root.alloc(annotation, f)

    def alloc_input(self, annotation, f):
        return self._root.alloc_input(annotation, f)

    def enforce(self, annotation, a, b, c):
        return self._root.enforce(annotation, a, b, c)

    def get_root(self):
        return self._root.get_root()

    def namespace(self, name):
        return self._root.get_root().namespace(name)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._root.get_root().pop_namespace()
        return False


class UInt32:
    """uint32.rs:14-18: 32 Booleans, least significant first, plus the known value if any."""

    __slots__ = ("bits", "value")

    def __init__(self, bits: List[Boolean], value: Optional[int]):
        self.bits, self.value = bits, value

    @staticmethod
    def constant(value: int) -> "UInt32":
        return UInt32([Boolean.constant((value >> i) & 1 == 1) for i in range(32)], value & 0xFFFFFFFF)

    @staticmethod
    def alloc(cs, value: Optional[int]) -> "UInt32":  # uint32.rs:43-72
        bits = []
        for i in range(32):
            with cs.namespace(f"allocated bit {i}") as ns:
                bits.append(Boolean.from_bit(AllocatedBit.alloc(ns, None if value is None else bool((value >> i) & 1))))
        return UInt32(bits, value)

    def into_bits_be(self) -> List[Boolean]:
        return list(reversed(self.bits))

    def into_bits(self) -> List[Boolean]:
        return list(self.bits)

    @staticmethod
    def _value_of(bits_le: List[Boolean]) -> Optional[int]:
        v = 0
        for i, b in enumerate(bits_le):
            bv = b.get_value()
            if bv is None:
                return None
            v |= int(bv) << i
        return v

    @staticmethod
    def from_bits_be(bits: List[Boolean]) -> "UInt32":  # uint32.rs:80-109
        assert len(bits) == 32
        le = list(reversed(bits))
        return UInt32(le, UInt32._value_of(le))

    @staticmethod
    def from_bits(bits: List[Boolean]) -> "UInt32":  # uint32.rs:118-163
        assert len(bits) == 32
        return UInt32(list(bits), UInt32._value_of(bits))

    def rotr(self, by: int) -> "UInt32":  # uint32.rs:165-181
        by %= 32
        v = None if self.value is None else ((self.value >> by) | (self.value << (32 - by))) & 0xFFFFFFFF
        return UInt32(self.bits[by:] + self.bits[:by], v)

    def shr(self, by: int) -> "UInt32":  # uint32.rs:183-201
        by %= 32
        v = None if self.value is None else self.value >> by
        return UInt32(self.bits[by:] + [Boolean.constant(False)] * by, v)

    def xor(self, cs, other: "UInt32") -> "UInt32":  # uint32.rs:281-303
        v = None if self.value is None or other.value is None else self.value ^ other.value
        bits = []
        for i, (a, b) in enumerate(zip(self.bits, other.bits)):
            with cs.namespace(f"xor of bit {i}") as ns:
                bits.append(Boolean.xor(ns, a, b))
        return UInt32(bits, v)

    @staticmethod
    def _triop(cs, field, a, b, c, tri_fn, circuit_fn, name):  # uint32.rs:203-236
        v = None if None in (a.value, b.value, c.value) else tri_fn(a.value, b.value, c.value) & 0xFFFFFFFF
        bits = []
        for i, (x, y, z) in enumerate(zip(a.bits, b.bits, c.bits)):
            with cs.namespace(f"{name} {i}") as ns:
                bits.append(circuit_fn(ns, field, x, y, z))
        return UInt32(bits, v)

    @staticmethod
    def sha256_maj(cs, field, a, b, c) -> "UInt32":
        return UInt32._triop(cs, field, a, b, c, lambda x, y, z: (x & y) ^ (x & z) ^ (y & z), Boolean.sha256_maj, "maj")

    @staticmethod
    def sha256_ch(cs, field, a, b, c) -> "UInt32":
        return UInt32._triop(cs, field, a, b, c, lambda x, y, z: (x & y) ^ ((~x) & z), Boolean.sha256_ch, "ch")

    @staticmethod
    def addmany(cs, field, operands: List["UInt32"]) -> "UInt32":  # uint32.rs:306-406; cs.get_root() is a MultiEq
        assert 2 <= len(operands) <= 10
        p = field.p
        max_value = len(operands) * 0xFFFFFFFF
        result_value: Optional[int] = 0
        lc = LinearCombination.zero(field)
        all_constants = True
        for op in operands:
            if op.value is None:
                result_value = None
            elif result_value is not None:
                result_value += op.value
            coeff = 1
            for bit in op.bits:
                lc = lc + bit.lc(field, ONE, coeff)
                all_constants &= bit.is_constant()
                coeff = (coeff * 2) % p
        modular_value = None if result_value is None else result_value & 0xFFFFFFFF
        if all_constants and modular_value is not None:
            return UInt32.constant(modular_value)
        result_bits = []
        result_lc = LinearCombination.zero(field)
        coeff, i = 1, 0
        while max_value != 0:
            with cs.namespace(f"result bit {i}") as ns:
                b = AllocatedBit.alloc(ns, None if result_value is None else bool((result_value >> i) & 1))
            result_lc = result_lc + (coeff, b.variable)
            result_bits.append(Boolean.from_bit(b))
            max_value >>= 1
            i += 1
            coeff = (coeff * 2) % p
        cs.get_root().enforce_equal(i, lc, result_lc)
        return UInt32(result_bits[:32], modular_value)


ROUND_CONSTANTS = [
    0x428A2F98, 0x71374491, 0xB5C0FBCF, 0xE9B5DBA5, 0x3956C25B, 0x59F111F1, 0x923F82A4, 0xAB1C5ED5,
    0xD807AA98, 0x12835B01, 0x243185BE, 0x550C7DC3, 0x72BE5D74, 0x80DEB1FE, 0x9BDC06A7, 0xC19BF174,
    0xE49B69C1, 0xEFBE4786, 0x0FC19DC6, 0x240CA1CC, 0x2DE92C6F, 0x4A7484AA, 0x5CB0A9DC, 0x76F988DA,
    0x983E5152, 0xA831C66D, 0xB00327C8, 0xBF597FC7, 0xC6E00BF3, 0xD5A79147, 0x06CA6351, 0x14292967,
    0x27B70A85, 0x2E1B2138, 0x4D2C6DFC, 0x53380D13, 0x650A7354, 0x766A0ABB, 0x81C2C92E, 0x92722C85,
    0xA2BFE8A1, 0xA81A664B, 0xC24B8B70, 0xC76C51A3, 0xD192E819, 0xD6990624, 0xF40E3585, 0x106AA070,
    0x19A4C116, 0x1E376C08, 0x2748774C, 0x34B0BCB5, 0x391C0CB3, 0x4ED8AA4A, 0x5B9CCA4F, 0x682E6FF3,
    0x748F82EE, 0x78A5636F, 0x84C87814, 0x8CC70208, 0x90BEFFFA, 0xA4506CEB, 0xBEF9A3F7, 0xC67178F2,
]
IV = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]


def sha256_compression_function(cs, field, input_bits: List[Boolean], current: List[UInt32]) -> List[UInt32]:
    """sha256.rs:83-272."""
    assert len(input_bits) == 512 and len(current) == 8
    w = [UInt32.from_bits_be(input_bits[32 * i: 32 * i + 32]) for i in range(16)]
    me = MultiEq(cs, field)
    for i in range(16, 64):
        with me.namespace(f"w extension {i}") as ns:
            s0 = w[i - 15].rotr(7)
            with ns.namespace("first xor for s0") as n2:
                s0 = s0.xor(n2, w[i - 15].rotr(18))
            with ns.namespace("second xor for s0") as n2:
                s0 = s0.xor(n2, w[i - 15].shr(3))
            s1 = w[i - 2].rotr(17)
            with ns.namespace("first xor for s1") as n2:
                s1 = s1.xor(n2, w[i - 2].rotr(19))
            with ns.namespace("second xor for s1") as n2:
                s1 = s1.xor(n2, w[i - 2].shr(10))
            with ns.namespace("computation of w[i]") as n2:
                w.append(UInt32.addmany(n2, field, [w[i - 16], s0, w[i - 7], s1]))

    def compute(maybe, ns, others):  # sha256.rs:128-148
        kind, v = maybe
        if kind == "concrete":
            return v
        return UInt32.addmany(ns, field, v + others)

    a = ("concrete", current[0])
    b, c, d = current[1], current[2], current[3]
    e = ("concrete", current[4])
    f, g, h = current[5], current[6], current[7]
    for i in range(64):
        with me.namespace(f"compression round {i}") as ns:
            with ns.namespace("deferred e computation") as n2:
                new_e = compute(e, n2, [])
            s1 = new_e.rotr(6)
            with ns.namespace("first xor for s1") as n2:
                s1 = s1.xor(n2, new_e.rotr(11))
            with ns.namespace("second xor for s1") as n2:
                s1 = s1.xor(n2, new_e.rotr(25))
            with ns.namespace("ch") as n2:
                ch = UInt32.sha256_ch(n2, field, new_e, f, g)
            temp1 = [h, s1, ch, UInt32.constant(ROUND_CONSTANTS[i]), w[i]]
            with ns.namespace("deferred a computation") as n2:
                new_a = compute(a, n2, [])
            s0 = new_a.rotr(2)
            with ns.namespace("first xor for s0") as n2:
                s0 = s0.xor(n2, new_a.rotr(13))
            with ns.namespace("second xor for s0") as n2:
                s0 = s0.xor(n2, new_a.rotr(22))
            with ns.namespace("maj") as n2:
                maj = UInt32.sha256_maj(n2, field, new_a, b, c)
            temp2 = [s0, maj]
            h, g, f = g, f, new_e
            e = ("deferred", temp1 + [d])
            d, c, b = c, b, new_a
            a = ("deferred", temp1 + temp2)

    def add2(name, x, y):
        with me.namespace(name) as ns:
            return UInt32.addmany(ns, field, [x, y])

    with me.namespace("deferred h0 computation") as ns:
        h0 = compute(a, ns, [current[0]])
    h1 = add2("new h1", current[1], b)
    h2 = add2("new h2", current[2], c)
    h3 = add2("new h3", current[3], d)
    with me.namespace("deferred h4 computation") as ns:
        h4 = compute(e, ns, [current[4]])
    h5 = add2("new h5", current[5], f)
    h6 = add2("new h6", current[6], g)
    h7 = add2("new h7", current[7], h)
    me.finish()  # MultiEq::drop (multieq.rs:61-67)
    return [h0, h1, h2, h3, h4, h5, h6, h7]


def sha256_iv() -> List[UInt32]:
    return [UInt32.constant(v) for v in IV]


def sha256_block_no_padding(cs, field, input_bits: List[Boolean]) -> List[Boolean]:  # sha256.rs:32-48
    out = sha256_compression_function(cs, field, input_bits, sha256_iv())
    return [b for word in out for b in word.into_bits_be()]


def sha256(cs, field, input_bits: List[Boolean]) -> List[Boolean]:  # sha256.rs:50-77
    assert len(input_bits) % 8 == 0
    padded = list(input_bits)
    plen = len(padded)
    padded.append(Boolean.constant(True))
    while (len(padded) + 64) % 512 != 0:
        padded.append(Boolean.constant(False))
    for i in range(63, -1, -1):
        padded.append(Boolean.constant((plen >> i) & 1 == 1))
    cur = sha256_iv()
    for i in range(len(padded) // 512):
        with cs.namespace(f"block {i}") as ns:
            cur = sha256_compression_function(ns, field, padded[512 * i: 512 * i + 512], cur)
    return [b for word in cur for b in word.into_bits_be()]


# ---- blake2s (crates/bellpepper/src/gadgets/blake2s.rs) -------------------------------------------------------------------
BLAKE2S_R = (16, 12, 8, 7)  # blake2s.rs:29-32
BLAKE2S_SIGMA = [  # blake2s.rs:50-61
    [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15],
    [14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3],
    [11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4],
    [7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8],
    [9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13],
    [2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9],
    [12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11],
    [13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10],
    [6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5],
    [10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0],
]


def blake2s_mixing_g(cs, field, v: List[UInt32], a: int, b: int, c: int, d: int, x: UInt32, y: UInt32):
    """blake2s.rs:86-121 (cs's root is a MultiEq)."""
    r1, r2, r3, r4 = BLAKE2S_R
    with cs.namespace("mixing step 1") as ns:
        v[a] = UInt32.addmany(ns, field, [v[a], v[b], x])
    with cs.namespace("mixing step 2") as ns:
        v[d] = v[d].xor(ns, v[a]).rotr(r1)
    with cs.namespace("mixing step 3") as ns:
        v[c] = UInt32.addmany(ns, field, [v[c], v[d]])
    with cs.namespace("mixing step 4") as ns:
        v[b] = v[b].xor(ns, v[c]).rotr(r2)
    with cs.namespace("mixing step 5") as ns:
        v[a] = UInt32.addmany(ns, field, [v[a], v[b], y])
    with cs.namespace("mixing step 6") as ns:
        v[d] = v[d].xor(ns, v[a]).rotr(r3)
    with cs.namespace("mixing step 7") as ns:
        v[c] = UInt32.addmany(ns, field, [v[c], v[d]])
    with cs.namespace("mixing step 8") as ns:
        v[b] = v[b].xor(ns, v[c]).rotr(r4)


BLAKE2S_G_ARGS = [(0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15), (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13),
                  (3, 4, 9, 14)]  # blake2s.rs:226-305


def blake2s_compression(cs, field, h: List[UInt32], m: List[UInt32], t: int, f: bool):
    """blake2s.rs:171-315; updates h in place."""
    assert len(h) == 8 and len(m) == 16
    v = list(h) + [UInt32.constant(x) for x in IV]
    with cs.namespace("first xor") as ns:
        v[12] = v[12].xor(ns, UInt32.constant(t & 0xFFFFFFFF))
    with cs.namespace("second xor") as ns:
        v[13] = v[13].xor(ns, UInt32.constant((t >> 32) & 0xFFFFFFFF))
    if f:
        with cs.namespace("third xor") as ns:
            v[14] = v[14].xor(ns, UInt32.constant(0xFFFFFFFF))
    me = MultiEq(cs, field)
    for i in range(10):
        with me.namespace(f"round {i}") as rns:
            s = BLAKE2S_SIGMA[i % 10]
            for j, (a, b, c, d) in enumerate(BLAKE2S_G_ARGS):
                with rns.namespace(f"mixing invocation {j + 1}") as ns:
                    blake2s_mixing_g(ns, field, v, a, b, c, d, m[s[2 * j]], m[s[2 * j + 1]])
    me.finish()  # MultiEq::drop at the end of the scope (blake2s.rs:224-306)
    for i in range(8):
        with cs.namespace(f"h[{i}] ^ v[{i}] ^ v[{i} + 8]") as ns:
            with ns.namespace("first xor") as n2:
                h[i] = h[i].xor(n2, v[i])
            with ns.namespace("second xor") as n2:
                h[i] = h[i].xor(n2, v[i + 8])


def blake2s(cs, field, input_bits: List[Boolean], personalization: bytes) -> List[Boolean]:
    """blake2s.rs:344-406; input bits little-endian per byte, output via into_bits (little-endian)."""
    assert len(personalization) == 8 and len(input_bits) % 8 == 0
    h = [UInt32.constant(x) for x in IV]
    h[0] = UInt32.constant(IV[0] ^ 0x01010000 ^ 32)
    h[6] = UInt32.constant(IV[6] ^ int.from_bytes(personalization[0:4], "little"))
    h[7] = UInt32.constant(IV[7] ^ int.from_bytes(personalization[4:8], "little"))
    blocks = []
    for b0 in range(0, len(input_bits), 512):
        block = input_bits[b0: b0 + 512]
        words = []
        for w0 in range(0, len(block), 32):
            tmp = list(block[w0: w0 + 32])
            tmp += [Boolean.constant(False)] * (32 - len(tmp))
            words.append(UInt32.from_bits(tmp))
        words += [UInt32.constant(0)] * (16 - len(words))
        blocks.append(words)
    if not blocks:
        blocks.append([UInt32.constant(0) for _ in range(16)])
    for i, block in enumerate(blocks[:-1]):
        with cs.namespace(f"block {i}") as ns:
            blake2s_compression(ns, field, h, block, (i + 1) * 64, False)
    with cs.namespace("final block") as ns:
        blake2s_compression(ns, field, h, blocks[-1], len(input_bits) // 8, True)
    return [b for word in h for b in word.into_bits()]


def xorshift_bytes(seed_bytes: bytes, n: int) -> bytes:
    """rand_xorshift 0.3 XorShiftRng::from_seed(seed).next_u32() as u8, n times (verified to
    regenerate the reference's pinned BLAKE2s digests)."""
    x, y, z, w = (int.from_bytes(seed_bytes[4 * i: 4 * i + 4], "little") for i in range(4))
    out = bytearray()
    for _ in range(n):
        t = (x ^ (x << 11)) & 0xFFFFFFFF
        x, y, z = y, z, w
        w = (w ^ (w >> 19) ^ t ^ (t >> 8)) & 0xFFFFFFFF
        out.append(w & 0xFF)
    return bytes(out)


SEED_3D = bytes([0x59, 0x62, 0xBE, 0x3D, 0x76, 0x3D, 0x31, 0x8D, 0x17, 0xDB, 0x37, 0x32, 0x54, 0x06, 0xBC, 0xE5])
SEED_5D = bytes([0x59, 0x62, 0xBE, 0x5D, 0x76, 0x3D, 0x31, 0x8D, 0x17, 0xDB, 0x37, 0x32, 0x54, 0x06, 0xBC, 0xE5])
