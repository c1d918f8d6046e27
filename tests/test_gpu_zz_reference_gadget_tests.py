"""GPU (-m gpu): the reference's own boolean / uint32 gadget tests (tests/kat_scenarios.py) with the DEVICE giving every
verdict.  Wherever a reference test asserts is_satisfied() / which_is_unsatisfied(), the rows and the witness of that moment
go through the C ABI (bp_cs_alloc / bp_cs_enforce / bp_cs_first_unsatisfied) and the first unsatisfied row must be the
oracle's -- the tests' own expectations (satisfied or not, WHICH constraint fails) then apply to the device's answer."""
import numpy as np
import pytest

import kat_scenarios as S
from oracle import c_api
from oracle.fields import FIELDS
from oracle.r1cs_py import TestConstraintSystem

pytestmark = pytest.mark.gpu

from gpu_util import Handle  # noqa: E402


def make(fid):
    F = FIELDS[fid]
    stats = {"device_checks": 0}

    def verdict(cs):
        want = cs.first_unsatisfied_row()
        if cs.num_constraints() == 0:  # nothing was enforced (all-constant operands): nothing to send
            return None
        lens, cols, coeffs, inputs, aux = cs.to_csr()
        lens, cols = np.asarray(lens, np.uint32), np.asarray(cols, np.uint32)
        coeffs, inputs, aux = c_api.ints_to_limbs(coeffs), c_api.ints_to_limbs(inputs), c_api.ints_to_limbs(aux)
        with Handle(fid) as h:
            h.load_instance(lens, cols, coeffs, inputs, aux)
            got = h.first_unsatisfied()
        assert got == want, (got, want)
        stats["device_checks"] += 1
        return None if got < 0 else cs.constraints[got][3]

    return F, (lambda: TestConstraintSystem(F)), verdict, stats


def test_allocated_bit_ops_enforce_equal_alloc_conditionally_on_device():
    F, new_cs, verdict, stats = make(0)
    S.allocated_bit_ops(new_cs, verdict)
    S.enforce_equal(new_cs, F, verdict)
    S.alloc_conditionally(new_cs, verdict)
    assert stats["device_checks"] == 48 + 48 + 4


def test_boolean_xor_and_or_over_operand_kinds_on_device():
    F, new_cs, verdict, stats = make(0)
    S.boolean_binops(new_cs, verdict)
    assert stats["device_checks"] == 192


def test_boolean_sha256_ch_maj_on_device():
    F, new_cs, verdict, stats = make(0)
    S.boolean_sha256_ch_maj(new_cs, F, verdict)
    assert stats["device_checks"] == 544


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_uint32_ops_on_device(fid):
    F, new_cs, verdict, stats = make(fid)
    S.uint32_xor(new_cs, verdict, 6)
    S.uint32_sha256_maj_ch(new_cs, F, verdict, 6)
    S.uint32_addmany(new_cs, F, verdict, 8)  # the flipped result bit breaks a MultiEq row (full-width coefficients 2^k)
    assert stats["device_checks"] == 6 + 12 + 24


# ---- the same boolean tests through the C++ front-end (csrc/host/gadgets.hpp) streaming into a device handle ----------------
def new_tcs():
    from bellpepper_b200 import fixtures

    return fixtures.Tcs(0, device=0, named=True)


@pytest.mark.parametrize("op", ["xor", "and", "or", "sha256_ch", "sha256_maj"])
def test_cpp_front_end_boolean_ops_on_device(op):
    """boolean.rs:1109-2003 with the reference's own flow: build, is_satisfied, get the result variable, set it wrong,
    which_is_unsatisfied names the gadget's constraint -- gadgets, paths and verdicts all on the product side."""
    three = op.startswith("sha256")
    n_flips = 0
    for a in S.VARIANTS:
        for b in S.VARIANTS:
            for c in (S.VARIANTS if three else ["True"]):
                va, vb, vc = S.val(a), S.val(b), S.val(c)
                expected = {"xor": va ^ vb, "and": va & vb, "or": va | vb, "sha256_ch": (va & vb) ^ ((not va) & vc),
                            "sha256_maj": (va & vb) ^ (va & vc) ^ (vb & vc)}[op]
                consts = [S.is_constant(x) for x in ((a, b, c) if three else (a, b))]
                with new_tcs() as t:
                    kind, value = t.boolean_op(op, a, b, c)
                    assert t.is_satisfied()
                    assert value == bool(expected), (op, a, b, c)
                    if all(consts):
                        assert t.num_constraints() == 0 and kind == "Constant"
                    if not any(consts):
                        if three:
                            var, con, v = op[7:], op[7:] + " computation", int(expected)
                        else:
                            var, con, v = S.binop_result(op, a, b)
                        assert t.get(var) == v
                        t.set(var, 1 - v)
                        assert t.which_is_unsatisfied() == con
                        t.set(var, v)
                        assert t.is_satisfied()
                        n_flips += 1
    assert n_flips == (64 if three else 16)


def test_cpp_front_end_enforce_equal_and_u64_bits_on_device():  # boolean.rs:935-1026, 1776-1794
    for a in S.VARIANTS:
        for b in S.VARIANTS:
            with new_tcs() as t:
                if S.is_constant(a) and S.is_constant(b) and S.val(a) != S.val(b):
                    with pytest.raises(RuntimeError, match="unsatisfiable"):
                        t.boolean_op("enforce_equal", a, b)
                    continue
                t.boolean_op("enforce_equal", a, b)
                assert t.is_satisfied() == (S.val(a) == S.val(b)), (a, b)
    with new_tcs() as t:
        bits = t.u64_bits(17234652694787248421)
        assert t.is_satisfied() and t.num_constraints() == 64
        assert sum(int(x) << i for i, x in enumerate(bits)) == 17234652694787248421
        t.set("bit 63/boolean", 2)
        assert t.which_is_unsatisfied() == "bit 63/boolean constraint"


def test_cpp_front_end_uint32_ops_on_device():
    """uint32.rs:492-780 through the C++ front-end on the device: satisfied, the expected word, and -- test_uint32_addmany's last
    step -- flipping "addition/result bit 0/boolean" leaves every bit boolean but breaks the MultiEq row."""
    rng = S.XorShift()
    for op in ("xor", "addmany", "sha256_maj", "sha256_ch"):
        for _ in range(6):
            a, b, c, d = (rng.next_u32() for _ in range(4))
            want = {"xor": a ^ b ^ c, "addmany": ((a ^ b) + c + d) & 0xFFFFFFFF, "sha256_maj": (a & b) ^ (a & c) ^ (b & c),
                    "sha256_ch": (a & b) ^ (~a & 0xFFFFFFFF & c)}[op]
            with new_tcs() as t:
                value, n_const = t.uint32_op(op, a, b, c, d)
                assert value == want
                assert t.is_satisfied()
                if op == "addmany":
                    assert n_const == 0
                    path = "addition/result bit 0/boolean"
                    old = t.get(path)
                    assert old == (want & 1)
                    t.set(path, 1 - old)
                    assert t.which_is_unsatisfied() == "multieq 0"
                    t.set(path, old)
                    assert t.is_satisfied()
