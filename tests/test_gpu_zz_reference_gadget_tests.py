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
