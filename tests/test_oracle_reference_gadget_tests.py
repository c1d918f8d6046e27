"""CPU: the reference's own boolean / uint32 gadget tests (tests/kat_scenarios.py) against the oracle -- the Python
restatement gives the verdicts, the C restatement must give the same first-unsatisfied row every time.  This pins the oracle's
check loop (test_cs.rs:239-253) and its gadget restatement on every expectation those tests hold."""
import pytest

import kat_scenarios as S
from oracle import c_api
from oracle.fields import FIELDS
from oracle.r1cs_py import TestConstraintSystem


def make(fid):
    F = FIELDS[fid]
    stats = {"checks": 0}

    def verdict(cs):
        py_row = cs.first_unsatisfied_row()
        if cs.num_constraints():
            inst = c_api.from_python_cs(cs)
            assert inst.check(1, True) == py_row and inst.check(2, False) == py_row
            inst.close()
        stats["checks"] += 1
        return None if py_row < 0 else cs.constraints[py_row][3]

    return F, (lambda: TestConstraintSystem(F)), verdict, stats


def test_allocated_bit_ops_enforce_equal_negation_alloc_conditionally():
    F, new_cs, verdict, stats = make(0)
    S.allocated_bit_ops(new_cs, verdict)
    S.enforce_equal(new_cs, F, verdict)
    S.boolean_negation(new_cs)
    S.alloc_conditionally(new_cs, verdict)
    assert stats["checks"] > 100


def test_boolean_xor_and_or_over_operand_kinds():
    F, new_cs, verdict, _ = make(0)
    S.boolean_binops(new_cs, verdict)


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_boolean_sha256_ch_maj(fid):
    F, new_cs, verdict, stats = make(fid)
    S.boolean_sha256_ch_maj(new_cs, F, verdict)
    assert stats["checks"] >= 2 * 216


def test_uint32_ops():
    F, new_cs, verdict, _ = make(0)
    S.uint32_rotr_shr()
    S.uint32_addmany_constants(new_cs, F, 100)
    S.uint32_xor(new_cs, verdict, 60)
    S.uint32_sha256_maj_ch(new_cs, F, verdict, 40)


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_uint32_addmany_and_the_flipped_result_bit(fid):
    F, new_cs, verdict, _ = make(fid)
    S.uint32_addmany(new_cs, F, verdict, 60)
