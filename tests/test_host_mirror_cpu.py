"""CPU: the host-side halves of the Python mirror (bellpepper_b200/cs.py) that need no device -- the LinearCombination algebra
(lc.rs:35-375) and the path rules (test_cs.rs:363-375) -- against the oracle's restatement, which the reference's KATs pin.
(TestConstraintSystem itself cannot be built here: there is no CPU evaluation path.)"""
import random

import pytest

from bellpepper_b200 import cs as M
from bellpepper_b200 import ffi
from bellpepper_b200.fields import MODULUS
from oracle import r1cs_py as O
from oracle.fields import FIELDS


def flat_oracle(lc):
    cols, coeffs = lc.flat()
    return list(cols), list(coeffs)


@pytest.mark.parametrize("fid", [0, 1, 2])
def test_linear_combination_algebra_matches_the_oracle(fid):
    p = MODULUS[fid]
    F = FIELDS[fid]
    rng = random.Random(fid)
    for trial in range(60):
        a, b = M.LinearCombination.zero(p), O.LinearCombination.zero(F)
        saved = []  # earlier LCs to add / subtract / scale in
        for step in range(rng.randrange(1, 40)):
            kind = rng.choice((M.INPUT, M.AUX))
            idx = rng.randrange(0, 12)  # few indices: same-key merges, out-of-order inserts, the key+1 fast path
            vm, vo = M.Variable(kind, idx), O.Variable(kind, idx)
            coeff = rng.choice((1, 2, p - 1, 0, rng.randrange(p)))
            op = rng.randrange(8)
            if op == 0:
                a, b = a + vm, b + vo
            elif op == 1:
                a, b = a - vm, b - vo  # e.g. x - x stays as a zero-coefficient term (lc.rs:74-113)
            elif op == 2:
                a, b = a + (coeff, vm), b + (coeff, vo)
            elif op == 3:
                a, b = a - (coeff, vm), b - (coeff, vo)
            elif saved and op == 4:
                sm, so = rng.choice(saved)
                a, b = a + sm, b + so
            elif saved and op == 5:
                sm, so = rng.choice(saved)
                a, b = a - sm, b - so
            elif saved and op == 6:
                sm, so = rng.choice(saved)
                a, b = a + (coeff, sm), b + (coeff, so)  # (S, &LC): every coefficient scaled (lc.rs:339-375)
            elif saved:
                sm, so = rng.choice(saved)
                a, b = a - (coeff, sm), b - (coeff, so)
            cols_m, coeffs_m = a.flat()
            cols_o, coeffs_o = flat_oracle(b)
            assert list(cols_m) == cols_o and list(coeffs_m) == coeffs_o, (trial, step, op)
            assert len(a) == len(cols_o) and a.is_empty() == (len(cols_o) == 0)
            if rng.random() < 0.3:
                saved.append((a, b))
        # iteration order: inputs first, then aux, ascending (lc.rs:155-160)
        order = [(v.kind, v.index) for v, _ in a.iter()]
        assert order == sorted(order)
        assert all(c & ffi.COL_AUX for c in a.flat()[0][len(list(a.iter_inputs())):])


def test_from_coeff_from_variable_and_bad_operands():
    p = MODULUS[0]
    v = M.Variable(M.AUX, 3)
    assert M.LinearCombination.from_variable(p, v).flat() == ([3 | ffi.COL_AUX], [1])
    assert M.LinearCombination.from_coeff(p, v, p + 5).flat() == ([3 | ffi.COL_AUX], [5])
    assert M.Variable.new_unchecked(M.INPUT, 0).get_unchecked() == (M.INPUT, 0) and M.ONE == M.Variable(M.INPUT, 0)
    with pytest.raises(TypeError):
        M.LinearCombination.zero(p) + 3


def test_compute_path_rules():  # test_cs.rs:363-375, 456-469
    assert M.compute_path([], "a") == O.compute_path([], "a") == "a"
    assert M.compute_path(["x", "y"], "z") == O.compute_path(["x", "y"], "z") == "x/y/z"
    with pytest.raises(AssertionError):
        M.compute_path(["x"], "a/b")
