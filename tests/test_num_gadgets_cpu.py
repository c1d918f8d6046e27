"""num gadgets of the C++ front-end (csrc/host/gadgets.hpp: AllocatedNum, to_bits_le[_strict], mul / square / add /
assert_nonzero / conditionally_reverse; reference: crates/bellpepper-core/src/gadgets/num.rs) recorded on the host and checked
by the ORACLE: structure (row and variable counts the gadget code implies), satisfaction, and the reference's own KAT
behaviours (num.rs:591-764) -- the GPU runs the same scenarios in tests/test_gpu_gadgets_num.py."""
import random

import numpy as np
import pytest

from bellpepper_b200 import fixtures
from oracle import c_api
from oracle.fields import FIELDS

FIDS = sorted(FIELDS)


def _record(fid, fn):
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        out = fn(t)
        return out, t.host_csr()


@pytest.mark.parametrize("fid", FIDS)
def test_to_bits_le_structure_and_values(fid):
    p = FIELDS[fid].p
    rng = random.Random(fid)
    for v in (0, 1, p - 1, rng.randrange(p), rng.randrange(p)):
        bits, (lens, cols, coeffs, inputs, aux) = _record(fid, lambda t: t.num_unpack(v))
        assert sum(int(b) << i for i, b in enumerate(bits)) == v                    # num.rs:737-746
        n_rows = lens.size // 3
        assert n_rows == 256 and aux.shape[0] == 256                                # 255 boolean rows + the unpacking row
        assert list(lens[-3:]) == [0, 0, 256]                                       # 0 * 0 = sum 2^i b_i - num  (num.rs:263-274)
        # coefficients of the unpacking row: -1 on "num" (aux 0) then 2^i on bit i
        k = int(lens[:-3].sum())
        assert c_api.limbs_to_ints(coeffs[k:k + 1])[0] == p - 1
        assert c_api.limbs_to_ints(coeffs[k + 1:k + 256]) == [pow(2, i, p) for i in range(255)]
        inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
        assert inst.check(1, True) == -1
        inst.set(True, 0, (v + 1) % p)                                              # cs.set("num", other) -> unsatisfied
        assert inst.check(1, True) == 255
        inst.set(True, 0, v)
        for i in (0, 1, 100, 253, 254):                                             # num.rs:753-762
            inst.set(True, 1 + i, 1 - int(bits[i]))
            assert inst.check(1, True) == 255                                       # the bit is still boolean: only the unpacking row fails
            inst.set(True, 1 + i, int(bits[i]))
        assert inst.check(1, True) == -1


@pytest.mark.parametrize("fid", FIDS)
def test_to_bits_le_strict(fid):
    p = FIELDS[fid].p
    rng = random.Random(10 + fid)
    for v in (p - 1, 0, rng.randrange(p)):
        bits, (lens, cols, coeffs, inputs, aux) = _record(fid, lambda t: t.num_unpack(v, strict=True))
        assert sum(int(b) << i for i, b in enumerate(bits)) == v
        inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
        assert inst.check(1, True) == -1
    # num.rs:696-714: with value p - 1, setting the top bit pattern to the characteristic breaks "bit 254/boolean constraint":
    # bit 254 is the LAST bit allocated (lowest), guarded by the AND of every run of ones above it
    bits, (lens, cols, coeffs, inputs, aux) = _record(fid, lambda t: t.num_unpack(p - 1, strict=True))
    with fixtures.Tcs(fid, device=-1, named=True) as t:
        t.num_unpack(p - 1, strict=True)
        paths = [t.row_path(r) for r in range(t.num_constraints())]
    assert paths[-1] == "unpacking constraint" and "bit 254/boolean constraint" in paths
    # p - 1 is even: its lowest bit is 0 and that position is conditionally allocated
    assert bits[0] == 0


@pytest.mark.parametrize("fid", FIDS)
def test_arith_gadgets(fid):
    p = FIELDS[fid].p
    a, b = 12, 10
    _, (lens, cols, coeffs, inputs, aux) = _record(fid, lambda t: t.num_arith(a, b))
    vals = c_api.limbs_to_ints(aux)
    # allocation order: a, b, product, squared, sum, inverse, condition bit, reversal results c, d
    assert vals[:5] == [12, 10, 120, 144, 264]
    assert (vals[5] * b) % p == 1
    assert vals[6] == 1 and vals[7:9] == [b, a]                                      # condition true: (b, a)  (num.rs:403-455)
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    assert inst.check(1, True) == -1
    inst.set(True, 2, 121)                                                           # num.rs:634-637
    assert inst.check(1, True) == 0
    inst.set(True, 2, 120)
    # (p - 1) + 1 wraps to 0 (num.rs:591-609)
    _, (lens, cols, coeffs, inputs, aux) = _record(fid, lambda t: t.num_arith(p - 1, 1))
    vals = c_api.limbs_to_ints(aux)
    assert vals[2] == p - 1 and vals[3] == 1 and vals[4] == 0
    assert c_api.Instance(fid, lens, cols, coeffs, inputs, aux).check(1, True) == -1
    # assert_nonzero on zero is a synthesis error (num.rs:676-693)
    with pytest.raises(RuntimeError):
        with fixtures.Tcs(fid, device=-1, named=False) as t:
            t.num_arith(5, 0)


def test_chain_is_product_heavy_and_satisfied():
    fid = 0
    p = FIELDS[fid].p
    _, (lens, cols, coeffs, inputs, aux) = _record(fid, lambda t: t.num_chain(200, 50, 3, p - 2))
    n_rows = lens.size // 3
    assert n_rows == 3 * 200 + 4 * 256
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    assert inst.check(2, False) == -1
    wide = sum(1 for v in c_api.limbs_to_ints(aux) if v >= (1 << 24))
    assert wide > 500  # full-width witness values: the full-width kernels' case
