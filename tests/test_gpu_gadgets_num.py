"""GPU (-m gpu): the reference's num KATs (crates/bellpepper-core/src/gadgets/num.rs:591-764) through the C++ front-end's
AllocatedNum gadgets on the device -- full-width witness values and 256-term unpacking rows, i.e. gadget circuits that drive
the full-width kernels -- with the reference's paths, plus oracle parity on a product-heavy gadget chain."""
import ctypes
import random

import numpy as np
import pytest

from bellpepper_b200 import ffi, fixtures
from oracle import c_api
from oracle.fields import FIELDS

pytestmark = pytest.mark.gpu
FIDS = sorted(FIELDS)


@pytest.mark.parametrize("fid", FIDS)
def test_into_bits(fid):
    """num.rs:717-764 test_into_bits: random r, to_bits_le / to_bits_le_strict alternately; satisfied; the bits are r's;
    `num` := another value -> unsatisfied; every single-bit flip -> unsatisfied, and satisfied again once restored."""
    p = FIELDS[fid].p
    rng = random.Random(0x5962 + fid)
    for i in range(4):
        r = rng.randrange(p)
        with fixtures.Tcs(fid, device=0, named=True) as cs:
            bits = cs.num_unpack(r, strict=bool(i % 2))
            assert cs.is_satisfied()
            assert sum(int(b) << k for k, b in enumerate(bits)) == r
            cs.set("num", rng.randrange(p))
            assert not cs.is_satisfied()
            cs.set("num", r)
            assert cs.is_satisfied()
            for k in range(255):
                name = f"bit {k}/boolean"
                cur = cs.get(name)
                cs.set(name, 1 - cur)
                assert not cs.is_satisfied(), (i, k)
                cs.set(name, cur)
            assert cs.is_satisfied()


@pytest.mark.parametrize("fid", FIDS)
def test_into_bits_strict_names_the_conditional_bit(fid):
    """num.rs:696-714: value -1; making the representation the characteristic breaks "bit 254/boolean constraint" FIRST."""
    p = FIELDS[fid].p
    with fixtures.Tcs(fid, device=0, named=True) as cs:
        cs.num_unpack(p - 1, strict=True)
        assert cs.is_satisfied()
        cs.set("bit 254/boolean", 1)
        assert cs.which_is_unsatisfied() == "bit 254/boolean constraint"


@pytest.mark.parametrize("fid", FIDS)
def test_num_arithmetic_kats(fid):
    """num.rs:591-693: addition wraps, squaring, multiplication, nonzero assertion; perturbing a result names its constraint."""
    p = FIELDS[fid].p
    with fixtures.Tcs(fid, device=0, named=True) as cs:
        cs.num_arith(12, 10)
        assert cs.is_satisfied()
        assert cs.get("product num") == 120 and cs.get("squared num") == 144 and cs.get("sum num") == 264
        cs.set("product num", 121)
        assert cs.which_is_unsatisfied() == "multiplication constraint"
        cs.set("product num", 120)
        cs.set("squared num", 10)
        assert cs.which_is_unsatisfied() == "squaring constraint"
        cs.set("squared num", 144)
        cs.set("nonzero/ephemeral inverse", 3)
        assert cs.which_is_unsatisfied() == "nonzero/nonzero assertion constraint"
    with fixtures.Tcs(fid, device=0, named=True) as cs:
        cs.num_arith(p - 1, 1)
        assert cs.is_satisfied()
        assert cs.get("sum num") == 0            # (p - 1)^2 = 1,  (p - 1) * 1 + 1 = 0  mod p
        cs.set("sum num", 1)
        assert not cs.is_satisfied()


@pytest.mark.parametrize("fid", FIDS)
def test_product_heavy_gadget_chain_matches_oracle(fid):
    """A gadget circuit whose witness is full-width (x <- x^2 y + x, unpacked every 64 steps): rows of the full-width kernels
    and 256-term rows; first-unsatisfied row and A.w / B.w / C.w against the oracle, with default and forced kernel choices."""
    p = FIELDS[fid].p
    rng = random.Random(77 + fid)
    x0, y0 = rng.randrange(p), rng.randrange(p)
    L = ffi.load()
    with fixtures.Tcs(fid, device=-1, named=False) as rec:
        rec.num_chain(3000, 64, x0, y0)
        lens, cols, coeffs, inputs, aux = rec.host_csr()
    inst = c_api.Instance(fid, lens, cols, coeffs, inputs, aux)
    n_rows = lens.size // 3
    bad, az_r, bz_r, cz_r = inst.eval(2)
    assert bad == -1
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.num_chain(3000, 64, x0, y0)
        h = ffi.vp(t.handle)
        assert t.first_unsatisfied_row() == -1
        for ft, variant in ((96, -1), (8, -1), (1000, -1), (96, 1), (8, 8)):
            assert L.bp_cs_set_option(h, b"fat_terms", ft) == 0 and L.bp_cs_set_option(h, b"variant", variant) == 0
            az, bz, cz = (np.zeros((n_rows, 4), np.uint64) for _ in range(3))
            assert L.bp_cs_eval(h, az.ctypes.data, bz.ctypes.data, cz.ctypes.data) == 0
            assert (az == az_r).all() and (bz == bz_r).all() and (cz == cz_r).all()
            for idx in (5, aux.shape[0] // 2, aux.shape[0] - 9):
                old = c_api.limbs_to_ints(aux[idx:idx + 1])[0]
                new = rng.randrange(p)
                v = c_api.ints_to_limbs([new])
                assert L.bp_cs_set(h, 1, idx, v.ctypes.data) == 0
                inst.set(True, idx, new)
                assert t.first_unsatisfied_row() == inst.check(2, False) >= 0
                v = c_api.ints_to_limbs([old])
                assert L.bp_cs_set(h, 1, idx, v.ctypes.data) == 0
                inst.set(True, idx, old)
            assert t.first_unsatisfied_row() == -1
