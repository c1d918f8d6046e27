"""GPU test (-m gpu) of bp_cs_recheck_scalars_mont: the witness handed over AS IT SITS IN MEMORY in the reference's Vec<Scalar>
(witness_cs.rs:45-57; blstrs::Scalar / pasta_curves::{Fp,Fq} = 4 x u64 limbs of x * 2^256 mod p) gives the verdicts, the stored
witness and the A.w, B.w, C.w of the same witness in canonical form -- and of the CPU oracle."""
import ctypes
import random

import numpy as np
import pytest

from oracle import c_api
from oracle.fields import FIELDS

pytestmark = pytest.mark.gpu

from gpu_util import Handle  # noqa: E402
from test_gpu_parity import _gadget_like_instance, _satisfiable  # noqa: E402


def to_mont(fid, limbs):
    """canonical (n, 4) u64 limbs -> limbs of x * 2^256 mod p (Python integers: the definition)."""
    p = FIELDS[fid].p
    vals = c_api.limbs_to_ints(limbs)
    return c_api.ints_to_limbs([(v << 256) % p for v in vals]) if len(vals) else np.zeros((0, 4), np.uint64)


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_recheck_scalars_mont_equals_canonical_and_oracle(fid):
    p = FIELDS[fid].p
    lens, cols, coeffs, inputs, aux, _ = _gadget_like_instance(fid, 77, 3000, 5000, 9)
    rng = random.Random(100 + fid)
    row = ctypes.c_int64()
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        a = np.zeros_like(aux)
        a[:, 0] = aux[:, 0] & np.uint64(1)  # a bit witness: the packer is what runs
        i = inputs.copy()
        i[1:, 0] &= np.uint64(1)
        i[1:, 1:] = 0
        for trial in range(3):
            for _ in range(5):
                k = rng.randrange(a.shape[0])
                a[k] = c_api.ints_to_limbs([1 - int(a[k][0]) if int(a[k][0]) in (0, 1) and not a[k][1:].any() else 0])[0]
            if trial >= 1:  # exceptions: small, full-width, and the two values whose limb patterns are the traps
                for v in (2, 255, p - 1, rng.randrange(p), (1 << 256) % p, pow(1 << 256, -1, p)):
                    a[rng.randrange(a.shape[0])] = c_api.ints_to_limbs([v])[0]
            if trial == 2:
                i[rng.randrange(1, i.shape[0])] = c_api.ints_to_limbs([rng.randrange(p)])[0]
            ref = c_api.Instance(fid, lens, cols, coeffs, i, a)
            want = ref.check(2, False)
            im, am = to_mont(fid, i), to_mont(fid, a)
            h.ok(h.L.bp_cs_recheck_scalars_mont(h.h, im.ctypes.data, am.ctypes.data, ctypes.byref(row)))
            assert row.value == want
            got = np.zeros_like(a)
            h.ok(h.L.bp_cs_witness(h.h, 1, 0, a.shape[0], got.ctypes.data))
            assert (got == a).all()  # stored canonical
            got_i = np.zeros_like(i)
            h.ok(h.L.bp_cs_witness(h.h, 0, 0, i.shape[0], got_i.ctypes.data))
            assert (got_i == i).all()
            _, az_r, bz_r, cz_r = ref.eval(2)
            az, bz, cz = h.eval(lens.size // 3)
            assert (az == az_r).all() and (bz == bz_r).all() and (cz == cz_r).all()
            # and the canonical entry point says the same
            h.ok(h.L.bp_cs_recheck_scalars(h.h, i.ctypes.data, a.ctypes.data, ctypes.byref(row)))
            assert row.value == want
        # limbs >= p are not a Scalar
        am2 = am.copy()
        am2[10] = np.frombuffer(int(p).to_bytes(32, "little"), dtype="<u8")
        assert h.L.bp_cs_recheck_scalars_mont(h.h, im.ctypes.data, am2.ctypes.data, ctypes.byref(row)) == -3
    # a full-width witness is converted on the host and sent as it is
    lens, cols, coeffs, inputs, aux = _satisfiable(fid, 4, 800, 1200)
    with Handle(fid) as h:
        h.load_instance(lens, cols, coeffs, inputs, aux)
        im, am = to_mont(fid, inputs), to_mont(fid, aux)
        h.ok(h.L.bp_cs_recheck_scalars_mont(h.h, im.ctypes.data, am.ctypes.data, ctypes.byref(row)))
        assert row.value == -1
        a = aux.copy()
        a[700] = c_api.ints_to_limbs([5])[0]
        am = to_mont(fid, a)
        h.ok(h.L.bp_cs_recheck_scalars_mont(h.h, None, am.ctypes.data, ctypes.byref(row)))
        assert row.value == c_api.Instance(fid, lens, cols, coeffs, inputs, a).check(2, False) >= 0
        got = np.zeros_like(a)
        h.ok(h.L.bp_cs_witness(h.h, 1, 0, a.shape[0], got.ctypes.data))
        assert (got == a).all()
