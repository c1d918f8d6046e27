"""Oracle self-consistency: C CIOS Montgomery vs Python `%`, constants, synthetic recipe (CPU only)."""
import ctypes
import random

import numpy as np
import pytest

from oracle import c_api, lib, synth
from oracle.fields import FIELDS

FIDS = sorted(FIELDS)


@pytest.mark.parametrize("fid", FIDS)
def test_constants(fid):
    f = FIELDS[fid]
    p = np.zeros(4, np.uint64)
    r2 = np.zeros(4, np.uint64)
    inv = ctypes.c_uint64()
    u64p = ctypes.POINTER(ctypes.c_uint64)
    assert lib().bpo_field_params(fid, p.ctypes.data_as(u64p), ctypes.byref(inv), r2.ctypes.data_as(u64p)) == 0
    assert c_api.limbs_to_ints(p)[0] == f.p
    assert inv.value == f.inv64
    assert c_api.limbs_to_ints(r2)[0] == pow(2, 512, f.p)
    assert f.p.bit_length() == 255 and (f.p - 1) % (1 << 32) == 0  # 2-adicity >= 32 -> inv32 == 0xffffffff
    assert f.inv32 == 0xFFFFFFFF


def _edge(p):
    R = (1 << 256) % p
    return [0, 1, 2, p - 1, p - 2, R, (R * R) % p, (1 << 254) % p, ((1 << 255) - 1) % p, p >> 1, 0xFFFFFFFF, 1 << 64]


@pytest.mark.parametrize("fid", FIDS)
def test_mul_add_vs_python(fid):
    p = FIELDS[fid].p
    rng = random.Random(1234 + fid)
    vals = _edge(p) + [rng.randrange(p) for _ in range(200)]
    for a in vals[:24]:
        for b in vals[:24]:
            assert c_api.mul(fid, a, b) == (a * b) % p
            assert c_api.add(fid, a, b) == (a + b) % p
    for _ in range(500):
        a, b = rng.choice(vals), rng.choice(vals)
        assert c_api.mul(fid, a, b) == (a * b) % p
        assert c_api.add(fid, a, b) == (a + b) % p


@pytest.mark.parametrize("fid", FIDS)
def test_synth_c_matches_python_spec(fid):
    seed, t, n_vars, n_inputs, n_rows, row0 = synth.SEED, 6, 1000, synth.N_INPUTS, 40, 7
    lens, cols, coeffs = c_api.synth_rows(fid, seed, t, n_vars, n_inputs, row0, n_rows)
    cvals = c_api.limbs_to_ints(coeffs)
    k = 0
    for r in range(n_rows):
        for lc in range(3):
            terms = synth.lc_terms(fid, seed, t, n_vars, n_inputs, row0 + r, lc)
            assert lens[3 * r + lc] == len(terms)
            last = -1
            for tagged, coeff in terms:
                assert int(cols[k]) == tagged and cvals[k] == coeff and coeff < FIELDS[fid].p
                unified = (tagged & 0x7FFFFFFF) + (n_inputs if tagged >> 31 else 0)
                assert unified > last and unified < n_vars  # ascending, unique, in range
                last = unified
                k += 1
    assert k == cols.size
    w = c_api.limbs_to_ints(c_api.synth_witness(fid, seed, 0, 50))
    assert w == [synth.witness(fid, seed, i) for i in range(50)] and w[0] == 1
    w2 = c_api.limbs_to_ints(c_api.synth_witness(fid, seed, 30, 10))
    assert w2 == w[30:40]


def test_synth_len_mean():
    t = 6
    ls = [synth.lc_len(synth.SEED, t, r, lc) for r in range(2000) for lc in range(3)]
    assert min(ls) == 1 and max(ls) == 2 * t - 1
    assert abs(sum(ls) / len(ls) - t) < 0.2
