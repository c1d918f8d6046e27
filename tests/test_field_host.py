"""Device field arithmetic (bellpepper_b200/csrc/field.cuh), host twins of the PTX chains, vs Python ints."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle.fields import FIELDS

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fh():
    src = os.path.join(HERE, "host", "field_host.cpp")
    so = os.path.join(HERE, "host", "libfield_host.so")
    hdr = os.path.join(HERE, "..", "bellpepper_b200", "csrc", "field.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", so], check=True)
    L = ctypes.CDLL(so)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.field_host_op.argtypes = [ctypes.c_int, ctypes.c_int, u32p, u32p, u32p]

    def call(field, op, a, b, nout, a_limbs=8, out_init=0):
        A = np.frombuffer(int(a).to_bytes(4 * a_limbs, "little"), np.uint32).copy()
        B = np.frombuffer(int(b).to_bytes(32, "little"), np.uint32).copy()
        O = np.frombuffer(int(out_init).to_bytes(4 * nout, "little"), np.uint32).copy()
        assert L.field_host_op(field, op, A.ctypes.data_as(u32p), B.ctypes.data_as(u32p), O.ctypes.data_as(u32p)) == 0
        return int.from_bytes(O.tobytes(), "little")

    return call


def edge(p):
    return [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (1 << 288) % p, 0xFFFFFFFF, 1 << 32, (1 << 254), p >> 1,
            0xFFFFFFFF_00000000_FFFFFFFF_00000000_FFFFFFFF_00000000_FFFFFFFF % p]


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_mul_wide_and_mac(fh, fid):
    p = FIELDS[fid].p
    rng = random.Random(fid)
    vals = edge(p) + [rng.randrange(p) for _ in range(40)] + [(1 << 256) - 1, (1 << 256) - (1 << 32)]
    for a in vals:
        for b in vals[:20]:
            assert fh(fid, 0, a, b, 16) == a * b
    acc = 0
    for _ in range(300):
        a, b = rng.choice(vals), rng.choice(vals)
        acc = fh(fid, 3, a, b, 17, out_init=acc)
        assert acc < (1 << 544)
    # and it matches the running integer sum
    rng = random.Random(fid)
    vals2 = vals
    chk, acc = 0, 0
    for _ in range(200):
        a, b = rng.choice(vals2), rng.choice(vals2)
        chk += a * b
        acc = fh(fid, 3, a, b, 17, out_init=acc)
    assert acc == chk


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_redc(fh, fid):
    p = FIELDS[fid].p
    rng = random.Random(100 + fid)
    inv288 = pow(2, -288, p)
    inv256 = pow(2, -256, p)
    cases = [0, 1, p, p * p, (p - 1) ** 2, ((1 << 32) - 1) * (p - 1) ** 2, (1 << 542) - 1, (1 << 288), (1 << 288) - 1,
             (1 << 256) - 1, p << 288, (p << 288) - 1]
    cases += [rng.randrange(1 << 542) for _ in range(300)]
    cases += [rng.randrange(1 << 40) * rng.randrange(p) ** 2 % (1 << 542) for _ in range(100)]
    for T in cases:
        u = fh(fid, 1, T, 0, 8, a_limbs=17)
        assert u < (1 << 256) and u % p == (T * inv288) % p
        if T < (1 << 32) * p * p:
            assert u < 2 * p
            assert fh(fid, 4, u, 0, 8) == (T * inv288) % p  # reduce_once -> canonical
    for T in [0, 1, p - 1, (1 << 256) - 1] + [rng.randrange(1 << 256) for _ in range(200)]:
        v = fh(fid, 7, T, 0, 8)
        assert v <= p and v % p == (T * inv256) % p


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_mont_mul_predicates(fh, fid):
    p = FIELDS[fid].p
    rng = random.Random(200 + fid)
    inv256 = pow(2, -256, p)
    vals = edge(p) + [rng.randrange(p) for _ in range(60)]
    for a in vals:
        for b in vals[:16]:
            assert fh(fid, 2, a, b, 8) == (a * b * inv256) % p
    assert fh(fid, 5, 0, 0, 1) == 1 and fh(fid, 5, p, 0, 1) == 1 and fh(fid, 5, 1, 0, 1) == 0 and fh(fid, 5, p - 1, 0, 1) == 0
    assert fh(fid, 6, p - 1, 0, 1) == 1 and fh(fid, 6, p, 0, 1) == 0 and fh(fid, 6, (1 << 256) - 1, 0, 1) == 0


@pytest.fixture(scope="module")
def fh2(fh):
    so = os.path.join(HERE, "host", "libfield_host.so")
    L = ctypes.CDLL(so)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    L.field_host_op2.argtypes = [ctypes.c_int, ctypes.c_int, u32p, u32p, u32p]

    def call(field, op, a, b, nout, a_limbs=8, out_init=0):
        A = np.frombuffer(int(a).to_bytes(4 * a_limbs, "little"), np.uint32).copy()
        B = np.frombuffer(int(b).to_bytes(32, "little"), np.uint32).copy()
        O = np.frombuffer(int(out_init).to_bytes(4 * nout, "little"), np.uint32).copy()
        assert L.field_host_op2(field, op, A.ctypes.data_as(u32p), B.ctypes.data_as(u32p), O.ctypes.data_as(u32p)) == 0
        return int.from_bytes(O.tobytes(), "little")

    return call


@pytest.mark.parametrize("fid", sorted(FIELDS))
def test_plain_primitives(fh2, fid):
    p = FIELDS[fid].p
    rng = random.Random(300 + fid)
    vals = edge(p) + [rng.randrange(p) for _ in range(50)]
    for x in vals:
        assert fh2(fid, 10, x, 0, 8) == p - x
        for acc in (0, 1, p - 1, 7 * p - 3, (1 << 256) - 1):
            assert fh2(fid, 11, x, 0, 9, out_init=acc) == acc + x
            big = acc + rng.randrange(1 << 530)
            assert fh2(fid, 12, x, 0, 17, out_init=big) == big + x
            for s in (0, 1, 3, 7, 0xFFFFFFFF, rng.randrange(1 << 32)):
                if acc + s * x < (1 << 288):
                    assert fh2(fid, 13, x, s, 9, out_init=acc) == acc + s * x
                assert fh2(fid, 14, x, s, 17, out_init=big) == big + s * x
    for v in [0, 1, p - 1, p, p + 1, 2 * p, 4 * p - 1, 4 * p, 7 * p + 5, 8 * p - 1] + [rng.randrange(8 * p) for _ in range(300)]:
        assert fh2(fid, 15, v, 0, 8, a_limbs=9) == v % p
    K = {"gen": 0, "p1": 1, "m1": 2, "p2": 3, "m2": 4, "pow2p": 5, "pow2m": 6, "zero": 7}
    one = lambda k: k | 0xFF00  # packed exponents of a single power of two
    cases = [(0, "zero", 0), (1, "p1", 0), (p - 1, "m1", 0), (2, "p2", 0), (p - 2, "m2", 0), (3, "pow2p", 0 | 1 << 8),
             (p - 3, "pow2m", 0 | 1 << 8), (7, "gen", 0), (p - 7, "gen", 0), (4, "pow2p", one(2)), (p - 4, "pow2m", one(2)),
             (0xFFFFFFFF, "gen", 0), (1 << 31, "pow2p", one(31)), (1 << 32, "pow2p", one(32)), (p - (1 << 32), "pow2m", one(32)),
             (p >> 1, "gen", 0), (1 << 200, "pow2p", one(200)), (p - (1 << 253), "pow2m", one(253)),
             ((1 << 200) + 1, "pow2p", 0 | 200 << 8), ((1 << 253) + (1 << 31), "pow2p", 31 | 253 << 8),
             (p - (1 << 100) - (1 << 37), "pow2m", 37 | 100 << 8), ((1 << 100) + (1 << 37) + 1, "gen", 0)]
    cases += [(1 << k, "pow2p", one(k)) for k in range(2, 254)] + [(p - (1 << k), "pow2m", one(k)) for k in range(2, 254)]
    cases += [((1 << k) + (1 << (k + 13)), "pow2p", k | (k + 13) << 8) for k in range(0, 240, 7)]
    for v, cls, s in cases:
        r = fh2(fid, 16, v, 0, 2)
        assert (r & 0xFFFFFFFF, r >> 32) == (K[cls], s), (hex(v), cls)
