"""The host packing pass of bp_cs_recheck_scalars (csrc/host/pack.cpp; exported as bp_pack_scalars): 32-byte canonical scalars,
the reference's witness format (witness_cs.rs:45-57), -> one bit per 0/1 value + an exception list.  Host only: runs without a
GPU, against numpy, on both kernels (AVX2 and the portable loop)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from bellpepper_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack(L, scal, exc_cap=None):
    n = scal.shape[0]
    bits = np.full((n + 7) // 8 + 3, 0xAB, np.uint8)  # three guard bytes behind the bit string
    cap = n if exc_cap is None else exc_cap
    idx = np.zeros(max(cap, 1), np.uint64)
    vals = np.zeros((max(cap, 1), 4), np.uint64)
    n_exc = ctypes.c_uint64(12345)
    rc = L.bp_pack_scalars(scal.ctypes.data if n else None, n, bits.ctypes.data, idx.ctypes.data if cap else None,
                           vals.ctypes.data if cap else None, cap, ctypes.byref(n_exc))
    assert (bits[(n + 7) // 8:] == 0xAB).all(), "wrote past the bit string"
    return rc, bits[: (n + 7) // 8], idx, vals, n_exc.value


def expected(scal):
    n = scal.shape[0]
    is_bit = (scal[:, 1:] == 0).all(axis=1) & (scal[:, 0] <= 1)
    want_bits = np.packbits(np.where(is_bit, scal[:, 0], 0).astype(np.uint8), bitorder="little") if n else np.zeros(0, np.uint8)
    exc = np.nonzero(~is_bit)[0]
    return want_bits, exc


def make(n, rng, n_exc):
    scal = np.zeros((n, 4), np.uint64)
    scal[:, 0] = rng.integers(0, 2, size=n, dtype=np.uint64)
    if n and n_exc:
        where = rng.choice(n, size=min(n, n_exc), replace=False)
        kinds = rng.integers(0, 5, size=where.size)
        for w, k in zip(where, kinds):
            if k == 0:
                scal[w, 0] = 2  # the smallest non-bit
            elif k == 1:
                scal[w, 0] = rng.integers(2, 1 << 63, dtype=np.uint64)
            elif k == 2:
                scal[w] = (int(scal[w, 0]), 0, 0, 1)  # low limb looks like a bit, a high limb does not
            elif k == 3:
                scal[w] = (int(scal[w, 0]), 1 << 63, 0, 0)
            else:
                scal[w] = rng.integers(0, 1 << 62, size=4, dtype=np.uint64)
    return scal


def run_cases(L):
    rng = np.random.default_rng(20261017)
    for n, n_exc in [(0, 0), (1, 0), (1, 1), (7, 2), (8, 0), (8, 8), (9, 1), (63, 5), (64, 0), (1000, 0), (1001, 37), (4099, 4099),
                     (262147, 1000), (262144 * 9 + 5, 3)]:
        scal = make(n, rng, n_exc)
        want_bits, exc = expected(scal)
        rc, bits, idx, vals, k = pack(L, scal)
        assert rc == 0 and k == exc.size, (n, rc, k, exc.size)
        assert (bits == want_bits).all(), n
        assert (idx[:k] == exc.astype(np.uint64)).all(), n  # ascending index order
        assert (vals[:k] == scal[exc]).all(), n
    # not enough room for the exceptions: the count is still reported, the first exc_cap are stored
    scal = make(5000, rng, 100)
    want_bits, exc = expected(scal)
    rc, bits, idx, vals, k = pack(L, scal, exc_cap=10)
    assert rc == ffi.BP_E_RANGE and k == exc.size
    assert (bits == want_bits).all() and (idx[:10] == exc[:10].astype(np.uint64)).all() and (vals[:10] == scal[exc[:10]]).all()
    rc, bits, idx, vals, k = pack(L, scal, exc_cap=0)
    assert rc == ffi.BP_E_RANGE and k == exc.size and (bits == want_bits).all()
    # an unaligned source (8-byte aligned only, as a Vec<Scalar> may be)
    raw = np.zeros(4 * 1003 + 1, np.uint64)
    view = raw[1:].reshape(1003, 4)
    view[:] = make(1003, rng, 9)
    want_bits, exc = expected(view)
    rc, bits, idx, vals, k = pack(L, view)
    assert rc == 0 and k == exc.size and (bits == want_bits).all() and (idx[:k] == exc.astype(np.uint64)).all()
    # argument errors
    one = ctypes.c_uint64()
    assert L.bp_pack_scalars(None, 5, None, None, None, 0, ctypes.byref(one)) == ffi.BP_E_ARG
    assert L.bp_pack_scalars(scal.ctypes.data, 5, bits.ctypes.data, None, None, 0, None) == ffi.BP_E_ARG
    assert L.bp_pack_scalars(scal.ctypes.data, 5, bits.ctypes.data, None, None, 4, ctypes.byref(one)) == ffi.BP_E_ARG


def test_pack_scalars_matches_numpy():
    L = ffi.load()
    assert L.bp_pack_kernel() in (b"avx2", b"portable")
    run_cases(L)


@pytest.mark.parametrize("env", [{"BP_PACK_SIMD": "0"}, {"BP_PACK_THREADS": "1"}, {"BP_PACK_THREADS": "5"}, {"BP_PACK_THREADS": "64"}])
def test_pack_scalars_other_kernels_and_thread_counts(env):
    """The kernel choice is made once per process: the portable loop and other thread counts run in a child process."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from bellpepper_b200 import ffi\nimport test_pack_cpu as t\nL = ffi.load()\n"
            "print(L.bp_pack_kernel().decode())\nt.run_cases(L)\nprint('ok')\n") % (ROOT, os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.split()
    assert lines[-1] == "ok"
    if env.get("BP_PACK_SIMD") == "0":
        assert lines[0] == "portable"


# ---- scalars in their in-memory Montgomery form (blstrs::Scalar / pasta_curves::{Fp,Fq}: x * 2^256 mod p) ---------------------
def _limbs(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def _to_int(l):
    return sum(int(v) << (64 * i) for i, v in enumerate(l))


def run_mont_cases(L):
    from bellpepper_b200.fields import MODULUS

    rng = np.random.default_rng(7)
    for fid in (0, 1, 2):
        p = MODULUS[fid]
        R = (1 << 256) % p
        for n, n_exc in [(0, 0), (1, 0), (5, 2), (8, 0), (9, 9), (1000, 0), (4099, 50), (262147, 300)]:
            vals = [int(b) for b in rng.integers(0, 2, size=n)]
            where = rng.choice(n, size=min(n, n_exc), replace=False) if n else []
            for k, w in enumerate(where):
                # values around the traps: 2, p - 1, the integer whose CANONICAL form has the limbs of R (a bit only as Montgomery),
                # the integer 1/R (its Montgomery form is the canonical 1), random
                vals[w] = [2, p - 1, R, pow(R, -1, p), int.from_bytes(rng.bytes(32), "little") % p][k % 5]
            mont = np.array([_limbs(v * R % p) for v in vals], np.uint64).reshape(n, 4)
            exc = [i for i, v in enumerate(vals) if v > 1]
            want_bits = np.packbits(np.array([v if v <= 1 else 0 for v in vals], np.uint8), bitorder="little") if n else np.zeros(0, np.uint8)
            bits = np.full((n + 7) // 8 + 2, 0xCD, np.uint8)
            cap = max(1, len(exc))
            idx = np.zeros(cap, np.uint64)
            ev = np.zeros((cap, 4), np.uint64)
            k = ctypes.c_uint64()
            rc = L.bp_pack_scalars_mont(fid, mont.ctypes.data if n else None, n, bits.ctypes.data, idx.ctypes.data, ev.ctypes.data, cap, ctypes.byref(k))
            assert rc == 0 and k.value == len(exc), (fid, n, rc, k.value, len(exc))
            assert (bits[: (n + 7) // 8] == want_bits).all() and (bits[(n + 7) // 8:] == 0xCD).all()
            assert [int(i) for i in idx[: len(exc)]] == exc
            assert [_to_int(ev[j]) for j in range(len(exc))] == [vals[i] for i in exc]  # canonical again
            # the whole array back to canonical form
            canon = np.zeros((max(n, 1), 4), np.uint64)
            assert L.bp_scalars_from_mont(fid, mont.ctypes.data if n else None, n, canon.ctypes.data) == 0
            assert [_to_int(canon[i]) for i in range(n)] == vals
        # a limb pattern >= p is not a scalar
        bad = np.array([_limbs(1 * R % p), _limbs(p), _limbs(0)], np.uint64)
        bits = np.zeros(1, np.uint8)
        idx, ev, k = np.zeros(4, np.uint64), np.zeros((4, 4), np.uint64), ctypes.c_uint64()
        assert L.bp_pack_scalars_mont(fid, bad.ctypes.data, 3, bits.ctypes.data, idx.ctypes.data, ev.ctypes.data, 4, ctypes.byref(k)) == ffi.BP_E_RANGE
        assert L.bp_scalars_from_mont(fid, bad.ctypes.data, 3, np.zeros((3, 4), np.uint64).ctypes.data) == ffi.BP_E_RANGE
    assert L.bp_pack_scalars_mont(3, bad.ctypes.data, 3, bits.ctypes.data, idx.ctypes.data, ev.ctypes.data, 4, ctypes.byref(k)) == ffi.BP_E_ARG
    assert L.bp_pack_scalars_mont(-1, bad.ctypes.data, 3, bits.ctypes.data, idx.ctypes.data, ev.ctypes.data, 4, ctypes.byref(k)) == ffi.BP_E_ARG


def test_pack_scalars_mont_matches_python_ints():
    run_mont_cases(ffi.load())


def test_pack_scalars_mont_portable_kernel():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from bellpepper_b200 import ffi\nimport test_pack_cpu as t\nL = ffi.load()\n"
            "assert L.bp_pack_kernel() == b'portable'\nt.run_mont_cases(L)\nprint('ok')\n") % (ROOT, os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-c", code], env={**os.environ, "BP_PACK_SIMD": "0", "BP_PACK_THREADS": "3"}, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.split()[-1] == "ok", out.stderr[-2000:]
