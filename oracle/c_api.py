"""TEST INFRASTRUCTURE (oracle) -- numpy-facing wrappers over oracle/bp_oracle.c."""

from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import lib

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)


def _p64(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_u64p)


def _p32(a: np.ndarray):
    return a.ctypes.data_as(_u32p)


def ints_to_limbs(values, out: Optional[np.ndarray] = None) -> np.ndarray:
    """list of python ints (canonical) -> uint64[n, 4] little-endian limbs."""
    n = len(values)
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    arr = np.frombuffer(buf, dtype="<u8").reshape(n, 4).copy() if n else np.zeros((0, 4), np.uint64)
    if out is not None:
        out[...] = arr
        return out
    return arr


def limbs_to_ints(arr: np.ndarray):
    a = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i : 32 * i + 32], "little") for i in range(a.shape[0])]


def mul(field: int, a: int, b: int) -> int:
    A, B, R = ints_to_limbs([a]), ints_to_limbs([b]), np.zeros((1, 4), np.uint64)
    assert lib().bpo_mul(field, _p64(A), _p64(B), _p64(R)) == 0
    return limbs_to_ints(R)[0]


def add(field: int, a: int, b: int) -> int:
    A, B, R = ints_to_limbs([a]), ints_to_limbs([b]), np.zeros((1, 4), np.uint64)
    assert lib().bpo_add(field, _p64(A), _p64(B), _p64(R)) == 0
    return limbs_to_ints(R)[0]


class Instance:
    """A prepared CSR instance held by the C oracle (Montgomery form inside, like blstrs)."""

    def __init__(self, field: int, lens: np.ndarray, cols: np.ndarray, coeffs: np.ndarray,
                 inputs: np.ndarray, aux: np.ndarray):
        lens = np.ascontiguousarray(lens, np.uint32)
        cols = np.ascontiguousarray(cols, np.uint32)
        coeffs = np.ascontiguousarray(coeffs, np.uint64).reshape(-1, 4)
        inputs = np.ascontiguousarray(inputs, np.uint64).reshape(-1, 4)
        aux = np.ascontiguousarray(aux, np.uint64).reshape(-1, 4)
        assert lens.size % 3 == 0 and int(lens.sum()) == cols.size == coeffs.shape[0]
        self.n_rows = lens.size // 3
        self.field = field
        self._h = lib().bpo_prepare(field, self.n_rows, _p32(lens), _p32(cols), _p64(coeffs),
                                    _p64(inputs), inputs.shape[0], _p64(aux), aux.shape[0])
        if not self._h:
            raise ValueError("bpo_prepare rejected the instance (non-canonical element or column out of range)")

    def set(self, is_aux: bool, idx: int, value: int):
        v = ints_to_limbs([value])
        if lib().bpo_set(self._h, int(is_aux), idx, _p64(v)) != 0:
            raise ValueError("bpo_set: out of range")

    def check(self, threads: int = 1, early_exit: bool = True) -> int:
        """First unsatisfied row or -1 (test_cs.rs:239-253)."""
        return int(lib().bpo_check(self._h, threads, int(early_exit), None, None, None))

    def eval(self, threads: int = 1) -> Tuple[int, np.ndarray, np.ndarray, np.ndarray]:
        az = np.zeros((self.n_rows, 4), np.uint64)
        bz = np.zeros_like(az)
        cz = np.zeros_like(az)
        bad = int(lib().bpo_check(self._h, threads, 0, _p64(az), _p64(bz), _p64(cz)))
        return bad, az, bz, cz

    def close(self):
        if self._h:
            lib().bpo_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def from_python_cs(cs) -> Instance:
    """oracle.r1cs_py.TestConstraintSystem -> C oracle instance (same CSR the product ingests)."""
    lens, cols, coeffs, inputs, aux = cs.to_csr()
    return Instance(cs.field.fid, np.asarray(lens, np.uint32), np.asarray(cols, np.uint32),
                    ints_to_limbs(coeffs), ints_to_limbs(inputs), ints_to_limbs(aux))


# ---- synthetic recipe (bulk, C) ------------------------------------------------------------------
def synth_rows(field: int, seed: int, t: int, n_vars: int, n_inputs: int, row0: int, n_rows: int):
    lens = np.zeros(3 * n_rows, np.uint32)
    nnz = int(lib().bpo_synth_lens(seed, t, row0, n_rows, _p32(lens)))
    cols = np.zeros(nnz, np.uint32)
    coeffs = np.zeros((nnz, 4), np.uint64)
    rc = lib().bpo_synth_fill(field, seed, t, n_vars, n_inputs, row0, n_rows, _p32(cols), _p64(coeffs))
    if rc != 0:
        raise ValueError("bpo_synth_fill: bad arguments")
    return lens, cols, coeffs


def synth_witness(field: int, seed: int, i0: int, n: int) -> np.ndarray:
    out = np.zeros((n, 4), np.uint64)
    assert lib().bpo_synth_witness(field, seed, i0, n, _p64(out)) == 0
    return out


def synth_instance(field: int, seed: int, t: int, n_vars: int, n_inputs: int, n_rows: int, row0: int = 0):
    """Full small synthetic instance: (lens, cols, coeffs, inputs, aux) as numpy arrays."""
    lens, cols, coeffs = synth_rows(field, seed, t, n_vars, n_inputs, row0, n_rows)
    w = synth_witness(field, seed, 0, n_vars)
    return lens, cols, coeffs, w[:n_inputs].copy(), w[n_inputs:].copy()


def synth_witness_at(field: int, seed: int, idx: np.ndarray) -> np.ndarray:
    idx = np.ascontiguousarray(idx, np.uint64)
    out = np.zeros((idx.size, 4), np.uint64)
    assert lib().bpo_synth_witness_at(field, seed, _p64(idx), idx.size, _p64(out)) == 0
    return out


def synth_sparse_instance(field: int, seed: int, t: int, n_vars: int, n_inputs: int, row0: int, n_rows: int) -> Instance:
    """Rows [row0, row0+n_rows) of the synthetic recipe as an oracle instance WITHOUT the whole witness: the columns the rows
    read are renumbered into a compact aux space and only those witness elements are generated (same values, same sums).  For
    sampled parity checks of instances whose witness is GiBs (BASELINE configs[4]: 2^27 variables)."""
    lens, cols, coeffs = synth_rows(field, seed, t, n_vars, n_inputs, row0, n_rows)
    unified = np.where(cols >> 31, (cols & np.uint32(0x7FFFFFFF)).astype(np.uint64) + np.uint64(n_inputs), cols.astype(np.uint64))
    uniq, inv = np.unique(unified, return_inverse=True)
    w = synth_witness_at(field, seed, uniq)
    one = np.asarray([[1, 0, 0, 0]], np.uint64)
    return Instance(field, lens, inv.astype(np.uint32) | np.uint32(0x80000000), coeffs, one, w)
