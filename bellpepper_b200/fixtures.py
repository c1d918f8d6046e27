"""ctypes binding of include/bp_fixtures.h: the C++ host front-end (B200-backed TestConstraintSystem mirror +
gadget circuits).  Used by tests and bench.py; production C++ users include csrc/host/*.hpp directly."""

from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

from . import ffi

vp = ctypes.c_void_p
u64p = ctypes.POINTER(ctypes.c_uint64)
_SIGS = {
    "bp_tcs_new": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(vp)]),
    "bp_tcs_free": (None, [vp]),
    "bp_tcs_last_error": (ctypes.c_char_p, [vp]),
    "bp_tcs_handle": (vp, [vp]),
    "bp_tcs_flush": (ctypes.c_int, [vp]),
    "bp_tcs_sha256_block": (ctypes.c_int, [vp, vp, vp]),
    "bp_tcs_sha256": (ctypes.c_int, [vp, vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, vp, u64p]),
    "bp_tcs_sha256_ranges": (ctypes.c_int, [vp, vp, ctypes.c_uint64, u64p, ctypes.c_uint64, vp, u64p, u64p]),
    "bp_tcs_blake2s": (ctypes.c_int, [vp, vp, ctypes.c_uint64, vp, vp]),
    "bp_tcs_boolean_op": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "bp_tcs_u64_bits": (ctypes.c_int, [vp, ctypes.c_uint64, vp]),
    "bp_tcs_uint32_op": (ctypes.c_int, [vp, ctypes.c_int] + [ctypes.c_uint32] * 4 + [ctypes.POINTER(ctypes.c_uint32)] * 2),
    "bp_tcs_num_unpack": (ctypes.c_int, [vp, vp, ctypes.c_int, vp]),
    "bp_tcs_num_arith": (ctypes.c_int, [vp, vp, vp]),
    "bp_tcs_num_chain": (ctypes.c_int, [vp, ctypes.c_uint64, ctypes.c_uint64, vp, vp]),
    "bp_tcs_record_witness_program": (ctypes.c_int, [vp, ctypes.c_int]),
    "bp_tcs_witness_program": (ctypes.c_int, [vp, ctypes.POINTER(vp), u64p]),
    "bp_sha256_chain_states": (ctypes.c_int, [vp, ctypes.c_uint64, vp, ctypes.c_uint64, u64p]),
    "bp_blake2s_chain_states": (ctypes.c_int, [vp, ctypes.c_uint64, vp, vp, ctypes.c_uint64, u64p]),
    "bp_wcs_selftest": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "bp_tcs_which_is_unsatisfied": (ctypes.c_int64, [vp, vp, ctypes.c_uint64]),
    "bp_tcs_set": (ctypes.c_int, [vp, ctypes.c_char_p, vp]),
    "bp_tcs_get": (ctypes.c_int, [vp, ctypes.c_char_p, vp]),
    "bp_tcs_num_constraints": (ctypes.c_uint64, [vp]),
    "bp_tcs_num_inputs": (ctypes.c_uint64, [vp]),
    "bp_tcs_num_aux": (ctypes.c_uint64, [vp]),
    "bp_tcs_row_path": (ctypes.c_int, [vp, ctypes.c_uint64, vp, ctypes.c_uint64]),
    "bp_tcs_host_csr": (ctypes.c_int, [vp] + [ctypes.POINTER(vp), u64p, ctypes.POINTER(vp), ctypes.POINTER(vp), u64p,
                                              ctypes.POINTER(vp), u64p, ctypes.POINTER(vp), u64p]),
}
_bound = False
_host = None
U64_MAX = (1 << 64) - 1
FRONTEND_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbp_frontend.so")


def _lib():
    global _bound
    L = ffi.load()
    if not _bound:
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def _host_lib():
    """libbp_frontend.so: the same front-end built on its own (g++, no CUDA) for HOST RECORDING -- what the CPU legs use
    (bench.py --impl reference, cpu_baseline samples, CPU tests), so that they never map the CUDA library."""
    global _host
    if _host is None:
        if not os.path.exists(FRONTEND_PATH):
            raise ffi.NativeLibraryMissing(f"{FRONTEND_PATH} not found -- run __graft_entry__.build()")
        L = ctypes.CDLL(FRONTEND_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _host = L
    return _host


def available() -> bool:
    try:
        return hasattr(_lib(), "bp_tcs_sha256")
    except Exception:
        return False


def xorshift_bytes(n: int, seed=(0x3DBE6259, 0x8D313D76, 0x3237DB17, 0xE5BC0654)) -> bytes:
    """The reference tests' XorShiftRng stream (`next_u32() as u8`); seed words are the LE u32s of the seed bytes
    [0x59,0x62,0xbe,0x3d, 0x76,0x3d,0x31,0x8d, 0x17,0xdb,0x37,0x32, 0x54,0x06,0xbc,0xe5] (sha256.rs:312-315)."""
    x, y, z, w = seed
    out = np.empty(n, np.uint8)
    for i in range(n):
        t = (x ^ (x << 11)) & 0xFFFFFFFF
        x, y, z = y, z, w
        w = (w ^ (w >> 19) ^ t ^ (t >> 8)) & 0xFFFFFFFF
        out[i] = w & 0xFF
    return out.tobytes()


class Tcs:
    """One C++ TestConstraintSystem (+ its device handle or host recorder)."""

    def __init__(self, field: int, device: int = 0, named: bool = True, reserve=(0, 0, 0)):
        self.L = _lib() if device >= 0 else _host_lib()  # host recording never touches the CUDA library
        self.t = vp()
        rc = self.L.bp_tcs_new(field, device, int(named), *reserve, ctypes.byref(self.t))
        if rc != 0:
            raise RuntimeError(f"bp_tcs_new -> {rc} (device {device}; there is no CPU evaluation path)")
        self.field = field

    def close(self, keep_handle: bool = False):
        if self.t:
            self.L.bp_tcs_free(self.t)
            self.t = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"bp_tcs error {rc}: {(self.L.bp_tcs_last_error(self.t) or b'').decode()}")

    @property
    def handle(self):
        return self.L.bp_tcs_handle(self.t)

    def sha256_block(self, block: bytes) -> bytes:
        assert len(block) == 64
        out = ctypes.create_string_buffer(32)
        self._ck(self.L.bp_tcs_sha256_block(self.t, block, out))
        return out.raw

    def sha256(self, msg: bytes, block_begin: int = 0, block_end: int = U64_MAX) -> Tuple[bytes, int]:
        out = ctypes.create_string_buffer(32)
        before = ctypes.c_uint64()
        self._ck(self.L.bp_tcs_sha256(self.t, msg, len(msg), block_begin, block_end, out, ctypes.byref(before)))
        return out.raw, before.value

    def sha256_ranges(self, msg: bytes, ranges):
        """Rows of the compression blocks in `ranges` (list of (begin, end), ascending, disjoint) only; returns (digest,
        [(global_rows_before, local_rows_before) per range])."""
        flat = (ctypes.c_uint64 * (2 * len(ranges)))(*[x for r in ranges for x in r])
        gb = (ctypes.c_uint64 * len(ranges))()
        lb = (ctypes.c_uint64 * len(ranges))()
        out = ctypes.create_string_buffer(32)
        self._ck(self.L.bp_tcs_sha256_ranges(self.t, msg, len(msg), flat, len(ranges), out, gb, lb))
        return out.raw, [(int(g), int(l)) for g, l in zip(gb, lb)]

    @staticmethod
    def _limbs(value: int) -> np.ndarray:
        return np.frombuffer(int(value).to_bytes(32, "little"), dtype="<u8").copy()

    BOOLEAN_OPS = {"xor": 0, "and": 1, "or": 2, "sha256_ch": 3, "sha256_maj": 4, "enforce_equal": 5}
    OPERAND_KINDS = ["True", "False", "AllocatedTrue", "AllocatedFalse", "NegatedAllocatedTrue", "NegatedAllocatedFalse"]

    def boolean_op(self, op: str, a: str, b: str, c: str = "True"):
        """Boolean::xor / and / or / sha256_ch / sha256_maj / enforce_equal over operands built like the reference tests'
        dyn_construct (boolean.rs:1109-2003); returns (result kind: 'Is' / 'Not' / 'Constant', result value)."""
        k, v = ctypes.c_int(), ctypes.c_int()
        kinds = [self.OPERAND_KINDS.index(x) for x in (a, b, c)]
        self._ck(self.L.bp_tcs_boolean_op(self.t, self.BOOLEAN_OPS[op], *kinds, ctypes.byref(k), ctypes.byref(v)))
        return ["Is", "Not", "Constant"][k.value], bool(v.value)

    UINT32_OPS = {"xor": 0, "addmany": 1, "sha256_maj": 2, "sha256_ch": 3}

    def uint32_op(self, op: str, a: int, b: int, c: int, d: int = 0):
        """The reference's uint32 test circuits (uint32.rs:492-780); returns (value of the result word, number of constant bits)."""
        r, k = ctypes.c_uint32(), ctypes.c_uint32()
        self._ck(self.L.bp_tcs_uint32_op(self.t, self.UINT32_OPS[op], a, b, c, d, ctypes.byref(r), ctypes.byref(k)))
        return r.value, k.value

    def u64_bits(self, value: int):
        """u64_into_boolean_vec_le (boolean.rs:274-304); returns the 64 bit values, little-endian."""
        bits = np.zeros(64, np.uint8)
        self._ck(self.L.bp_tcs_u64_bits(self.t, value, bits.ctypes.data))
        return bits

    def num_unpack(self, value: int, strict: bool = False):
        """AllocatedNum::alloc + to_bits_le[_strict] (num.rs:128-274); returns the 255 bits, little-endian."""
        v, bits = self._limbs(value), np.zeros(255, np.uint8)
        self._ck(self.L.bp_tcs_num_unpack(self.t, v.ctypes.data, int(strict), bits.ctypes.data))
        return bits

    def num_arith(self, a: int, b: int):
        va, vb = self._limbs(a), self._limbs(b)
        self._ck(self.L.bp_tcs_num_arith(self.t, va.ctypes.data, vb.ctypes.data))

    def num_chain(self, n_steps: int, unpack_every: int, x0: int, y0: int):
        vx, vy = self._limbs(x0), self._limbs(y0)
        self._ck(self.L.bp_tcs_num_chain(self.t, n_steps, unpack_every, vx.ctypes.data, vy.ctypes.data))

    def record_witness_program(self, on: bool = True):
        self._ck(self.L.bp_tcs_record_witness_program(self.t, int(on)))

    def witness_program(self) -> np.ndarray:
        """The device witness program recorded by the last sha256 synthesis (uint32 words; include/bp_r1cs.h)."""
        p, n = vp(), ctypes.c_uint64()
        self._ck(self.L.bp_tcs_witness_program(self.t, ctypes.byref(p), ctypes.byref(n)))
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), shape=(n.value,)).copy()

    def blake2s(self, msg: bytes, personalization: bytes = b"12345678") -> bytes:
        assert len(personalization) == 8
        out = ctypes.create_string_buffer(32)
        self._ck(self.L.bp_tcs_blake2s(self.t, msg, len(msg), personalization, out))
        return out.raw

    def which_is_unsatisfied(self) -> Optional[str]:
        buf = ctypes.create_string_buffer(512)
        row = self.L.bp_tcs_which_is_unsatisfied(self.t, buf, 512)
        if row < -1:
            self._ck(int(row) + 1)
        return None if row < 0 else (buf.value.decode() or str(row))

    def first_unsatisfied_row(self) -> int:
        row = self.L.bp_tcs_which_is_unsatisfied(self.t, None, 0)
        if row < -1:
            self._ck(int(row) + 1)
        return int(row)

    def is_satisfied(self) -> bool:
        return self.first_unsatisfied_row() < 0

    def set(self, path: str, value: int):
        v = np.frombuffer(int(value).to_bytes(32, "little"), dtype="<u8").copy()
        self._ck(self.L.bp_tcs_set(self.t, path.encode(), v.ctypes.data))

    def get(self, path: str) -> int:
        v = np.zeros(4, np.uint64)
        self._ck(self.L.bp_tcs_get(self.t, path.encode(), v.ctypes.data))
        return int.from_bytes(v.tobytes(), "little")

    def num_constraints(self) -> int:
        return self.L.bp_tcs_num_constraints(self.t)

    def num_inputs(self) -> int:
        return self.L.bp_tcs_num_inputs(self.t)

    def num_aux(self) -> int:
        return self.L.bp_tcs_num_aux(self.t)

    def row_path(self, row: int) -> str:
        buf = ctypes.create_string_buffer(512)
        self._ck(self.L.bp_tcs_row_path(self.t, row, buf, 512))
        return buf.value.decode()

    def host_csr(self):
        """(lens, cols, coeffs[nnz,4], inputs[n,4], aux[n,4]) copies of a host-recording system."""
        p = [vp() for _ in range(5)]
        n = [ctypes.c_uint64() for _ in range(4)]
        self._ck(self.L.bp_tcs_host_csr(self.t, ctypes.byref(p[0]), ctypes.byref(n[0]), ctypes.byref(p[1]), ctypes.byref(p[2]),
                                         ctypes.byref(n[1]), ctypes.byref(p[3]), ctypes.byref(n[2]), ctypes.byref(p[4]), ctypes.byref(n[3])))
        n_rows, nnz, n_in, n_aux = (x.value for x in n)

        def arr(ptr, count, ctype, dtype):
            if count == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(count,)).copy()

        lens = arr(p[0], 3 * n_rows, ctypes.c_uint32, np.uint32)
        cols = arr(p[1], nnz, ctypes.c_uint32, np.uint32)
        coeffs = arr(p[2], 4 * nnz, ctypes.c_uint64, np.uint64).reshape(-1, 4)
        inputs = arr(p[3], 4 * n_in, ctypes.c_uint64, np.uint64).reshape(-1, 4)
        aux = arr(p[4], 4 * n_aux, ctypes.c_uint64, np.uint64).reshape(-1, 4)
        return lens, cols, coeffs, inputs, aux


def chain_message(blocks: int) -> bytes:
    """Message whose sha256() circuit has exactly `blocks` compression calls (the last one holds the padding)."""
    return xorshift_bytes(64 * blocks - 9)


def witness_cs_selftest(field: int, device: int = 0) -> int:
    return int(_lib().bp_wcs_selftest(field, device))


def blake2s_into_new_handle(field: int, device: int, n_bytes: int, record_witness_program: bool = False):
    """BASELINE configs[2]: blake2s gadget over an n_bytes preimage (xorshift bytes), streamed into a fresh device handle."""
    blocks = max(1, (n_bytes + 63) // 64)
    t = Tcs(field, device, named=False, reserve=(blocks * 21600 + 4096, blocks * 140000 + 65536, blocks * 21700 + 4096))
    if record_witness_program:
        t.record_witness_program()
    digest = t.blake2s(xorshift_bytes(n_bytes))
    info = {"rows_total": t.num_constraints(), "row0": 0, "bytes": n_bytes, "digest": digest.hex(), "tcs": t}
    if record_witness_program:
        info["witness_program"] = t.witness_program()
    return vp(t.handle), info


def blake2s_host_csr(field: int, n_bytes: int):
    with Tcs(field, device=-1, named=False) as t:
        t.blake2s(xorshift_bytes(n_bytes))
        return t.host_csr()


def sha256_chain_host_csr(field: int, blocks: int):
    with Tcs(field, device=-1, named=False) as t:
        t.sha256(chain_message(blocks))
        return t.host_csr()


def sha256_chain_into_new_handle(field: int, device: int, blocks: int, rank: int = 0, world: int = 1, record_witness_program: bool = False):
    """BASELINE configs[1]: sha256 gadget over `blocks` chained compression blocks, rows of this rank's block range
    streamed into a fresh device handle.  Returns (bp_cs handle as c_void_p, info); the caller frees via `info['tcs']`."""
    # rows are what is balanced: the shard holding block 0 also holds the 8 * len(msg) boolean rows of the input bits
    from .sharding import split_units_by_rows

    b0, b1 = split_units_by_rows(blocks, 26192, 8 * (64 * blocks - 9), rank, world)
    per_block_rows, per_block_terms, per_block_vars = 26400, 170000, 26500
    t = Tcs(field, device, named=False,
            reserve=((b1 - b0) * per_block_rows + 4096, (b1 - b0) * per_block_terms + 65536, blocks * per_block_vars + 4096))
    if record_witness_program:
        t.record_witness_program()
    digest, before = t.sha256(chain_message(blocks), b0, b1)
    L = ffi.load()
    h = vp(t.handle)
    assert L.bp_cs_set_row_base(h, before) == 0
    info = {"rows_total": t.num_constraints(), "row0": before, "blocks": blocks, "digest": digest.hex(), "tcs": t}
    if record_witness_program:
        info["witness_program"] = t.witness_program()
    return h, info


def sha256_chain_states(msg: bytes) -> np.ndarray:
    """uint32[blocks, 8]: the hash state before each compression block (plain SHA-256 on the host)."""
    L = _host_lib()
    n = ctypes.c_uint64()
    assert L.bp_sha256_chain_states(msg, len(msg), None, 0, ctypes.byref(n)) == 0
    out = np.zeros((n.value, 8), np.uint32)
    assert L.bp_sha256_chain_states(msg, len(msg), out.ctypes.data, n.value, ctypes.byref(n)) == 0
    return out


def blake2s_chain_states(msg: bytes, personalization: bytes = b"12345678") -> np.ndarray:
    """uint32[blocks, 8]: the BLAKE2s chaining value before each compression of the gadget's hash (plain BLAKE2s on the host)."""
    assert len(personalization) == 8
    L = _host_lib()
    n = ctypes.c_uint64()
    assert L.bp_blake2s_chain_states(msg, len(msg), personalization, None, 0, ctypes.byref(n)) == 0
    out = np.zeros((n.value, 8), np.uint32)
    assert L.bp_blake2s_chain_states(msg, len(msg), personalization, out.ctypes.data, n.value, ctypes.byref(n)) == 0
    return out
