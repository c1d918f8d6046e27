"""Host-side mirror of the reference's constraint-system interface over the C ABI.

Same names, argument meaning and error behaviour as the reference, so tests read like the reference's
own (paths relative to /root/reference):

* `Variable`, `Index`, `LinearCombination`      crates/bellpepper-core/src/lc.rs
* `ConstraintSystem` protocol, `Namespace`      crates/bellpepper-core/src/constraint_system.rs:61-333
* `TestConstraintSystem`                        crates/bellpepper-core/src/util_cs/test_cs.rs  (B200-backed)
* `WitnessCS`                                   crates/bellpepper/src/util_cs/witness_cs.rs    (B200-backed storage)

What stays on the host: names, namespaces, closures, LC construction algebra (`lc + (coeff, var)` ...).
What crosses the ABI: flat canonical field elements, tagged u32 columns, per-row LC lengths.  All
evaluation -- `which_is_unsatisfied`, `is_satisfied`, `LinearCombination.eval`, `eval_all` -- runs on the
GPU through libbp_r1cs.so; there is no CPU evaluation path in this package (it fails loudly if the
library or a device is missing).

The production host front-end for large circuits is the C++ one (csrc/host/); this Python mirror exists
for API-level parity tests and small circuits.
"""

from __future__ import annotations

import bisect
import ctypes
from typing import Callable, Iterable, List, Optional, Tuple

import numpy as np

from . import ffi
from .fields import MODULUS

INPUT, AUX = 0, 1


class SynthesisError(Exception):
    """constraint_system.rs:21-57."""


class AssignmentMissing(SynthesisError):
    def __str__(self):
        return "an assignment for a variable could not be computed"


class DivisionByZero(SynthesisError):
    def __str__(self):
        return "division by zero"


class Unsatisfiable(SynthesisError):
    def __str__(self):
        return "unsatisfiable constraint system"


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"bp_r1cs error {code}: {msg}")
        self.code = code


class Index:
    """lc.rs:27-30."""

    Input = INPUT
    Aux = AUX


class Variable:
    """lc.rs:8-22."""

    __slots__ = ("kind", "index")

    def __init__(self, kind: int, index: int):
        self.kind, self.index = kind, index

    @staticmethod
    def new_unchecked(kind: int, index: int) -> "Variable":
        return Variable(kind, index)

    def get_unchecked(self) -> Tuple[int, int]:
        return self.kind, self.index

    def __eq__(self, o):
        return isinstance(o, Variable) and (self.kind, self.index) == (o.kind, o.index)

    def __hash__(self):
        return hash((self.kind, self.index))

    def __repr__(self):
        return f"Variable({'Aux' if self.kind == AUX else 'Input'}({self.index}))"

    def tagged(self) -> int:
        return self.index | (ffi.COL_AUX if self.kind == AUX else 0)


def _limbs(values: Iterable[int]) -> np.ndarray:
    vals = list(values)
    if not vals:
        return np.zeros((0, 4), np.uint64)
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype="<u8").reshape(-1, 4).copy()


def _ints(arr: np.ndarray) -> List[int]:
    raw = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [int.from_bytes(raw[i : i + 32], "little") for i in range(0, len(raw), 32)]


class LinearCombination:
    """lc.rs:35-38 and the operator algebra at lc.rs:270-375: two index-sorted term lists (inputs, aux),
    unique keys, same-key insertions add coefficients, zero coefficients are retained (lc.rs:74-113)."""

    __slots__ = ("p", "_k", "_v")

    def __init__(self, p: int):
        self.p = p
        self._k = ([], [])  # keys per index space
        self._v = ([], [])  # coefficients per index space

    @classmethod
    def zero(cls, p: int) -> "LinearCombination":
        return cls(p)

    @classmethod
    def from_coeff(cls, p: int, var: Variable, coeff: int) -> "LinearCombination":
        lc = cls(p)
        lc._insert(var.kind, var.index, coeff % p)
        return lc

    @classmethod
    def from_variable(cls, p: int, var: Variable) -> "LinearCombination":
        return cls.from_coeff(p, var, 1)

    def _insert(self, kind: int, key: int, coeff: int):
        ks, vs = self._k[kind], self._v[kind]
        if ks and ks[-1] < key:  # the common append case
            ks.append(key)
            vs.append(coeff)
            return
        i = bisect.bisect_left(ks, key)
        if i < len(ks) and ks[i] == key:
            vs[i] = (vs[i] + coeff) % self.p
        else:
            ks.insert(i, key)
            vs.insert(i, coeff)

    def _clone(self) -> "LinearCombination":
        c = LinearCombination(self.p)
        c._k = (self._k[0][:], self._k[1][:])
        c._v = (self._v[0][:], self._v[1][:])
        return c

    def iter(self):
        """lc.rs:155-160: inputs first, then aux."""
        for kind in (INPUT, AUX):
            for k, v in zip(self._k[kind], self._v[kind]):
                yield Variable(kind, k), v

    def iter_inputs(self):
        return zip(self._k[INPUT], self._v[INPUT])

    def iter_aux(self):
        return zip(self._k[AUX], self._v[AUX])

    def __len__(self):
        return len(self._k[0]) + len(self._k[1])

    def is_empty(self) -> bool:
        return len(self) == 0

    def _combine(self, other, sign: int) -> "LinearCombination":
        p = self.p
        out = self._clone()
        if isinstance(other, Variable):
            out._insert(other.kind, other.index, sign % p)
        elif isinstance(other, LinearCombination):
            for var, c in other.iter():
                out._insert(var.kind, var.index, (sign * c) % p)
        elif isinstance(other, tuple) and len(other) == 2 and isinstance(other[1], Variable):
            out._insert(other[1].kind, other[1].index, (sign * other[0]) % p)
        elif isinstance(other, tuple) and len(other) == 2 and isinstance(other[1], LinearCombination):
            for var, c in other[1].iter():
                out._insert(var.kind, var.index, (sign * other[0] * c) % p)
        else:
            raise TypeError(f"unsupported LinearCombination operand: {other!r}")
        return out

    def __add__(self, other):
        return self._combine(other, 1)

    def __sub__(self, other):
        return self._combine(other, -1)

    def flat(self) -> Tuple[List[int], List[int]]:
        cols = self._k[INPUT] + [k | ffi.COL_AUX for k in self._k[AUX]]
        return cols, self._v[INPUT] + self._v[AUX]

    def eval(self, cs: "TestConstraintSystem") -> int:
        """LinearCombination::eval (lc.rs:245-267) against a B200-backed system's current witness."""
        return cs.eval_lc(self)


def compute_path(ns: List[str], this: str) -> str:
    """test_cs.rs:363-375 (panics -> AssertionError)."""
    assert "/" not in this, "'/' is not allowed in names"
    return this if not ns else "/".join(ns) + "/" + this


def _s(annotation) -> str:
    return annotation() if callable(annotation) else annotation


class Namespace:
    """constraint_system.rs:242-333: forwards to the root; pops the namespace when it goes out of scope
    (use as a context manager; `close()` is the explicit Drop)."""

    def __init__(self, root):
        self._root = root
        self._open = True

    def one(self):
        return self._root.one()

    def alloc(self, annotation, f):
        return self._root.alloc(annotation, f)

    def alloc_input(self, annotation, f):
        return self._root.alloc_input(annotation, f)

    def enforce(self, annotation, a, b, c):
        return self._root.enforce(annotation, a, b, c)

    def push_namespace(self, _name):
        raise AssertionError("push_namespace is not forwarded by Namespace (constraint_system.rs:285-291)")

    def pop_namespace(self):
        raise AssertionError("pop_namespace is not forwarded by Namespace (constraint_system.rs:293-295)")

    def get_root(self):
        return self._root

    def namespace(self, name):
        return self._root.namespace(name)

    def is_witness_generator(self):
        return self._root.is_witness_generator()

    def extend_inputs(self, v):
        return self._root.extend_inputs(v)

    def extend_aux(self, v):
        return self._root.extend_aux(v)

    def allocate_empty(self, aux_n, inputs_n):
        return self._root.allocate_empty(aux_n, inputs_n)

    def inputs_slice(self):
        return self._root.inputs_slice()

    def aux_slice(self):
        return self._root.aux_slice()

    def close(self):
        if self._open:
            self._open = False
            self._root.pop_namespace()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _Device:
    """Owns one `bp_cs*` and batches host-side appends before they cross the ABI."""

    FLUSH_TERMS = 1 << 16

    def __init__(self, field: int, device: int = 0, reserve_rows: int = 0, reserve_nnz: int = 0, reserve_vars: int = 0):
        self.L = ffi.load()
        self.field = field
        self.p = MODULUS[field]
        h = ffi.vp()
        rc = self.L.bp_cs_new(field, device, reserve_rows, reserve_nnz, reserve_vars, ctypes.byref(h))
        if rc != ffi.BP_OK:
            raise NativeError(rc, "bp_cs_new failed (no CUDA device? this package has no CPU path)")
        self.h = h
        self._pend_vals = ([], [])  # pending allocs per index space
        self._count = [1, 0]  # committed + pending element counts (inputs start with ONE)
        self._lens: List[int] = []
        self._cols: List[int] = []
        self._coeffs: List[int] = []
        self.n_rows = 0
        self._structure = bytearray()  # serialised LCs for TestConstraintSystem.hash (test_cs.rs:64-115)
        self._row_digests: List[bytes] = []  # one digest per constraint over its three raw LCs, for `delta`

    def close(self):
        if getattr(self, "h", None):
            self.L.bp_cs_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != ffi.BP_OK:
            raise NativeError(rc, (self.L.bp_cs_last_error(self.h) or b"").decode())

    # -- appends ------------------------------------------------------------------------------------
    def push_var(self, kind: int, value: int) -> int:
        idx = self._count[kind]
        self._pend_vals[kind].append(value % self.p)
        self._count[kind] += 1
        return idx

    def push_row(self, a: LinearCombination, b: LinearCombination, c: LinearCombination) -> int:
        import hashlib

        raw = hashlib.blake2s(digest_size=16)
        for lc in (a, b, c):
            cols, coeffs = lc.flat()
            # LinearCombination == (derived PartialEq, lc.rs:34-46): the term lists as they are, zero coefficients included
            raw.update(len(cols).to_bytes(4, "little") + b"".join(col.to_bytes(4, "little") + co.to_bytes(32, "little")
                                                                  for col, co in zip(cols, coeffs)))
            self._lens.append(len(cols))
            self._cols.extend(cols)
            self._coeffs.extend(coeffs)
            # hash_lc (test_cs.rs:89-115): proc_lc drops zero coefficients; inputs sort before aux, ascending -- the flat order
            kept = [(col, co) for col, co in zip(cols, coeffs) if co != 0]
            self._structure += len(kept).to_bytes(8, "big")
            for col, co in kept:
                self._structure += (b"A" if col & ffi.COL_AUX else b"I") + (col & 0x7FFFFFFF).to_bytes(8, "big") + co.to_bytes(32, "big")
        self._row_digests.append(raw.digest())
        self.n_rows += 1
        if len(self._cols) >= self.FLUSH_TERMS:
            self.flush()
        return self.n_rows - 1

    def flush(self):
        for kind in (INPUT, AUX):
            vals = self._pend_vals[kind]
            if vals:
                # packed staging (include/bp_r1cs.h: bp_cs_alloc_u8): one byte per value that fits a byte, the others
                # are patched afterwards -- gadget witnesses are almost entirely bits
                first = ctypes.c_uint64()
                packed = np.fromiter((v if v < 256 else 0 for v in vals), np.uint8, len(vals))
                self._ck(self.L.bp_cs_alloc_u8(self.h, kind, packed.ctypes.data, len(vals), ctypes.byref(first)))
                for pos, v in enumerate(vals):
                    if v >= 256:
                        arr = _limbs([v])
                        self._ck(self.L.bp_cs_set(self.h, kind, first.value + pos, arr.ctypes.data))
                vals.clear()
        if self._lens:
            lens = np.asarray(self._lens, np.uint32)
            cols = np.asarray(self._cols, np.uint32)
            coeffs = _limbs(self._coeffs)
            self._ck(self.L.bp_cs_enforce(self.h, lens.size // 3, lens.ctypes.data, cols.ctypes.data, coeffs.ctypes.data))
            self._lens.clear()
            self._cols.clear()
            self._coeffs.clear()

    # -- element access -----------------------------------------------------------------------------
    def set(self, kind: int, idx: int, value: int):
        self.flush()
        arr = _limbs([value % self.p])
        self._ck(self.L.bp_cs_set(self.h, kind, idx, arr.ctypes.data))

    def get(self, kind: int, idx: int) -> int:
        self.flush()
        out = np.zeros(4, np.uint64)
        self._ck(self.L.bp_cs_get(self.h, kind, idx, out.ctypes.data))
        return _ints(out)[0]

    def witness(self, kind: int) -> List[int]:
        self.flush()
        n = self._count[kind]
        out = np.zeros((n, 4), np.uint64)
        self._ck(self.L.bp_cs_witness(self.h, kind, 0, n, out.ctypes.data))
        return _ints(out)

    def set_range(self, kind: int, first: int, values: List[int]):
        self.flush()
        arr = _limbs([v % self.p for v in values])
        self._ck(self.L.bp_cs_set_range(self.h, kind, first, len(values), arr.ctypes.data))

    # -- evaluation (GPU) ---------------------------------------------------------------------------
    def first_unsatisfied(self) -> int:
        self.flush()
        row = ctypes.c_int64()
        self._ck(self.L.bp_cs_first_unsatisfied(self.h, ctypes.byref(row)))
        return row.value

    def eval_all(self):
        self.flush()
        n = self.n_rows
        az, bz, cz = (np.zeros((n, 4), np.uint64) for _ in range(3))
        self._ck(self.L.bp_cs_eval(self.h, az.ctypes.data, bz.ctypes.data, cz.ctypes.data))
        return _ints(az), _ints(bz), _ints(cz)

    def eval_lc(self, lc: LinearCombination) -> int:
        self.flush()
        cols, coeffs = lc.flat()
        c = np.asarray(cols, np.uint32)
        v = _limbs(coeffs)
        out = np.zeros(4, np.uint64)
        self._ck(self.L.bp_cs_eval_lc(self.h, c.ctypes.data, v.ctypes.data, len(cols), out.ctypes.data))
        return _ints(out)[0]

    def counts(self):
        self.flush()
        a, b, c, d = (ctypes.c_uint64() for _ in range(4))
        self._ck(self.L.bp_cs_counts(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return a.value, b.value, c.value, d.value

    def set_option(self, key: str, value: int):
        self._ck(self.L.bp_cs_set_option(self.h, key.encode(), value))

    def get_option(self, key: str) -> int:
        v = ctypes.c_int64()
        self._ck(self.L.bp_cs_get_option(self.h, key.encode(), ctypes.byref(v)))
        return v.value


ONE = Variable(INPUT, 0)


class _ConstraintSystemBase:
    """Provided methods of `trait ConstraintSystem` (constraint_system.rs:61-237)."""

    @staticmethod
    def one() -> Variable:
        return ONE  # constraint_system.rs:73-75

    def namespace(self, name) -> Namespace:
        self.get_root().push_namespace(name)
        return Namespace(self.get_root())

    def get_root(self):
        return self

    @staticmethod
    def is_extensible() -> bool:
        return False

    def extend(self, _other):
        raise NotImplementedError("ConstraintSystem::extend must be implemented for types implementing ConstraintSystem")

    def is_witness_generator(self) -> bool:
        return False

    def extend_inputs(self, _v):
        assert self.is_witness_generator()
        raise NotImplementedError

    def extend_aux(self, _v):
        assert self.is_witness_generator()
        raise NotImplementedError

    def allocate_empty(self, _aux_n, _inputs_n):
        assert self.is_witness_generator()
        raise NotImplementedError

    def inputs_slice(self):
        assert self.is_witness_generator()
        raise NotImplementedError

    def aux_slice(self):
        assert self.is_witness_generator()
        raise NotImplementedError


class TestConstraintSystem(_ConstraintSystemBase):
    """B200-backed replacement for `TestConstraintSystem` (test_cs.rs:19-447).

    Names, namespaces and the row -> path table live here; witness values and the three CSR matrices
    live in HBM behind the C ABI.  `which_is_unsatisfied` maps the GPU's first failing row back to the
    path the reference would return.
    """

    __test__ = False

    def __init__(self, field: int = 0, device: int = 0, **reserve):
        self.field = field
        self.p = MODULUS[field]
        self.dev = _Device(field, device, **reserve)
        self.named_objects = {"ONE": ("var", ONE)}
        self.current_namespace: List[str] = []
        self.constraint_paths: List[str] = []
        self.input_names: List[str] = ["ONE"]
        self.aux_names: List[str] = []

    @classmethod
    def new(cls, field: int = 0, device: int = 0):
        return cls(field, device)

    def close(self):
        self.dev.close()

    # ---- trait ConstraintSystem (test_cs.rs:377-447) --------------------------------------------
    def _set_named_obj(self, path: str, obj):
        assert path not in self.named_objects, f"tried to create object at existing path: {path}"
        self.named_objects[path] = obj

    def alloc(self, annotation, f: Callable[[], int]) -> Variable:
        path = compute_path(self.current_namespace, _s(annotation))
        value = f()  # may raise SynthesisError: nothing is registered in that case (test_cs.rs:388)
        index = self.dev.push_var(AUX, value)
        self.aux_names.append(path)
        var = Variable(AUX, index)
        self._set_named_obj(path, ("var", var))
        return var

    def alloc_input(self, annotation, f: Callable[[], int]) -> Variable:
        path = compute_path(self.current_namespace, _s(annotation))
        value = f()
        index = self.dev.push_var(INPUT, value)
        self.input_names.append(path)
        var = Variable(INPUT, index)
        self._set_named_obj(path, ("var", var))
        return var

    def enforce(self, annotation, a, b, c) -> None:
        path = compute_path(self.current_namespace, _s(annotation))
        self._set_named_obj(path, ("constraint", len(self.constraint_paths)))
        z = LinearCombination.zero
        self.dev.push_row(a(z(self.p)), b(z(self.p)), c(z(self.p)))
        self.constraint_paths.append(path)

    def push_namespace(self, name) -> None:
        name = _s(name)
        self._set_named_obj(compute_path(self.current_namespace, name), ("namespace", None))
        self.current_namespace.append(name)

    def pop_namespace(self) -> None:
        assert self.current_namespace
        self.current_namespace.pop()

    # ---- the hot path (test_cs.rs:239-264) ------------------------------------------------------
    def which_is_unsatisfied(self) -> Optional[str]:
        row = self.dev.first_unsatisfied()
        return None if row < 0 else self.constraint_paths[row]

    def is_satisfied(self) -> bool:
        b = self.which_is_unsatisfied()
        if b is not None:
            print(f'fail: "{b}"')  # test_cs.rs:258
            return False
        return True

    def eval_all(self):
        """Batched LinearCombination::eval over every row: (Az, Bz, Cz) lists of canonical ints."""
        return self.dev.eval_all()

    def eval_lc(self, lc: LinearCombination) -> int:
        return self.dev.eval_lc(lc)

    # ---- accessors (test_cs.rs:175-334) ---------------------------------------------------------
    def scalar_inputs(self) -> List[int]:
        return self.dev.witness(INPUT)

    def scalar_aux(self) -> List[int]:
        return self.dev.witness(AUX)

    def hash(self) -> str:
        """`TestConstraintSystem::hash` (test_cs.rs:214-237): Blake2s over the counts and every constraint's three LCs."""
        import hashlib

        h = hashlib.blake2s()
        h.update(self.dev._count[0].to_bytes(8, "big") + self.dev._count[1].to_bytes(8, "big") + self.dev.n_rows.to_bytes(8, "big"))
        h.update(bytes(self.dev._structure))
        return h.hexdigest()

    def delta(self, other: "TestConstraintSystem", ignore_counts: bool = False):
        """`Comparable::delta` (util_cs/mod.rs:39-76): ("Equal",), ("InputCountMismatch", a, b), ("ConstraintCountMismatch", a, b),
        ("ConstraintMismatch", index, path_here, path_there) for the first constraint that differs (LCs or path), or ("Different",)."""
        mine = [d + p.encode() for d, p in zip(self.dev._row_digests, self.constraint_paths)]
        theirs = [d + p.encode() for d, p in zip(other.dev._row_digests, other.constraint_paths)]
        input_count_matches = self.num_inputs() == other.num_inputs()
        constraint_count_matches = self.num_constraints() == other.num_constraints()
        inputs_match = self.input_names == other.input_names
        constraints_match = mine == theirs
        if not ignore_counts and not input_count_matches:
            return ("InputCountMismatch", self.num_inputs(), other.num_inputs())
        if not ignore_counts and not constraint_count_matches:
            return ("ConstraintCountMismatch", self.num_constraints(), other.num_constraints())
        if not constraints_match:
            for i, (x, y) in enumerate(zip(mine, theirs)):
                if x != y:
                    return ("ConstraintMismatch", i, self.constraint_paths[i], other.constraint_paths[i])
            raise IndexError("constraint lists differ only in length")  # the reference unwraps a None here (mod.rs:66)
        if input_count_matches and constraint_count_matches and inputs_match:
            return ("Equal",)
        return ("Different",)

    def num_constraints(self) -> int:
        return len(self.constraint_paths)

    def num_inputs(self) -> int:
        return len(self.input_names)

    def _var_at(self, path: str, verb: str) -> Variable:
        obj = self.named_objects.get(path)
        if obj is None:
            raise KeyError(f"no variable exists at path: {path}")
        if obj[0] != "var":
            raise KeyError(f"tried to {verb} path `{path}`, but `{obj[0]}` exists there (not a variable)")
        return obj[1]

    def set(self, path: str, to: int) -> None:
        v = self._var_at(path, "set")
        self.dev.set(v.kind, v.index, to)

    def get(self, path: str) -> int:
        v = self._var_at(path, "get value of")
        return self.dev.get(v.kind, v.index)

    def get_input(self, index: int, path: str) -> int:
        assert self.input_names[index] == path
        return self.dev.get(INPUT, index)

    def get_inputs(self) -> List[Tuple[int, str]]:
        return list(zip(self.scalar_inputs(), self.input_names))

    def verify(self, expected: List[int]) -> bool:
        assert len(expected) + 1 == len(self.input_names)
        got = self.scalar_inputs()[1:]
        return all(a == b % self.p for a, b in zip(got, expected))

    def pretty_print_list(self) -> List[str]:
        return [f"INPUT {n}" for n in self.input_names] + [f"AUX {n}" for n in self.aux_names] + list(self.constraint_paths)

    def pretty_print(self) -> str:
        return "\n".join(self.pretty_print_list())

    def pretty_print_equations(self) -> str:
        """`MetricCS::pretty_print` (crates/bellpepper/src/util_cs/metric_cs.rs:130-195): the inputs, then one line per constraint
        `path: (A) * (B) = (C)` with every LC normalised by `proc_lc` (merged, zero coefficients dropped, inputs before aux).
        Coefficients: -1 prints as " - ", 1 as nothing, a power of two as "2^i . " FOLLOWED by the scalar's Debug form (the
        reference breaks out of its search loop and then prints the scalar as well), anything else as the Debug form alone:
        `Scalar(0x...)` for BLS12-381 (blstrs), `0x...` for the pasta fields, 64 hex digits big-endian."""
        p = self.p
        pow2 = {pow(2, i, p): i for i in range(255)}
        dbg = (lambda c: f"Scalar(0x{c:064x})") if self.dev.field == 0 else (lambda c: f"0x{c:064x}")
        out = [f"INPUT {n}\n" for n in self.input_names]
        buf, pos = bytes(self.dev._structure), 0

        def pp() -> str:
            nonlocal pos
            n = int.from_bytes(buf[pos:pos + 8], "big")
            pos += 8
            parts, first = ["("], True
            for _ in range(n):
                kind, idx, co = buf[pos:pos + 1], int.from_bytes(buf[pos + 1:pos + 9], "big"), int.from_bytes(buf[pos + 9:pos + 41], "big")
                pos += 41
                if co == p - 1:
                    parts.append(" - ")
                elif not first:
                    parts.append(" + ")
                first = False
                if co != 1 and co != p - 1:
                    if co in pow2:
                        parts.append(f"2^{pow2[co]} . ")
                    parts.append(f"{dbg(co)} . ")
                parts.append(f"`I{self.input_names[idx]}`" if kind == b"I" else f"`A{self.aux_names[idx]}`")
            if first:
                parts.append("0")
            parts.append(")")
            return "".join(parts)

        for path in self.constraint_paths:
            a, b, c = pp(), pp(), pp()
            out.append(f"\n{path}: {a} * {b} = {c}")
        out.append("\n")
        return "".join(out)


class SizedWitness:
    """`SizedWitness` (witness_cs.rs:7-41): a bulk witness producer of known size.  Subclasses give the three counts and
    `generate_witness_into(aux, inputs) -> result`, which fills the two lists in place (the reference's `&mut [Scalar]`)."""

    def num_constraints(self) -> int:
        raise NotImplementedError

    def num_inputs(self) -> int:
        raise NotImplementedError

    def num_aux(self) -> int:
        raise NotImplementedError

    def generate_witness_into(self, aux: List[int], inputs: List[int]) -> int:
        raise NotImplementedError

    def generate_witness(self):  # witness_cs.rs:13-26
        aux, inputs = [0] * self.num_aux(), [0] * self.num_inputs()
        result = self.generate_witness_into(aux, inputs)
        return aux, inputs, result

    def generate_witness_into_cs(self, cs) -> int:  # witness_cs.rs:28-40
        assert cs.is_witness_generator()
        aux_count, inputs_count = self.num_aux(), self.num_inputs()
        a0, i0 = cs.allocate_empty(aux_count, inputs_count)
        aux, inputs = [0] * aux_count, [0] * inputs_count
        result = self.generate_witness_into(aux, inputs)
        cs.fill_aux(a0, aux)        # one bulk upload each: the flat-buffer equivalent of writing through the slices
        cs.fill_inputs(i0, inputs)
        return result


class WitnessCS(_ConstraintSystemBase):
    """`WitnessCS` (witness_cs.rs:45-201) with its two flat assignment vectors held in HBM.

    `enforce` is a no-op exactly as in the reference (witness_cs.rs:125-134).  `allocate_empty` cannot hand
    out `&mut [Scalar]` windows into device memory; it returns the start indices and the caller fills the
    window with `fill_aux` / `fill_inputs` (one bulk H2D each), which is the flat-buffer equivalent.
    """

    def __init__(self, field: int = 0, device: int = 0, **reserve):
        self.field = field
        self.p = MODULUS[field]
        self.dev = _Device(field, device, **reserve)

    @classmethod
    def new(cls, field: int = 0, device: int = 0):
        return cls(field, device)

    @classmethod
    def from_assignments(cls, field: int, input_assignment: List[int], aux_assignment: List[int], device: int = 0):
        w = cls(field, device)
        assert input_assignment, "input_assignment must start with ONE"
        w.dev.set(INPUT, 0, input_assignment[0])
        w.extend_inputs(input_assignment[1:])
        w.extend_aux(aux_assignment)
        return w

    def to_assignments(self):
        return self.input_assignment(), self.aux_assignment()

    def close(self):
        self.dev.close()

    def input_assignment(self) -> List[int]:
        return self.dev.witness(INPUT)

    def aux_assignment(self) -> List[int]:
        return self.dev.witness(AUX)

    def alloc(self, _annotation, f) -> Variable:
        return Variable(AUX, self.dev.push_var(AUX, f()))

    def alloc_input(self, _annotation, f) -> Variable:
        return Variable(INPUT, self.dev.push_var(INPUT, f()))

    def enforce(self, _annotation, _a, _b, _c) -> None:
        pass

    def push_namespace(self, _name) -> None:
        pass

    def pop_namespace(self) -> None:
        pass

    @staticmethod
    def is_extensible() -> bool:
        return True

    def extend(self, other: "WitnessCS") -> None:
        self.extend_inputs(other.input_assignment()[1:])  # skip other's ONE (witness_cs.rs:158-161)
        self.extend_aux(other.aux_assignment())

    def is_witness_generator(self) -> bool:
        return True

    def extend_inputs(self, new_inputs) -> None:
        for v in new_inputs:
            self.dev.push_var(INPUT, v)

    def extend_aux(self, new_aux) -> None:
        for v in new_aux:
            self.dev.push_var(AUX, v)

    def allocate_empty(self, aux_n: int, inputs_n: int) -> Tuple[int, int]:
        a0, i0 = self.dev._count[AUX], self.dev._count[INPUT]
        self.extend_aux([0] * aux_n)  # aux first (witness_cs.rs:179-193)
        self.extend_inputs([0] * inputs_n)
        return a0, i0

    def fill_aux(self, first: int, values: List[int]) -> None:
        self.dev.set_range(AUX, first, values)

    def fill_inputs(self, first: int, values: List[int]) -> None:
        self.dev.set_range(INPUT, first, values)

    def inputs_slice(self) -> List[int]:
        return self.input_assignment()

    def aux_slice(self) -> List[int]:
        return self.aux_assignment()

    def scalar_inputs(self) -> List[int]:  # deprecated in the reference (witness_cs.rs:204-211)
        return self.input_assignment()

    def scalar_aux(self) -> List[int]:
        return self.aux_assignment()
