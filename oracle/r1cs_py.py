"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Pure-Python big-int restatement of the reference's R1CS checker.  Slow, authoritative, small
cases only.  Nothing under `bellpepper_b200/` may import this module.

What it restates (all paths relative to /root/reference):

* `LinearCombination` / `Indexer` / `Variable` / `Index`   crates/bellpepper-core/src/lc.rs:8-46, 74-113, 131-267, 270-375
* `eval_lc`, `which_is_unsatisfied`, `is_satisfied`       crates/bellpepper-core/src/util_cs/test_cs.rs:137-155, 239-264
* `TestConstraintSystem` ingest + accessors + `hash`      crates/bellpepper-core/src/util_cs/test_cs.rs:64-115, 157-334, 363-447
* `WitnessCS`                                             crates/bellpepper/src/util_cs/witness_cs.rs:45-201
* `Namespace` (pop on scope exit)                         crates/bellpepper-core/src/constraint_system.rs:242-333

Field elements are Python ints in [0, p); all arithmetic is `% p`.

Pinning: BLS12-381 Fr results are checked in tests/test_oracle_kat.py against the reference's own
known-answer tests (test_cs.rs:472-510 and the gadget KATs listed in SURVEY.md 8c).  Pallas Fr and
Vesta Fr are **parity unpinned** (the reference has no test and no dependency for them).
"""

from __future__ import annotations

import bisect
import hashlib
import struct
from typing import Callable, Iterable, List, Optional, Tuple

from .fields import Field

INPUT = 0
AUX = 1
AUX_TAG = 1 << 31  # device/ABI column encoding: bit 31 set = aux index space


class SynthesisError(Exception):
    """constraint_system.rs:21-57 -- only the variants that gadgets on this path raise."""


class AssignmentMissing(SynthesisError):
    pass


class DivisionByZero(SynthesisError):
    pass


class Unsatisfiable(SynthesisError):
    pass


class Variable:
    """lc.rs:8-30: a column identifier in one of two index spaces."""

    __slots__ = ("kind", "index")

    def __init__(self, kind: int, index: int):
        self.kind = kind
        self.index = index

    def __eq__(self, other):
        return isinstance(other, Variable) and self.kind == other.kind and self.index == other.index

    def __hash__(self):
        return hash((self.kind, self.index))

    def __repr__(self):
        return f"{'Aux' if self.kind == AUX else 'Input'}({self.index})"

    def tagged(self) -> int:
        return self.index | (AUX_TAG if self.kind == AUX else 0)


ONE = Variable(INPUT, 0)  # constraint_system.rs:73-75


class _SortedTerms:
    """lc.rs:40-129 `Indexer`: index-sorted, unique keys, same-key inserts add coefficients.
    Zero coefficients are RETAINED (only `proc_lc` drops them)."""

    __slots__ = ("keys", "vals")

    def __init__(self):
        self.keys: List[int] = []
        self.vals: List[int] = []

    def copy(self):
        c = _SortedTerms()
        c.keys = self.keys[:]
        c.vals = self.vals[:]
        return c

    def add(self, key: int, coeff: int, p: int):
        i = bisect.bisect_left(self.keys, key)
        if i < len(self.keys) and self.keys[i] == key:
            self.vals[i] = (self.vals[i] + coeff) % p
        else:
            self.keys.insert(i, key)
            self.vals.insert(i, coeff % p)


class LinearCombination:
    """lc.rs:35-38 + the operator algebra at lc.rs:270-375.

    Operators are non-mutating here (`lc + x` returns a new LC) -- Rust moves `self`, so the
    observable result is the same.
    """

    __slots__ = ("field", "inputs", "aux")

    def __init__(self, field: Field):
        self.field = field
        self.inputs = _SortedTerms()
        self.aux = _SortedTerms()

    @classmethod
    def zero(cls, field: Field):
        return cls(field)

    @classmethod
    def from_coeff(cls, field: Field, var: Variable, coeff: int):
        return cls(field)._added(var, coeff)

    @classmethod
    def from_variable(cls, field: Field, var: Variable):
        return cls.from_coeff(field, var, 1)

    def copy(self):
        c = LinearCombination(self.field)
        c.inputs = self.inputs.copy()
        c.aux = self.aux.copy()
        return c

    def _added(self, var: Variable, coeff: int):
        (self.aux if var.kind == AUX else self.inputs).add(var.index, coeff, self.field.p)
        return self

    # lc.rs:155-160: inputs first, then aux
    def iter(self) -> Iterable[Tuple[Variable, int]]:
        for k, v in zip(self.inputs.keys, self.inputs.vals):
            yield Variable(INPUT, k), v
        for k, v in zip(self.aux.keys, self.aux.vals):
            yield Variable(AUX, k), v

    def __len__(self):
        return len(self.inputs.keys) + len(self.aux.keys)

    def is_empty(self):
        return len(self) == 0

    def _combine(self, other, sign: int):
        p = self.field.p
        out = self.copy()
        if isinstance(other, Variable):  # lc.rs:291-309
            return out._added(other, sign % p)
        if isinstance(other, LinearCombination):  # lc.rs:311-337
            for var, c in other.iter():
                out._added(var, (sign * c) % p)
            return out
        if isinstance(other, tuple) and len(other) == 2:
            coeff, what = other
            if isinstance(what, Variable):  # lc.rs:270-289
                return out._added(what, (sign * coeff) % p)
            if isinstance(what, LinearCombination):  # lc.rs:339-375: scale every coefficient
                for var, c in what.iter():
                    out._added(var, (sign * coeff * c) % p)
                return out
        raise TypeError(f"cannot combine LinearCombination with {other!r}")

    def __add__(self, other):
        return self._combine(other, 1)

    def __sub__(self, other):
        return self._combine(other, -1)

    def eval(self, input_assignment: List[int], aux_assignment: List[int]) -> int:
        """lc.rs:245-267.  (Skipping the multiply for coeff == 1 cannot change the value.)"""
        p = self.field.p
        acc = 0
        for k, c in zip(self.inputs.keys, self.inputs.vals):
            acc = (acc + input_assignment[k] * c) % p
        for k, c in zip(self.aux.keys, self.aux.vals):
            acc = (acc + aux_assignment[k] * c) % p
        return acc

    def flat(self) -> Tuple[List[int], List[int]]:
        """(tagged u32 columns, canonical coefficients) in the reference's iteration order."""
        cols, coeffs = [], []
        for var, c in self.iter():
            cols.append(var.tagged())
            coeffs.append(c)
        return cols, coeffs


def compute_path(ns: List[str], this: str) -> str:
    """test_cs.rs:363-375."""
    assert "/" not in this, "'/' is not allowed in names"
    return this if not ns else "/".join(ns) + "/" + this


def _name(annotation) -> str:
    return annotation() if callable(annotation) else annotation


class _NamespaceGuard:
    """constraint_system.rs:242-333: forwards to the root, pops on scope exit."""

    def __init__(self, root):
        self._root = root
        self._open = True

    def __getattr__(self, item):
        return getattr(self._root, item)

    def namespace(self, name):
        return self._root.namespace(name)

    def close(self):
        if self._open:
            self._open = False
            self._root.pop_namespace()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class TestConstraintSystem:
    """test_cs.rs:19-33, 157-447 -- stores named inputs/aux/constraints and checks them."""

    __test__ = False  # not a pytest class

    def __init__(self, field: Field):
        self.field = field
        self.named_objects = {"ONE": ("var", ONE)}
        self.current_namespace: List[str] = []
        self.constraints: List[Tuple[LinearCombination, LinearCombination, LinearCombination, str]] = []
        self.inputs: List[List] = [[1, "ONE"]]  # test_cs.rs:169 -- mutable via set("ONE", ..)
        self.aux: List[List] = []

    @staticmethod
    def one() -> Variable:
        return ONE

    # ---- ingest (test_cs.rs:377-447) -------------------------------------------------------
    def _set_named_obj(self, path: str, obj):
        assert path not in self.named_objects, f"tried to create object at existing path: {path}"
        self.named_objects[path] = obj

    def alloc(self, annotation, f: Callable[[], int]) -> Variable:
        index = len(self.aux)
        path = compute_path(self.current_namespace, _name(annotation))
        value = f() % self.field.p  # a raising closure leaves no variable behind (test_cs.rs:388)
        self.aux.append([value, path])
        var = Variable(AUX, index)
        self._set_named_obj(path, ("var", var))
        return var

    def alloc_input(self, annotation, f: Callable[[], int]) -> Variable:
        index = len(self.inputs)
        path = compute_path(self.current_namespace, _name(annotation))
        value = f() % self.field.p
        self.inputs.append([value, path])
        var = Variable(INPUT, index)
        self._set_named_obj(path, ("var", var))
        return var

    def enforce(self, annotation, a, b, c):
        path = compute_path(self.current_namespace, _name(annotation))
        self._set_named_obj(path, ("constraint", len(self.constraints)))
        z = LinearCombination.zero
        self.constraints.append((a(z(self.field)), b(z(self.field)), c(z(self.field)), path))

    def push_namespace(self, name):
        name = _name(name)
        self._set_named_obj(compute_path(self.current_namespace, name), ("namespace", None))
        self.current_namespace.append(name)

    def pop_namespace(self):
        assert self.current_namespace
        self.current_namespace.pop()

    def namespace(self, name) -> _NamespaceGuard:
        self.push_namespace(name)
        return _NamespaceGuard(self)

    def get_root(self):
        return self

    # ---- the hot path (test_cs.rs:137-155, 239-264) ---------------------------------------
    def _eval_lc(self, lc: LinearCombination) -> int:
        p = self.field.p
        acc = 0
        for var, coeff in lc.iter():
            tmp = (self.aux if var.kind == AUX else self.inputs)[var.index][0]
            acc = (acc + tmp * coeff) % p
        return acc

    def eval_row(self, i: int) -> Tuple[int, int, int]:
        a, b, c, _ = self.constraints[i]
        return self._eval_lc(a), self._eval_lc(b), self._eval_lc(c)

    def first_unsatisfied_row(self) -> int:
        p = self.field.p
        for i in range(len(self.constraints)):
            az, bz, cz = self.eval_row(i)
            if (az * bz) % p != cz:
                return i
        return -1

    def which_is_unsatisfied(self) -> Optional[str]:
        i = self.first_unsatisfied_row()
        return None if i < 0 else self.constraints[i][3]

    def is_satisfied(self) -> bool:
        return self.which_is_unsatisfied() is None

    # ---- accessors (test_cs.rs:175-334) ----------------------------------------------------
    def scalar_inputs(self) -> List[int]:
        return [v for v, _ in self.inputs]

    def scalar_aux(self) -> List[int]:
        return [v for v, _ in self.aux]

    def num_constraints(self) -> int:
        return len(self.constraints)

    def num_inputs(self) -> int:
        return len(self.inputs)

    def _var_at(self, path: str) -> Variable:
        obj = self.named_objects.get(path)
        if obj is None:
            raise KeyError(f"no variable exists at path: {path}")
        if obj[0] != "var":
            raise KeyError(f"path `{path}` holds `{obj[0]}`, not a variable")
        return obj[1]

    def set(self, path: str, to: int):
        v = self._var_at(path)
        (self.aux if v.kind == AUX else self.inputs)[v.index][0] = to % self.field.p

    def get(self, path: str) -> int:
        v = self._var_at(path)
        return (self.aux if v.kind == AUX else self.inputs)[v.index][0]

    def get_input(self, index: int, path: str) -> int:
        value, name = self.inputs[index]
        assert path == name
        return value

    def verify(self, expected: List[int]) -> bool:
        assert len(expected) + 1 == len(self.inputs)
        return all(a[0] == b % self.field.p for a, b in zip(self.inputs[1:], expected))

    def pretty_print_list(self) -> List[str]:
        return (
            [f"INPUT {n}" for _, n in self.inputs]
            + [f"AUX {n}" for _, n in self.aux]
            + [path for *_, path in self.constraints]
        )

    # ---- structure fingerprint (test_cs.rs:64-115, 214-237) --------------------------------
    def hash(self) -> str:
        h = hashlib.blake2s()
        h.update(struct.pack(">QQQ", len(self.inputs), len(self.aux), len(self.constraints)))
        for a, b, c, _ in self.constraints:
            for lc in (a, b, c):
                terms = [(v, co) for v, co in lc.iter() if co != 0]  # proc_lc: merged, zeros dropped
                h.update(struct.pack(">Q", len(terms)))
                for v, co in terms:
                    h.update((b"A" if v.kind == AUX else b"I") + struct.pack(">Q", v.index))
                    h.update(co.to_bytes(32, "big"))
        return h.hexdigest()

    # ---- flat export: what crosses the C ABI ----------------------------------------------
    def to_csr(self):
        """lens[3N], cols[nnz] (tagged), coeffs[nnz] (python ints), inputs, aux."""
        lens, cols, coeffs = [], [], []
        for a, b, c, _ in self.constraints:
            for lc in (a, b, c):
                cc, vv = lc.flat()
                lens.append(len(cc))
                cols.extend(cc)
                coeffs.extend(vv)
        return lens, cols, coeffs, self.scalar_inputs(), self.scalar_aux()


class WitnessCS:
    """witness_cs.rs:45-201: flat witness vectors, `enforce` is a no-op."""

    def __init__(self, field: Field, input_assignment=None, aux_assignment=None):
        self.field = field
        self.input_assignment: List[int] = [1] if input_assignment is None else list(input_assignment)
        self.aux_assignment: List[int] = [] if aux_assignment is None else list(aux_assignment)

    @classmethod
    def from_assignments(cls, field, input_assignment, aux_assignment):
        return cls(field, input_assignment, aux_assignment)

    def to_assignments(self):
        return self.input_assignment, self.aux_assignment

    @staticmethod
    def one() -> Variable:
        return ONE

    def alloc(self, _annotation, f) -> Variable:
        self.aux_assignment.append(f() % self.field.p)
        return Variable(AUX, len(self.aux_assignment) - 1)

    def alloc_input(self, _annotation, f) -> Variable:
        self.input_assignment.append(f() % self.field.p)
        return Variable(INPUT, len(self.input_assignment) - 1)

    def enforce(self, _annotation, _a, _b, _c):
        pass

    def push_namespace(self, _name):
        pass

    def pop_namespace(self):
        pass

    def namespace(self, name):
        return _NamespaceGuard(self)

    def get_root(self):
        return self

    @staticmethod
    def is_extensible() -> bool:
        return True

    def extend(self, other: "WitnessCS"):
        self.input_assignment.extend(other.input_assignment[1:])  # skip other's ONE
        self.aux_assignment.extend(other.aux_assignment)

    def is_witness_generator(self) -> bool:
        return True

    def extend_inputs(self, new_inputs):
        self.input_assignment.extend(x % self.field.p for x in new_inputs)

    def extend_aux(self, new_aux):
        self.aux_assignment.extend(x % self.field.p for x in new_aux)

    def allocate_empty(self, aux_n: int, inputs_n: int):
        """Returns (aux_start, inputs_start): the caller fills `[start:]` -- aux FIRST (witness_cs.rs:179-193)."""
        a0, i0 = len(self.aux_assignment), len(self.input_assignment)
        self.aux_assignment.extend([0] * aux_n)
        self.input_assignment.extend([0] * inputs_n)
        return a0, i0

    def inputs_slice(self):
        return self.input_assignment

    def aux_slice(self):
        return self.aux_assignment
