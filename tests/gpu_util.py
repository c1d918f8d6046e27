"""Helpers shared by the GPU parity tests: raw C-ABI handle wrapper (ctypes + numpy only)."""
import ctypes

import numpy as np

from bellpepper_b200 import ffi


class Handle:
    def __init__(self, field, device=0, reserve=(0, 0, 0)):
        self.L = ffi.load()
        self.h = ffi.vp()
        rc = self.L.bp_cs_new(field, device, *reserve, ctypes.byref(self.h))
        assert rc == 0, f"bp_cs_new -> {rc}"
        self.field = field

    def err(self):
        return (self.L.bp_cs_last_error(self.h) or b"").decode()

    def ok(self, rc):
        assert rc == 0, f"rc={rc}: {self.err()}"

    def close(self):
        if self.h:
            self.L.bp_cs_free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def load_instance(self, lens, cols, coeffs, inputs, aux):
        first = ctypes.c_uint64()
        if inputs.shape[0] > 1:
            self.ok(self.L.bp_cs_alloc(self.h, 0, inputs[1:].ctypes.data, inputs.shape[0] - 1, ctypes.byref(first)))
        if int(inputs[0][0]) != 1 or inputs[0][1:].any():
            self.ok(self.L.bp_cs_set(self.h, 0, 0, inputs[0].ctypes.data))
        if aux.shape[0]:
            self.ok(self.L.bp_cs_alloc(self.h, 1, aux.ctypes.data, aux.shape[0], ctypes.byref(first)))
        self.ok(self.L.bp_cs_enforce(self.h, lens.size // 3, lens.ctypes.data, cols.ctypes.data, coeffs.ctypes.data))

    def first_unsatisfied(self):
        row = ctypes.c_int64()
        self.ok(self.L.bp_cs_first_unsatisfied(self.h, ctypes.byref(row)))
        return row.value

    def eval(self, n_rows):
        az, bz, cz = (np.zeros((n_rows, 4), np.uint64) for _ in range(3))
        self.ok(self.L.bp_cs_eval(self.h, az.ctypes.data, bz.ctypes.data, cz.ctypes.data))
        return az, bz, cz

    def counts(self):
        a, b, c, d = (ctypes.c_uint64() for _ in range(4))
        self.ok(self.L.bp_cs_counts(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return a.value, b.value, c.value, d.value

    def opt(self, key, value=None):
        if value is None:
            v = ctypes.c_int64()
            self.ok(self.L.bp_cs_get_option(self.h, key.encode(), ctypes.byref(v)))
            return v.value
        self.ok(self.L.bp_cs_set_option(self.h, key.encode(), value))
