"""ctypes binding of include/bp_r1cs.h -- the same stub a reference-side maintainer would write
(INTEGRATION.md shows the Rust `extern "C"` twin).  No torch types cross this boundary."""

from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbp_r1cs.so")

BP_OK, BP_E_CUDA, BP_E_OOM, BP_E_RANGE, BP_E_STATE, BP_E_ARG = 0, -1, -2, -3, -4, -5
FIELD_BLS12_381_FR, FIELD_PALLAS_FR, FIELD_VESTA_FR = 0, 1, 2
COL_AUX = 0x80000000

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)
i64p = ctypes.POINTER(ctypes.c_int64)
vp = ctypes.c_void_p

# name -> (restype, argtypes): every symbol include/bp_r1cs.h declares
SIGNATURES = {
    "bp_abi_version": (ctypes.c_int, []),
    "bp_cs_new": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(vp)]),
    "bp_cs_free": (None, [vp]),
    "bp_cs_last_error": (ctypes.c_char_p, [vp]),
    "bp_cs_alloc": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_uint64, u64p]),
    "bp_cs_alloc_u8": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_uint64, u64p]),
    "bp_cs_set_range_u8": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp]),
    "bp_cs_set": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, vp]),
    "bp_cs_get": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, vp]),
    "bp_cs_set_range": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp]),
    "bp_cs_witness": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp]),
    "bp_cs_enforce": (ctypes.c_int, [vp, ctypes.c_uint64, vp, vp, vp]),
    "bp_cs_counts": (ctypes.c_int, [vp, u64p, u64p, u64p, u64p]),
    "bp_cs_first_unsatisfied": (ctypes.c_int, [vp, i64p]),
    "bp_cs_recheck_u8": (ctypes.c_int, [vp, vp, vp, i64p]),
    "bp_cs_recheck_u8_async": (ctypes.c_int, [vp, vp, vp, vp]),
    "bp_cs_set_range_bits": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp]),
    "bp_cs_recheck_bits": (ctypes.c_int, [vp, vp, vp, i64p]),
    "bp_cs_recheck_bits_async": (ctypes.c_int, [vp, vp, vp, vp]),
    "bp_cs_check_async": (ctypes.c_int, [vp, vp]),
    "bp_cs_eval": (ctypes.c_int, [vp, vp, vp, vp]),
    "bp_cs_eval_async": (ctypes.c_int, [vp, vp, vp, vp]),
    "bp_cs_eval_lc": (ctypes.c_int, [vp, vp, vp, ctypes.c_uint32, vp]),
    "bp_cs_save": (ctypes.c_int, [vp, ctypes.c_char_p]),
    "bp_cs_load": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(vp)]),
    "bp_cs_export": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_int]),
    "bp_cs_set_stream": (ctypes.c_int, [vp, vp]),
    "bp_cs_set_row_base": (ctypes.c_int, [vp, ctypes.c_uint64]),
    "bp_cs_sync": (ctypes.c_int, [vp]),
    "bp_cs_set_option": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_int64]),
    "bp_cs_get_option": (ctypes.c_int, [vp, ctypes.c_char_p, i64p]),
    "bp_cs_set_many": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_uint64, vp, vp]),
    "bp_cs_recheck_scalars": (ctypes.c_int, [vp, vp, vp, i64p]),
    "bp_cs_recheck_scalars_async": (ctypes.c_int, [vp, vp, vp, vp]),
    "bp_structure_hash": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, vp, vp, vp, ctypes.c_char_p]),
    "bp_pack_scalars": (ctypes.c_int, [vp, ctypes.c_uint64, vp, vp, vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]),
    "bp_pack_kernel": (ctypes.c_char_p, []),
    "bp_cs_recheck_scalars_mont": (ctypes.c_int, [vp, vp, vp, i64p]),
    "bp_cs_recheck_scalars_mont_async": (ctypes.c_int, [vp, vp, vp, vp]),
    "bp_pack_scalars_mont": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_uint64, vp, vp, vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]),
    "bp_scalars_from_mont": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_uint64, vp]),
    "bp_cs_set_witness_program": (ctypes.c_int, [vp, vp, ctypes.c_uint64]),
    "bp_cs_generate_witness_async": (ctypes.c_int, [vp, vp, ctypes.c_uint64, vp, ctypes.c_uint64]),
    "bp_group_unique_id": (ctypes.c_int, [vp]),
    "bp_group_init": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]),
    "bp_group_free": (None, [vp]),
    "bp_group_info": (ctypes.c_int, [vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "bp_group_check": (ctypes.c_int, [vp, i64p]),
    "bp_group_check_async": (ctypes.c_int, [vp, vp]),
    "bp_group_reduce_async": (ctypes.c_int, [vp, vp]),
    "bp_group_broadcast_witness": (ctypes.c_int, [vp, ctypes.c_int]),
    "bp_group_set_witness_sharded": (ctypes.c_int, [vp, ctypes.c_int, vp]),
    "bp_split_rows_by_nnz": (ctypes.c_int, [vp, ctypes.c_uint64, ctypes.c_int, u64p]),
    "bp_cs_synth_rows": (ctypes.c_int, [vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]),
    "bp_cs_synth_witness": (ctypes.c_int, [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libbp_r1cs.so.  There is no fallback: a missing or incomplete library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found -- run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
            "bellpepper_b200 has no CPU path."
        )
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(L, name)
        except AttributeError as e:
            raise NativeLibraryMissing(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
