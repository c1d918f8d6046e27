//! `extern "C"` twin of include/bp_r1cs.h (keep in sync with BP_ABI_VERSION = 1).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct bp_cs {
    _private: [u8; 0],
}

/// A group of row shards on several GPUs of one node (one process per GPU).
#[repr(C)]
pub struct bp_group {
    _private: [u8; 0],
}
pub const BP_GROUP_ID_BYTES: usize = 128;

pub const BP_OK: c_int = 0;
pub const BP_E_CUDA: c_int = -1;
pub const BP_E_OOM: c_int = -2;
pub const BP_E_RANGE: c_int = -3;
pub const BP_E_STATE: c_int = -4;
pub const BP_E_ARG: c_int = -5;

pub const BP_FIELD_BLS12_381_FR: c_int = 0;
pub const BP_FIELD_PALLAS_FR: c_int = 1;
pub const BP_FIELD_VESTA_FR: c_int = 2;
pub const BP_COL_AUX: u32 = 0x8000_0000;

extern "C" {
    pub fn bp_abi_version() -> c_int;
    pub fn bp_cs_new(field: c_int, device: c_int, reserve_rows: u64, reserve_nnz: u64, reserve_vars: u64, out: *mut *mut bp_cs) -> c_int;
    pub fn bp_cs_free(cs: *mut bp_cs);
    pub fn bp_cs_last_error(cs: *const bp_cs) -> *const c_char;
    pub fn bp_cs_alloc(cs: *mut bp_cs, is_aux: c_int, vals_le: *const u64, n: u64, first_index: *mut u64) -> c_int;
    pub fn bp_cs_alloc_u8(cs: *mut bp_cs, is_aux: c_int, vals: *const u8, n: u64, first_index: *mut u64) -> c_int;
    pub fn bp_cs_set_range_u8(cs: *mut bp_cs, is_aux: c_int, first: u64, n: u64, vals: *const u8) -> c_int;
    pub fn bp_cs_set(cs: *mut bp_cs, is_aux: c_int, idx: u64, v: *const u64) -> c_int;
    pub fn bp_cs_get(cs: *mut bp_cs, is_aux: c_int, idx: u64, v: *mut u64) -> c_int;
    pub fn bp_cs_set_range(cs: *mut bp_cs, is_aux: c_int, first: u64, n: u64, vals_le: *const u64) -> c_int;
    pub fn bp_cs_witness(cs: *mut bp_cs, is_aux: c_int, first: u64, n: u64, out_le: *mut u64) -> c_int;
    pub fn bp_cs_enforce(cs: *mut bp_cs, n_rows: u64, lens: *const u32, cols: *const u32, coeffs_le: *const u64) -> c_int;
    pub fn bp_cs_counts(cs: *mut bp_cs, n_inputs: *mut u64, n_aux: *mut u64, n_rows: *mut u64, nnz: *mut u64) -> c_int;
    pub fn bp_cs_first_unsatisfied(cs: *mut bp_cs, row: *mut i64) -> c_int;
    pub fn bp_cs_recheck_u8(cs: *mut bp_cs, inputs_u8: *const u8, aux_u8: *const u8, row: *mut i64) -> c_int;
    pub fn bp_cs_recheck_u8_async(cs: *mut bp_cs, inputs_u8: *const u8, aux_u8: *const u8, dev_result: *mut i64) -> c_int;
    pub fn bp_cs_set_range_bits(cs: *mut bp_cs, is_aux: c_int, first: u64, n: u64, bits: *const u8) -> c_int;
    pub fn bp_cs_recheck_bits(cs: *mut bp_cs, inputs_bits: *const u8, aux_bits: *const u8, row: *mut i64) -> c_int;
    pub fn bp_cs_recheck_bits_async(cs: *mut bp_cs, inputs_bits: *const u8, aux_bits: *const u8, dev_result: *mut i64) -> c_int;
    pub fn bp_cs_set_many(cs: *mut bp_cs, is_aux: c_int, n: u64, idx: *const u64, vals_le: *const u64) -> c_int;
    pub fn bp_cs_recheck_scalars(cs: *mut bp_cs, inputs_le: *const u64, aux_le: *const u64, row: *mut i64) -> c_int;
    pub fn bp_cs_recheck_scalars_async(cs: *mut bp_cs, inputs_le: *const u64, aux_le: *const u64, dev_result: *mut i64) -> c_int;
    pub fn bp_structure_hash(field: c_int, n_inputs: u64, n_aux: u64, n_rows: u64, lens: *const u32, cols: *const u32, coeffs_le: *const u64, out_hex: *mut c_char) -> c_int;
    pub fn bp_pack_scalars(scalars_le: *const u64, n: u64, bits: *mut u8, exc_idx: *mut u64, exc_vals_le: *mut u64, exc_cap: u64, n_exc: *mut u64) -> c_int;
    pub fn bp_pack_kernel() -> *const c_char;
    pub fn bp_cs_recheck_scalars_mont(cs: *mut bp_cs, inputs_mont: *const u64, aux_mont: *const u64, row: *mut i64) -> c_int;
    pub fn bp_cs_recheck_scalars_mont_async(cs: *mut bp_cs, inputs_mont: *const u64, aux_mont: *const u64, dev_result: *mut i64) -> c_int;
    pub fn bp_pack_scalars_mont(field: c_int, scalars_mont: *const u64, n: u64, bits: *mut u8, exc_idx: *mut u64, exc_vals_le: *mut u64, exc_cap: u64, n_exc: *mut u64) -> c_int;
    pub fn bp_scalars_from_mont(field: c_int, scalars_mont: *const u64, n: u64, scalars_le: *mut u64) -> c_int;
    pub fn bp_cs_check_async(cs: *mut bp_cs, dev_result: *mut i64) -> c_int;
    pub fn bp_cs_eval(cs: *mut bp_cs, az: *mut u64, bz: *mut u64, cz: *mut u64) -> c_int;
    pub fn bp_cs_eval_async(cs: *mut bp_cs, dev_az: *mut u64, dev_bz: *mut u64, dev_cz: *mut u64) -> c_int;
    pub fn bp_cs_eval_lc(cs: *mut bp_cs, cols: *const u32, coeffs_le: *const u64, n_terms: u32, out: *mut u64) -> c_int;
    pub fn bp_cs_save(cs: *mut bp_cs, path: *const c_char) -> c_int;
    pub fn bp_cs_load(path: *const c_char, device: c_int, out: *mut *mut bp_cs) -> c_int;
    pub fn bp_cs_export(cs: *mut bp_cs, path: *const c_char, with_products: c_int) -> c_int;
    pub fn bp_cs_set_stream(cs: *mut bp_cs, cuda_stream: *mut c_void) -> c_int;
    pub fn bp_cs_set_row_base(cs: *mut bp_cs, row_base: u64) -> c_int;
    pub fn bp_cs_sync(cs: *mut bp_cs) -> c_int;
    pub fn bp_cs_set_option(cs: *mut bp_cs, key: *const c_char, value: i64) -> c_int;
    pub fn bp_cs_get_option(cs: *mut bp_cs, key: *const c_char, value: *mut i64) -> c_int;
    pub fn bp_cs_set_witness_program(cs: *mut bp_cs, words: *const u32, n_words: u64) -> c_int;
    pub fn bp_cs_generate_witness_async(cs: *mut bp_cs, msg: *const u8, msg_len: u64, states: *const u32, n_state_words: u64) -> c_int;
    pub fn bp_group_unique_id(id: *mut u8) -> c_int;
    pub fn bp_group_init(cs: *mut bp_cs, id: *const u8, rank: c_int, world: c_int, out: *mut *mut bp_group) -> c_int;
    pub fn bp_group_free(g: *mut bp_group);
    pub fn bp_group_info(g: *mut bp_group, rank: *mut c_int, world: *mut c_int, transport: *mut c_int) -> c_int;
    pub fn bp_group_check(g: *mut bp_group, row: *mut i64) -> c_int;
    pub fn bp_group_check_async(g: *mut bp_group, dev_result: *mut i64) -> c_int;
    pub fn bp_group_reduce_async(g: *mut bp_group, dev_result: *mut i64) -> c_int;
    pub fn bp_group_broadcast_witness(g: *mut bp_group, root: c_int) -> c_int;
    pub fn bp_group_set_witness_sharded(g: *mut bp_group, is_aux: c_int, slice_le: *const u64) -> c_int;
    pub fn bp_split_rows_by_nnz(lens: *const u32, n_rows: u64, world: c_int, bounds: *mut u64) -> c_int;
    pub fn bp_cs_synth_rows(cs: *mut bp_cs, seed: u64, t: u32, n_vars: u64, n_inputs: u64, row0: u64, n_rows: u64) -> c_int;
    pub fn bp_cs_synth_witness(cs: *mut bp_cs, seed: u64, n_vars: u64, n_inputs: u64) -> c_int;
}
