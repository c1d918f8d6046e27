/* bp_fixtures.h -- C entry points of the C++ host front-end (bellpepper_b200/csrc/host/): the B200-backed
 * TestConstraintSystem mirror and the gadget circuits that produce BASELINE configs 1-3.
 *
 * The front-end itself is header-only C++ (cs.hpp, lc.hpp, gadgets.hpp) and is what a C++ user includes; these
 * C wrappers exist so that pytest (ctypes) and bench.py can drive it.  They replace nothing in the reference on
 * their own: the reference-side equivalents are `TestConstraintSystem` (crates/bellpepper-core/src/util_cs/
 * test_cs.rs) and the gadget functions cited at each entry point.
 */
#ifndef BP_FIXTURES_H
#define BP_FIXTURES_H

#include <stdint.h>
#include "bp_r1cs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bp_tcs bp_tcs;

/* device >= 0: rows/witness stream into a fresh bp_cs on that device (evaluation available).
 * device <  0: host recording only (structure export for CPU-side tests and the CPU baseline; no evaluation).
 * named != 0 keeps the reference's path bookkeeping (test_cs.rs:325-375); named == 0 skips every annotation. */
int bp_tcs_new(int field, int device, int named, uint64_t reserve_rows, uint64_t reserve_nnz, uint64_t reserve_vars, bp_tcs** out);
void bp_tcs_free(bp_tcs* t);
const char* bp_tcs_last_error(const bp_tcs* t);
bp_cs* bp_tcs_handle(bp_tcs* t);   /* NULL for host recording */
int bp_tcs_flush(bp_tcs* t);

/* sha256_compression_function over 512 freshly allocated bits ("input bit i") with the IV
 * (crates/bellpepper/src/gadgets/sha256.rs:310-336 test_full_block).  out32 = the 8 output words, big-endian. */
int bp_tcs_sha256_block(bp_tcs* t, const uint8_t block[64], uint8_t out32[32]);

/* sha256() gadget over `len` message bytes, each bit allocated as "input bit <byte> <bit>" most significant
 * first (sha256.rs:50-77, 365-417).  Only rows of compression blocks [block_begin, block_end) are kept (row
 * sharding for multi-GPU; pass 0, UINT64_MAX for all); every witness element is always kept.  *rows_before
 * receives the number of rows that precede the first kept row (the shard's global row base). */
int bp_tcs_sha256(bp_tcs* t, const uint8_t* msg, uint64_t len, uint64_t block_begin, uint64_t block_end, uint8_t digest[32],
                  uint64_t* rows_before);

/* Same circuit, rows kept for several block ranges at once: ranges = n_ranges pairs [begin, end), ascending and disjoint.
 * global_before[i] = rows of the WHOLE circuit that precede range i's first row (a range starting at block 0 also owns the
 * rows of the input bits and gets 0), local_before[i] = kept rows that precede it.  For sampled parity checks of a long
 * chain: one host recording gives the oracle the rows of a few blocks anywhere in the chain, under their global numbers. */
int bp_tcs_sha256_ranges(bp_tcs* t, const uint8_t* msg, uint64_t len, const uint64_t* ranges, uint64_t n_ranges, uint8_t digest[32],
                         uint64_t* global_before, uint64_t* local_before);

/* Witness program (include/bp_r1cs.h: bp_cs_set_witness_program).  With recording on, the next bp_tcs_sha256[_ranges] / bp_tcs_blake2s also
 * records how every aux variable of the circuit follows from the message bits (csrc/host/wtape.hpp) and builds the device
 * program, one unit per compression block; bp_tcs_witness_program lends it out (valid until the next synthesis on `t`).
 * bp_sha256_chain_states: the 8-word hash state BEFORE each of the (len + 9 + 63) / 64 blocks, block 0 = the IV -- the
 * per-unit chaining states bp_cs_generate_witness_async takes (states == NULL: just the block count). */
int bp_tcs_record_witness_program(bp_tcs* t, int on);
int bp_tcs_witness_program(bp_tcs* t, const uint32_t** words, uint64_t* n_words);
int bp_sha256_chain_states(const uint8_t* msg, uint64_t len, uint32_t* states, uint64_t max_blocks, uint64_t* n_blocks);
/* The same for the blake2s gadget (bp_tcs_blake2s with recording on: one unit per compression, message bits least significant
 * first): the 8 words of the BLAKE2s chaining value h before each of the max(1, ceil(len / 64)) compressions. */
int bp_blake2s_chain_states(const uint8_t* msg, uint64_t len, const uint8_t personalization[8], uint32_t* states, uint64_t max_blocks,
                            uint64_t* n_blocks);

/* num gadgets (crates/bellpepper-core/src/gadgets/num.rs), driven the way the reference's tests drive them:
 *  bp_tcs_num_unpack: AllocatedNum::alloc ("num") then to_bits_le (num.rs:263-274) or to_bits_le_strict (:128-247) at the
 *    root -- "bit i/boolean", "unpacking constraint" (one 256-term row over full-width values); bits_out (nullable)
 *    receives the NUM_BITS = 255 bits, little-endian.  KATs: num.rs:696-764.
 *  bp_tcs_num_arith: "a/num", "b/num", mul / square / add (num.rs:276-370: "product num", "squared num", "sum num"),
 *    "nonzero/..." assert_nonzero (:372-401) and "swap/..." conditionally_reverse (:403-455).  KATs: num.rs:591-693.
 *  bp_tcs_num_chain: a product-heavy gadget circuit, x <- x^2 * y + x n times with x unpacked every `unpack_every`
 *    steps: rows whose witness values are full-width field elements (the full-width kernels' case). */
int bp_tcs_num_unpack(bp_tcs* t, const uint64_t value[4], int strict, uint8_t* bits_out);
int bp_tcs_num_arith(bp_tcs* t, const uint64_t a[4], const uint64_t b[4]);
int bp_tcs_num_chain(bp_tcs* t, uint64_t n_steps, uint64_t unpack_every, const uint64_t x0[4], const uint64_t y0[4]);

/* Boolean gadgets the way the reference's tests drive them (boolean.rs:1109-2003): operands "a", "b" (and "c" for op 3, 4), each
 * built in its own namespace like the tests' dyn_construct -- kind 0 = Constant(true), 1 = Constant(false), 2 / 3 = an allocated
 * true / false bit ("a/boolean"), 4 / 5 = the same, negated -- then op 0 = Boolean::xor (boolean.rs:472-491), 1 = and (:494-516),
 * 2 = or (:519-533), 3 = sha256_ch (:536-641), 4 = sha256_maj (:644-759), 5 = enforce_equal (:383-427; two unequal constants:
 * BP_E_STATE, "unsatisfiable constraint system").  *result_kind = 0 Is / 1 Not / 2 Constant, *result_value = its value
 * (both nullable; op 5 reports Constant(false)).
 * bp_tcs_u64_bits: u64_into_boolean_vec_le (boolean.rs:274-304) at the root: "bit i/boolean". */
int bp_tcs_boolean_op(bp_tcs* t, int op, int kind_a, int kind_b, int kind_c, int* result_kind, int* result_value);
int bp_tcs_u64_bits(bp_tcs* t, uint64_t value, uint8_t bits_out[64]);

/* UInt32 gadgets the way the reference's tests drive them (crates/bellpepper/src/gadgets/uint32.rs:492-780): a allocated in
 * "a_bit", b constant; op 0 = xor: c allocated in "c_bit", (a ^ b) in "first xor", ^ c in "second xor" (:492-535); op 1 = addmany:
 * c constant, d allocated in "d_bit", (a ^ b) in "xor", then addmany [r, c, d] in "addition" under a MultiEq (:581-635: flipping
 * "addition/result bit 0/boolean" breaks the MultiEq row); op 2 / 3 = sha256_maj / sha256_ch(a, b, c) with c in "c_bit"
 * (:694-780).  *result = the value of the result word, *n_constant_bits = how many of its bits are constants. */
int bp_tcs_uint32_op(bp_tcs* t, int op, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t* result, uint32_t* n_constant_bits);

/* blake2s() gadget over `len` message bytes, each bit allocated as "input bit <byte> <bit>" least significant first,
 * with an 8-byte personalization (crates/bellpepper/src/gadgets/blake2s.rs:344-406, tests :498-555).  digest = the 32
 * output bytes (the gadget's output bits are little-endian per byte). */
int bp_tcs_blake2s(bp_tcs* t, const uint8_t* msg, uint64_t len, const uint8_t personalization[8], uint8_t digest[32]);

/* Scripted self-test of the C++ `WitnessCS` mirror (csrc/host/cs.hpp; witness_cs.rs:94-201: alloc / alloc_input / no-op
 * enforce / extend / allocate_empty / slices) on a device.  Returns 0, or the number of the first expectation that failed. */
int bp_wcs_selftest(int field, int device);

/* TestConstraintSystem surface (test_cs.rs:239-323).  which_is_unsatisfied: returns the row (>= 0), -1 when
 * satisfied, < -1 on error; `path` (cap bytes) receives the constraint's path when named. */
int64_t bp_tcs_which_is_unsatisfied(bp_tcs* t, char* path, uint64_t cap);
int bp_tcs_set(bp_tcs* t, const char* path, const uint64_t v[4]);
int bp_tcs_get(bp_tcs* t, const char* path, uint64_t v[4]);
uint64_t bp_tcs_num_constraints(const bp_tcs* t);
uint64_t bp_tcs_num_inputs(const bp_tcs* t);
uint64_t bp_tcs_num_aux(const bp_tcs* t);
int bp_tcs_row_path(const bp_tcs* t, uint64_t row, char* path, uint64_t cap);

/* Host recording: borrow the recorded CSR (valid until the next call on `t`). */
int bp_tcs_host_csr(bp_tcs* t, const uint32_t** lens, uint64_t* n_rows, const uint32_t** cols, const uint64_t** coeffs, uint64_t* nnz,
                    const uint64_t** inputs, uint64_t* n_inputs, const uint64_t** aux, uint64_t* n_aux);

#ifdef __cplusplus
}
#endif
#endif
