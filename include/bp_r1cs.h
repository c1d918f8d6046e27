/* bp_r1cs.h -- C ABI of the B200 R1CS evaluation engine (libbp_r1cs.so).
 *
 * This is the drop-in boundary for ONE path of argumentcomputer/bellpepper: batch evaluation of
 * LinearCombination<F> over the witness (three CSR sparse-matrix x vector products over a 255-bit
 * prime field) followed by the element-wise (A.w) o (B.w) == C.w check, i.e. the work of
 * TestConstraintSystem::which_is_unsatisfied / eval_lc / LinearCombination::eval.
 *
 * The reference has no FFI: its boundary is the Rust trait `ConstraintSystem<Scalar>`
 * (crates/bellpepper-core/src/constraint_system.rs:61-237).  A replacement backend is a Rust type that
 * implements the trait, keeps names/namespaces/closures on the host, and forwards FLAT data to these
 * entry points (binding shown in INTEGRATION.md and rust/bellpepper-b200/).  Each entry point cites the
 * reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - Field elements cross the boundary as 4 x uint64_t little-endian limbs of the CANONICAL residue
 *    (what `PrimeField::to_repr()` yields; test_cs.rs:108-111).  Values >= p are rejected (BP_E_RANGE).
 *    Montgomery / pre-scaled forms are device-internal.
 *  - Columns are uint32_t: bit 31 clear = Index::Input(i), bit 31 set = Index::Aux(i)  (lc.rs:27-30).
 *    Input 0 is the constant ONE (constraint_system.rs:73-75) and is an ordinary, mutable witness slot
 *    (test_cs.rs:160-169, 270-275).
 *  - Every function returns BP_OK or a negative error class; nothing throws or aborts across the
 *    boundary.  bp_cs_last_error() gives the message for the handle's last failure.
 *  - A handle is thread-compatible, not thread-safe (`ConstraintSystem: Send`, all mutation through
 *    `&mut self`).  Every call selects the handle's device itself; no reliance on the caller's current
 *    CUDA device.  The caller owns every buffer it passes; the library copies before returning (the *_async and
 *    bp_cs_recheck_* entry points say where a buffer must instead stay valid until the stream has consumed it).
 *  - There is no CPU fallback: without a usable CUDA device bp_cs_new fails with BP_E_CUDA.
 */
#ifndef BP_R1CS_H
#define BP_R1CS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_ABI_VERSION 1

enum {
    BP_OK = 0,
    BP_E_CUDA = -1,  /* CUDA runtime failure (no device, launch error, ...) */
    BP_E_OOM = -2,   /* host or device allocation failed */
    BP_E_RANGE = -3, /* index >= count, value >= p, nnz >= 2^32 */
    BP_E_STATE = -4, /* call not valid in the handle's current state */
    BP_E_ARG = -5    /* null pointer / unknown field / bad option */
};

enum {
    BP_FIELD_BLS12_381_FR = 0, /* blstrs::Scalar -- the only field the reference's tests use */
    BP_FIELD_PALLAS_FR = 1,    /* pasta Fq */
    BP_FIELD_VESTA_FR = 2      /* pasta Fp */
};

#define BP_COL_AUX 0x80000000u

typedef struct bp_cs bp_cs;

/* ---- lifetime --------------------------------------------------------------------------------------
 * TestConstraintSystem::new (test_cs.rs:157-178) / WitnessCS::new (witness_cs.rs:94-101): the new
 * system holds inputs = [ONE], no aux, no constraints.  `reserve_*` are capacity hints (0 = grow on
 * demand) that avoid device reallocation while rows stream in. */
int bp_cs_new(int field, int device, uint64_t reserve_rows, uint64_t reserve_nnz, uint64_t reserve_vars, bp_cs** out);
void bp_cs_free(bp_cs* cs);
const char* bp_cs_last_error(const bp_cs* cs);
int bp_abi_version(void);

/* ---- witness ingest --------------------------------------------------------------------------------
 * alloc / alloc_input (test_cs.rs:380-408, witness_cs.rs:103-123), extend_inputs / extend_aux
 * (witness_cs.rs:171-177) and the fill of allocate_empty (witness_cs.rs:179-193): append `n` values
 * to one index space; *first_index receives the index of the first appended element. */
int bp_cs_alloc(bp_cs* cs, int is_aux, const uint64_t* vals_le, uint64_t n, uint64_t* first_index);

/* Packed form of bp_cs_alloc for witnesses whose values fit one byte: element i = vals[i] (0..255).  Gadget
 * circuits allocate almost nothing but AllocatedBit values (boolean.rs:68-97: `alloc(|| .., || Ok(if b {ONE} else
 * {ZERO}))`), so the Rust wrapper's staging buffer holds one byte per such value instead of 32 and flushes it here;
 * a value that does not fit (rare) is patched afterwards with bp_cs_set.  Same index assignment as bp_cs_alloc. */
int bp_cs_alloc_u8(bp_cs* cs, int is_aux, const uint8_t* vals, uint64_t n, uint64_t* first_index);

/* set (test_cs.rs:270-282): overwrite one element.  get (test_cs.rs:311-323). */
int bp_cs_set(bp_cs* cs, int is_aux, uint64_t idx, const uint64_t v[4]);
int bp_cs_get(bp_cs* cs, int is_aux, uint64_t idx, uint64_t v[4]);

/* Bulk overwrite of an existing range [first, first+n): SizedWitness::generate_witness_into
 * (witness_cs.rs:12, 28-40) writing into the slices returned by allocate_empty.  Up to 4096 elements are validated on
 * the host first (a rejected value leaves the witness untouched); larger batches are validated on the device after the
 * copy, and a batch holding a value >= p is rejected with BP_E_RANGE and leaves the whole range ZERO. */
int bp_cs_set_range(bp_cs* cs, int is_aux, uint64_t first, uint64_t n, const uint64_t* vals_le);

/* Packed form of bp_cs_set_range (see bp_cs_alloc_u8). */
int bp_cs_set_range_u8(bp_cs* cs, int is_aux, uint64_t first, uint64_t n, const uint8_t* vals);

/* scalar_inputs / scalar_aux (test_cs.rs:180-189), inputs_slice / aux_slice (witness_cs.rs:195-201). */
int bp_cs_witness(bp_cs* cs, int is_aux, uint64_t first, uint64_t n, uint64_t* out_le);

/* ---- constraint ingest -----------------------------------------------------------------------------
 * enforce (test_cs.rs:410-427) for a batch of rows.  For row r: lens[3r], lens[3r+1], lens[3r+2] are
 * |A|, |B|, |C|; the rows' terms follow each other in `cols` / `coeffs_le` in A, B, C order, each LC in
 * the reference's iteration order (inputs then aux, ascending index: lc.rs:155-160).  Zero-length LCs
 * and zero coefficients are legal.  Row indices are assigned in call order (test_cs.rs:419). */
int bp_cs_enforce(bp_cs* cs, uint64_t n_rows, const uint32_t* lens, const uint32_t* cols, const uint64_t* coeffs_le);

/* num_inputs / num_constraints (test_cs.rs:266, 295) and sizes. */
int bp_cs_counts(bp_cs* cs, uint64_t* n_inputs, uint64_t* n_aux, uint64_t* n_rows, uint64_t* nnz);

/* TestConstraintSystem::hash (test_cs.rs:64-115, 214-237) from flat rows -- the arrays bp_cs_enforce takes / bp_cs_export
 * writes: BLAKE2s-256 over (n_inputs, n_aux, n_rows) and, per LC, its terms with same-variable coefficients added and zero
 * coefficients dropped (proc_lc), inputs before aux, ascending index; coefficients as big-endian to_repr().  HOST ONLY, no
 * handle.  out_hex receives 64 hex digits + NUL: two front-ends that emit the same matrices get the same string. */
int bp_structure_hash(int field, uint64_t n_inputs, uint64_t n_aux, uint64_t n_rows, const uint32_t* lens, const uint32_t* cols,
                      const uint64_t* coeffs_le, char out_hex[65]);

/* ---- the hot path ----------------------------------------------------------------------------------
 * which_is_unsatisfied / is_satisfied (test_cs.rs:239-264): *row = index of the FIRST row with
 * (A.w)(B.w) != C.w, or -1 when every row holds.  The host side maps the row to its path.
 * A term whose column is out of range makes the call fail with BP_E_RANGE (the reference panics on the
 * slice index, test_cs.rs:146-147). */
int bp_cs_first_unsatisfied(bp_cs* cs, int64_t* row);

/* Re-check with a NEW witness given in the packed form (see bp_cs_alloc_u8): replaces all inputs (inputs_u8 may be NULL to
 * keep them; otherwise n_inputs bytes including ONE) and all aux values (n_aux bytes), then does which_is_unsatisfied.
 * The SizedWitness / WitnessCS flow (witness_cs.rs:7-41): same circuit, next witness.  When aux_u8 is pinned host memory the
 * upload is pipelined with the check: rows are checked as soon as the variables they read have arrived. */
int bp_cs_recheck_u8(bp_cs* cs, const uint8_t* inputs_u8, const uint8_t* aux_u8, int64_t* row);
/* Same, leaving the first failing GLOBAL row in DEVICE memory like bp_cs_check_async (row-sharded multi-GPU use: every rank
 * uploads the new witness, then one min-all-reduce).  The host buffers must stay valid until the stream has consumed them. */
int bp_cs_recheck_u8_async(bp_cs* cs, const uint8_t* inputs_u8, const uint8_t* aux_u8, int64_t* dev_result);

/* The densest form, for witnesses made of bits only (Boolean / AllocatedBit values: what sha256, blake2s, uint32 circuits
 * allocate): value i = bit i of the byte string, least significant bit of a byte first (the gadgets' own little-endian bit
 * order, boolean.rs:307-366 / blake2s.rs:520).  A value that is not 0 or 1 is patched afterwards with bp_cs_set. */
int bp_cs_set_range_bits(bp_cs* cs, int is_aux, uint64_t first, uint64_t n, const uint8_t* bits);
int bp_cs_recheck_bits(bp_cs* cs, const uint8_t* inputs_bits, const uint8_t* aux_bits, int64_t* row);
int bp_cs_recheck_bits_async(bp_cs* cs, const uint8_t* inputs_bits, const uint8_t* aux_bits, int64_t* dev_result);

/* Batched `set` (test_cs.rs:270-282; the 255 x 2 x 200 set-and-recheck loop of num.rs:753-762): element idx[i] of one index
 * space becomes vals[4i..4i+4) for i < n, by one copy and one scatter kernel.  Indices and values are validated first: a
 * rejected batch (BP_E_RANGE) changes nothing.  Later entries win when an index repeats. */
int bp_cs_set_many(bp_cs* cs, int is_aux, uint64_t n, const uint64_t* idx, const uint64_t* vals_le);

/* Same circuit, next witness, in the REFERENCE's own format: 32-byte canonical scalars, n_inputs (incl. ONE; NULL keeps the
 * inputs) and n_aux of them -- what WitnessCS holds (witness_cs.rs:45-57: input_assignment / aux_assignment).  The library
 * packs on the host before it sends (every 0/1 value becomes one bit, anything else an (index, value) exception applied by a
 * scatter kernel), so a gadget witness of 10^8 bits costs 14 MB of PCIe traffic instead of 3.5 GB; a witness that is not
 * mostly bits is sent as it is.  The packing pass uses every hardware thread; several processes on one host (one per GPU)
 * should share them: environment variable BP_PACK_THREADS = threads per process.  Then which_is_unsatisfied.  The _async form
 * leaves the first failing GLOBAL row in DEVICE memory (INT64_MAX = satisfied); the scalars are consumed before either form
 * returns. */
int bp_cs_recheck_scalars(bp_cs* cs, const uint64_t* inputs_le, const uint64_t* aux_le, int64_t* row);
int bp_cs_recheck_scalars_async(bp_cs* cs, const uint64_t* inputs_le, const uint64_t* aux_le, int64_t* dev_result);

/* The packing pass of bp_cs_recheck_scalars alone: HOST ONLY, no handle and no device -- for a caller that packs once and
 * re-checks many times, or overlaps the pass with its own work, and then calls bp_cs_recheck_bits (+ bp_cs_set_many for the
 * exceptions).  n scalars (32-byte canonical little-endian, witness_cs.rs:45-57) -> bits[(n + 7) / 8]: element i is bit
 * (i & 7) of bits[i >> 3] when its value is 0 or 1; any other element leaves a 0 bit and is reported as exception
 * (exc_idx[k], exc_vals_le[4k .. 4k + 4)), in ascending index order.  *n_exc = the number found; when it exceeds exc_cap only
 * the first exc_cap are stored and BP_E_RANGE is returned (call again with more room).  Values are NOT compared with p here
 * (bp_cs_set_many does that).  Threads as above (BP_PACK_THREADS); one sequential read of the scalars, AVX2 when the CPU has
 * it (bp_pack_kernel: "avx2" / "portable"; BP_PACK_SIMD=0 forces the portable loop). */
int bp_pack_scalars(const uint64_t* scalars_le, uint64_t n, uint8_t* bits, uint64_t* exc_idx, uint64_t* exc_vals_le, uint64_t exc_cap,
                    uint64_t* n_exc);
const char* bp_pack_kernel(void);

/* The same two calls for scalars AS THEY SIT IN MEMORY in the reference's Vec<Scalar> (WitnessCS::input_assignment /
 * aux_assignment, witness_cs.rs:45-57): blstrs::Scalar (blstrs 0.7, Cargo.toml:10) and pasta_curves::{Fp, Fq} keep 4 x u64
 * little-endian limbs of the MONTGOMERY form x * 2^256 mod p, so the binding can hand over `vec.as_ptr()` without a to_repr()
 * pass over 10^8 elements: the packing pass looks for the limb patterns of 0 and of 2^256 mod p, and converts the (few) other
 * elements on the host.  An element >= p is not a Scalar: BP_E_RANGE.  A witness that is not mostly bits is converted on the
 * host threads (bp_scalars_from_mont) and sent like bp_cs_set_range.  bp_pack_scalars_mont reports its exceptions in
 * CANONICAL form (ready for bp_cs_set_many). */
int bp_cs_recheck_scalars_mont(bp_cs* cs, const uint64_t* inputs_mont, const uint64_t* aux_mont, int64_t* row);
int bp_cs_recheck_scalars_mont_async(bp_cs* cs, const uint64_t* inputs_mont, const uint64_t* aux_mont, int64_t* dev_result);
int bp_pack_scalars_mont(int field, const uint64_t* scalars_mont, uint64_t n, uint8_t* bits, uint64_t* exc_idx, uint64_t* exc_vals_le,
                         uint64_t exc_cap, uint64_t* n_exc);
/* Host only: n Montgomery-form scalars -> canonical little-endian limbs (x * 2^-256 mod p), all host threads. */
int bp_scalars_from_mont(int field, const uint64_t* scalars_mont, uint64_t n, uint64_t* scalars_le);

/* ---- witness generation on the device (the "witness evaluation" half: SizedWitness::generate_witness_into, witness_cs.rs:7-41)
 * For a bit-logic gadget circuit (boolean / uint32 / sha256: every aux value is a bit that follows from earlier bits) the NEXT
 * witness of an already synthesized circuit need not be produced by re-running the gadgets' host closures
 * (sha256.rs:83-272 through boolean.rs:68-272, 536-759 and uint32.rs:306-406) and uploading n_aux values: the front-end
 * records, once, HOW each aux variable follows from earlier ones (csrc/host/wtape.hpp; bp_tcs_witness_program in
 * bp_fixtures.h) and the device replays that program from the message bytes.
 *
 * Program = uint32 words: a 16-word header {magic "BPWP", version 1, n_units, n_tapes, aux index of message bit 0,
 * n_msg_bits, bit order (1 = most significant bit of a byte first), n_aux, max variables / max sums of a tape, offsets},
 * a unit table {tape, first aux index, first message bit, chaining-state index} -- a unit is e.g. one compression block --
 * a tape table, and per tape its dependency levels, entries {result | op << 28, a, b, c} (XOR, AND, AND_NOT, NOR, CH, MAJ of
 * operands that are constants, variables of the unit, message bits or chaining-state bits, possibly negated; or bit j of an
 * integer sum of weighted bits, uint32.rs:306-406) and sums.  bp_cs_set_witness_program validates every index (BP_E_ARG).
 *
 * bp_cs_generate_witness_async: msg = the message bytes (msg_len * 8 must equal the program's n_msg_bits), states = 8 words of
 * chaining state per unit (for a sha256 chain: the hash state BEFORE each block; bp_sha256_chain_states computes them on the
 * host in microseconds; for a blake2s circuit: the BLAKE2s chaining value before each compression, bp_blake2s_chain_states).  Writes every aux value (inputs are untouched); follow with bp_cs_first_unsatisfied / check_async.
 * 64 bytes of H2D per block instead of 26 000 witness values. */
int bp_cs_set_witness_program(bp_cs* cs, const uint32_t* words, uint64_t n_words);
int bp_cs_generate_witness_async(bp_cs* cs, const uint8_t* msg, uint64_t msg_len, const uint32_t* states, uint64_t n_state_words);

/* Same check, asynchronous: enqueue on the handle's stream and leave the result in DEVICE memory as one
 * int64 (first failing GLOBAL row = row_base + local row; INT64_MAX when satisfied) so that a row-sharded
 * multi-GPU caller can min-all-reduce it without a host round trip.  No host synchronisation. */
int bp_cs_check_async(bp_cs* cs, int64_t* dev_result);

/* Batched LinearCombination::eval over all rows (lc.rs:245-267): canonical A.w, B.w, C.w, n_rows x 4
 * limbs each, into HOST buffers (any may be NULL). */
int bp_cs_eval(bp_cs* cs, uint64_t* az, uint64_t* bz, uint64_t* cz);
/* Same into DEVICE buffers, asynchronous on the handle's stream. */
int bp_cs_eval_async(bp_cs* cs, uint64_t* dev_az, uint64_t* dev_bz, uint64_t* dev_cz);

/* LinearCombination::eval for one ad-hoc LC (lc.rs:245-267). */
int bp_cs_eval_lc(bp_cs* cs, const uint32_t* cols, const uint64_t* coeffs_le, uint32_t n_terms, uint64_t out[4]);

/* ---- checkpoint / resume ---------------------------------------------------------------------------
 * Not a reference interface (LinearCombination is not serialisable there, lc.rs:34): write an ingested system -- matrices in
 * their device-internal form, canonical witness, row base -- to a file, and create a handle from such a file, so that a
 * 10^8-row synthesis is paid once.  The header records its own size, the writer's byte order and the counts; the file
 * ends with a 64-bit checksum of the payload.  bp_cs_load returns BP_E_ARG for a file that is not one of ours (magic /
 * layout / ABI version / byte order / counts out of range), BP_E_STATE for an unreadable, truncated or corrupted one
 * (checksum mismatch, row offsets that are not a non-decreasing sequence from 0 to nnz), BP_E_RANGE when its witness is
 * not canonical. */
int bp_cs_save(bp_cs* cs, const char* path);
int bp_cs_load(const char* path, int device, bp_cs** out);

/* ---- prover hand-off ------------------------------------------------------------------------------------
 * Not a reference interface either (the reference has no on-disk R1CS).  What the step AFTER the check needs -- a Nova /
 * Spartan style shape (A, B, C) + witness, optionally with A.w, B.w, C.w -- in a documented, implementation-independent
 * layout: canonical little-endian coefficients, no class bits, no internal scaling; exactly the arrays bp_cs_enforce /
 * bp_cs_alloc take, so another handle (or the CPU oracle) ingests the file as it is.
 *   char magic[8] = "BPR1CSX\1"; u32 version = 1; u32 field; u64 n_rows, n_inputs, n_aux, nnz, row_base;
 *   u32 flags (bit 0: products follow); u32 reserved;
 *   u32 lens[3 n_rows]  (|A_i|, |B_i|, |C_i|);  u32 cols[nnz]  (bit 31 = aux);  u64 coeffs[nnz][4];
 *   u64 inputs[n_inputs][4];  u64 aux[n_aux][4];  [u64 az[n_rows][4], bz[..], cz[..]];  u64 checksum (as bp_cs_save). */
int bp_cs_export(bp_cs* cs, const char* path, int with_products);

/* ---- execution control -----------------------------------------------------------------------------*/
/* Use an existing CUDA stream (cudaStream_t as void*) for all work.  NULL = the handle's own non-blocking
 * stream; the legacy default stream is cudaStreamLegacy, i.e. (void*)0x1. */
int bp_cs_set_stream(bp_cs* cs, void* cuda_stream);
/* Row-sharded use: global index of this handle's row 0 (default 0). */
int bp_cs_set_row_base(bp_cs* cs, uint64_t row_base);
/* Block until all enqueued work of this handle is complete. */
int bp_cs_sync(bp_cs* cs);
/* "sparse_upload" (0/1): bp_cs_recheck_u8[_async] uploads only the 2^16-element chunks of the aux witness that some row of
 * THIS handle reads -- for row shards, whose rows read a fraction of the witness; elements no row of the handle reads keep
 * their previous values.
 * Tuning / introspection knobs: "fat_terms" (rows with more terms use the warp-per-row kernel; default 96);
 * "variant" (-1 = default; else bit 0: no small-operand kernel, bit 1: no witness shadows in the warp-per-row kernel,
 * bit 2: park A.w/B.w in shared memory, bit 3: no integer pass over the fat rows -- every variant returns the same
 * results, the parity tests run them all); "graph" (default 1: a check is replayed as one captured CUDA graph while its
 * arguments are unchanged; read-only "graph_replays" / "graph_captures");
 * read-only: "launches" (kernels launched so far), "plain_rows" / "generic_rows" / "fat_rows" (rows per kernel),
 * "deferred_rows" / "fat_undecided_rows" (plain / fat rows the last check sent on to the full-width kernels), "gen_terms",
 * "sm_count". */
int bp_cs_set_option(bp_cs* cs, const char* key, int64_t value);
int bp_cs_get_option(bp_cs* cs, const char* key, int64_t* value);

/* ---- multi-GPU: row shards on several GPUs of one node (one process per GPU) -------------------------------
 * Not a reference interface (the reference is single-threaded); the hooks it would sit behind are `extend`
 * (constraint_system.rs:138-148: sub-systems synthesized independently, concatenated in order) and the bulk witness
 * ingress of `WitnessCS` (witness_cs.rs:171-193).  Rows are independent (test_cs.rs:240-250), so every rank holds a
 * contiguous row range (bp_cs_set_row_base) and the WHOLE witness; the only exchange of a check is one MIN over the
 * ranks' first-unsatisfied GLOBAL rows.  Inside the library that MIN does not launch a collective: the kernels leave the
 * word in device memory and a one-warp kernel publishes it with system-scope stores into every peer's mailbox
 * (peer memory mapped with CUDA IPC over NVLink) and reads the other ranks' words from its own; check + exchange replay as
 * ONE captured CUDA graph.  NCCL (bound at run time) is used to create the group, to move witnesses, and as the
 * fallback transport when peer mapping is unavailable.
 *
 * All bp_group_* calls are collective: every rank of the group makes the same calls in the same order. */
typedef struct bp_group bp_group;
#define BP_GROUP_ID_BYTES 128
/* Rank 0 creates the id (an ncclUniqueId) and hands it to the other ranks by any channel it likes. */
int bp_group_unique_id(uint8_t id[BP_GROUP_ID_BYTES]);
/* Join `cs` (this rank's row shard) to the group.  world == 1 is valid (id may be NULL) and needs no NCCL. */
int bp_group_init(bp_cs* cs, const uint8_t id[BP_GROUP_ID_BYTES], int rank, int world, bp_group** out);
void bp_group_free(bp_group* g); /* collective too: a rank's mailbox is freed only after every peer has unmapped it */
/* *transport: 0 = single rank, 1 = NCCL all-reduce, 2 = peer-memory mailboxes. */
int bp_group_info(bp_group* g, int* rank, int* world, int* transport);
/* which_is_unsatisfied over ALL shards (test_cs.rs:239-253): *row = the first unsatisfied GLOBAL row or -1, the same
 * value on every rank. */
int bp_group_check(bp_group* g, int64_t* row);
/* Same, asynchronous: the reduced word (INT64_MAX = satisfied) is left in DEVICE memory on the handle's stream. */
int bp_group_check_async(bp_group* g, int64_t* dev_result);
/* Just the exchange: MIN over the ranks of the int64 at dev_result, in place (after any *_async producer). */
int bp_group_reduce_async(bp_group* g, int64_t* dev_result);
/* "Witness broadcast once": the witness held by `root` (inputs, aux) replaces every rank's, over NVLink. */
int bp_group_broadcast_witness(bp_group* g, int root);
/* New witness from host memory with every PCIe link carrying 1/world of it: rank r passes the elements
 * [r*cnt/world, (r+1)*cnt/world) of the index space (`slice` = the first of them; cnt = current size of the space); the
 * slices are all-gathered over NVLink and validated everywhere.  A value >= p: BP_E_RANGE, the index space is zeroed. */
int bp_group_set_witness_sharded(bp_group* g, int is_aux, const uint64_t* slice_le);
/* Contiguous row shards balanced by TERMS, not rows (gadget circuits are skewed by their MultiEq rows): lens as for
 * bp_cs_enforce; bounds[r] = first row of rank r, bounds[world] = n_rows.  Host-side helper, no device needed. */
int bp_split_rows_by_nnz(const uint32_t* lens, uint64_t n_rows, int world, uint64_t* bounds);

/* ---- measurement fixture: synthetic instances generated in HBM ----------------------------------------
 * Not a reference interface.  Appends rows [row0, row0+n_rows) of the counter-based synthetic R1CS
 * (recipe: DESIGN.md "Synthetic instances"; CPU twin: oracle/bp_oracle.c bpo_synth_*) and, with
 * bp_cs_synth_witness, REPLACES the witness by the recipe's n_vars elements (n_inputs of them inputs). */
int bp_cs_synth_rows(bp_cs* cs, uint64_t seed, uint32_t t, uint64_t n_vars, uint64_t n_inputs, uint64_t row0, uint64_t n_rows);
int bp_cs_synth_witness(bp_cs* cs, uint64_t seed, uint64_t n_vars, uint64_t n_inputs);

#ifdef __cplusplus
}
#endif
#endif /* BP_R1CS_H */
