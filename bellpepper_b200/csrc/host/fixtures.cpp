// C wrappers over the C++ host front-end (see include/bp_fixtures.h).
#include "../../../include/bp_fixtures.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#include "blake2s_host.hpp"
#include "gadgets.hpp"

using namespace bph;

namespace {

// Forwards witness always, rows only while enabled (row sharding by compression block).
struct FilterSink : Sink {
    Sink* inner;
    bool rows_enabled = true;
    uint64_t rows_dropped = 0;
    explicit FilterSink(Sink* s) : inner(s) {}
    void alloc(int is_aux, const uint64_t* v, uint64_t n) override { inner->alloc(is_aux, v, n); }
    void alloc_packed(int is_aux, const uint8_t* b, uint64_t n, const uint64_t* wp, const uint64_t* wv, uint64_t nw) override {
        inner->alloc_packed(is_aux, b, n, wp, wv, nw);
    }
    void enforce(uint64_t n_rows, const uint32_t* l, const uint32_t* c, const uint64_t* v, uint64_t nnz) override {
        if (rows_enabled) inner->enforce(n_rows, l, c, v, nnz);
        else rows_dropped += n_rows;
    }
    bp_cs* handle() override { return inner->handle(); }
};

}  // namespace

struct bp_tcs {
    int field;
    bool named;
    bp_cs* h = nullptr;
    std::unique_ptr<Sink> base;
    std::unique_ptr<FilterSink> filter;
    std::unique_ptr<TestConstraintSystem> named_cs;
    std::unique_ptr<BulkConstraintSystem> bulk_cs;
    std::string err;
    bool record_tape = false;        // bp_tcs_record_witness_program: the next sha256 synthesis also records its witness tape
    std::vector<uint32_t> wprog;     // the program built from it (wtape.hpp)
};

namespace {

template <class F> int guarded(bp_tcs* t, F&& f) {
    try {
        f();
        return BP_OK;
    } catch (const SynthesisError& e) {
        t->err = e.what();
        return e.kind == SynthesisError::Native ? BP_E_CUDA : BP_E_STATE;
    } catch (const std::out_of_range& e) {
        t->err = e.what();
        return BP_E_RANGE;
    } catch (const std::exception& e) {
        t->err = e.what();
        return BP_E_STATE;
    }
}

template <class CS> void alloc_message_bits(CS& cs, const uint8_t* msg, uint64_t len, std::vector<Boolean>& bits) {
    bits.reserve(len * 8);
    for (uint64_t i = 0; i < len; ++i)
        for (int j = 7; j >= 0; --j) {
            auto ns = cs.ns([&] { return "input bit " + std::to_string(i) + " " + std::to_string(j); });
            bits.push_back(Boolean::from(AllocatedBit::alloc(ns, (OptBool)((msg[i] >> j) & 1))));
        }
}

// The reference's boolean tests build each operand with `dyn_construct` (boolean.rs:1120-1148): a constant, or a bit allocated
// in the operand's own namespace, possibly negated.
template <class CS> Boolean boolean_operand(CS& cs, int kind, const char* name) {
    auto ns = cs.ns([&] { return std::string(name); });
    switch (kind) {
        case 0: return Boolean::constant(true);
        case 1: return Boolean::constant(false);
        case 2: return Boolean::from(AllocatedBit::alloc(ns, (OptBool)1));
        case 3: return Boolean::from(AllocatedBit::alloc(ns, (OptBool)0));
        case 4: return Boolean::from(AllocatedBit::alloc(ns, (OptBool)1)).not_();
        case 5: return Boolean::from(AllocatedBit::alloc(ns, (OptBool)0)).not_();
    }
    throw std::out_of_range("operand kind");
}

void bits_to_bytes_be(const std::vector<Boolean>& bits, uint8_t* out) {
    for (size_t i = 0; i < bits.size() / 8; ++i) {
        uint8_t b = 0;
        for (int j = 0; j < 8; ++j) {
            const OptBool v = bits[8 * i + j].get_value();
            if (v < 0) throw SynthesisError::assignment_missing();
            b = (uint8_t)((b << 1) | v);
        }
        out[i] = b;
    }
}

// sha256() with row filtering per compression block; same circuit as gadgets.hpp `sha256`.  Rows of the blocks inside any of the
// n_ranges half-open block ranges [ranges[2i], ranges[2i+1]) (ascending, disjoint) are kept, all others dropped; for each range
// global_before[i] = rows that precede its first row in the WHOLE circuit, local_before[i] = kept rows that precede it.
template <class CS>
void sha256_ranges(CS& cs, FilterSink& fs, const uint8_t* msg, uint64_t len, const uint64_t* ranges, uint64_t n_ranges, uint8_t digest[32],
                   uint64_t* global_before, uint64_t* local_before, WitnessTape* tape = nullptr) {
    struct TapeScope {  // recording is on exactly while this synthesis runs
        explicit TapeScope(WitnessTape* t) { g_tape = t; }
        ~TapeScope() { g_tape = nullptr; }
    } tape_scope(tape);
    if (tape) {
        tape->aux_base = cs.num_aux();
        tape->n_msg_bits = 8 * len;
    }
    auto keep_block = [&](uint64_t blk, uint64_t* which) {
        for (uint64_t i = 0; i < n_ranges; ++i)
            if (blk >= ranges[2 * i] && blk < ranges[2 * i + 1]) {
                *which = i;
                return true;
            }
        return false;
    };
    // the input-bit rows belong to "block -1": kept by whoever holds block 0
    cs.flush();
    uint64_t which = 0;
    fs.rows_enabled = keep_block(0, &which);
    const uint64_t dropped0 = fs.rows_dropped;
    uint64_t kept = 0, seen = 0;  // rows kept / rows produced so far by THIS call
    auto account = [&](uint64_t produced, bool was_kept) {
        seen += produced;
        if (was_kept) kept += produced;
    };
    uint64_t rows_at = cs.num_constraints();
    std::vector<Boolean> input;
    alloc_message_bits(cs, msg, len, input);
    cs.flush();
    account(cs.num_constraints() - rows_at, fs.rows_enabled);
    rows_at = cs.num_constraints();
    std::vector<Boolean> padded = input;
    const uint64_t plen = padded.size();
    padded.push_back(Boolean::constant(true));
    while ((padded.size() + 64) % 512 != 0) padded.push_back(Boolean::constant(false));
    for (int i = 63; i >= 0; --i) padded.push_back(Boolean::constant((plen >> i) & 1));
    std::vector<UInt32> cur = sha256_iv();
    for (uint64_t blk = 0; blk < padded.size() / 512; ++blk) {
        const bool keep = keep_block(blk, &which);
        if (keep && blk == ranges[2 * which]) {  // first block of a range: what precedes it
            // (a range that starts at block 0 also owns the input-bit rows)
            if (global_before) global_before[which] = blk == 0 ? 0 : seen;
            if (local_before) local_before[which] = blk == 0 ? 0 : kept;
        }
        fs.rows_enabled = keep;
        if (tape) {  // a unit of the witness program: its message window and which variables carry the chaining state
            tape->begin_unit(512 * blk);
            for (uint32_t j = 0; j < 8; ++j)
                for (uint32_t i = 0; i < 32; ++i) {
                    const Boolean& b = cur[j].bits[i];
                    if (!b.is_constant()) tape->units.back().state[b.bit.variable.index()] = (32 * j + i) | (b.kind == Boolean::Not ? 0x80000000u : 0u);
                }
        }
        auto ns = cs.ns([&] { return "block " + std::to_string(blk); });
        cur = sha256_compression_function(ns, padded.data() + 512 * blk, cur);
        cs.flush();
        account(cs.num_constraints() - rows_at, keep);
        rows_at = cs.num_constraints();
    }
    (void)dropped0;
    fs.rows_enabled = true;
    std::vector<Boolean> out;
    for (auto& wd : cur) wd.into_bits_be(out);
    bits_to_bytes_be(out, digest);
}

template <class CS>
void sha256_sharded(CS& cs, FilterSink& fs, const uint8_t* msg, uint64_t len, uint64_t bb, uint64_t be, uint8_t digest[32], uint64_t* rows_before) {
    const uint64_t r[2] = {bb, be};
    uint64_t g = 0, l = 0;
    sha256_ranges(cs, fs, msg, len, r, 1, digest, &g, &l);
    if (rows_before) *rows_before = g;
}

}  // namespace

extern "C" {

int bp_tcs_new(int field, int device, int named, uint64_t reserve_rows, uint64_t reserve_nnz, uint64_t reserve_vars, bp_tcs** out) {
    if (!out || field < 0 || field > 2) return BP_E_ARG;
    *out = nullptr;
    std::unique_ptr<bp_tcs> t(new bp_tcs());
    t->field = field;
    t->named = named != 0;
    if (device >= 0) {
        const int rc = bp_cs_new(field, device, reserve_rows, reserve_nnz, reserve_vars, &t->h);
        if (rc != BP_OK) return rc;
        t->base.reset(new DeviceSink(t->h));
    } else {
        t->base.reset(new HostSink());
    }
    t->filter.reset(new FilterSink(t->base.get()));
    if (t->named) t->named_cs.reset(new TestConstraintSystem(field, t->filter.get()));
    else t->bulk_cs.reset(new BulkConstraintSystem(field, t->filter.get()));
    *out = t.release();
    return BP_OK;
}

void bp_tcs_free(bp_tcs* t) {
    if (!t) return;
    t->named_cs.reset();
    t->bulk_cs.reset();
    if (t->h) bp_cs_free(t->h);
    delete t;
}

const char* bp_tcs_last_error(const bp_tcs* t) { return t ? t->err.c_str() : "null"; }
bp_cs* bp_tcs_handle(bp_tcs* t) { return t ? t->h : nullptr; }

int bp_tcs_flush(bp_tcs* t) {
    if (!t) return BP_E_ARG;
    return guarded(t, [&] {
        if (t->named) t->named_cs->flush();
        else t->bulk_cs->flush();
    });
}

int bp_tcs_sha256_block(bp_tcs* t, const uint8_t block[64], uint8_t out32[32]) {
    if (!t || !block || !out32) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            std::vector<Boolean> bits;
            for (int i = 0; i < 512; ++i) {
                auto ns = cs.ns([&] { return "input bit " + std::to_string(i); });
                bits.push_back(Boolean::from(AllocatedBit::alloc(ns, (OptBool)((block[i / 8] >> (7 - i % 8)) & 1))));
            }
            std::vector<UInt32> r = sha256_compression_function(cs, bits.data(), sha256_iv());
            std::vector<Boolean> ob;
            for (auto& wd : r) wd.into_bits_be(ob);
            bits_to_bytes_be(ob, out32);
            cs.flush();
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

int bp_tcs_sha256(bp_tcs* t, const uint8_t* msg, uint64_t len, uint64_t bb, uint64_t be, uint8_t digest[32], uint64_t* rows_before) {
    if (!t || (!msg && len) || !digest) return BP_E_ARG;
    const uint64_t r[2] = {bb, be};
    uint64_t g = 0, l = 0;
    const int rc = bb < be ? bp_tcs_sha256_ranges(t, msg, len, r, 1, digest, &g, &l) : BP_E_ARG;
    if (rc == BP_OK && rows_before) *rows_before = g;
    return rc;
}

int bp_tcs_sha256_ranges(bp_tcs* t, const uint8_t* msg, uint64_t len, const uint64_t* ranges, uint64_t n_ranges, uint8_t digest[32],
                         uint64_t* global_before, uint64_t* local_before) {
    if (!t || (!msg && len) || !digest || !ranges || !n_ranges) return BP_E_ARG;
    for (uint64_t i = 0; i < n_ranges; ++i)
        if (ranges[2 * i] >= ranges[2 * i + 1] || (i && ranges[2 * i] < ranges[2 * i - 1])) return BP_E_ARG;
    return guarded(t, [&] {
        WitnessTape tape;
        WitnessTape* tp = t->record_tape ? &tape : nullptr;
        if (t->named) sha256_ranges(*t->named_cs, *t->filter, msg, len, ranges, n_ranges, digest, global_before, local_before, tp);
        else sha256_ranges(*t->bulk_cs, *t->filter, msg, len, ranges, n_ranges, digest, global_before, local_before, tp);
        if (tp) {
            t->wprog = tape.build_program(t->named ? t->named_cs->num_aux() : t->bulk_cs->num_aux(), /*msb_first=*/true);
            if (!tape.ok) throw std::runtime_error("witness tape: " + tape.why);
        }
    });
}

int bp_tcs_record_witness_program(bp_tcs* t, int on) {
    if (!t) return BP_E_ARG;
    t->record_tape = on != 0;
    return BP_OK;
}

int bp_tcs_witness_program(bp_tcs* t, const uint32_t** words, uint64_t* n_words) {
    if (!t || !words || !n_words) return BP_E_ARG;
    if (t->wprog.empty()) {
        t->err = "no witness program recorded (bp_tcs_record_witness_program before the synthesis)";
        return BP_E_STATE;
    }
    *words = t->wprog.data();
    *n_words = t->wprog.size();
    return BP_OK;
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
// One SHA-256 compression with the x86 SHA extensions (two rounds per SHA256RNDS2), used when the CPU has them: the chaining
// states of a 4096-block message in ~0.2 ms instead of 1.7 ms with the portable code below.
__attribute__((target("sha,sse4.1,ssse3"))) static void sha256_compress_shani(uint32_t state[8], const uint8_t* data) {
    using sha256_detail::K;
    const __m128i MASK = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
    __m128i TMP = _mm_loadu_si128((const __m128i*)&state[0]);     // DCBA
    __m128i STATE1 = _mm_loadu_si128((const __m128i*)&state[4]);  // HGFE
    TMP = _mm_shuffle_epi32(TMP, 0xB1);                           // CDAB
    STATE1 = _mm_shuffle_epi32(STATE1, 0x1B);                     // EFGH
    __m128i STATE0 = _mm_alignr_epi8(TMP, STATE1, 8);             // ABEF
    STATE1 = _mm_blend_epi16(STATE1, TMP, 0xF0);                  // CDGH
    const __m128i ABEF_SAVE = STATE0, CDGH_SAVE = STATE1;
    __m128i M[4];
    for (int i = 0; i < 4; ++i) M[i] = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i*)(data + 16 * i)), MASK);
    for (int r = 0; r < 16; ++r) {  // four rounds each
        __m128i MSG = _mm_add_epi32(M[r & 3], _mm_loadu_si128((const __m128i*)&K[4 * r]));
        STATE1 = _mm_sha256rnds2_epu32(STATE1, STATE0, MSG);
        MSG = _mm_shuffle_epi32(MSG, 0x0E);
        STATE0 = _mm_sha256rnds2_epu32(STATE0, STATE1, MSG);
        if (r < 12) {  // schedule W[4(r+4) .. 4(r+4)+3] into M[r & 3]
            __m128i t = _mm_sha256msg1_epu32(M[r & 3], M[(r + 1) & 3]);
            t = _mm_add_epi32(t, _mm_alignr_epi8(M[(r + 3) & 3], M[(r + 2) & 3], 4));
            M[r & 3] = _mm_sha256msg2_epu32(t, M[(r + 3) & 3]);
        }
    }
    STATE0 = _mm_add_epi32(STATE0, ABEF_SAVE);
    STATE1 = _mm_add_epi32(STATE1, CDGH_SAVE);
    TMP = _mm_shuffle_epi32(STATE0, 0x1B);        // FEBA
    STATE1 = _mm_shuffle_epi32(STATE1, 0xB1);     // DCHG
    STATE0 = _mm_blend_epi16(TMP, STATE1, 0xF0);  // DCBA
    STATE1 = _mm_alignr_epi8(STATE1, TMP, 8);     // HGFE
    _mm_storeu_si128((__m128i*)&state[0], STATE0);
    _mm_storeu_si128((__m128i*)&state[4], STATE1);
}
static bool have_shani() {
    static const bool ok = __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3");
    return ok;
}
#else
static bool have_shani() { return false; }
static void sha256_compress_shani(uint32_t*, const uint8_t*) {}
#endif

// Chaining values of the blake2s gadget's hash (blake2s.rs:344-406: digest length 32, no key, 8-byte personalization) over `len`
// message bytes: the 8 words of h BEFORE each of the max(1, ceil(len / 64)) compressions -- what bp_cs_generate_witness_async
// takes per unit of a recorded blake2s circuit.  Plain BLAKE2s on the host.
int bp_blake2s_chain_states(const uint8_t* msg, uint64_t len, const uint8_t personalization[8], uint32_t* states, uint64_t max_blocks,
                            uint64_t* n_blocks) {
    if ((!msg && len) || !personalization || !n_blocks) return BP_E_ARG;
    const uint64_t blocks = len ? (len + 63) / 64 : 1;
    *n_blocks = blocks;
    if (!states) return BP_OK;
    if (max_blocks < blocks) return BP_E_RANGE;
    auto le32 = [](const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) h[i] = kBlake2sIV[i];
    h[0] ^= 0x01010000u ^ 32u;
    h[6] ^= le32(personalization);
    h[7] ^= le32(personalization + 4);
    for (uint64_t b = 0; b < blocks; ++b) {
        std::memcpy(states + 8 * b, h, sizeof h);
        uint8_t blk[64] = {0};
        const uint64_t have = len > 64 * b ? std::min<uint64_t>(64, len - 64 * b) : 0;
        if (have) std::memcpy(blk, msg + 64 * b, have);
        const bool last = b + 1 == blocks;
        blake2s_compress(h, blk, last ? len : 64 * (b + 1), last);
    }
    return BP_OK;
}

// Chaining states of sha256 over `len` message bytes: 8 words BEFORE each compression block (block 0: the IV), blocks as the
// gadget pads them (sha256.rs:50-77).  Plain SHA-256 on the host: what bp_cs_generate_witness_async wants per unit.
int bp_sha256_chain_states(const uint8_t* msg, uint64_t len, uint32_t* states, uint64_t max_blocks, uint64_t* n_blocks) {
    if ((!msg && len) || !n_blocks) return BP_E_ARG;
    using sha256_detail::IV;
    using sha256_detail::K;
    const uint64_t blocks = (len + 9 + 63) / 64;
    *n_blocks = blocks;
    if (!states) return BP_OK;
    if (max_blocks < blocks) return BP_E_RANGE;
    uint32_t h[8];
    std::memcpy(h, IV, sizeof h);
    auto rotr = [](uint32_t x, unsigned r) { return (x >> r) | (x << (32 - r)); };
    for (uint64_t b = 0; b < blocks; ++b) {
        std::memcpy(states + 8 * b, h, sizeof h);
        uint8_t blk[64];
        if (64 * (b + 1) <= len) std::memcpy(blk, msg + 64 * b, 64);
        else for (unsigned i = 0; i < 64; ++i) {
            const uint64_t p = 64 * b + i;
            uint8_t v = 0;
            if (p < len) v = msg[p];
            else if (p == len) v = 0x80;
            else if (p >= 64 * blocks - 8) v = (uint8_t)((8 * len) >> (8 * (64 * blocks - 1 - p)));
            blk[i] = v;
        }
        if (have_shani() && !getenv("BP_NO_SHANI")) {
            sha256_compress_shani(h, blk);
            continue;
        }
        uint32_t w[64];
        for (unsigned i = 0; i < 16; ++i) w[i] = (uint32_t)blk[4 * i] << 24 | (uint32_t)blk[4 * i + 1] << 16 | (uint32_t)blk[4 * i + 2] << 8 | blk[4 * i + 3];
        for (unsigned i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (unsigned i = 0; i < 64; ++i) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & bb) ^ (a & c) ^ (bb & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
        }
        h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    return BP_OK;
}

// ---- num gadgets (crates/bellpepper-core/src/gadgets/num.rs) -----------------------------------------------------------------
int bp_tcs_num_unpack(bp_tcs* t, const uint64_t value[4], int strict, uint8_t* bits_out) {
    if (!t || !value) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            Fr v;
            std::memcpy(v.l, value, 32);
            if (!cs.field()->is_canonical(v)) throw std::out_of_range("value >= p");
            const AllocatedNum n = AllocatedNum::alloc(cs, [&] { return v; });
            const std::vector<Boolean> bits = strict ? n.to_bits_le_strict(cs) : n.to_bits_le(cs);
            if (bits_out)
                for (size_t i = 0; i < bits.size(); ++i) bits_out[i] = (uint8_t)bits[i].get_value();
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

int bp_tcs_boolean_op(bp_tcs* t, int op, int kind_a, int kind_b, int kind_c, int* result_kind, int* result_value) {
    if (!t || op < 0 || op > 5) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            const Boolean a = boolean_operand(cs, kind_a, "a"), b = boolean_operand(cs, kind_b, "b");
            Boolean r = Boolean::constant(false);
            switch (op) {
                case 0: r = Boolean::xor_(cs, a, b); break;
                case 1: r = Boolean::and_(cs, a, b); break;
                case 2: r = Boolean::or_(cs, a, b); break;
                case 3: r = Boolean::sha256_ch(cs, a, b, boolean_operand(cs, kind_c, "c")); break;
                case 4: r = Boolean::sha256_maj(cs, a, b, boolean_operand(cs, kind_c, "c")); break;
                case 5: Boolean::enforce_equal(cs, a, b); break;
            }
            if (result_kind) *result_kind = r.kind == Boolean::Is ? 0 : r.kind == Boolean::Not ? 1 : 2;
            if (result_value) *result_value = (int)r.get_value();
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

// The reference's uint32 tests (uint32.rs:492-780): a = UInt32::alloc in "a_bit", b (and c for op 1) constants, the last operand
// allocated in "c_bit" / "d_bit"; op 0: (a xor b) in "first xor", then xor c in "second xor"; op 1: (a xor b) in "xor", then
// addmany [r, c, d] in "addition" under a MultiEq; op 2 / 3: sha256_maj / sha256_ch(a, b, c) at the root.
int bp_tcs_uint32_op(bp_tcs* t, int op, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t* result, uint32_t* n_constant_bits) {
    if (!t || op < 0 || op > 3) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            using CSv = std::remove_reference_t<decltype(cs)>;
            auto alloc_in = [&](const char* name, uint32_t v) {
                auto ns = cs.ns([&] { return std::string(name); });
                return UInt32::alloc(ns, v);
            };
            const UInt32 a_bit = alloc_in("a_bit", a);
            const UInt32 b_bit = UInt32::constant(b);
            UInt32 r = UInt32::constant(0);
            if (op == 0) {
                const UInt32 c_bit = alloc_in("c_bit", c);
                { auto ns = cs.ns([] { return std::string("first xor"); }); r = a_bit.xor_(ns, b_bit); }
                { auto ns = cs.ns([] { return std::string("second xor"); }); r = r.xor_(ns, c_bit); }
            } else if (op == 1) {
                const UInt32 c_bit = UInt32::constant(c);
                const UInt32 d_bit = alloc_in("d_bit", d);
                { auto ns = cs.ns([] { return std::string("xor"); }); r = a_bit.xor_(ns, b_bit); }
                {
                    MultiEq<CSv> me(cs);
                    auto ns = me.ns([] { return std::string("addition"); });
                    const UInt32 ops[3] = {r, c_bit, d_bit};
                    r = UInt32::addmany(ns, ops, 3);
                }  // the MultiEq drops here: its wide row is emitted
            } else {
                const UInt32 c_bit = alloc_in("c_bit", c);
                r = op == 2 ? UInt32::sha256_maj(cs, a_bit, b_bit, c_bit) : UInt32::sha256_ch(cs, a_bit, b_bit, c_bit);
            }
            uint32_t v = 0, n_const = 0;
            for (unsigned i = 0; i < 32; ++i) {
                const OptBool bv = r.bits[i].get_value();
                if (bv < 0) throw SynthesisError::assignment_missing();
                v |= (uint32_t)bv << i;
                n_const += r.bits[i].is_constant();
            }
            if (result) *result = v;
            if (n_constant_bits) *n_constant_bits = n_const;
            cs.flush();
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

int bp_tcs_u64_bits(bp_tcs* t, uint64_t value, uint8_t bits_out[64]) {
    if (!t) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            const std::vector<Boolean> bits = u64_into_boolean_vec_le(cs, &value);
            if (bits_out)
                for (size_t i = 0; i < bits.size(); ++i) bits_out[i] = (uint8_t)bits[i].get_value();
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

int bp_tcs_num_arith(bp_tcs* t, const uint64_t a4[4], const uint64_t b4[4]) {
    if (!t || !a4 || !b4) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            Fr av, bv;
            std::memcpy(av.l, a4, 32);
            std::memcpy(bv.l, b4, 32);
            AllocatedNum a{std::nullopt, Variable{0}}, b{std::nullopt, Variable{0}};
            {
                auto ns = cs.ns([] { return std::string("a"); });
                a = AllocatedNum::alloc(ns, [&] { return av; });
            }
            {
                auto ns = cs.ns([] { return std::string("b"); });
                b = AllocatedNum::alloc(ns, [&] { return bv; });
            }
            const AllocatedNum prod = a.mul(cs, b);     // "product num", "multiplication constraint"
            const AllocatedNum sq = a.square(cs);       // "squared num", "squaring constraint"
            const AllocatedNum sum = prod.add(cs, sq);  // "sum num", "addition constraint"
            (void)sum;
            {
                auto ns = cs.ns([] { return std::string("nonzero"); });
                b.assert_nonzero(ns);                   // "nonzero/ephemeral inverse", "nonzero/nonzero assertion constraint"
            }
            {
                auto ns = cs.ns([] { return std::string("swap"); });
                const AllocatedBit c = AllocatedBit::alloc(ns.ns([] { return std::string("condition"); }), (OptBool)1);
                AllocatedNum::conditionally_reverse(ns, a, b, Boolean::from(c));
            }
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

// A product-heavy GADGET circuit for the full-width kernels: x <- x^2 * y + x, n times, x unpacked into bits every
// `unpack_every` steps (0 = never): multiplication / squaring / addition rows over full-width values and 256-term rows.
int bp_tcs_num_chain(bp_tcs* t, uint64_t n_steps, uint64_t unpack_every, const uint64_t x0[4], const uint64_t y0[4]) {
    if (!t || !x0 || !y0) return BP_E_ARG;
    return guarded(t, [&] {
        auto run = [&](auto& cs) {
            Fr xv, yv;
            std::memcpy(xv.l, x0, 32);
            std::memcpy(yv.l, y0, 32);
            AllocatedNum x{std::nullopt, Variable{0}}, y{std::nullopt, Variable{0}};
            {
                auto ns = cs.ns([] { return std::string("x0"); });
                x = AllocatedNum::alloc(ns, [&] { return xv; });
            }
            {
                auto ns = cs.ns([] { return std::string("y"); });
                y = AllocatedNum::alloc(ns, [&] { return yv; });
            }
            for (uint64_t i = 0; i < n_steps; ++i) {
                auto ns = cs.ns([&] { return "step " + std::to_string(i); });
                const AllocatedNum s = x.square(ns);
                AllocatedNum m{std::nullopt, Variable{0}};
                {
                    auto n2 = ns.ns([] { return std::string("times y"); });
                    m = s.mul(n2, y);
                }
                {
                    auto n2 = ns.ns([] { return std::string("plus x"); });
                    x = m.add(n2, x);
                }
                if (unpack_every && (i + 1) % unpack_every == 0) {
                    auto n2 = ns.ns([] { return std::string("unpack"); });
                    x.to_bits_le(n2);
                }
            }
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
    });
}

int bp_tcs_blake2s(bp_tcs* t, const uint8_t* msg, uint64_t len, const uint8_t personalization[8], uint8_t digest[32]) {
    if (!t || (!msg && len) || !personalization || !digest) return BP_E_ARG;
    return guarded(t, [&] {
        WitnessTape tape;
        auto run = [&](auto& cs) {
            struct TapeScope {  // recording is on exactly while this synthesis runs
                explicit TapeScope(WitnessTape* tp) { g_tape = tp; }
                ~TapeScope() { g_tape = nullptr; }
            } tape_scope(t->record_tape ? &tape : nullptr);
            if (t->record_tape) {
                tape.aux_base = cs.num_aux();
                tape.n_msg_bits = 8 * len;
            }
            std::vector<Boolean> bits;
            bits.reserve(len * 8);
            for (uint64_t i = 0; i < len; ++i)
                for (int j = 0; j < 8; ++j) {  // little-endian bit order per byte (blake2s.rs:520)
                    auto ns = cs.ns([&] { return "input bit " + std::to_string(i) + " " + std::to_string(j); });
                    bits.push_back(Boolean::from(AllocatedBit::alloc(ns, (OptBool)((msg[i] >> j) & 1))));
                }
            std::vector<Boolean> out = blake2s(cs, bits, personalization);
            for (size_t i = 0; i < 32; ++i) {
                uint8_t b = 0;
                for (int j = 0; j < 8; ++j) {
                    const OptBool v = out[8 * i + j].get_value();
                    if (v < 0) throw SynthesisError::assignment_missing();
                    b = (uint8_t)(b | (v << j));
                }
                digest[i] = b;
            }
            cs.flush();
        };
        if (t->named) run(*t->named_cs);
        else run(*t->bulk_cs);
        if (t->record_tape) {  // one unit per compression; message bits least significant first within a byte
            t->wprog = tape.build_program(t->named ? t->named_cs->num_aux() : t->bulk_cs->num_aux(), /*msb_first=*/false);
            if (!tape.ok) throw std::runtime_error("witness tape: " + tape.why);
        }
    });
}

// Scripted exercise of the C++ WitnessCS against the reference's own test (witness_cs.rs tests + :94-201); returns 0 or the
// number of the first failed expectation.
int bp_wcs_selftest(int field, int device) {
    bp_cs *h1 = nullptr, *h2 = nullptr;
    if (bp_cs_new(field, device, 0, 0, 0, &h1) != BP_OK || bp_cs_new(field, device, 0, 0, 0, &h2) != BP_OK) return -1;
    int bad = 0;
    try {
        const FieldParams& fp = field_params(field);
        Fr pm1{{fp.p[0] - 1, fp.p[1], fp.p[2], fp.p[3]}};
        WitnessCS w(field, h1), other(field, h2);
        auto nm = [] { return std::string(); };
        Variable a = w.alloc(nm, [] { return Fr::from_u64(2); });
        Variable b = w.alloc_input(nm, [&] { return pm1; });
        w.enforce(nm, [](LinearCombination lc) { return lc; }, [](LinearCombination lc) { return lc; }, [](LinearCombination lc) { return lc; });
        if (!bad && !(a == Variable::aux(0) && b == Variable::input(1))) bad = 1;
        other.alloc_input(nm, [] { return Fr::from_u64(7); });
        other.alloc(nm, [] { return Fr::from_u64(300); });
        other.alloc(nm, [] { return Fr::from_u64(1); });
        w.extend(other);  // inputs: [1, p-1, 7]; aux: [2, 300, 1]
        auto fresh = w.allocate_empty(3, 1);  // aux first
        if (!bad && !(fresh.first == 3 && fresh.second == 3)) bad = 2;
        Fr fill_a[3] = {Fr::from_u64(10), pm1, Fr::from_u64(12)}, fill_i[1] = {Fr::from_u64(99)};
        w.fill_aux(fresh.first, fill_a, 3);
        w.fill_inputs(fresh.second, fill_i, 1);
        const std::vector<Fr> in = w.inputs_slice(), ax = w.aux_slice();
        const Fr want_in[4] = {Fr::one(), pm1, Fr::from_u64(7), Fr::from_u64(99)};
        const Fr want_ax[6] = {Fr::from_u64(2), Fr::from_u64(300), Fr::one(), Fr::from_u64(10), pm1, Fr::from_u64(12)};
        if (!bad && !(in.size() == 4 && ax.size() == 6)) bad = 3;
        for (size_t i = 0; !bad && i < 4; ++i) if (in[i] != want_in[i]) bad = 10 + (int)i;
        for (size_t i = 0; !bad && i < 6; ++i) if (ax[i] != want_ax[i]) bad = 20 + (int)i;
        if (!bad && !(WitnessCS::is_extensible() && WitnessCS::is_witness_generator())) bad = 4;
    } catch (const std::exception&) {
        bad = -2;
    }
    bp_cs_free(h1);
    bp_cs_free(h2);
    return bad;
}

int64_t bp_tcs_which_is_unsatisfied(bp_tcs* t, char* path, uint64_t cap) {
    if (!t) return -5;
    int64_t row = -1;
    const int rc = guarded(t, [&] {
        row = t->named ? t->named_cs->first_unsatisfied_row() : t->bulk_cs->first_unsatisfied_row();
        if (path && cap) {
            path[0] = 0;
            if (row >= 0 && t->named) {
                const std::string& p = t->named_cs->row_path((uint64_t)row);
                std::strncpy(path, p.c_str(), cap - 1);
                path[cap - 1] = 0;
            }
        }
    });
    return rc == BP_OK ? row : (int64_t)rc - 1;
}

int bp_tcs_set(bp_tcs* t, const char* path, const uint64_t v[4]) {
    if (!t || !path || !v) return BP_E_ARG;
    if (!t->named) return BP_E_STATE;
    return guarded(t, [&] {
        Fr x;
        std::memcpy(x.l, v, 32);
        t->named_cs->set(path, x);
    });
}

int bp_tcs_get(bp_tcs* t, const char* path, uint64_t v[4]) {
    if (!t || !path || !v) return BP_E_ARG;
    if (!t->named) return BP_E_STATE;
    return guarded(t, [&] {
        const Fr x = t->named_cs->get(path);
        std::memcpy(v, x.l, 32);
    });
}

uint64_t bp_tcs_num_constraints(const bp_tcs* t) { return t->named ? t->named_cs->num_constraints() : t->bulk_cs->num_constraints(); }
uint64_t bp_tcs_num_inputs(const bp_tcs* t) { return t->named ? t->named_cs->num_inputs() : t->bulk_cs->num_inputs(); }
uint64_t bp_tcs_num_aux(const bp_tcs* t) { return t->named ? t->named_cs->num_aux() : t->bulk_cs->num_aux(); }

int bp_tcs_row_path(const bp_tcs* t, uint64_t row, char* path, uint64_t cap) {
    if (!t || !path || !cap) return BP_E_ARG;
    if (!t->named) return BP_E_STATE;
    if (row >= t->named_cs->num_constraints()) return BP_E_RANGE;
    const std::string& p = t->named_cs->row_path(row);
    std::strncpy(path, p.c_str(), cap - 1);
    path[cap - 1] = 0;
    return BP_OK;
}

int bp_tcs_host_csr(bp_tcs* t, const uint32_t** lens, uint64_t* n_rows, const uint32_t** cols, const uint64_t** coeffs, uint64_t* nnz,
                    const uint64_t** inputs, uint64_t* n_inputs, const uint64_t** aux, uint64_t* n_aux) {
    if (!t) return BP_E_ARG;
    HostSink* hs = dynamic_cast<HostSink*>(t->base.get());
    if (!hs) return BP_E_STATE;
    const int rc = bp_tcs_flush(t);
    if (rc != BP_OK) return rc;
    *lens = hs->lens.data();
    *n_rows = hs->lens.size() / 3;
    *cols = hs->cols.data();
    *coeffs = hs->coeffs.data();
    *nnz = hs->cols.size();
    *inputs = hs->inputs.data();
    *n_inputs = hs->inputs.size() / 4;
    *aux = hs->aux.data();
    *n_aux = hs->aux.size() / 4;
    return BP_OK;
}

}  // extern "C"
