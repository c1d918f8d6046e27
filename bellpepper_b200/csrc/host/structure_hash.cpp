// TestConstraintSystem::hash (crates/bellpepper-core/src/util_cs/test_cs.rs:64-115, 214-237) computed from the flat rows that
// cross the C ABI (what bp_cs_enforce takes, what bp_cs_export writes): a circuit's structural fingerprint for C / C++ / Rust
// callers -- two front-ends that emit the same matrices get the same 64 hex digits, whatever produced them.  Host only.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/bp_r1cs.h"
#include "blake2s_host.hpp"
#include "fr.hpp"

namespace {

// BLAKE2s-256, unkeyed, default parameters (RFC 7693) -- what `Blake2s::new()` of blake2s_simd is (test_cs.rs:215).
struct Blake2s {
    uint32_t h[8];
    uint64_t t = 0;
    uint8_t buf[64];
    size_t fill = 0;

    Blake2s() {
        for (int i = 0; i < 8; ++i) h[i] = bph::kBlake2sIV[i];
        h[0] ^= 0x01010020u;  // digest length 32, no key, fanout 1, depth 1
    }
    void compress(const uint8_t* block, bool last) { bph::blake2s_compress(h, block, t, last); }
    void update(const uint8_t* p, size_t n) {
        while (n) {
            if (fill == 64) {  // (a full buffer is only compressed when more input follows: the last block is special)
                t += 64;
                compress(buf, false);
                fill = 0;
            }
            const size_t k = std::min(n, 64 - fill);
            std::memcpy(buf + fill, p, k);
            fill += k;
            p += k;
            n -= k;
        }
    }
    void finish(uint8_t out[32]) {
        t += fill;
        std::memset(buf + fill, 0, 64 - fill);
        compress(buf, true);
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(h[i] >> (8 * j));
    }
};

void be64(uint8_t* p, uint64_t v) {
    for (int i = 0; i < 8; ++i) p[i] = (uint8_t)(v >> (56 - 8 * i));
}

}  // namespace

extern "C" int bp_structure_hash(int field, uint64_t n_inputs, uint64_t n_aux, uint64_t n_rows, const uint32_t* lens, const uint32_t* cols,
                                 const uint64_t* coeffs_le, char out_hex[65]) {
    if (field < 0 || field > 2 || !out_hex || (n_rows && !lens)) return BP_E_ARG;
    try {
        const bph::Field f(field);
        Blake2s h;
        uint8_t buf[9 + 32];
        be64(buf, n_inputs);
        be64(buf + 8, n_aux);
        be64(buf + 16, n_rows);
        h.update(buf, 24);
        struct T { uint32_t col; bph::Fr c; };
        std::vector<T> lc;
        uint64_t k = 0;
        for (uint64_t i = 0; i < 3 * n_rows; ++i) {
            // proc_lc (test_cs.rs:64-87): a map ordered inputs-before-aux, ascending index -- the order of the tagged column --
            // same-variable coefficients added, zero coefficients dropped
            lc.clear();
            const uint32_t len = lens[i];
            if (len && (!cols || !coeffs_le)) return BP_E_ARG;
            for (uint32_t j = 0; j < len; ++j, ++k) {
                T t;
                t.col = cols[k];
                std::memcpy(t.c.l, coeffs_le + 4 * k, 32);
                if (!f.is_canonical(t.c)) return BP_E_RANGE;
                lc.push_back(t);
            }
            if (!std::is_sorted(lc.begin(), lc.end(), [](const T& a, const T& b) { return a.col < b.col; }))
                std::stable_sort(lc.begin(), lc.end(), [](const T& a, const T& b) { return a.col < b.col; });
            size_t w = 0;
            for (size_t r = 0; r < lc.size();) {
                T m = lc[r++];
                while (r < lc.size() && lc[r].col == m.col) m.c = f.add(m.c, lc[r++].c);
                if (!m.c.is_zero()) lc[w++] = m;
            }
            be64(buf, w);
            h.update(buf, 8);
            for (size_t r = 0; r < w; ++r) {
                buf[0] = (lc[r].col & BP_COL_AUX) ? 'A' : 'I';
                be64(buf + 1, lc[r].col & ~BP_COL_AUX);
                for (int b = 0; b < 32; ++b) buf[9 + b] = (uint8_t)(lc[r].c.l[3 - b / 8] >> (56 - 8 * (b % 8)));  // to_repr(), reversed
                h.update(buf, 41);
            }
        }
        uint8_t d[32];
        h.finish(d);
        static const char* hex = "0123456789abcdef";
        for (int i = 0; i < 32; ++i) {
            out_hex[2 * i] = hex[d[i] >> 4];
            out_hex[2 * i + 1] = hex[d[i] & 15];
        }
        out_hex[64] = 0;
        return BP_OK;
    } catch (...) {
        return BP_E_OOM;
    }
}
