// TestConstraintSystem::hash (crates/bellpepper-core/src/util_cs/test_cs.rs:64-115, 214-237) computed from the flat rows that
// cross the C ABI (what bp_cs_enforce takes, what bp_cs_export writes): a circuit's structural fingerprint for C / C++ / Rust
// callers -- two front-ends that emit the same matrices get the same 64 hex digits, whatever produced them.  Host only.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/bp_r1cs.h"
#include "fr.hpp"

namespace {

// BLAKE2s-256, unkeyed, default parameters (RFC 7693) -- what `Blake2s::new()` of blake2s_simd is (test_cs.rs:215).
struct Blake2s {
    uint32_t h[8];
    uint64_t t = 0;
    uint8_t buf[64];
    size_t fill = 0;

    static constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};

    Blake2s() {
        for (int i = 0; i < 8; ++i) h[i] = IV[i];
        h[0] ^= 0x01010020u;  // digest length 32, no key, fanout 1, depth 1
    }
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(const uint8_t* block, bool last) {
        static const uint8_t S[10][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        uint32_t m[16], v[16];
        for (int i = 0; i < 16; ++i)
            m[i] = (uint32_t)block[4 * i] | (uint32_t)block[4 * i + 1] << 8 | (uint32_t)block[4 * i + 2] << 16 | (uint32_t)block[4 * i + 3] << 24;
        for (int i = 0; i < 8; ++i) {
            v[i] = h[i];
            v[8 + i] = IV[i];
        }
        v[12] ^= (uint32_t)t;
        v[13] ^= (uint32_t)(t >> 32);
        if (last) v[14] = ~v[14];
        auto G = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
            v[a] = v[a] + v[b] + x;
            v[d] = rotr(v[d] ^ v[a], 16);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 12);
            v[a] = v[a] + v[b] + y;
            v[d] = rotr(v[d] ^ v[a], 8);
            v[c] = v[c] + v[d];
            v[b] = rotr(v[b] ^ v[c], 7);
        };
        for (int r = 0; r < 10; ++r) {
            const uint8_t* s = S[r];
            G(0, 4, 8, 12, m[s[0]], m[s[1]]);
            G(1, 5, 9, 13, m[s[2]], m[s[3]]);
            G(2, 6, 10, 14, m[s[4]], m[s[5]]);
            G(3, 7, 11, 15, m[s[6]], m[s[7]]);
            G(0, 5, 10, 15, m[s[8]], m[s[9]]);
            G(1, 6, 11, 12, m[s[10]], m[s[11]]);
            G(2, 7, 8, 13, m[s[12]], m[s[13]]);
            G(3, 4, 9, 14, m[s[14]], m[s[15]]);
        }
        for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[8 + i];
    }
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
constexpr uint32_t Blake2s::IV[8];

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
