// Host-side field element for the C++ front-end: canonical residue in 4 x u64 limbs.
// Mirrors the slice of `ff::PrimeField` the reference's gadgets and LC algebra use:
// ZERO, ONE, from(u64), add/sub/neg/double/mul, is_zero, ==, pow2.  Coefficient algebra only -- witness
// evaluation never happens here (that is the GPU's job).
#pragma once
#include <cstdint>
#include <cstring>

namespace bph {

struct FieldParams {
    uint64_t p[4];
    uint64_t inv;    // -p^-1 mod 2^64
    uint64_t r2[4];  // 2^512 mod p
};

inline const FieldParams& field_params(int f) {
    static const FieldParams P[3] = {
        {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
         0xfffffffeffffffffULL,
         {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}},
        {{0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0ULL, 0x4000000000000000ULL},
         0x8c46eb20ffffffffULL,
         {0xfc9678ff0000000fULL, 0x67bb433d891a16e3ULL, 0x7fae231004ccf590ULL, 0x096d41af7ccfdaa9ULL}},
        {{0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0ULL, 0x4000000000000000ULL},
         0x992d30ecffffffffULL,
         {0x8c78ecb30000000fULL, 0xd7d30dbd8b0de0e7ULL, 0x7797a99bc3c95d18ULL, 0x096d41af7b9cb714ULL}},
    };
    return P[f];
}

struct Fr {
    uint64_t l[4];

    static Fr zero() { return Fr{{0, 0, 0, 0}}; }
    static Fr one() { return Fr{{1, 0, 0, 0}}; }
    static Fr from_u64(uint64_t v) { return Fr{{v, 0, 0, 0}}; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fr& o) const { return std::memcmp(l, o.l, 32) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
};

typedef unsigned __int128 u128;

inline bool geq(const uint64_t* a, const uint64_t* b) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
inline void sub_raw(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - br;
        r[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}

// A field "context": the modulus travels with the constraint system, not with each element.
struct Field {
    int id;
    const FieldParams* fp;
    explicit Field(int f) : id(f), fp(&field_params(f)) {}

    Fr add(const Fr& a, const Fr& b) const {
        Fr t;
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)a.l[i] + b.l[i];
            t.l[i] = (uint64_t)c;
            c >>= 64;
        }
        if (geq(t.l, fp->p)) sub_raw(t.l, t.l, fp->p);  // p < 2^255: no carry out
        return t;
    }
    Fr neg(const Fr& a) const {
        if (a.is_zero()) return a;
        Fr t;
        sub_raw(t.l, fp->p, a.l);
        return t;
    }
    Fr sub(const Fr& a, const Fr& b) const { return add(a, neg(b)); }
    Fr dbl(const Fr& a) const { return add(a, a); }

    // CIOS Montgomery product a*b/2^256 mod p
    Fr mont(const Fr& a, const Fr& b) const {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)a.l[j] * b.l[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * fp->inv;
            c = ((u128)m * fp->p[0] + t[0]) >> 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)m * fp->p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fr r;
        if (t[4] || geq(t, fp->p)) sub_raw(r.l, t, fp->p);
        else std::memcpy(r.l, t, 32);
        return r;
    }
    // canonical a*b mod p
    Fr mul(const Fr& a, const Fr& b) const {
        Fr r2;
        std::memcpy(r2.l, fp->r2, 32);
        return mont(mont(a, b), r2);
    }
    // 2^k mod p, k < 2^16 (MultiEq shift coefficients: Scalar::from(2).pow_vartime([bits_used]))
    Fr pow2(unsigned k) const {
        if (k < 254) {  // below the modulus' bit length: a plain shift
            Fr r = Fr::zero();
            r.l[k / 64] = 1ULL << (k % 64);
            return r;
        }
        Fr r = pow2(253);
        for (unsigned i = 253; i < k; ++i) r = dbl(r);
        return r;
    }
    // a * 2^k mod p.  Gadget coefficients are almost always +-(small) so the product is a plain shift; anything
    // else takes the general multiplication.
    Fr mul_pow2(const Fr& a, unsigned k) const {
        auto shl_small = [](uint64_t v, unsigned sh) {
            Fr r = Fr::zero();
            const unsigned w = sh / 64, b = sh % 64;
            r.l[w] = v << b;
            if (b && w + 1 < 4) r.l[w + 1] = v >> (64 - b);
            return r;
        };
        if (k + 64 < 254) {
            if ((a.l[1] | a.l[2] | a.l[3]) == 0) return shl_small(a.l[0], k);
            Fr n;
            sub_raw(n.l, fp->p, a.l);  // p - a
            if ((n.l[1] | n.l[2] | n.l[3]) == 0) return neg(shl_small(n.l[0], k));
        }
        return mul(a, pow2(k));
    }
    bool is_canonical(const Fr& a) const { return !geq(a.l, fp->p); }
    Fr square(const Fr& a) const { return mul(a, a); }
    // a^(p-2): the inverse of a non-zero element (`invert()` in num.rs:378-389; witness values of the num gadgets only)
    Fr invert(const Fr& a) const {
        uint64_t e[4];
        const uint64_t two[4] = {2, 0, 0, 0};
        sub_raw(e, fp->p, two);
        Fr r = Fr::one(), b = a;
        for (int i = 0; i < 256; ++i) {
            if ((e[i / 64] >> (i % 64)) & 1) r = mul(r, b);
            b = mul(b, b);
        }
        return r;
    }
    // bit i of the canonical representation (to_le_bits)
    static bool bit(const Fr& a, unsigned i) { return (a.l[i / 64] >> (i % 64)) & 1; }
};

}  // namespace bph
