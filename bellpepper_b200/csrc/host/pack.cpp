// See pack.hpp.  Host only: no CUDA, no handle.
#include "pack.hpp"
#include "fr.hpp"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "../../../include/bp_r1cs.h"

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define BP_PACK_HAVE_AVX2 1
#endif

namespace bp {
namespace {

void pack_portable(const uint64_t* src, uint64_t n, uint8_t* dst, uint64_t b0, uint64_t b1, std::vector<PackExc>& out) {
    for (uint64_t b = b0; b < b1; ++b) {
        uint8_t acc = 0;
        const uint64_t e0 = 8 * b, e1 = std::min<uint64_t>(e0 + 8, n);
        for (uint64_t i = e0; i < e1; ++i) {
            const uint64_t* v = src + 4 * i;
            if ((v[1] | v[2] | v[3]) == 0 && v[0] <= 1) {
                acc |= (uint8_t)(v[0] << (i - e0));
            } else {
                out.push_back(PackExc{i, {v[0], v[1], v[2], v[3]}});
            }
        }
        dst[b] = acc;
    }
}

#ifdef BP_PACK_HAVE_AVX2
// One output byte per step: eight 256-bit loads (8 elements, four cache lines), one OR tree and one VPTEST say whether all
// eight are 0 or 1 (every bit but bit 0 of limb 0 clear); then the byte is the OR of the elements shifted into place.  A group
// with an exception takes the portable loop.  The pass is bound by how many cache misses one core keeps in flight (measured:
// 8 GB/s per thread for the portable loop, 12 GB/s for this one), hence the prefetch 2 KiB ahead of the loads.
__attribute__((target("avx2"))) void pack_avx2(const uint64_t* src, uint64_t n, uint8_t* dst, uint64_t b0, uint64_t b1,
                                               std::vector<PackExc>& out) {
    const __m256i not_bit = _mm256_set_epi64x(-1, -1, -1, ~1ll);
    const uint64_t full = std::min(b1, n / 8);  // bytes below `full` have all eight elements
    uint64_t b = b0;
    for (; b < full; ++b) {
        const __m256i* p = (const __m256i*)(src + 32 * b);
        const char* ahead = (const char*)p + 2048;  // (a prefetch never faults: running past the array is harmless)
        _mm_prefetch(ahead, _MM_HINT_T0);
        _mm_prefetch(ahead + 64, _MM_HINT_T0);
        _mm_prefetch(ahead + 128, _MM_HINT_T0);
        _mm_prefetch(ahead + 192, _MM_HINT_T0);
        const __m256i v0 = _mm256_loadu_si256(p), v1 = _mm256_loadu_si256(p + 1), v2 = _mm256_loadu_si256(p + 2),
                      v3 = _mm256_loadu_si256(p + 3), v4 = _mm256_loadu_si256(p + 4), v5 = _mm256_loadu_si256(p + 5),
                      v6 = _mm256_loadu_si256(p + 6), v7 = _mm256_loadu_si256(p + 7);
        const __m256i any = _mm256_or_si256(_mm256_or_si256(_mm256_or_si256(v0, v1), _mm256_or_si256(v2, v3)),
                                            _mm256_or_si256(_mm256_or_si256(v4, v5), _mm256_or_si256(v6, v7)));
        if (__builtin_expect(_mm256_testz_si256(any, not_bit), 1)) {
            const __m256i lo = _mm256_or_si256(_mm256_or_si256(v0, _mm256_slli_epi64(v1, 1)),
                                               _mm256_or_si256(_mm256_slli_epi64(v2, 2), _mm256_slli_epi64(v3, 3)));
            const __m256i hi = _mm256_or_si256(_mm256_or_si256(_mm256_slli_epi64(v4, 4), _mm256_slli_epi64(v5, 5)),
                                               _mm256_or_si256(_mm256_slli_epi64(v6, 6), _mm256_slli_epi64(v7, 7)));
            dst[b] = (uint8_t)_mm_cvtsi128_si32(_mm256_castsi256_si128(_mm256_or_si256(lo, hi)));
        } else {
            pack_portable(src, n, dst, b, b + 1, out);
        }
    }
    if (b < b1) pack_portable(src, n, dst, b, b1, out);  // the ragged last byte
}
#endif

void pack_mont_portable(const uint64_t* src, uint64_t n, const uint64_t* one, uint8_t* dst, uint64_t b0, uint64_t b1,
                        std::vector<PackExc>& out) {
    for (uint64_t b = b0; b < b1; ++b) {
        uint8_t acc = 0;
        const uint64_t e0 = 8 * b, e1 = std::min<uint64_t>(e0 + 8, n);
        for (uint64_t i = e0; i < e1; ++i) {
            const uint64_t* v = src + 4 * i;
            if ((v[0] | v[1] | v[2] | v[3]) == 0) continue;
            if (v[0] == one[0] && v[1] == one[1] && v[2] == one[2] && v[3] == one[3]) {
                acc |= (uint8_t)(1u << (i - e0));
            } else {
                out.push_back(PackExc{i, {v[0], v[1], v[2], v[3]}});
            }
        }
        dst[b] = acc;
    }
}

#ifdef BP_PACK_HAVE_AVX2
// As pack_avx2, with two 64-bit-lane compares per element (all-zero / equal to `one`) instead of the mask test.
__attribute__((target("avx2"))) void pack_mont_avx2(const uint64_t* src, uint64_t n, const uint64_t* one, uint8_t* dst, uint64_t b0,
                                                    uint64_t b1, std::vector<PackExc>& out) {
    const __m256i zero = _mm256_setzero_si256();
    const __m256i pat = _mm256_loadu_si256((const __m256i*)one);
    const uint64_t full = std::min(b1, n / 8);
    uint64_t b = b0;
    for (; b < full; ++b) {
        const __m256i* p = (const __m256i*)(src + 32 * b);
        const char* ahead = (const char*)p + 2048;
        _mm_prefetch(ahead, _MM_HINT_T0);
        _mm_prefetch(ahead + 64, _MM_HINT_T0);
        _mm_prefetch(ahead + 128, _MM_HINT_T0);
        _mm_prefetch(ahead + 192, _MM_HINT_T0);
        unsigned ones = 0, known = 0;
#pragma GCC unroll 8
        for (int k = 0; k < 8; ++k) {
            const __m256i v = _mm256_loadu_si256(p + k);
            const unsigned is0 = _mm256_movemask_pd(_mm256_castsi256_pd(_mm256_cmpeq_epi64(v, zero))) == 15u;
            const unsigned is1 = _mm256_movemask_pd(_mm256_castsi256_pd(_mm256_cmpeq_epi64(v, pat))) == 15u;
            ones |= is1 << k;
            known |= (is0 | is1) << k;
        }
        if (__builtin_expect(known == 0xffu, 1)) {
            dst[b] = (uint8_t)ones;
        } else {
            pack_mont_portable(src, n, one, dst, b, b + 1, out);
        }
    }
    if (b < b1) pack_mont_portable(src, n, one, dst, b, b1, out);
}
#endif

bool use_avx2() {
#ifdef BP_PACK_HAVE_AVX2
    static const bool ok = [] {
        const char* e = std::getenv("BP_PACK_SIMD");
        return !(e && e[0] == '0') && __builtin_cpu_supports("avx2");
    }();
    return ok;
#else
    return false;
#endif
}

}  // namespace

void pack_bit_bytes(const uint64_t* src, uint64_t n, uint8_t* dst, uint64_t b0, uint64_t b1, std::vector<PackExc>& out) {
    b1 = std::min<uint64_t>(b1, (n + 7) / 8);
    if (b0 >= b1) return;
#ifdef BP_PACK_HAVE_AVX2
    if (use_avx2()) return pack_avx2(src, n, dst, b0, b1, out);
#endif
    pack_portable(src, n, dst, b0, b1, out);
}

void pack_bit_bytes_mont(const uint64_t* src, uint64_t n, const uint64_t one[4], uint8_t* dst, uint64_t b0, uint64_t b1,
                         std::vector<PackExc>& out) {
    b1 = std::min<uint64_t>(b1, (n + 7) / 8);
    if (b0 >= b1) return;
#ifdef BP_PACK_HAVE_AVX2
    if (use_avx2()) return pack_mont_avx2(src, n, one, dst, b0, b1, out);
#endif
    pack_mont_portable(src, n, one, dst, b0, b1, out);
}

void mont_one(int field, uint64_t one[4]) {
    const bph::Field f(field);
    bph::Fr r2;
    std::memcpy(r2.l, f.fp->r2, 32);
    const bph::Fr r = f.mont(bph::Fr::one(), r2);  // 1 * 2^512 / 2^256
    std::memcpy(one, r.l, 32);
}

bool from_mont(int field, const uint64_t x[4], uint64_t out[4]) {
    const bph::Field f(field);
    bph::Fr a;
    std::memcpy(a.l, x, 32);
    if (!f.is_canonical(a)) return false;
    const bph::Fr r = f.mont(a, bph::Fr::one());  // x * 1 / 2^256
    std::memcpy(out, r.l, 32);
    return true;
}

unsigned pack_threads() {
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("BP_PACK_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    return std::max(1u, std::min(nt, 64u));
}

const char* pack_kernel_name() { return use_avx2() ? "avx2" : "portable"; }

}  // namespace bp

namespace {

// field < 0: canonical scalars; else Montgomery scalars of that field (exception values are converted to canonical)
int pack_all(int field, const uint64_t* scalars, uint64_t n, uint8_t* bits, uint64_t* exc_idx, uint64_t* exc_vals_le, uint64_t exc_cap,
             uint64_t* n_exc) {
    if (!n_exc || (n && (!scalars || !bits)) || (exc_cap && (!exc_idx || !exc_vals_le)) || field > 2) return BP_E_ARG;
    *n_exc = 0;
    if (!n) return BP_OK;
    try {
        uint64_t one[4] = {1, 0, 0, 0};
        if (field >= 0) bp::mont_one(field, one);
        const uint64_t n_bytes = (n + 7) / 8;
        const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(bp::pack_threads(), n_bytes / 4096 + 1));
        std::vector<std::vector<bp::PackExc>> exc(nt);
        std::atomic<bool> failed{false};
        auto work = [&](unsigned t) {
            try {
                const uint64_t b0 = n_bytes * t / nt, b1 = n_bytes * (t + 1) / nt;
                if (field < 0) bp::pack_bit_bytes(scalars, n, bits, b0, b1, exc[t]);
                else bp::pack_bit_bytes_mont(scalars, n, one, bits, b0, b1, exc[t]);
            } catch (...) {  // nothing may leave a thread
                failed = true;
            }
        };
        bp::run_on_threads(nt, work);
        if (failed) return BP_E_OOM;
        uint64_t k = 0;
        bool not_a_scalar = false;
        for (unsigned t = 0; t < nt; ++t)  // thread t packed a lower byte range than thread t + 1: already ascending
            for (const bp::PackExc& e : exc[t]) {
                if (k < exc_cap) {
                    exc_idx[k] = e.idx;
                    if (field < 0) std::memcpy(exc_vals_le + 4 * k, e.v, 32);
                    else if (!bp::from_mont(field, e.v, exc_vals_le + 4 * k)) not_a_scalar = true;
                }
                ++k;
            }
        *n_exc = k;
        if (not_a_scalar) return BP_E_RANGE;
        return k > exc_cap ? BP_E_RANGE : BP_OK;
    } catch (const std::bad_alloc&) {
        return BP_E_OOM;
    } catch (...) {
        return BP_E_STATE;
    }
}

}  // namespace

extern "C" {

int bp_pack_scalars(const uint64_t* scalars_le, uint64_t n, uint8_t* bits, uint64_t* exc_idx, uint64_t* exc_vals_le, uint64_t exc_cap,
                    uint64_t* n_exc) {
    return pack_all(-1, scalars_le, n, bits, exc_idx, exc_vals_le, exc_cap, n_exc);
}

int bp_pack_scalars_mont(int field, const uint64_t* scalars_mont, uint64_t n, uint8_t* bits, uint64_t* exc_idx, uint64_t* exc_vals_le,
                         uint64_t exc_cap, uint64_t* n_exc) {
    if (field < 0) return BP_E_ARG;
    return pack_all(field, scalars_mont, n, bits, exc_idx, exc_vals_le, exc_cap, n_exc);
}

int bp_scalars_from_mont(int field, const uint64_t* scalars_mont, uint64_t n, uint64_t* scalars_le) {
    if (field < 0 || field > 2 || (n && (!scalars_mont || !scalars_le))) return BP_E_ARG;
    const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(bp::pack_threads(), n / 65536 + 1));
    std::atomic<bool> bad{false};
    auto work = [&](unsigned t) {
        for (uint64_t i = n * t / nt, e = n * (t + 1) / nt; i < e; ++i)
            if (!bp::from_mont(field, scalars_mont + 4 * i, scalars_le + 4 * i)) bad = true;
    };
    try {
        bp::run_on_threads(nt, work);
    } catch (...) {
        return BP_E_OOM;
    }
    return bad ? BP_E_RANGE : BP_OK;
}

const char* bp_pack_kernel(void) { return bp::pack_kernel_name(); }

}  // extern "C"
