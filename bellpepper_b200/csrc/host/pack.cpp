// See pack.hpp.  Host only: no CUDA, no handle.
#include "pack.hpp"

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

unsigned pack_threads() {
    unsigned nt = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("BP_PACK_THREADS")) nt = (unsigned)std::max(1, std::atoi(e));
    return std::max(1u, std::min(nt, 64u));
}

const char* pack_kernel_name() { return use_avx2() ? "avx2" : "portable"; }

}  // namespace bp

extern "C" {

int bp_pack_scalars(const uint64_t* scalars_le, uint64_t n, uint8_t* bits, uint64_t* exc_idx, uint64_t* exc_vals_le, uint64_t exc_cap,
                    uint64_t* n_exc) {
    if (!n_exc || (n && (!scalars_le || !bits)) || (exc_cap && (!exc_idx || !exc_vals_le))) return BP_E_ARG;
    *n_exc = 0;
    if (!n) return BP_OK;
    try {
        const uint64_t n_bytes = (n + 7) / 8;
        const unsigned nt = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(bp::pack_threads(), n_bytes / 4096 + 1));
        std::vector<std::vector<bp::PackExc>> exc(nt);
        std::atomic<bool> failed{false};
        auto work = [&](unsigned t) {
            try {
                bp::pack_bit_bytes(scalars_le, n, bits, n_bytes * t / nt, n_bytes * (t + 1) / nt, exc[t]);
            } catch (...) {  // nothing may leave a thread
                failed = true;
            }
        };
        {
            std::vector<std::thread> th;
            for (unsigned t = 1; t < nt; ++t) th.emplace_back(work, t);
            work(0);
            for (auto& x : th) x.join();
        }
        if (failed) return BP_E_OOM;
        uint64_t k = 0;
        for (unsigned t = 0; t < nt; ++t)  // thread t packed a lower byte range than thread t + 1: already ascending
            for (const bp::PackExc& e : exc[t]) {
                if (k < exc_cap) {
                    exc_idx[k] = e.idx;
                    std::memcpy(exc_vals_le + 4 * k, e.v, 32);
                }
                ++k;
            }
        *n_exc = k;
        return k > exc_cap ? BP_E_RANGE : BP_OK;
    } catch (const std::bad_alloc&) {
        return BP_E_OOM;
    } catch (...) {
        return BP_E_STATE;
    }
}

const char* bp_pack_kernel(void) { return bp::pack_kernel_name(); }

}  // extern "C"
