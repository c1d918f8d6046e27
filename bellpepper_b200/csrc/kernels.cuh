// CUDA kernels of the R1CS evaluation engine (sm_100a).  DESIGN.md has the layout and the per-kernel rooflines.
//
//   K1 check_rows<F,false> / check_fat_rows<F,false>   which_is_unsatisfied      test_cs.rs:239-253 (+ eval_lc :137-155)
//   K2 check_rows<F,true>  / check_fat_rows<F,true>    batched LinearCombination::eval   lc.rs:245-267
//   K3 classify_lcs<F> + convert_terms<F>              canonical coefficient -> device-internal form + class
//      validate_canonical<F>                           value < p check for witness uploads
//   K5 synth_*                                         synthetic instance generator (measurement fixture)
//      eval_lc_kernel<F>                               one ad-hoc LinearCombination::eval
//
// Device layout (per handle == per row shard):
//   row_ptr : u32[3N+1]   LC offsets; LC 3i, 3i+1, 3i+2 are A_i, B_i, C_i; a row's terms are contiguous
//   cols    : u32[nnz]    bit 31 = aux index space, bits 30..28 = coefficient class, bits 27..0 = index
//   vals    : uint4[2nnz] coefficient, 8 x u32 limbs, INTERNAL form (below)
//   inputs  : uint4[2*n_inputs], aux : uint4[2*n_aux]   canonical witness
//
// Internal coefficient form.  Evaluation never reduces per term:
//   * general A/B LC  : stored = c * 2^288 mod p, class GEN.  acc (17 limbs) += stored * w ; value = redc(acc) in [0,2p)
//   * plain   A/B LC  : every coefficient is a small signed integer and sum|c| <= 7: classes P1/M1/P2/M2/PS/MS/ZERO,
//                       acc (9 limbs) += |c| * (w or p-w) ; value = acc mod p by three conditional subtractions
//   * C terms         : the NEGATED coefficient, unscaled: stored = p - c (class GEN) or its small class.
// Row check:  X = Az*Bz + sum_C (-c)*w  (unreduced, 17 limbs);  satisfied  <=>  redc(X) in {0, p}
// (the zero test is invariant under redc's 2^-288 factor).  Per row: T_gen + 1 products, <= 3 reductions.
//
// Rows with more than `fat_terms` terms (MultiEq rows, num.rs unpacking rows) are skipped by the thread-per-row
// kernel and done by check_fat_rows, one warp per row: lanes stride the terms (coalesced 1 KB segments), partial
// sums are combined with a 17-limb xor-shuffle butterfly.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "field.cuh"
#include "synth.cuh"

namespace bp {

struct FieldConsts {
    uint32_t k288m[8];  // 2^544 mod p : mont_mul(c, k288m) = c * 2^288
    uint32_t k576[8];   // 2^576 mod p : redc(u * k576) = u * 2^288   (emit mode: undo redc's scaling on C)
};

struct CsrView {
    const uint32_t* __restrict__ row_ptr;
    const uint32_t* __restrict__ cols;
    const uint4* __restrict__ vals;
    const uint4* __restrict__ inputs;
    const uint4* __restrict__ aux;
    uint32_t n_rows;
    uint32_t n_inputs;
    uint32_t n_aux;
    uint32_t fat_terms;  // rows with more terms than this belong to check_fat_rows
    unsigned long long row_base;
};

struct CheckOut {
    long long* first_bad;  // global row index, atomicMin; INT64_MAX = satisfied
    unsigned int* err;     // bit 0: a column was out of range
    uint4* az;             // emit mode only (nullable)
    uint4* bz;
    uint4* cz;
};

// Plan: what kind of row is it (one byte per row, built by classify_rows)?
enum RowKind : uint8_t {
    kRowGeneric = 0,   // check_rows (full arithmetic)
    kRowPlain = 1,     // every term is a small-coefficient class and sum|c| of C is <= 5: check_rows_plain can decide it
    kRowFat = 2,       // check_fat_rows
    kRowDeferred = 3   // a plain row whose VALUES needed the full product this time: check_rows picks it up and resets it
};

__device__ __forceinline__ void ld8(uint32_t* x, const uint4* p) {
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    x[0] = lo.x; x[1] = lo.y; x[2] = lo.z; x[3] = lo.w;
    x[4] = hi.x; x[5] = hi.y; x[6] = hi.z; x[7] = hi.w;
}
__device__ __forceinline__ void st8(uint4* p, const uint32_t* x) {
    p[0] = make_uint4(x[0], x[1], x[2], x[3]);
    p[1] = make_uint4(x[4], x[5], x[6], x[7]);
}
template <int N> __device__ __forceinline__ void zeron(uint32_t* a) {
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = 0;
}

// The latency-critical half of a term -- its class and the gathered witness element -- is fetched one term ahead of
// the arithmetic; the coefficient (sequential, cache-friendly) is read when the term is folded.
struct TermW {
    uint32_t cls;
    uint32_t w[8];
};

__device__ __forceinline__ void load_term(TermW& t, uint32_t k, bool valid, const CsrView& m, unsigned int& err) {
    t.cls = kClsZero;
    if (!valid) return;
    const uint32_t col = __ldg(m.cols + k);
    const uint32_t cls = (col >> kColClsShift) & 7u;
    if (cls == kClsZero) return;
    const uint32_t idx = col & kColIdxMask;
    const bool is_aux = (col & kColAux) != 0;
    if (idx >= (is_aux ? m.n_aux : m.n_inputs)) { err = 1; return; }
    t.cls = cls;
    ld8(t.w, (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx);
}

// Fold term k (already loaded into t) into acc.  RIPPLE = 9: A/B accumulator (plain sums stay below 2^288); 17: the
// row-check accumulator.  `gen` is set when a full product was folded; `mag` accumulates the plain magnitudes.
template <int F, int RIPPLE>
__device__ __forceinline__ void apply_term(uint32_t* acc /*17*/, TermW& t, uint32_t k, const CsrView& m, uint32_t& gen, uint32_t& mag) {
    const uint32_t cls = t.cls;
    if (cls == kClsZero) return;
    if (cls == kClsGen) {
        uint32_t c[8];
        ld8(c, m.vals + 2 * (size_t)k);
        mac_wide(acc, c, t.w);
        gen = 1;
        return;
    }
    if (cls == kClsM1 || cls == kClsM2 || cls == kClsMS) {
        uint32_t n[8];
        neg_mod<F>(n, t.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) t.w[i] = n[i];
    }
    if (cls == kClsP1 || cls == kClsM1) {
        acc_add8<RIPPLE>(acc, t.w);
        mag += 1;
    } else if (cls == kClsP2 || cls == kClsM2) {
        acc_add8<RIPPLE>(acc, t.w);
        acc_add8<RIPPLE>(acc, t.w);
        mag += 2;
    } else {
        const uint32_t sm = __ldg(reinterpret_cast<const uint32_t*>(m.vals + 2 * (size_t)k));
        acc_mad_small<RIPPLE>(acc, t.w, sm);
        mag += sm > 8u ? 8u : sm;
    }
}

// Kernel feature bits (template parameter V), kept switchable so that each can be measured on the GPU (DESIGN.md "Variants").
constexpr int kVPipe = 1;      // fetch the next term's column + witness while the current term is folded
constexpr int kVMagSkip = 2;   // plain A/B sums below 2p are used unreduced
constexpr int kVBitRow = 4;    // rows whose Az or Bz is 0/1 (and C is plain) are decided without any multiplication
constexpr int kVPark = 8;      // az / bz wait in shared memory while C is folded (fewer live registers)
constexpr int kVPrefetch = 16; // L1 prefetch of every witness line of the row before the first term is folded

// Terms k0, k0+STEP, ... < k1.  STEP = 1 for the thread-per-row kernel, 32 for a lane of the warp-per-row kernel.
template <int F, int RIPPLE, uint32_t STEP, bool PIPE>
__device__ __forceinline__ void fold_range(uint32_t* acc, uint32_t k0, uint32_t k1, const CsrView& m, unsigned int& err, uint32_t& gen,
                                           uint32_t& mag) {
    if (k0 >= k1) return;
    if (PIPE) {
        TermW cur;
        load_term(cur, k0, true, m, err);
#pragma unroll 1
        for (uint32_t k = k0; k < k1; k += STEP) {
            TermW nxt;
            load_term(nxt, k + STEP, k + STEP < k1, m, err);
            apply_term<F, RIPPLE>(acc, cur, k, m, gen, mag);
            cur = nxt;
        }
    } else {
#pragma unroll 1
        for (uint32_t k = k0; k < k1; k += STEP) {
            TermW cur;
            load_term(cur, k, true, m, err);
            apply_term<F, RIPPLE>(acc, cur, k, m, gen, mag);
        }
    }
}

// L1 prefetch of the witness elements the terms [k0,k1) will gather (one 32-byte element never straddles a line).
__device__ __forceinline__ void prefetch_witness(uint32_t k0, uint32_t k1, const CsrView& m) {
#pragma unroll 1
    for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t col = __ldg(m.cols + k);
        const uint32_t idx = col & kColIdxMask;
        const bool is_aux = (col & kColAux) != 0;
        if (((col >> kColClsShift) & 7u) != kClsZero && idx < (is_aux ? m.n_aux : m.n_inputs)) {
            const uint4* a = (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
        }
    }
}

// acc (17 limbs, holding one A/B LC) -> its value, 8 limbs, < 2^256 and = the LC mod p.
// general: redc -> [0,2p).  plain: the sum is < (mag+..)*p; up to 2p it already fits 8 limbs and is used as is
// (the row check only needs the value mod p and < 2^256); otherwise three conditional subtractions -> [0,p).
template <int F, bool MAGSKIP> __device__ __forceinline__ void finish_ab(uint32_t* out, const uint32_t* acc, uint32_t any_gen, uint32_t mag) {
    if (any_gen) {
        redc_acc<F>(out, acc);
    } else if (MAGSKIP && mag <= 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = acc[i];
    } else {
        reduce_8p<F>(out, acc);
    }
}

// Canonical form of an A/B value for emit mode: plain sums may sit anywhere in [0, 2p] (e.g. 2*(p-0)), redc outputs
// in [0, 2p); three conditional subtractions cover both.
template <int F> __device__ __forceinline__ void canon_ab(uint32_t* v /*8, in/out*/) {
    uint32_t t[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = v[i];
    t[8] = 0;
    reduce_8p<F>(v, t);
}

// 8-limb value <= 1 ?  (returns 0, 1, or 2 for "something else")
__device__ __forceinline__ uint32_t small01(const uint32_t* v) {
    uint32_t hi = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) hi |= v[i];
    return (hi == 0 && v[0] <= 1u) ? v[0] : 2u;
}

// The row test.  acc holds sum_C (-c)*w.  General path: acc += Az*Bz, one lazy reduction, zero test.
// Shortcut (bit-valued rows, i.e. nearly every row of a boolean-gadget circuit with an honest witness): when Az or Bz
// is 0 or 1 and C had no full products, Az*Bz is 0 or the other operand and the whole left-hand side stays below 8p,
// so  == 0 (mod p)  is decided with additions and three conditional subtractions -- no multiplication at all.
// The shortcut is taken on VALUES, so any witness gives the same verdict as the general path.
template <int F, bool BITROW>
__device__ __forceinline__ bool row_satisfied(uint32_t* acc /*17*/, const uint32_t* az, const uint32_t* bz, uint32_t gen_c, uint32_t mag_c) {
    const uint32_t sa = BITROW ? small01(az) : 2u, sb = BITROW ? small01(bz) : 2u;
    if (BITROW && !gen_c && mag_c <= 5u && (sa < 2u || sb < 2u)) {
        uint32_t prod[8];
        const bool zero = (sa == 0u) || (sb == 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i) prod[i] = zero ? 0u : (sa == 1u ? bz[i] : az[i]);
        acc_add8<9>(acc, prod);  // < 5p + 2p
        uint32_t r[8], nz = 0;
        reduce_8p<F>(r, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) nz |= r[i];
        return nz == 0;
    }
    uint32_t y[8];
    mac_wide(acc, az, bz);
    redc_acc<F>(y, acc);
    return is_zero_mod_p<F>(y);
}


// Canonical C.w from the unreduced negated sum Xc:  -(Xc mod p)
template <int F> __device__ __forceinline__ void finish_c_canonical(uint32_t* out, const uint32_t* xc /*17*/, const FieldConsts& fc) {
    uint32_t u[8], t[17], k[8];
    redc_acc<F>(u, xc);  // Xc * 2^-288
#pragma unroll
    for (int i = 0; i < 8; ++i) k[i] = fc.k576[i];
    mul_wide(t, u, k);
    t[16] = 0;
    redc_acc<F>(u, t);  // Xc mod p, in [0, 2p)
    reduce_once<F>(u);
    uint32_t n[8], nz = 0;
    neg_mod<F>(n, u);
#pragma unroll
    for (int i = 0; i < 8; ++i) nz |= u[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = nz ? n[i] : 0u;
}

// az / bz are parked in shared memory while the C terms are folded (their 16 registers are what pushes the kernel over
// the 5-blocks-per-SM register budget); column-major so that a warp's accesses are conflict-free.
__device__ __forceinline__ void park8(uint32_t (*slot)[128], const uint32_t* v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) slot[i][threadIdx.x] = v[i];
}
__device__ __forceinline__ void unpark8(uint32_t* v, uint32_t (*slot)[128]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = slot[i][threadIdx.x];
}

// Block-level min of per-thread candidate rows -> one atomicMin per block.
__device__ __forceinline__ void publish_first_bad(uint32_t my_bad, const CsrView& m, const CheckOut& o, unsigned int my_err) {
    __shared__ unsigned int s_min;
    __shared__ unsigned int s_err;
    if (threadIdx.x == 0) { s_min = 0xffffffffu; s_err = 0; }
    __syncthreads();
    const unsigned int wmin = __reduce_min_sync(0xffffffffu, my_bad);
    const unsigned int werr = __reduce_or_sync(0xffffffffu, my_err);
    if ((threadIdx.x & 31) == 0) {
        if (wmin != 0xffffffffu) atomicMin(&s_min, wmin);
        if (werr) atomicOr(&s_err, werr);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_min != 0xffffffffu) atomicMin(o.first_bad, (long long)(m.row_base + s_min));
        if (s_err) atomicOr(o.err, s_err);
    }
}

// ---- K1 / K2, thin rows: one thread per constraint -------------------------------------------------------------------
template <int F, bool EMIT, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_rows(CsrView m, CheckOut o, FieldConsts fc, uint8_t* __restrict__ kind = nullptr) {
    constexpr bool PIPE = (V & kVPipe) != 0, PARK = (V & kVPark) != 0;
    __shared__ uint32_t s_az[PARK ? 8 : 1][128], s_bz[PARK ? 8 : 1][128];
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < m.n_rows; row += gridDim.x * blockDim.x) {
        if (kind) {  // rows already decided by check_rows_plain (or owned by check_fat_rows) are not ours
            const uint8_t kd = kind[row];
            if (kd == kRowPlain || kd == kRowFat) continue;
            if (kd == kRowDeferred) kind[row] = kRowPlain;  // picked up: back to its static kind for the next check
        }
        const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                       p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        if (p3 - p0 > m.fat_terms) continue;  // check_fat_rows does it
        if (V & kVPrefetch) prefetch_witness(p0, p3, m);
        uint32_t acc[17], az[8], bz[8];
        uint32_t gen = 0, mag = 0;
        zeron<17>(acc);
        fold_range<F, 9, 1, PIPE>(acc, p0, p1, m, my_err, gen, mag);
        finish_ab<F, (V & kVMagSkip) != 0>(az, acc, gen, mag);
        if (PARK) park8(s_az, az);
        gen = 0; mag = 0;
        zeron<17>(acc);
        fold_range<F, 9, 1, PIPE>(acc, p1, p2, m, my_err, gen, mag);
        finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, gen, mag);
        if (PARK) park8(s_bz, bz);
        gen = 0; mag = 0;
        zeron<17>(acc);
        fold_range<F, 17, 1, PIPE>(acc, p2, p3, m, my_err, gen, mag);
        if (PARK) {
            unpark8(az, s_az);
            unpark8(bz, s_bz);
        }
        if (EMIT) {
            uint32_t v[8];
            if (o.cz) {
                finish_c_canonical<F>(v, acc, fc);
                st8(o.cz + 2 * (size_t)row, v);
            }
            if (o.az) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = az[i];
                canon_ab<F>(v);
                st8(o.az + 2 * (size_t)row, v);
            }
            if (o.bz) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = bz[i];
                canon_ab<F>(v);
                st8(o.bz + 2 * (size_t)row, v);
            }
        }
        if (!row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, gen, mag) && row < my_bad) my_bad = row;
    }
    publish_first_bad(my_bad, m, o, my_err);
}

// ---- plan: what kind of row is it? -------------------------------------------------------------------------------------

__global__ void classify_rows(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ cols, const uint4* __restrict__ vals,
                              uint32_t n_rows, uint32_t fat_terms, uint8_t* __restrict__ kind) {
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        const uint32_t p0 = row_ptr[3 * (size_t)row], p2 = row_ptr[3 * (size_t)row + 2], p3 = row_ptr[3 * (size_t)row + 3];
        uint8_t k = kRowPlain;
        if (p3 - p0 > fat_terms) {
            k = kRowFat;
        } else {
            uint32_t mag_c = 0;
            for (uint32_t t = p0; t < p3 && k == kRowPlain; ++t) {
                const uint32_t cls = (__ldg(cols + t) >> kColClsShift) & 7u;
                if (cls == kClsGen) k = kRowGeneric;
                if (t >= p2) {
                    if (cls == kClsP1 || cls == kClsM1) mag_c += 1;
                    else if (cls == kClsP2 || cls == kClsM2) mag_c += 2;
                    else if (cls == kClsPS || cls == kClsMS) {
                        const uint32_t sm = __ldg(reinterpret_cast<const uint32_t*>(vals + 2 * (size_t)t));
                        mag_c += sm > 8u ? 8u : sm;
                    }
                    if (mag_c > 5u) k = kRowGeneric;
                }
            }
        }
        kind[row] = k;
    }
}

// Fold the plain terms [k0,k1) into a 9-limb accumulator (sum < 8p by construction of the row kind).
template <int F>
__device__ __forceinline__ void fold_plain(uint32_t* acc /*9*/, uint32_t k0, uint32_t k1, const CsrView& m, unsigned int& err, uint32_t& mag) {
#pragma unroll 1
    for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t col = __ldg(m.cols + k);
        const uint32_t cls = (col >> kColClsShift) & 7u;
        if (cls == kClsZero) continue;
        const uint32_t idx = col & kColIdxMask;
        const bool is_aux = (col & kColAux) != 0;
        if (idx >= (is_aux ? m.n_aux : m.n_inputs)) { err = 1; continue; }
        uint32_t w[8];
        ld8(w, (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx);
        if (cls == kClsM1 || cls == kClsM2 || cls == kClsMS) {
            uint32_t n[8];
            neg_mod<F>(n, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = n[i];
        }
        if (cls == kClsP1 || cls == kClsM1) {
            acc_add8<9>(acc, w);
            mag += 1;
        } else if (cls == kClsP2 || cls == kClsM2) {
            acc_add8<9>(acc, w);
            acc_add8<9>(acc, w);
            mag += 2;
        } else {
            const uint32_t sm = __ldg(reinterpret_cast<const uint32_t*>(m.vals + 2 * (size_t)k));
            acc_mad_small<9>(acc, w, sm);
            mag += sm > 8u ? 8u : sm;
        }
    }
}

template <int F> __device__ __forceinline__ void finish_plain(uint32_t* out /*8*/, const uint32_t* acc /*9*/, uint32_t mag) {
    if (mag <= 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = acc[i];
    } else {
        reduce_8p<F>(out, acc);
    }
}

// ---- K1, plain rows: additions only.  Small register footprint -> 8 CTAs/SM; rows whose values need the full product
// (neither Az nor Bz is 0/1) are handed to check_rows through the kind byte.
template <int F, int MB>
__global__ void __launch_bounds__(128, MB) check_rows_plain(CsrView m, CheckOut o, uint8_t* __restrict__ kind) {
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < m.n_rows; row += gridDim.x * blockDim.x) {
        if (kind[row] != kRowPlain) continue;
        const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                       p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        uint32_t acc[9], az[8], bz[8];
        uint32_t mag = 0;
        zeron<9>(acc);
        fold_plain<F>(acc, p0, p1, m, my_err, mag);
        finish_plain<F>(az, acc, mag);
        mag = 0;
        zeron<9>(acc);
        fold_plain<F>(acc, p1, p2, m, my_err, mag);
        finish_plain<F>(bz, acc, mag);
        const uint32_t sa = small01(az), sb = small01(bz);
        if (sa == 2u && sb == 2u) {  // needs Az*Bz: not ours
            kind[row] = kRowDeferred;
            continue;
        }
        mag = 0;
        zeron<9>(acc);
        fold_plain<F>(acc, p2, p3, m, my_err, mag);  // < 5p by the row kind
        const bool zero = (sa == 0u) || (sb == 0u);
        uint32_t prod[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) prod[i] = zero ? 0u : (sa == 1u ? bz[i] : az[i]);
        acc_add8<9>(acc, prod);  // < 5p + 2p
        uint32_t r[8], nz = 0;
        reduce_8p<F>(r, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) nz |= r[i];
        if (nz != 0 && row < my_bad) my_bad = row;
    }
    publish_first_bad(my_bad, m, o, my_err);
}

// 17-limb sum across the warp; every lane ends with the total.
__device__ __forceinline__ void warp_sum17(uint32_t* acc) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        uint32_t t[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) t[i] = __shfl_xor_sync(0xffffffffu, acc[i], off);
        (void)addn<17>(acc, acc, t);
    }
}

// One LC [k0,k1) by a whole warp: lanes stride the terms (two terms' loads in flight per lane), then the partial sums
// are combined.  Every lane returns with the total, the "a full product was folded" flag and the plain magnitude.
template <int F, int RIPPLE, bool PIPE>
__device__ __forceinline__ void warp_fold_lc(uint32_t* acc, uint32_t k0, uint32_t k1, const CsrView& m, unsigned int& err, uint32_t& any_gen,
                                             uint32_t& mag) {
    const uint32_t lane = threadIdx.x & 31u;
    zeron<17>(acc);
    uint32_t g = 0, mg = 0;
    if (k1 - k0 <= 4) {  // short LC: lane 0 alone
        if (lane == 0) fold_range<F, RIPPLE, 1, false>(acc, k0, k1, m, err, g, mg);
    } else {
        fold_range<F, 17, 32, PIPE>(acc, k0 + lane, k1, m, err, g, mg);
    }
    warp_sum17(acc);
    any_gen = __any_sync(0xffffffffu, g != 0) ? 1u : 0u;
    mag = __reduce_add_sync(0xffffffffu, mg > 8u ? 8u : mg);
}

// ---- K1 / K2, fat rows: one warp per constraint -------------------------------------------------------------------------
template <int F, bool EMIT, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_fat_rows(CsrView m, CheckOut o, FieldConsts fc, const uint32_t* __restrict__ fat_rows,
                                                          const uint32_t* __restrict__ n_fat) {
    constexpr bool PIPE = (V & kVPipe) != 0, PARK = (V & kVPark) != 0;
    __shared__ uint32_t s_az[PARK ? 8 : 1][128], s_bz[PARK ? 8 : 1][128];
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    const uint32_t n = *n_fat;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const uint32_t row = fat_rows[i];
        const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                       p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        uint32_t acc[17], az[8], bz[8];
        uint32_t any_gen, mag;
        warp_fold_lc<F, 9, PIPE>(acc, p0, p1, m, my_err, any_gen, mag);
        finish_ab<F, (V & kVMagSkip) != 0>(az, acc, any_gen, mag);
        if (PARK) park8(s_az, az);
        warp_fold_lc<F, 9, PIPE>(acc, p1, p2, m, my_err, any_gen, mag);
        finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, any_gen, mag);
        if (PARK) park8(s_bz, bz);
        warp_fold_lc<F, 17, PIPE>(acc, p2, p3, m, my_err, any_gen, mag);
        if (PARK) {
            unpark8(az, s_az);
            unpark8(bz, s_bz);
        }
        if (EMIT && (threadIdx.x & 31u) == 0) {
            uint32_t v[8];
            if (o.cz) {
                finish_c_canonical<F>(v, acc, fc);
                st8(o.cz + 2 * (size_t)row, v);
            }
            if (o.az) {
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2) v[i2] = az[i2];
                canon_ab<F>(v);
                st8(o.az + 2 * (size_t)row, v);
            }
            if (o.bz) {
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2) v[i2] = bz[i2];
                canon_ab<F>(v);
                st8(o.bz + 2 * (size_t)row, v);
            }
        }
        if (!row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, any_gen, mag) && row < my_bad) my_bad = row;
    }
    publish_first_bad(my_bad, m, o, my_err);
}

// One fat row by the whole warp, outlined so that its register needs do not leak into the thin path's allocation.
// Returns (to every lane) whether the row is satisfied; emits canonical Az/Bz/Cz from lane 0 when asked to.
template <int F, bool EMIT, int V>
__device__ __noinline__ bool fat_row_by_warp(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3, uint32_t frow, const CsrView& m,
                                             const CheckOut& o, const FieldConsts& fc, unsigned int* err_out) {
    unsigned int err = 0;
    uint32_t acc[17], az[8], bz[8];
    uint32_t any_gen, mag;
    warp_fold_lc<F, 9, false>(acc, q0, q1, m, err, any_gen, mag);
    finish_ab<F, (V & kVMagSkip) != 0>(az, acc, any_gen, mag);
    warp_fold_lc<F, 9, false>(acc, q1, q2, m, err, any_gen, mag);
    finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, any_gen, mag);
    warp_fold_lc<F, 17, false>(acc, q2, q3, m, err, any_gen, mag);
    if (EMIT && (threadIdx.x & 31u) == 0) {
        uint32_t v[8];
        if (o.cz) {
            finish_c_canonical<F>(v, acc, fc);
            st8(o.cz + 2 * (size_t)frow, v);
        }
        if (o.az) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = az[i];
            canon_ab<F>(v);
            st8(o.az + 2 * (size_t)frow, v);
        }
        if (o.bz) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = bz[i];
            canon_ab<F>(v);
            st8(o.bz + 2 * (size_t)frow, v);
        }
    }
    *err_out |= err;
    return row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, any_gen, mag);
}

// One thin row by its own thread, outlined for the same reason as fat_row_by_warp.  Returns "satisfied".
template <int F, bool EMIT, int V>
__device__ __noinline__ bool thin_row_by_thread(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t row, const CsrView& m,
                                                const CheckOut& o, const FieldConsts& fc, unsigned int* err_out) {
    unsigned int err = 0;
    uint32_t acc[17], az[8], bz[8];
    uint32_t gen = 0, mag = 0;
    zeron<17>(acc);
    fold_range<F, 9, 1, false>(acc, p0, p1, m, err, gen, mag);
    finish_ab<F, (V & kVMagSkip) != 0>(az, acc, gen, mag);
    gen = 0; mag = 0;
    zeron<17>(acc);
    fold_range<F, 9, 1, false>(acc, p1, p2, m, err, gen, mag);
    finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, gen, mag);
    gen = 0; mag = 0;
    zeron<17>(acc);
    fold_range<F, 17, 1, false>(acc, p2, p3, m, err, gen, mag);
    if (EMIT) {
        uint32_t v[8];
        if (o.cz) {
            finish_c_canonical<F>(v, acc, fc);
            st8(o.cz + 2 * (size_t)row, v);
        }
        if (o.az) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = az[i];
            canon_ab<F>(v);
            st8(o.az + 2 * (size_t)row, v);
        }
        if (o.bz) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = bz[i];
            canon_ab<F>(v);
            st8(o.bz + 2 * (size_t)row, v);
        }
    }
    *err_out |= err;
    return row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, gen, mag);
}

// ---- K1 / K2, fused: thin rows by their own thread, fat rows by the whole warp, in the SAME pass over the rows --------
// A separate fat-row kernel gathers from a witness region that has long left the caches (measured: L2 hit 8 %, and it
// runs at the HBM random-access rate); handled in place, a fat row finds its operands in L2/L1 because the neighbouring
// thin rows have just touched the same variables.  A warp owns 32 consecutive rows: each lane evaluates its row if it is
// thin; then the lanes that hold a fat row are served one after the other by all 32 lanes together.
template <int F, bool EMIT, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_rows_fused(CsrView m, CheckOut o, FieldConsts fc) {
    constexpr bool PARK = (V & kVPark) != 0;
    __shared__ uint32_t s_az[PARK ? 8 : 1][128], s_bz[PARK ? 8 : 1][128];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    const uint32_t n_blocks = (m.n_rows + 31u) / 32u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < n_blocks; blk += n_warps) {
        const uint32_t row = blk * 32u + lane;
        const bool live = row < m.n_rows;
        uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
        if (live) {
            p0 = __ldg(m.row_ptr + 3 * (size_t)row);
            p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1);
            p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2);
            p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        }
        const bool fat = live && (p3 - p0 > m.fat_terms);
        if (live && !fat) {
            uint32_t acc[17], az[8], bz[8];
            uint32_t gen = 0, mag = 0;
            zeron<17>(acc);
            fold_range<F, 9, 1, false>(acc, p0, p1, m, my_err, gen, mag);
            finish_ab<F, (V & kVMagSkip) != 0>(az, acc, gen, mag);
            if (PARK) park8(s_az, az);
            gen = 0; mag = 0;
            zeron<17>(acc);
            fold_range<F, 9, 1, false>(acc, p1, p2, m, my_err, gen, mag);
            finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, gen, mag);
            if (PARK) park8(s_bz, bz);
            gen = 0; mag = 0;
            zeron<17>(acc);
            fold_range<F, 17, 1, false>(acc, p2, p3, m, my_err, gen, mag);
            if (PARK) {
                unpark8(az, s_az);
                unpark8(bz, s_bz);
            }
            if (EMIT) {
                uint32_t v[8];
                if (o.cz) {
                    finish_c_canonical<F>(v, acc, fc);
                    st8(o.cz + 2 * (size_t)row, v);
                }
                if (o.az) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = az[i];
                    canon_ab<F>(v);
                    st8(o.az + 2 * (size_t)row, v);
                }
                if (o.bz) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = bz[i];
                    canon_ab<F>(v);
                    st8(o.bz + 2 * (size_t)row, v);
                }
            }
            if (!row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, gen, mag) && row < my_bad) my_bad = row;
        }
        // fat rows of this block: the whole warp serves them one at a time
        uint32_t fat_mask = __ballot_sync(0xffffffffu, fat);
        while (fat_mask) {
            const int src = __ffs(fat_mask) - 1;
            fat_mask &= fat_mask - 1;
            const uint32_t q0 = __shfl_sync(0xffffffffu, p0, src), q1 = __shfl_sync(0xffffffffu, p1, src),
                           q2 = __shfl_sync(0xffffffffu, p2, src), q3 = __shfl_sync(0xffffffffu, p3, src);
            const uint32_t frow = blk * 32u + (uint32_t)src;
            if (!fat_row_by_warp<F, EMIT, V>(q0, q1, q2, q3, frow, m, o, fc, &my_err) && frow < my_bad) my_bad = frow;
        }
    }
    publish_first_bad(my_bad, m, o, my_err);
}

// Same as check_rows_fused with BOTH paths outlined: the kernel body is only the dispatcher.
template <int F, bool EMIT, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_rows_fused2(CsrView m, CheckOut o, FieldConsts fc) {
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_blocks = (m.n_rows + 31u) / 32u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < n_blocks; blk += n_warps) {
        const uint32_t row = blk * 32u + lane;
        const bool live = row < m.n_rows;
        uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
        if (live) {
            p0 = __ldg(m.row_ptr + 3 * (size_t)row);
            p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1);
            p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2);
            p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        }
        const bool fat = live && (p3 - p0 > m.fat_terms);
        if (live && !fat) {
            if (!thin_row_by_thread<F, EMIT, V>(p0, p1, p2, p3, row, m, o, fc, &my_err) && row < my_bad) my_bad = row;
        }
        uint32_t fat_mask = __ballot_sync(0xffffffffu, fat);
        while (fat_mask) {
            const int src = __ffs(fat_mask) - 1;
            fat_mask &= fat_mask - 1;
            const uint32_t q0 = __shfl_sync(0xffffffffu, p0, src), q1 = __shfl_sync(0xffffffffu, p1, src),
                           q2 = __shfl_sync(0xffffffffu, p2, src), q3 = __shfl_sync(0xffffffffu, p3, src);
            const uint32_t frow = blk * 32u + (uint32_t)src;
            if (!fat_row_by_warp<F, EMIT, V>(q0, q1, q2, q3, frow, m, o, fc, &my_err) && frow < my_bad) my_bad = frow;
        }
    }
    publish_first_bad(my_bad, m, o, my_err);
}

// List the rows that belong to check_fat_rows (order is irrelevant: the result is a minimum).
__global__ void collect_fat_rows(const uint32_t* __restrict__ row_ptr, uint32_t n_rows, uint32_t fat_terms, uint32_t* fat_rows,
                                 uint32_t* n_fat) {
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        if (row_ptr[3 * (size_t)row + 3] - row_ptr[3 * (size_t)row] > fat_terms) fat_rows[atomicAdd(n_fat, 1u)] = row;
    }
}

__global__ void init_result(long long* first_bad, unsigned int* err) {
    *first_bad = 0x7fffffffffffffffLL;
    *err = 0;
}

// ---- K3: ingest conversion ------------------------------------------------------------------------------------------------
// Pass 1, one thread per LC of the chunk: an A/B LC is "plain" when every coefficient is a small signed integer and the
// magnitudes sum to <= 7 (so its value stays below 8p); C LCs are classed per term.  kind: 0 = general, 1 = plain.
template <int F>
__global__ void classify_lcs(const uint4* __restrict__ vals, const uint32_t* __restrict__ row_ptr, uint32_t lc0, uint32_t n_lc,
                             uint8_t* __restrict__ kind) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lc; i += gridDim.x * blockDim.x) {
        const uint32_t lc = lc0 + i;
        if (lc % 3u == 2u) { kind[i] = 1; continue; }
        const uint32_t k0 = row_ptr[lc], k1 = row_ptr[lc + 1];
        uint32_t plain = 1, mag = 0;
        if (k1 - k0 > 64) plain = 0;  // can only be plain if almost all coefficients are zero: not worth the scan
        for (uint32_t k = k0; plain && k < k1; ++k) {
            uint32_t c[8], s;
            ld8(c, vals + 2 * (size_t)k);
            const uint32_t cls = classify_coeff<F>(c, &s);
            if (cls == kClsGen) plain = 0;
            else if (cls == kClsP1 || cls == kClsM1) mag += 1;
            else if (cls == kClsP2 || cls == kClsM2) mag += 2;
            else if (cls == kClsPS || cls == kClsMS) mag = s > 7 ? 8 : mag + s;
            if (mag > 7) plain = 0;
        }
        kind[i] = (uint8_t)plain;
    }
}

// Pass 2, one thread per term: canonical -> internal form in place, class bits into the column word.
// err bit 1: coefficient >= p; bit 2: variable index does not fit 28 bits.
template <int F>
__global__ void convert_terms(uint4* vals, uint32_t* cols, const uint32_t* __restrict__ row_ptr, const uint8_t* __restrict__ kind,
                              uint32_t lc0, uint32_t n_lc, uint32_t k0, uint32_t n, FieldConsts fc, unsigned int* err) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = k0 + i;
        uint32_t lo = lc0, hi = lc0 + n_lc;  // largest lc with row_ptr[lc] <= k
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (__ldg(row_ptr + mid) <= k) lo = mid; else hi = mid;
        }
        const uint32_t type = lo % 3u;
        const bool plain = kind[lo - lc0] != 0;
        uint32_t c[8], r[8], s = 0, cls;
        ld8(c, vals + 2 * (size_t)k);
        if (!is_canonical<F>(c)) atomicOr(err, 2u);
        const uint32_t col = cols[k];
        if ((col & 0x7fffffffu) > kColIdxMask) atomicOr(err, 4u);
        if (type == 2u) {  // C: the negated coefficient, unscaled
            uint32_t nz = 0, n8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) nz |= c[j];
            neg_mod<F>(n8, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = nz ? n8[j] : 0u;
        }
        cls = classify_coeff<F>(c, &s);
        if (type != 2u && !plain && cls != kClsZero) cls = kClsGen;
        if (cls == kClsGen) {
            if (type == 2u) {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = c[j];
            } else {
                uint32_t kk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) kk[j] = fc.k288m[j];
                mont_mul<F>(r, c, kk);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = 0;
            r[0] = s;
        }
        st8(vals + 2 * (size_t)k, r);
        cols[k] = (col & (kColAux | kColIdxMask)) | (cls << kColClsShift);
        // instance statistic for the launch heuristic: how many terms need a full product (err[1] is the counter)
        const unsigned int gen_mask = __ballot_sync(__activemask(), cls == kClsGen);
        if (cls == kClsGen && (threadIdx.x & 31u) == (unsigned)(__ffs(gen_mask) - 1)) atomicAdd(err + 1, (unsigned int)__popc(gen_mask));
    }
}

template <int F> __global__ void validate_canonical(const uint4* __restrict__ v, uint64_t n, unsigned int* err) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x[8];
        ld8(x, v + 2 * i);
        if (!is_canonical<F>(x)) atomicOr(err, 2u);
    }
}

// row_ptr[i] += base for i in [0, n)   (after an exclusive scan of the chunk's lens)
__global__ void add_base(uint32_t* row_ptr, uint32_t n, uint32_t base) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row_ptr[i] += base;
}

// ---- one ad-hoc LC (lc.rs:245-267): a single warp; coefficients arrive canonical and are scaled on the fly -----------------
template <int F>
__global__ void eval_lc_kernel(const uint32_t* __restrict__ cols, const uint4* __restrict__ coeffs_canonical, uint32_t n, CsrView m,
                               FieldConsts fc, uint4* out, unsigned int* err) {
    uint32_t acc[17], kk[8];
    zeron<17>(acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) kk[j] = fc.k288m[j];
    unsigned int my_err = 0;
    for (uint32_t k = threadIdx.x; k < n; k += 32) {
        const uint32_t col = cols[k], idx = col & 0x7fffffffu;
        const bool is_aux = (col >> 31) != 0;
        if (idx >= (is_aux ? m.n_aux : m.n_inputs)) { my_err |= 1; continue; }
        uint32_t c[8], cm[8], w[8];
        ld8(c, coeffs_canonical + 2 * (size_t)k);
        if (!is_canonical<F>(c)) { my_err |= 2; continue; }
        mont_mul<F>(cm, c, kk);
        ld8(w, (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx);
        mac_wide(acc, cm, w);
    }
    warp_sum17(acc);
    my_err = __reduce_or_sync(0xffffffffu, my_err);
    if (threadIdx.x == 0) {
        if (my_err) atomicOr(err, my_err);
        uint32_t v[8];
        redc_acc<F>(v, acc);
        reduce_once<F>(v);
        st8(out, v);
    }
}

// ---- K5: synthetic generator ---------------------------------------------------------------------------------------------
__global__ void synth_lens(uint32_t* lens, uint64_t seed, uint32_t t, uint64_t lc0, uint32_t n_lc) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lc; i += gridDim.x * blockDim.x)
        lens[i] = sm_len(seed, t, lc0 + i);
}

// One thread per LC: writes its columns (class GEN) and INTERNAL-form coefficients.
template <int F>
__global__ void synth_fill(uint32_t* cols, uint4* vals, const uint32_t* __restrict__ row_ptr, uint32_t lc_first /*index into row_ptr*/,
                           uint32_t n_lc, uint64_t seed, uint64_t lcid0 /*global LC id of lc_first*/, uint64_t n_vars,
                           uint64_t n_inputs, FieldConsts fc) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lc; i += gridDim.x * blockDim.x) {
        const uint32_t k0 = row_ptr[lc_first + i], len = row_ptr[lc_first + i + 1] - k0;
        const uint64_t lcid = lcid0 + i;
        const uint32_t type = (uint32_t)(lcid % 3ull);
        uint32_t kk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) kk[j] = fc.k288m[j];
        for (uint32_t k = 0; k < len; ++k) {
            cols[k0 + k] = sm_col(seed, lcid, k, len, n_vars, n_inputs);  // class bits 0 = GEN
            uint32_t c[8], r[8];
            sm_sample<F>(sm_key(seed, 3, lcid, k), c);
            if (type == 2u) {
                uint32_t nz = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) nz |= c[j];
                neg_mod<F>(r, c);
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = nz ? r[j] : 0u;
            } else {
                mont_mul<F>(r, c, kk);
            }
            st8(vals + 2 * (size_t)(k0 + k), r);
        }
    }
}

template <int F>
__global__ void synth_witness(uint4* inputs, uint4* aux, uint64_t seed, uint64_t n_vars, uint64_t n_inputs) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_vars; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v[8];
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = j == 0 ? 1u : 0u;
        } else {
            sm_sample<F>(sm_key(seed, 4, i, 0), v);
        }
        st8((i < n_inputs ? inputs + 2 * i : aux + 2 * (i - n_inputs)), v);
    }
}

}  // namespace bp
