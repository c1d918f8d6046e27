// CUDA kernels of the R1CS evaluation engine (sm_100a).  DESIGN.md has the layout and the per-kernel rooflines.
//
//   K1  which_is_unsatisfied (test_cs.rs:239-253 + eval_lc :137-155), one check =
//         check_small            plain rows whose operands are small: 64-bit integer arithmetic on the witness shadows
//         check_rows<LIST>       generic rows + rows check_small deferred: full-width lazy-reduction arithmetic, thread per row
//         check_fat_int          fat rows (> fat_terms terms): exact integers from per-lane buckets, warp per row
//         check_fat_rows         fat rows check_fat_int left undecided: full-width, warp per row
//   K2  batched LinearCombination::eval (lc.rs:245-267): the same four kernels with EMIT = true
//   K3  classify_lcs<F> + convert_terms<F>      ingest: canonical coefficient -> class + device-internal form (+ exponents)
//       validate_canonical<F>, widen_u8/_bits   witness upload: value < p, shadow; packed forms -> shadows
//       build_row_meta, build_fat_words, ...    the plan: row kinds, term words of the integer kernels, readiness of rows
//   K5  synth_*                                 synthetic instance generator (measurement fixture)
//       eval_lc_kernel<F>                       one ad-hoc LinearCombination::eval
//
// Device layout (per handle == per row shard):
//   row_ptr : u32[3N+1]   LC offsets; LC 3i, 3i+1, 3i+2 are A_i, B_i, C_i; a row's terms are contiguous
//   cols    : u32[nnz]    bit 31 = aux index space, bits 30..28 = coefficient class, bits 27..0 = index
//   vals    : uint4[2nnz] coefficient, 8 x u32 limbs, INTERNAL form (below)
//   kexp    : u16[nnz]    exponents k1 | k2 << 8 of the terms whose coefficient is +-(2^k1 [+ 2^k2])
//   inputs  : uint4[2*n_inputs], aux : uint4[2*n_aux]   canonical witness;  shadow : u32 per element (see ld_witness)
//   row_meta, scols, fat_rows, gen_rows                  the plan (see build_row_meta / build_fat_words)
//
// Internal coefficient form.  Evaluation never reduces per term:
//   * general A/B LC  : stored = c * 2^288 mod p, class GEN (or POW2P/POW2M: same storage, the class is a hint).
//                       acc (17 limbs) += stored * w ; value = redc(acc) in [0,2p)
//   * plain   A/B LC  : every coefficient is +-1, +-2 or 0 and sum|c| <= 7: classes P1/M1/P2/M2/ZERO,
//                       acc (9 limbs) += |c| * (w or p-w) ; value = acc mod p by three conditional subtractions
//   * C terms         : the NEGATED coefficient, unscaled: stored = p - c, with the class of that value.
// Row check:  X = Az*Bz + sum_C (-c)*w  (unreduced, 17 limbs);  satisfied  <=>  redc(X) in {0, p}
// (the zero test is invariant under redc's 2^-288 factor).  Per row: T_gen + 1 products, <= 3 reductions.
// The integer kernels decide a row from 4-byte shadows whenever every operand it touches is < 2^24; any other row goes
// to the full-width kernels in the same check, so every witness gets the verdict of the general arithmetic.
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
    const uint16_t* __restrict__ kexp;  // packed exponents of the terms whose class is POW2P / POW2M (others: unspecified)
    const uint4* __restrict__ inputs;
    const uint4* __restrict__ aux;
    const uint32_t* __restrict__ shadow;    // witness shadows (the value when it is < 2^24, else kShadowBig): inputs at
    uint32_t aux_off;                       // [0, n_inputs), aux at [aux_off, aux_off + n_aux)
    uint32_t wide_valid;                    // 0 after a packed upload: inputs/aux are then current only where the shadow says "big"
    const uint32_t* __restrict__ row_meta;  // plan: per row, see meta_pack
    const uint32_t* __restrict__ scols;     // plan: per term, the word check_small works from (small_word)
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

// Plan: what kind of row is it (top byte of row_meta, built by build_row_meta)?
enum RowKind : uint32_t {
    kRowGeneric = 0,  // check_rows (full arithmetic)
    kRowPlain = 1,    // every term has a small-coefficient class, sum|c| <= 7 in A and in B, <= 5 in C, every length <= 254:
                      // check_small decides it from the witness shadows when every value it touches is small
    kRowFat = 2       // check_fat_rows
};
// row_meta word: bits 0..9 = offset of the row's first term from the first term of its 64-row block (saturating),
// bits 10..17 = |A|, bits 18..25 = |B| (both saturating at 255; a plain row has every length <= 254), bits 26..27 = RowKind.
// |C| of a row is the next row's offset minus its own offset, |A| and |B|.
constexpr uint32_t kSmallRows = 64;  // rows per warp and round of check_small
constexpr uint32_t kSmallCap = 640;  // words per staging buffer of check_small; a block with more terms is done row by row
constexpr uint32_t kMetaOffMask = 1023u;
__host__ __device__ __forceinline__ uint32_t meta_pack(uint32_t off, uint32_t la, uint32_t lb, uint32_t kind) {
    return (off < kMetaOffMask ? off : kMetaOffMask) | ((la < 255u ? la : 255u) << 10) | ((lb < 255u ? lb : 255u) << 18) | (kind << 26);
}
__host__ __device__ __forceinline__ uint32_t meta_kind(uint32_t meta) { return (meta >> 26) & 3u; }

// Witness shadow: a second, 4-byte copy of every witness element, maintained wherever the witness is written.
// Gadget circuits (sha256, blake2s, boolean, uint32) have bit- or byte-valued witnesses; a row whose operands are all
// small is decided in 64-bit integer arithmetic from 4-byte gathers, and a full-width coefficient times a small value is
// one 1x8 product instead of 8x8.  The shortcut is taken on VALUES: any element >= 2^24 sends its rows down the
// full-width path, so every witness gets the verdict of the general arithmetic.
constexpr uint32_t kSmallBits = 24;
constexpr uint32_t kShadowBig = 0xffffffffu;
__host__ __device__ __forceinline__ uint32_t shadow_of(const uint32_t* x /*8*/) {
    const uint32_t hi = x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7];
    return (hi == 0u && x[0] < (1u << kSmallBits)) ? x[0] : kShadowBig;
}
// A witness element as 8 limbs.  Packed uploads (bp_cs_*_u8) write only the shadows; from then on (wide_valid = 0) the
// 32-byte arrays are current exactly where the shadow says "big", and a small value is its zero-extended shadow.
__device__ __forceinline__ void ld_witness(uint32_t* w, const CsrView& m, bool is_aux, uint32_t idx) {
    if (!m.wide_valid) {
        const uint32_t s = __ldg(m.shadow + (is_aux ? m.aux_off : 0u) + idx);
        if (s != kShadowBig) {
            w[0] = s;
#pragma unroll
            for (int i = 1; i < 8; ++i) w[i] = 0;
            return;
        }
    }
    // one 256-bit load; the L2::64B hint keeps an L2 miss of this random gather at 64 bytes of DRAM traffic instead of 128
    // (microbench3 under ncu)
    const uint4* p = (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx;
    asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}

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
    ld_witness(t.w, m, is_aux, idx);
}

// Fold term k (already loaded into t) into acc.  RIPPLE = 9: A/B accumulator (plain sums stay below 2^288); 17: the
// row-check accumulator.  `gen` is set when a full product was folded; `mag` accumulates the plain magnitudes.
template <int F, int RIPPLE>
__device__ __forceinline__ void apply_term(uint32_t* acc /*17*/, TermW& t, uint32_t k, const CsrView& m, uint32_t& gen, uint32_t& mag) {
    const uint32_t cls = t.cls;
    if (cls == kClsZero) return;
    if (cls == kClsGen || cls == kClsPow2P || cls == kClsPow2M) {  // (+-2^k is stored like any full-width coefficient)
        uint32_t c[8];
        ld8(c, m.vals + 2 * (size_t)k);
        mac_wide(acc, c, t.w);
        gen = 1;
        return;
    }
    if (cls == kClsM1 || cls == kClsM2) {
        uint32_t n[8];
        neg_mod<F>(n, t.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) t.w[i] = n[i];
    }
    acc_add8<RIPPLE>(acc, t.w);
    mag += 1;
    if (cls == kClsP2 || cls == kClsM2) {
        acc_add8<RIPPLE>(acc, t.w);
        mag += 1;
    }
}

// Kernel feature bits (template parameter V), kept switchable so that each can be measured on the GPU (DESIGN.md "Variants").
constexpr int kVPipe = 1;      // fetch the next term's column + witness while the current term is folded
constexpr int kVMagSkip = 2;   // plain A/B sums below 2p are used unreduced
constexpr int kVBitRow = 4;    // rows whose Az or Bz is 0/1 (and C is plain) are decided without any multiplication
constexpr int kVPark = 8;      // az / bz wait in shared memory while C is folded (fewer live registers)
constexpr int kVPrefetch = 16; // L1 prefetch of every witness line of the row before the first term is folded
constexpr int kVShadow = 32;   // warp-per-row kernel: use the witness shadows (small operand => 1x8 product / integer add)

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
// LIST: the rows to do are list_a[0..n_a) (static: the plan's generic rows) followed by list_b[0..*n_b) (dynamic: plain
// rows that check_small deferred because an operand was not small); otherwise every row that is not fat.
template <int F, bool EMIT, int V, int MB, bool LIST = false>
__global__ void __launch_bounds__(128, MB) check_rows(CsrView m, CheckOut o, FieldConsts fc, const uint32_t* __restrict__ list_a = nullptr,
                                                      uint32_t n_a = 0, const uint32_t* __restrict__ list_b = nullptr,
                                                      const uint32_t* __restrict__ n_b = nullptr) {
    constexpr bool PIPE = (V & kVPipe) != 0, PARK = (V & kVPark) != 0;
    __shared__ uint32_t s_az[PARK ? 8 : 1][128], s_bz[PARK ? 8 : 1][128];
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    const uint32_t n_todo = LIST ? n_a + *n_b : m.n_rows;
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n_todo; it += gridDim.x * blockDim.x) {
        const uint32_t row = LIST ? (it < n_a ? __ldg(list_a + it) : list_b[it - n_a]) : it;
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

// ---- plan: row_meta and the term words of the plain rows, one thread per row ---------------------------------------------
// scols[k] for a term of a plain row: bits 28..31 = signed multiplier (+-1, +-2), bits 0..27 = index of its variable in
// the shadow array; every other term (and the padding) holds the null word: multiplier 0, index of the slot that is always 0.
__global__ void fill_u32(uint32_t* p, size_t n, uint32_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void build_row_meta(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ cols, uint32_t n_rows, uint32_t fat_terms,
                               uint32_t n_inputs, uint32_t n_aux, uint32_t aux_off, uint32_t* __restrict__ row_meta,
                               uint32_t* __restrict__ scols, uint32_t* __restrict__ counts /*4*/,
                               unsigned long long* __restrict__ term_counts /*3: terms per RowKind*/) {
    uint32_t n_kind[3] = {0, 0, 0};
    unsigned long long t_kind[3] = {0, 0, 0};
    bool oob = false;
    const bool idx_fits = (uint64_t)aux_off + n_aux <= (1ull << 28);  // shadow indices must fit 28 bits
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += gridDim.x * blockDim.x) {
        const uint32_t p0 = row_ptr[3 * (size_t)row], p1 = row_ptr[3 * (size_t)row + 1], p2 = row_ptr[3 * (size_t)row + 2],
                       p3 = row_ptr[3 * (size_t)row + 3];
        const uint32_t pb = row_ptr[3 * (size_t)(row & ~(kSmallRows - 1u))];
        const uint32_t la = p1 - p0, lb = p2 - p1, lc = p3 - p2;
        uint32_t k = kRowPlain;
        if (p3 - p0 > fat_terms) {
            k = kRowFat;
        } else if (la > 254u || lb > 254u || lc > 254u || !idx_fits) {
            k = kRowGeneric;
        } else {
            uint32_t mag[3] = {0, 0, 0};
            for (uint32_t t = p0; t < p3 && k == kRowPlain; ++t) {
                const uint32_t col = __ldg(cols + t);
                const uint32_t cls = (col >> kColClsShift) & 7u;
                const int lc_i = t < p1 ? 0 : (t < p2 ? 1 : 2);
                if (cls == kClsGen || cls == kClsPow2P || cls == kClsPow2M) k = kRowGeneric;  // check_small knows +-1, +-2 and 0
                else if (cls == kClsP1 || cls == kClsM1) mag[lc_i] += 1;
                else if (cls == kClsP2 || cls == kClsM2) mag[lc_i] += 2;
                if (mag[0] > 7u || mag[1] > 7u || mag[2] > 5u) k = kRowGeneric;
                // check_small does not bound-check its gathers: a plain row only has columns that exist
                if (cls != kClsZero && (col & kColIdxMask) >= ((col & kColAux) ? n_aux : n_inputs)) {
                    oob = true;
                    k = kRowGeneric;
                }
            }
            if (k == kRowPlain) {
                for (uint32_t t = p0; t < p3; ++t) {
                    const uint32_t col = __ldg(cols + t);
                    const uint32_t cls = (col >> kColClsShift) & 7u;
                    const uint32_t nib = (0x000E2F10u >> (4u * cls)) & 15u;  // P1 +1, M1 -1, P2 +2, M2 -2, else 0
                    if (nib) scols[t] = (nib << 28) | ((col & kColIdxMask) + ((col & kColAux) ? aux_off : 0u));
                }
            }
        }
        row_meta[row] = meta_pack(p0 - pb, la, lb, k);
        n_kind[0] += k == kRowGeneric;
        n_kind[1] += k == kRowPlain;
        n_kind[2] += k == kRowFat;
        t_kind[k] += p3 - p0;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // all threads are back together here
        const uint32_t tot = __reduce_add_sync(0xffffffffu, n_kind[i]);
        if ((threadIdx.x & 31u) == 0 && tot) atomicAdd(counts + i, tot);
        unsigned long long tt = t_kind[i];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) tt += __shfl_xor_sync(0xffffffffu, tt, d);
        if ((threadIdx.x & 31u) == 0 && tt) atomicAdd(term_counts + i, tt);
    }
    if (__any_sync(0xffffffffu, oob) && (threadIdx.x & 31u) == 0) atomicOr(counts + 3, 1u);
}

// ---- plan for the pipelined re-check (bp_cs_recheck_u8): which rows only read aux variables below a given index? ---------
// row_max[r] = largest aux index row r reads (0 when none); after an inclusive max-scan the rows that are ready once the
// aux elements [0, bound) have arrived are a PREFIX of the rows (circuits allocate a variable before they constrain it, so
// the prefix is long); ready_rows finds its length for every chunk boundary.
__global__ void row_max_aux(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ cols, uint32_t n_rows,
                            uint32_t* __restrict__ row_max) {
    const uint32_t lane = threadIdx.x & 31u, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
        const uint32_t k0 = row_ptr[3 * (size_t)row], k1 = row_ptr[3 * (size_t)row + 3];
        uint32_t mx = 0;
        for (uint32_t k = k0 + lane; k < k1; k += 32u) {
            const uint32_t col = cols[k];
            if ((col & kColAux) && ((col >> kColClsShift) & 7u) != kClsZero) mx = max(mx, col & kColIdxMask);
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) row_max[row] = mx;
    }
}
// needed[c] = 1 when some row of this handle reads an aux element of chunk c (2^16 elements): a row shard only has to
// receive those chunks of a new witness (bp_cs_set_option "sparse_upload").
constexpr uint32_t kNeedChunkLog2 = 16;
__global__ void mark_needed_aux(const uint32_t* __restrict__ cols, size_t nnz, uint8_t* __restrict__ needed) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nnz; k += (size_t)gridDim.x * blockDim.x) {
        const uint32_t col = cols[k];
        if ((col & kColAux) && ((col >> kColClsShift) & 7u) != kClsZero) needed[(col & kColIdxMask) >> kNeedChunkLog2] = 1;
    }
}
// out[3*i + 0] = number of leading rows whose aux reads are all below bounds[i]; [1], [2] = how many entries of the
// (ascending) fat / generic row lists lie below that row.
__global__ void ready_rows(const uint32_t* __restrict__ prefix_max, uint32_t n_rows, const uint32_t* __restrict__ fat_rows, uint32_t n_fat,
                           const uint32_t* __restrict__ gen_rows, uint32_t n_gen, const uint32_t* __restrict__ bounds, uint32_t n_bounds,
                           uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bounds) return;
    auto lower = [](const uint32_t* a, uint32_t n, uint32_t v) {  // first index with a[idx] >= v
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (a[mid] < v) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const uint32_t rows = lower(prefix_max, n_rows, bounds[i]);
    out[3 * i] = rows;
    out[3 * i + 1] = lower(fat_rows, n_fat, rows);
    out[3 * i + 2] = lower(gen_rows, n_gen, rows);
}

// ---- K1, plain rows with small operands: 64-bit integer arithmetic on the witness shadows ---------------------------------
// A warp owns 64 consecutive rows, two per lane.  The block's term words (scols) are brought into shared memory by one TMA
// bulk copy; they are handled TERM-parallel -- four consecutive words per lane (one 16-byte shared load), one 4-byte
// shadow gather each, four gathers in flight per lane -- and replaced in place by the EXCLUSIVE PREFIX SUM of the signed
// contributions c*w over the block (4 adds per lane + one shuffle scan per 128 terms).  A row's three sums are then four
// shared loads and three subtractions, and the row holds iff  Az*Bz + sum_C(-c)w == 0  as integers
// (|Az|,|Bz|,|Cz| < 2^27: nothing wraps, and |X| < p, so X = 0 mod p iff X = 0).
// A row with an operand that is not small is appended to `deferred` and decided by check_rows<LIST>.
//
// Per-warp software pipeline over its blocks b, b+W, b+2W, ...:
//   two blocks ahead   the block's term range [row_ptr[192b'], row_ptr[192(b'+1)]) is loaded into registers,
//   one block ahead    its term words are copied to shared memory (cp.async.bulk + mbarrier, double-buffered) and its
//                      row_meta words are loaded into registers,
//   current block      contributions + prefix sums, row verdicts.
// Only the shadow gathers of the current block are exposed latency.  A block whose range does not fit the buffer (it
// contains a fat row) is done thread-per-row straight from global memory.
constexpr int kSmallThreads = 256;
constexpr uint32_t kPoisonBit = 1u << 28;  // contribution of a term whose operand is not small (see small_poisoned)

__device__ __forceinline__ uint32_t ldg_early(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg_early2(const uint32_t* p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
// TMA bulk copy global -> shared (bytes and both addresses multiples of 16), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    }
}

// Column word -> index into the shadow array (inputs at 0, aux at aux_off).
__device__ __forceinline__ uint32_t shadow_index(uint32_t col, const CsrView& m) {
    return (col & kColIdxMask) + (((int32_t)col >> 31) & m.aux_off);
}
// Shadow of a term's variable, with a bounds check (index 0 = ONE stands in when there is nothing to load); branch-free so
// that the gathers of several terms are issued back to back.
__device__ __forceinline__ uint32_t small_gather(uint32_t col, const CsrView& m, unsigned int& err) {
    const uint32_t cls = (col >> kColClsShift) & 7u;
    const bool oob = (col & kColIdxMask) >= ((col & kColAux) ? m.n_aux : m.n_inputs);
    if (oob && cls != kClsZero) err = 1;
    return __ldg(m.shadow + ((oob || cls == kClsZero) ? 0u : shadow_index(col, m)));
}
// Contribution of a term word: multiplier (top nibble, signed) times the shadow; kPoisonBit when the operand is not small.
// (The null word's variable is the always-zero slot, so it is never poisoned.)
__device__ __forceinline__ uint32_t small_contrib(uint32_t word, uint32_t s) {
    return s == kShadowBig ? kPoisonBit : (uint32_t)(((int32_t)word >> 28) * (int32_t)s);
}
// A plain LC has at most 7 non-zero terms and |sum of the clean ones| < 2^27, so whether one was poisoned can be read off
// the total: bits 28.. of (sum + 2^27) are zero iff none was.
__device__ __forceinline__ bool small_poisoned(uint32_t a) { return ((a + (1u << 27)) >> 28) != 0u; }

// One row from its three sums: 0 = holds, 1 = fails, 2 = needs the full-width path.
__device__ __forceinline__ uint32_t small_verdict(uint32_t a, uint32_t b, uint32_t c) {
    if (small_poisoned(a) | small_poisoned(b) | small_poisoned(c)) return 2u;
    return ((long long)(int32_t)a * (long long)(int32_t)b + (long long)(int32_t)c) != 0ll ? 1u : 0u;
}

// Canonical residue of a small signed integer: v, or p - |v|.
template <int F> __device__ __forceinline__ void st_small_canonical(uint4* dst, int32_t v) {
    uint32_t x[8] = {(uint32_t)(v < 0 ? -v : v), 0, 0, 0, 0, 0, 0, 0};
    if (v < 0) {
        uint32_t r[8];
        neg_mod<F>(r, x);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = r[i];
    }
    st8(dst, x);
}
// Emit mode: A.w, B.w, C.w of a row the integer kernel decided (c is the sum over the NEGATED C coefficients).
template <int F> __device__ __forceinline__ void emit_small(const CheckOut& o, uint32_t row, uint32_t a, uint32_t b, uint32_t c) {
    if (o.az) st_small_canonical<F>(o.az + 2 * (size_t)row, (int32_t)a);
    if (o.bz) st_small_canonical<F>(o.bz + 2 * (size_t)row, (int32_t)b);
    if (o.cz) st_small_canonical<F>(o.cz + 2 * (size_t)row, -(int32_t)c);
}

// One plain row by its own thread straight from global memory (blocks that do not fit the staging buffer).
__device__ __forceinline__ uint32_t small_row_direct(uint32_t row, const CsrView& m, uint32_t* sums = nullptr) {
    const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                   p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
    uint32_t sum[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint32_t k0 = i == 0 ? p0 : (i == 1 ? p1 : p2), k1 = i == 0 ? p1 : (i == 1 ? p2 : p3);
#pragma unroll 1
        for (uint32_t k = k0; k < k1; ++k) {
            const uint32_t w = __ldg(m.scols + k);
            sum[i] += small_contrib(w, __ldg(m.shadow + (w & kColIdxMask)));
        }
    }
    if (sums) { sums[0] = sum[0]; sums[1] = sum[1]; sums[2] = sum[2]; }
    return small_verdict(sum[0], sum[1], sum[2]);
}

// EMIT: also write the canonical A.w, B.w, C.w of every row decided here (F is only used for that).
template <int F, bool EMIT>
__global__ void __launch_bounds__(kSmallThreads, 5) check_small(CsrView m, CheckOut o, uint32_t* __restrict__ deferred,
                                                                uint32_t* __restrict__ n_deferred, uint32_t blk_lo, uint32_t blk_hi) {
    extern __shared__ __align__(16) unsigned char small_smem[];
    constexpr uint32_t kStageWords = kSmallCap + 4u;  // + one group for the prefix value past the last term
    uint32_t(*s_buf)[2][kStageWords] = reinterpret_cast<uint32_t(*)[2][kStageWords]>(small_smem);
    __shared__ __align__(8) unsigned long long s_bar[kSmallThreads / 32][2];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    if (lane == 0) {
        mbar_init(&s_bar[wib][0], 1);
        mbar_init(&s_bar[wib][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t my_bad = 0xffffffffu;
    const uint32_t n_blocks = min(blk_hi, (m.n_rows + kSmallRows - 1u) / kSmallRows);  // this launch: blocks [blk_lo, blk_hi)
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t b0 = blk_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);

    // (volatile loads: issued HERE, one iteration before they are needed, not sunk to their first use)
    auto range_lo = [&](uint32_t b) { return ldg_early(m.row_ptr + 3 * (size_t)kSmallRows * b); };
    auto range_hi = [&](uint32_t b) { return ldg_early(m.row_ptr + 3 * (size_t)min(kSmallRows * (b + 1u), m.n_rows)); };
    auto load_meta = [&](uint32_t b) {  // the lane's two rows; row_meta is padded to an even number of words
        const uint32_t r = kSmallRows * b + 2u * lane;
        uint2 v = make_uint2(kMetaOffMask, kMetaOffMask);  // kind Generic
        if (r < m.n_rows) v = ldg_early2(m.row_meta + r);  // (v.y of a last odd row is whatever the padding holds: see plain1)
        return v;
    };
    // copy the words [kb & ~3, roundup4(ke)) of scols into a stage; false when they do not fit
    auto issue_copy = [&](uint32_t kb, uint32_t ke, uint32_t stage) {
        const uint32_t w0 = kb & ~3u, w1 = (ke + 3u) & ~3u;
        if (w1 - w0 > kSmallCap) return false;
        if (ke > kb && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // our generic-proxy accesses of the stage are done
            bulk_g2s(s_buf[wib][stage], m.scols + w0, (w1 - w0) * 4u, &s_bar[wib][stage]);
        }
        return true;
    };

    if (b0 < n_blocks) {
        // prologue: block b0 staged, block b0+W's range in registers
        uint32_t kb_cur = range_lo(b0), ke_cur = range_hi(b0);
        uint2 meta_cur = load_meta(b0);
        bool copied_cur = issue_copy(kb_cur, ke_cur, 0);
        uint32_t kb_nxt = 0, ke_nxt = 0;
        if (b0 + n_warps < n_blocks) { kb_nxt = range_lo(b0 + n_warps); ke_nxt = range_hi(b0 + n_warps); }
        uint32_t it = 0, phase = 0;  // phase bit s: parity the next wait on stage s expects (copies and waits pair up per stage)
        for (uint32_t blk = b0; blk < n_blocks; blk += n_warps, ++it) {
            const uint32_t stage = it & 1u;
            const uint32_t b_nxt = blk + n_warps, b_nn = blk + 2u * n_warps;
            // 1. loads for the blocks ahead (consumed in the next iteration)
            uint2 meta_nxt = make_uint2(kMetaOffMask, kMetaOffMask);
            uint32_t kb_nn = 0, ke_nn = 0;
            if (b_nxt < n_blocks) meta_nxt = load_meta(b_nxt);
            if (b_nn < n_blocks) { kb_nn = range_lo(b_nn); ke_nn = range_hi(b_nn); }
            // 2. the next block's term words: the other stage is free (its rows were finished before the last __syncwarp)
            bool copied_nxt = false;
            if (b_nxt < n_blocks) copied_nxt = issue_copy(kb_nxt, ke_nxt, stage ^ 1u);
            // 3. the current block
            const uint32_t row = blk * kSmallRows + 2u * lane;
            const bool plain0 = meta_kind(meta_cur.x) == kRowPlain, plain1 = meta_kind(meta_cur.y) == kRowPlain && row + 1u < m.n_rows;
            const bool any_plain = __any_sync(0xffffffffu, plain0 || plain1);
            uint32_t verdict0 = 0, verdict1 = 0;
            if (copied_cur) {
                if (ke_cur > kb_cur) {
                    mbar_wait(&s_bar[wib][stage], (phase >> stage) & 1u);
                    phase ^= 1u << stage;
                }
                if (any_plain) {
                    uint32_t* st = s_buf[wib][stage];
                    const uint32_t head = kb_cur & 3u, nt = ke_cur - kb_cur;
                    // staged words (a multiple of 4); nothing was copied for a block without terms (issue_copy), and the stage
                    // must not be read then: its stale words would be used as shadow indices
                    const uint32_t len = nt ? ((ke_cur + 3u) & ~3u) - (kb_cur & ~3u) : 0u;
                    // term words -> exclusive prefix sums of the contributions, in place; one more group holds the total
                    uint32_t carry = 0;
                    for (uint32_t t4 = 4u * lane; t4 - 4u * lane <= len; t4 += 128u) {
                        uint4 w = make_uint4(0, 0, 0, 0);
                        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                        if (t4 < len) {
                            w = *reinterpret_cast<const uint4*>(st + t4);
                            const uint32_t s0 = __ldg(m.shadow + (w.x & kColIdxMask)), s1 = __ldg(m.shadow + (w.y & kColIdxMask)),
                                           s2 = __ldg(m.shadow + (w.z & kColIdxMask)), s3 = __ldg(m.shadow + (w.w & kColIdxMask));
                            c0 = small_contrib(w.x, s0);
                            c1 = small_contrib(w.y, s1);
                            c2 = small_contrib(w.z, s2);
                            c3 = small_contrib(w.w, s3);
                        }
                        const uint32_t tot = c0 + c1 + c2 + c3;
                        uint32_t incl = tot;
#pragma unroll
                        for (uint32_t d = 1; d < 32; d <<= 1) {
                            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                            if (lane >= d) incl += up;
                        }
                        const uint32_t ex = carry + incl - tot;
                        if (t4 <= len) *reinterpret_cast<uint4*>(st + t4) = make_uint4(ex, ex + c0, ex + c0 + c1, ex + c0 + c1 + c2);
                        carry += __shfl_sync(0xffffffffu, incl, 31);
                    }
                    __syncwarp();
                    // offsets of the lane's two rows and of the row after them
                    const uint32_t o0 = meta_cur.x & kMetaOffMask, o1 = meta_cur.y & kMetaOffMask;
                    uint32_t o2 = __shfl_down_sync(0xffffffffu, o0, 1);
                    if (lane == 31u || row + 2u >= m.n_rows) o2 = nt;
                    if (row + 1u >= m.n_rows) o2 = nt;  // (row + 1 does not exist: row's terms end the block)
                    const uint32_t e1 = (row + 1u < m.n_rows) ? o1 : nt;
                    const uint32_t* q = st + head;
                    if (plain0) {
                        const uint32_t la = (meta_cur.x >> 10) & 255u, lb = (meta_cur.x >> 18) & 255u;
                        const uint32_t pa = q[o0], pb = q[o0 + la], pc = q[o0 + la + lb], pd = q[e1];
                        verdict0 = small_verdict(pb - pa, pc - pb, pd - pc);
                        if (EMIT && verdict0 != 2u) emit_small<F>(o, row, pb - pa, pc - pb, pd - pc);
                    }
                    if (plain1) {
                        const uint32_t la = (meta_cur.y >> 10) & 255u, lb = (meta_cur.y >> 18) & 255u;
                        const uint32_t pa = q[o1], pb = q[o1 + la], pc = q[o1 + la + lb], pd = q[o2];
                        verdict1 = small_verdict(pb - pa, pc - pb, pd - pc);
                        if (EMIT && verdict1 != 2u) emit_small<F>(o, row + 1u, pb - pa, pc - pb, pd - pc);
                    }
                }
            } else if (any_plain) {
                uint32_t sums[3];
                if (plain0) {
                    verdict0 = small_row_direct(row, m, sums);
                    if (EMIT && verdict0 != 2u) emit_small<F>(o, row, sums[0], sums[1], sums[2]);
                }
                if (plain1) {
                    verdict1 = small_row_direct(row + 1u, m, sums);
                    if (EMIT && verdict1 != 2u) emit_small<F>(o, row + 1u, sums[0], sums[1], sums[2]);
                }
            }
            if (verdict0 == 1u && row < my_bad) my_bad = row;
            if (verdict1 == 1u && row + 1u < my_bad) my_bad = row + 1u;
            // deferred rows, appended warp-aggregated
            const uint32_t d0 = __ballot_sync(0xffffffffu, verdict0 == 2u), d1 = __ballot_sync(0xffffffffu, verdict1 == 2u);
            if (d0 | d1) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(n_deferred, (uint32_t)(__popc(d0) + __popc(d1)));
                base = __shfl_sync(0xffffffffu, base, 0);
                const uint32_t below = (1u << lane) - 1u;
                if (verdict0 == 2u) deferred[base + (uint32_t)__popc(d0 & below)] = row;
                if (verdict1 == 2u) deferred[base + (uint32_t)__popc(d0) + (uint32_t)__popc(d1 & below)] = row + 1u;
            }
            __syncwarp();  // every lane is done with this stage before it is refilled
            // 4. rotate
            kb_cur = kb_nxt; ke_cur = ke_nxt; meta_cur = meta_nxt; copied_cur = copied_nxt;
            kb_nxt = kb_nn; ke_nxt = ke_nn;
        }
    }
    publish_first_bad(my_bad, m, o, 0u);
}
constexpr size_t kSmallSmem = (size_t)(kSmallThreads / 32) * 2 * (kSmallCap + 4) * 4;

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

// Small positive / negative plain contributions of one lane to a C LC, kept as integers until the LC ends:
//   sum_pos = sum mag*w (class P*),   sum_neg = sum mag*w and cnt = sum mag (class M*: the term is mag*(p - w)).
struct SmallSums {
    uint64_t pos_lo = 0, neg_lo = 0, cnt = 0;
    uint32_t pos_hi = 0, neg_hi = 0;
    uint32_t used = 0;
};

// acc (17 limbs) += pos + cnt*p - neg      (cnt*p >= neg because every w < p)
template <int F> __device__ __forceinline__ void fold_small_sums(uint32_t* acc, const SmallSums& ss) {
    uint32_t t[12];
    zeron<12>(t);
    uint32_t pl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pl[i] = PL<F>(i);
    acc_mad_small<12>(t, pl, (uint32_t)ss.cnt);
    acc_mad_small<11>(t + 1, pl, (uint32_t)(ss.cnt >> 32));
    uint32_t x[12];
    zeron<12>(x);
    x[0] = (uint32_t)ss.neg_lo; x[1] = (uint32_t)(ss.neg_lo >> 32); x[2] = ss.neg_hi;
    (void)subn<12>(t, t, x);
    x[0] = (uint32_t)ss.pos_lo; x[1] = (uint32_t)(ss.pos_lo >> 32); x[2] = ss.pos_hi;
    (void)addn<12>(t, t, x);
    uint32_t c = addn<12>(acc, acc, t);
#pragma unroll
    for (int i = 12; i < 17; ++i) {
        const uint32_t v = acc[i] + c;
        c = v < c ? 1u : 0u;
        acc[i] = v;
    }
}

// One lane's terms k0, k0+32, ... < k1 of a fat LC, with the witness shadows: a small operand turns a full-width
// coefficient into a 1x8 product and a plain coefficient into two 64-bit additions; anything else takes the general path.
// IS_C: C LCs take every class; A/B LCs only their GEN terms (a plain A/B LC has at most 7 non-zero terms and a value
// bound the general path maintains).
// A term whose coefficient is stored full-width (GEN, or the +-2^k hint classes)?
__device__ __forceinline__ bool is_product_class(uint32_t cls) { return cls == kClsGen || cls == kClsPow2P || cls == kClsPow2M; }
// Is term (class, shadow) one the shadow pass cannot take?  (operand not small; or a +-1 / +-2 term of a plain A/B LC)
template <bool IS_C> __device__ __forceinline__ bool shadow_slow(uint32_t cls, uint32_t sh) {
    return cls != kClsZero && (sh == kShadowBig || (!IS_C && !is_product_class(cls)));
}

constexpr int kFatU = 8;  // terms per lane and round in the shadow pass of the warp-per-row kernel
struct FatStage {
    uint4 c[kFatU][2][32];  // [term of the round][half][lane]: conflict-free 16-byte accesses
};
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// Pass 1 over a lane's terms k0, k0+32, ... < k1: everything whose operand is small.  Products go to a 10-limb
// accumulator (c*s < 2^280, so 2^40 of them fit), plain terms to `ss`.  Returns true when a term was left for pass 2.
// kFatU terms per lane and round: their column words (coalesced), then their shadows (4-byte gathers), then the
// coefficients of the terms that need one (operand not zero; cp.async straight into shared memory, no registers held)
// are each fetched together, so a round costs three memory round trips for 256 terms of the warp.
template <int F, bool IS_C>
__device__ __forceinline__ bool fold_lane_shadow(uint32_t* acc10, uint32_t k0, uint32_t k1, const CsrView& m, FatStage& fs, unsigned int& err,
                                                 uint32_t& gen, SmallSums& ss) {
    const uint32_t lane = threadIdx.x & 31u;
    bool any_slow = false;
#pragma unroll 1
    for (uint32_t kb = k0; kb < k1; kb += 32u * kFatU) {
        uint32_t col[kFatU], sh[kFatU];
#pragma unroll
        for (int j = 0; j < kFatU; ++j) col[j] = kb + 32u * j < k1 ? __ldg(m.cols + kb + 32u * j) : (kClsZero << kColClsShift);
#pragma unroll
        for (int j = 0; j < kFatU; ++j) {  // branch-free: the gathers are issued back to back
            const uint32_t idx = col[j] & kColIdxMask;
            if (idx >= ((col[j] & kColAux) ? m.n_aux : m.n_inputs) && ((col[j] >> kColClsShift) & 7u) != kClsZero) {
                err = 1;
                col[j] = kClsZero << kColClsShift;  // reported; contributes nothing
            }
            sh[j] = small_gather(col[j], m, err);
        }
        uint32_t need_c = 0;
#pragma unroll
        for (int j = 0; j < kFatU; ++j) {
            const uint32_t cls = (col[j] >> kColClsShift) & 7u;
            if (shadow_slow<IS_C>(cls, sh[j])) { any_slow = true; continue; }
            if (is_product_class(cls)) {
                gen = 1;
                if (sh[j] != 0u) {
                    need_c |= 1u << j;
                    const uint4* src = m.vals + 2 * (size_t)(kb + 32u * j);
                    cp_async16(&fs.c[j][0][lane], src);
                    cp_async16(&fs.c[j][1][lane], src + 1);
                }
            }
        }
        cp_async_wait_all();  // each lane reads back only what it copied itself: no cross-lane hazard
#pragma unroll
        for (int j = 0; j < kFatU; ++j) {
            const uint32_t cls = (col[j] >> kColClsShift) & 7u;
            if ((need_c >> j) & 1u) {
                uint32_t c[8];
                const uint4 lo = fs.c[j][0][lane], hi = fs.c[j][1][lane];
                c[0] = lo.x; c[1] = lo.y; c[2] = lo.z; c[3] = lo.w;
                c[4] = hi.x; c[5] = hi.y; c[6] = hi.z; c[7] = hi.w;
                acc_mad_small<10>(acc10, c, sh[j]);
            } else if (IS_C && cls != kClsZero && !is_product_class(cls) && sh[j] != kShadowBig && sh[j] != 0u) {
                const uint32_t mg = (cls == kClsP1 || cls == kClsM1) ? 1u : 2u;
                const uint64_t v = (uint64_t)mg * sh[j];
                if (cls == kClsP1 || cls == kClsP2) {
                    ss.pos_lo += v;
                    ss.pos_hi += ss.pos_lo < v ? 1u : 0u;
                } else {
                    ss.neg_lo += v;
                    ss.neg_hi += ss.neg_lo < v ? 1u : 0u;
                    ss.cnt += mg;
                }
                ss.used = 1;
            }
        }
    }
    return any_slow;
}

// Pass 2 (only when pass 1 left something): the terms pass 1 skipped, by the general path.
template <int F, int RIPPLE, bool IS_C>
__device__ __forceinline__ void fold_lane_slow(uint32_t* acc, uint32_t k0, uint32_t k1, const CsrView& m, unsigned int& err, uint32_t& gen,
                                               uint32_t& mag) {
#pragma unroll 1
    for (uint32_t k = k0; k < k1; k += 32u) {
        const uint32_t col = __ldg(m.cols + k);
        const uint32_t cls = (col >> kColClsShift) & 7u;
        if (cls == kClsZero) continue;
        const uint32_t idx = col & kColIdxMask;
        const bool is_aux = (col & kColAux) != 0;
        if (idx >= (is_aux ? m.n_aux : m.n_inputs)) continue;  // reported by pass 1
        if (!shadow_slow<IS_C>(cls, __ldg(m.shadow + (is_aux ? m.aux_off : 0u) + idx))) continue;
        TermW t;
        t.cls = cls;
        ld_witness(t.w, m, is_aux, idx);
        apply_term<F, RIPPLE>(acc, t, k, m, gen, mag);
    }
}

// ---- integer pass of the warp-per-row kernel -----------------------------------------------------------------------------
// MultiEq rows (multieq.rs:25-67: lhs * 1 = rhs with coefficients 2^k over bit- or word-valued variables) and their like
// are decided without any modular arithmetic: a term +-2^k * w with w < 2^24 is the exact integer  w << k, added into one
// of eight 64-bit buckets (limb position k/32) per sign, private to the lane and kept in shared memory (the index is
// dynamic); per LC the 32 lanes' buckets are summed and carried into two 10-limb integers P (positive part) and N
// (negative part).  With b = the LC of A/B that is a non-negative 32-bit integer and Y = the other one:
//     the row holds  <=>  Y_P * b + C_P == Y_N * b + C_N        (C holds the NEGATED coefficients)
// as integers whenever they are equal (an integer identity implies the congruence); when they differ and both sides are
// below 2^253 the difference is non-zero and smaller than p, so the row fails; anything else -- a full-width coefficient
// that is not +-2^k, an operand >= 2^24, two wide LCs -- falls back to the modular path.  Per term: two coalesced loads
// (column word, exponent byte), one 4-byte gather, ~20 instructions; the 32-byte coefficient is never read.
// Per LC at most 4096 terms, i.e. 128 per lane.  One term adds less than 2^56 to a 64-bit bucket (two shifted operands
// < 2^24 << 31 when both exponents fall into the same limb; a full-width coefficient adds < 2^32 per bucket and < 2^56 + 2^32
// to bucket 7), so a lane's bucket stays below 128 * (2^56 + 2^32) < 2^64: no wrap, the sums are the exact integers.
constexpr uint32_t kIntMaxTerms = 4096;
__device__ __forceinline__ uint32_t bk_slot(uint32_t j, uint32_t lane) { return j * 33u + lane; }  // (33: conflict-free column reads)
constexpr uint32_t kBucketWords = 16u * 33u;  // u64 per warp

// A full-width coefficient that is not +-2^k (e.g. the constant term of a UInt32 sum: the sum of 2^k over the set bits of
// a round constant, on ONE) times a small operand: recover the canonical coefficient (A/B LCs store c * 2^288: one
// Montgomery reduction undoes it; C LCs store p - c as it is) and add the 9 limbs of c * s into the positive buckets.
template <int F, bool IS_C>
__device__ __noinline__ void fat_gen_term(uint32_t k, uint32_t s, const CsrView& m, unsigned long long* bk) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t c[8];
    ld8(c, m.vals + 2 * (size_t)k);
    if (!IS_C) {
        uint32_t t[17];
#pragma unroll
        for (int i = 0; i < 17; ++i) t[i] = i < 8 ? c[i] : 0u;
        redc_acc<F>(c, t);  // stored * 2^-288 = c, in [0, 2p)
        reduce_once<F>(c);
    }
    unsigned long long carry = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        carry += (unsigned long long)c[i] * s;
        bk[bk_slot(i, lane)] += (uint32_t)carry;
        carry >>= 32;
    }
    bk[bk_slot(7, lane)] += carry << 32;  // (bucket 7 also carries limb 8)
}

// Term word of a fat row (plan, build_fat_words): bits 0..27 = index of the variable in the shadow array, bit 28 = negative,
// bits 29..30 = kind: 0 nothing to add (zero coefficient; padding), 1 = +-(2^e1 [+ 2^e2]) with the exponents in kexp,
// 2 = full-width coefficient.
constexpr uint32_t kFatNeg = 1u << 28, kFatKindShift = 29, kFatPow2 = 1u, kFatGen = 2u;

// Plan: term words (and exponents of the +-1 / +-2 classes) of the fat rows, one warp per row.
// flags[0] |= 1 when a column does not exist or a shadow index does not fit: the integer pass is then not used at all.
__global__ void build_fat_words(const uint32_t* __restrict__ fat_rows, uint32_t n_fat, const uint32_t* __restrict__ row_ptr,
                                const uint32_t* __restrict__ cols, uint32_t n_inputs, uint32_t n_aux, uint32_t aux_off,
                                uint32_t* __restrict__ scols, uint16_t* __restrict__ kexp, uint32_t* __restrict__ flags) {
    const uint32_t lane = threadIdx.x & 31u, warps = (gridDim.x * blockDim.x) >> 5;
    const bool idx_fits = (uint64_t)aux_off + n_aux <= (1ull << 28);
    bool bad = !idx_fits;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_fat; i += warps) {
        const uint32_t row = fat_rows[i];
        const uint32_t k0 = row_ptr[3 * (size_t)row], k1 = row_ptr[3 * (size_t)row + 3];
        for (uint32_t k = k0 + lane; k < k1; k += 32u) {
            const uint32_t col = cols[k];
            const uint32_t cls = (col >> kColClsShift) & 7u;
            if (cls == kClsZero) continue;  // (stays the null word)
            if ((col & kColIdxMask) >= ((col & kColAux) ? n_aux : n_inputs)) { bad = true; continue; }
            const uint32_t uidx = (col & kColIdxMask) + ((col & kColAux) ? aux_off : 0u);
            const uint32_t neg = (cls == kClsM1 || cls == kClsM2 || cls == kClsPow2M) ? kFatNeg : 0u;
            if (idx_fits) scols[k] = uidx | neg | ((cls == kClsGen ? kFatGen : kFatPow2) << kFatKindShift);
            if (cls == kClsP1 || cls == kClsM1) kexp[k] = (uint16_t)(0u | (kNoExp << 8));
            if (cls == kClsP2 || cls == kClsM2) kexp[k] = (uint16_t)(1u | (kNoExp << 8));
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 1u);
}

template <int F, bool IS_C>
__device__ __forceinline__ bool fat_lc_integer(uint32_t* P /*10*/, uint32_t* N /*10*/, uint32_t k0, uint32_t k1, const CsrView& m,
                                               unsigned long long* bk, uint32_t null_word) {
    const uint32_t lane = threadIdx.x & 31u;
    zeron<10>(P);
    zeron<10>(N);
    if (k1 == k0) return true;
    if (k1 - k0 > kIntMaxTerms) return false;
    if (k1 - k0 == 1u) {  // e.g. the "* 1" of a MultiEq row: +2^0 times a small operand
        const uint32_t w = __ldg(m.scols + k0), s = __ldg(m.shadow + (w & kColIdxMask));
        if ((w >> kFatKindShift) == 0u) return true;
        if ((w >> kFatKindShift) == kFatPow2 && !(w & kFatNeg) && __ldg(m.kexp + k0) == (uint16_t)(kNoExp << 8) && s != kShadowBig) {
            P[0] = s;
            return true;
        }
    }
    unsigned long long* mine = bk + lane;
#pragma unroll
    for (uint32_t j = 0; j < 16u; ++j) mine[33u * j] = 0ull;
    bool bad = false;
    // kFatU terms per lane and round.  The term words and exponents of the NEXT round are fetched (coalesced) while this
    // round's shadows are gathered, so a round exposes one memory round trip.
    uint32_t w_n[kFatU], kx_n[kFatU];
#pragma unroll
    for (int j = 0; j < kFatU; ++j) {
        const uint32_t k = k0 + lane + 32u * j;
        w_n[j] = k < k1 ? __ldg(m.scols + k) : null_word;
        kx_n[j] = k < k1 ? (uint32_t)__ldg(m.kexp + k) : 0u;
    }
#pragma unroll 1
    for (uint32_t kb = k0 + lane; kb < k1; kb += 32u * kFatU) {
        uint32_t w[kFatU], kx[kFatU], sh[kFatU];
#pragma unroll
        for (int j = 0; j < kFatU; ++j) {
            w[j] = w_n[j];
            kx[j] = kx_n[j];
        }
#pragma unroll
        for (int j = 0; j < kFatU; ++j) sh[j] = __ldg(m.shadow + (w[j] & kColIdxMask));  // (the null word's variable is always 0)
#pragma unroll
        for (int j = 0; j < kFatU; ++j) {
            const uint32_t k = kb + 32u * (kFatU + j);
            w_n[j] = k < k1 ? __ldg(m.scols + k) : null_word;
            kx_n[j] = k < k1 ? (uint32_t)__ldg(m.kexp + k) : 0u;
        }
        uint32_t gen_mask = 0;
#pragma unroll
        for (int j = 0; j < kFatU; ++j) {
            const uint32_t kind = w[j] >> kFatKindShift;
            if (sh[j] == 0u || kind == 0u) continue;  // (0 times any coefficient)
            if (sh[j] == kShadowBig) { bad = true; continue; }
            if (kind == kFatGen) { gen_mask |= 1u << j; continue; }
            unsigned long long* q = mine + ((w[j] & kFatNeg) ? 8u * 33u : 0u);
            const uint32_t e1 = kx[j] & 255u, e2 = kx[j] >> 8;
            q[33u * (e1 >> 5)] += (unsigned long long)sh[j] << (e1 & 31u);
            if (e2 != kNoExp) q[33u * (e2 >> 5)] += (unsigned long long)sh[j] << (e2 & 31u);
        }
#pragma unroll 1
        for (; gen_mask; gen_mask &= gen_mask - 1u) {  // rare: outlined, one at a time
            const uint32_t k = kb + 32u * (uint32_t)(__ffs((int)gen_mask) - 1);
            const uint32_t wk = __ldg(m.scols + k);
            fat_gen_term<F, IS_C>(k, __ldg(m.shadow + (wk & kColIdxMask)), m, bk);
        }
    }
    if (__any_sync(0xffffffffu, bad)) return false;
    __syncwarp();
    // lanes 0..7 sum the positive buckets 0..7 over the 32 lanes, lanes 16..23 the negative ones (96-bit totals)
    const uint32_t j = (lane & 7u) + ((lane & 16u) >> 1);
    unsigned long long lo = 0;
    uint32_t hi = 0;
#pragma unroll 8
    for (uint32_t i = 0; i < 32u; ++i) {
        const unsigned long long x = bk[bk_slot(j, i)];
        lo += x;
        hi += lo < x ? 1u : 0u;
    }
    __syncwarp();  // the buckets may be cleared by the next LC
    // limb i of the sum receives word 0 of bucket i, word 1 of bucket i-1, word 2 of bucket i-2
    const uint32_t li = lane & 15u;  // limb index handled by this lane (0..9 meaningful)
    const uint32_t w0 = li < 8u ? (uint32_t)lo : 0u;
    uint32_t w1 = __shfl_up_sync(0xffffffffu, (uint32_t)(lo >> 32), 1), w2 = __shfl_up_sync(0xffffffffu, hi, 2);
    if (li < 1u || li > 8u) w1 = 0;
    if (li < 2u || li > 9u) w2 = 0;
    const unsigned long long t = (unsigned long long)w0 + w1 + w2;  // < 2^34
    unsigned long long cp = 0, cn = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned long long tp = ((unsigned long long)__shfl_sync(0xffffffffu, (uint32_t)(t >> 32), i) << 32) |
                                      __shfl_sync(0xffffffffu, (uint32_t)t, i);
        const unsigned long long tn = ((unsigned long long)__shfl_sync(0xffffffffu, (uint32_t)(t >> 32), 16 + i) << 32) |
                                      __shfl_sync(0xffffffffu, (uint32_t)t, 16 + i);
        cp += tp;
        cn += tn;
        P[i] = (uint32_t)cp;
        N[i] = (uint32_t)cn;
        cp >>= 32;
        cn >>= 32;
    }
    return true;
}

// x (10 limbs) * b + y (10 limbs) -> r (12 limbs)
__device__ __forceinline__ void mad10(uint32_t* r, const uint32_t* x, uint32_t b, const uint32_t* y) {
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        c += (unsigned long long)x[i] * b + y[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    r[10] = (uint32_t)c;
    r[11] = (uint32_t)(c >> 32);
}

// (P - N) mod p for P, N < p, canonical; false when either is not below p.
template <int F> __device__ __forceinline__ bool int_residue(uint32_t* out /*8*/, const uint32_t* P /*10*/, const uint32_t* N /*10*/) {
    if ((P[8] | P[9] | N[8] | N[9]) != 0u || !is_canonical<F>(P) || !is_canonical<F>(N)) return false;
    uint32_t d[8], pl[8];
    const uint32_t borrow = subn<8>(d, P, N);
#pragma unroll
    for (int i = 0; i < 8; ++i) pl[i] = borrow ? PL<F>(i) : 0u;
    (void)addn<8>(out, d, pl);
    return true;
}

// One fat row as integers: 0 = holds, 1 = fails, 2 = undecided (take the modular path).  EMIT: lane 0 also writes the
// canonical A.w, B.w, C.w (a row whose sums are not below p is left undecided).
template <int F, bool EMIT>
__device__ __forceinline__ uint32_t fat_row_integer(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, const CsrView& m,
                                                    unsigned long long* bk, uint32_t null_word, const CheckOut& o, uint32_t row) {
    uint32_t Pa[10], Na[10], Pb[10], Nb[10];
    if (!fat_lc_integer<F, false>(Pa, Na, p0, p1, m, bk, null_word)) return 2u;
    if (!fat_lc_integer<F, false>(Pb, Nb, p1, p2, m, bk, null_word)) return 2u;
    uint32_t ea[8], eb[8];
    if (EMIT && (!int_residue<F>(ea, Pa, Na) || !int_residue<F>(eb, Pb, Nb))) return 2u;
    uint32_t a_wide = 0, b_wide = 0;  // non-zero when the LC is not a non-negative 32-bit integer
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        a_wide |= Na[i] | (i ? Pa[i] : 0u);
        b_wide |= Nb[i] | (i ? Pb[i] : 0u);
    }
    if (a_wide && b_wide) return 2u;
    const uint32_t b = b_wide ? Pa[0] : Pb[0];
    uint32_t YP[10], YN[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        YP[i] = b_wide ? Pb[i] : Pa[i];
        YN[i] = b_wide ? Nb[i] : Na[i];
    }
    uint32_t L[12], R[12];
    {
        uint32_t Pc[10], Nc[10];
        if (!fat_lc_integer<F, true>(Pc, Nc, p2, p3, m, bk, null_word)) return 2u;
        mad10(L, YP, b, Pc);
        mad10(R, YN, b, Nc);
        if (EMIT) {
            uint32_t ec[8];
            if (!int_residue<F>(ec, Nc, Pc)) return 2u;  // C holds the negated coefficients
            if ((threadIdx.x & 31u) == 0) {
                if (o.az) st8(o.az + 2 * (size_t)row, ea);
                if (o.bz) st8(o.bz + 2 * (size_t)row, eb);
                if (o.cz) st8(o.cz + 2 * (size_t)row, ec);
            }
        }
    }
    uint32_t diff = 0, high = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        diff |= L[i] ^ R[i];
        if (i >= 8) high |= L[i] | R[i];
    }
    if (diff == 0u) return 0u;
    // both below 2^253: the difference is non-zero and below p.  (Emit mode has written the row: it is decided either way;
    // the first-failure word is not read back by bp_cs_eval.)
    if (EMIT || (high == 0u && ((L[7] | R[7]) >> 29) == 0u)) return 1u;
    return 2u;
}

// One LC [k0,k1) by a whole warp: lanes stride the terms, then the partial sums are combined.  Every lane returns with
// the total, the "a full product was folded" flag and the plain magnitude.  SH: use the witness shadows.
template <int F, int RIPPLE, bool PIPE, bool SH = false>
__device__ __forceinline__ void warp_fold_lc(uint32_t* acc, uint32_t k0, uint32_t k1, const CsrView& m, unsigned int& err, uint32_t& any_gen,
                                             uint32_t& mag, FatStage* fs = nullptr) {
    const uint32_t lane = threadIdx.x & 31u;
    zeron<17>(acc);
    uint32_t g = 0, mg = 0;
    if (k1 - k0 <= 4) {  // short LC: lane 0 alone
        if (lane == 0) fold_range<F, RIPPLE, 1, false>(acc, k0, k1, m, err, g, mg);
    } else if (SH) {
        SmallSums ss;
        uint32_t acc10[10];
        zeron<10>(acc10);
        const bool any_slow = fold_lane_shadow<F, RIPPLE == 17>(acc10, k0 + lane, k1, m, *fs, err, g, ss);
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = acc10[i];
        if (ss.used) {
            fold_small_sums<F>(acc, ss);
            mg = 8;  // the sum is no longer a small multiple of p: the row takes the general test
        }
        if (any_slow) fold_lane_slow<F, 17, RIPPLE == 17>(acc, k0 + lane, k1, m, err, g, mg);
    } else {
        fold_range<F, 17, 32, PIPE>(acc, k0 + lane, k1, m, err, g, mg);
    }
    warp_sum17(acc);
    any_gen = __any_sync(0xffffffffu, g != 0) ? 1u : 0u;
    mag = __reduce_add_sync(0xffffffffu, mg > 8u ? 8u : mg);
}

// ---- K1, fat rows, integer pass: one warp per constraint; rows it cannot decide are listed for check_fat_rows -------------
template <int F, bool EMIT>
__global__ void __launch_bounds__(128, 6) check_fat_int(CsrView m, CheckOut o, const uint32_t* __restrict__ fat_rows, uint32_t n_fat,
                                                        uint32_t* __restrict__ undecided, uint32_t* __restrict__ n_undecided) {
    __shared__ unsigned long long s_bk[4][kBucketWords];
    unsigned long long* bk = s_bk[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t my_bad = 0xffffffffu;
    const uint32_t null_word = m.aux_off - 1u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_fat; i += warps) {
        const uint32_t row = __ldg(fat_rows + i);
        const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                       p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        const uint32_t v = fat_row_integer<F, EMIT>(p0, p1, p2, p3, m, bk, null_word, o, row);
        if (v == 1u && row < my_bad) my_bad = row;
        if (v == 2u && lane == 0) undecided[atomicAdd(n_undecided, 1u)] = row;
        __syncwarp();
    }
    publish_first_bad(my_bad, m, o, 0u);
}

// ---- K1 / K2, fat rows: one warp per constraint -------------------------------------------------------------------------
template <int F, bool EMIT, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_fat_rows(CsrView m, CheckOut o, FieldConsts fc, const uint32_t* __restrict__ fat_rows,
                                                          const uint32_t* __restrict__ n_fat) {
    constexpr bool PIPE = (V & kVPipe) != 0, PARK = (V & kVPark) != 0, SH = (V & kVShadow) != 0;
    __shared__ uint32_t s_az[PARK ? 8 : 1][128], s_bz[PARK ? 8 : 1][128];
    __shared__ FatStage s_fs[SH ? 4 : 1];
    FatStage* fs = &s_fs[SH ? (threadIdx.x >> 5) : 0];
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
        warp_fold_lc<F, 9, PIPE, SH>(acc, p0, p1, m, my_err, any_gen, mag, fs);
        finish_ab<F, (V & kVMagSkip) != 0>(az, acc, any_gen, mag);
        if (PARK) park8(s_az, az);
        warp_fold_lc<F, 9, PIPE, SH>(acc, p1, p2, m, my_err, any_gen, mag, fs);
        finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, any_gen, mag);
        if (PARK) park8(s_bz, bz);
        warp_fold_lc<F, 17, PIPE, SH>(acc, p2, p3, m, my_err, any_gen, mag, fs);
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

// ---- K1 / K2, product-heavy instances: the streaming full-width kernel over LC TILES ------------------------------------------
// A thread per row reading CSR makes every lane of a warp walk its own stream (32 distinct lines per load instruction, the L1
// tag stage is the bottleneck: measured 8 cycles per term and SM even with an L2-resident witness) and makes a warp wait for its
// longest row.  The plan therefore re-lays the rows out once, in tiles of 256 rows = 768 LCs:
//   * the tile's LCs are SORTED by length (descending) and cut into 24 slices of 32 -- one lane per LC, all lanes of a slice
//     (almost) equally long: no divergence, ~4 % padding;
//   * a slice stores its terms term-major: group j holds the j-th term of its 32 LCs, so a warp reads 32 x 32 B of coefficients
//     and 32 x 4 B of column words with ONE fully coalesced load each, streamed once (L1 no-allocate, L2 evict-first), while
//     each lane gathers its witness element (256-bit load, L2 evict-last) one term ahead of the arithmetic;
//   * an LC ends in shared memory (A.w, B.w reduced; the C sum unreduced, 17 limbs); after a block barrier a thread per row does
//     the Hadamard product, the one lazy reduction and the zero test exactly like check_rows.
// lct_lc[tile*768 + p] = (len << 16) | lc-in-tile (3*row_in_tile + type) for sorted position p; lct_slice_off[tile*24 + q] =
// first 32-slot group of slice q; lct_cols / lct_vals are indexed by slot = 32*group + lane.  Rows longer than fat_terms (done
// by check_fat_rows) and rows past the end have length-0 LCs.
constexpr uint32_t kLctRows = 256, kLctLcs = 768, kLctSlices = 24, kLctThreads = 256;
constexpr uint32_t kLctNullCol = kClsZero << kColClsShift;

__device__ __forceinline__ void ld256_stream(uint32_t* x, const uint4* p) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "l"(p));
}
// A random 32-byte gather that misses L2 costs 128 bytes of DRAM traffic by default and 64 with the L2::64B prefetch-size
// hint (measured with microbench3 under ncu: 107 -> 58 B per gather over a 512 MiB witness, 125 -> 63 B over 4 GiB; the
// cudaLimitMaxL2FetchGranularity device limit changes nothing): the hint halves the traffic of a kernel that is DRAM-bound.
__device__ __forceinline__ void ld256_keep(uint32_t* x, const uint4* p) {
    asm volatile("ld.global.nc.L2::evict_last.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "l"(p));
}
__device__ __forceinline__ uint32_t ld32_stream(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// Plan, pass 1: sort a tile's LCs by length, record (len, lc) per sorted position, the slice lengths (groups) and their sum.
__global__ void __launch_bounds__(256) lct_plan(const uint32_t* __restrict__ row_ptr, uint32_t n_rows, uint32_t fat_terms, uint32_t n_tiles,
                                               uint32_t* __restrict__ lct_lc, uint32_t* __restrict__ slice_rel /*n_tiles*24*/,
                                               uint32_t* __restrict__ tile_groups /*n_tiles*/) {
    __shared__ uint32_t key[1024];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < 1024u; i += 256u) {
            uint32_t k = 0;
            if (i < kLctLcs) {
                const uint64_t lc = (uint64_t)tile * kLctLcs + i;
                const uint32_t row = (uint32_t)(lc / 3u);
                uint32_t len = 0;
                if (row < n_rows && row_ptr[3 * (size_t)row + 3] - row_ptr[3 * (size_t)row] <= fat_terms) len = row_ptr[lc + 1] - row_ptr[lc];
                k = (len << 10) | (1023u - i);
            }
            key[i] = k;
        }
        __syncthreads();
        for (uint32_t size = 2; size <= 1024u; size <<= 1) {  // bitonic sort, DESCENDING
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = threadIdx.x; t < 512u; t += 256u) {
                    const uint32_t lo = 2 * t - (t & (stride - 1));  // index with the `stride` bit clear
                    const uint32_t hi = lo + stride;
                    const bool desc = (lo & size) == 0;
                    const uint32_t a = key[lo], b = key[hi];
                    if ((a < b) == desc) { key[lo] = b; key[hi] = a; }
                }
                __syncthreads();
            }
        }
        for (uint32_t p = threadIdx.x; p < kLctLcs; p += 256u) {
            const uint32_t k = key[p];
            lct_lc[(size_t)tile * kLctLcs + p] = ((k >> 10) << 16) | (1023u - (k & 1023u));
        }
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            for (uint32_t q = 0; q < kLctSlices; ++q) {
                slice_rel[(size_t)tile * kLctSlices + q] = acc;
                acc += key[32u * q] >> 10;  // the slice's longest LC (sorted: its first)
            }
            tile_groups[tile] = acc;
        }
        __syncthreads();
    }
}

// Plan, pass 2: copy the terms into their slots (term-major within a slice); slice_off = tile_base + slice_rel, in place.
__global__ void __launch_bounds__(256) lct_fill(const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ cols, const uint4* __restrict__ vals,
                                               uint32_t n_tiles, const uint32_t* __restrict__ lct_lc, const uint32_t* __restrict__ tile_base,
                                               uint32_t* __restrict__ slice_off /*in: rel, out: absolute*/, uint32_t* __restrict__ lct_cols,
                                               uint4* __restrict__ lct_vals) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t q = warp; q < kLctSlices; q += 8u) {
            const uint32_t w = lct_lc[(size_t)tile * kLctLcs + 32u * q + lane];
            const uint32_t len = w >> 16, L = __shfl_sync(0xffffffffu, len, 0);
            const uint32_t g0 = tile_base[tile] + slice_off[(size_t)tile * kLctSlices + q];
            const uint32_t k0 = row_ptr[(size_t)tile * kLctLcs + (w & 0xffffu)];  // (only read when len > 0: then the LC exists)
            for (uint32_t j = 0; j < L; ++j) {
                const size_t slot = ((size_t)g0 + j) * 32u + lane;
                if (j < len) {
                    lct_cols[slot] = cols[k0 + j];
                    lct_vals[2 * slot] = vals[2 * (size_t)(k0 + j)];
                    lct_vals[2 * slot + 1] = vals[2 * (size_t)(k0 + j) + 1];
                } else {
                    lct_cols[slot] = kLctNullCol;
                    lct_vals[2 * slot] = make_uint4(0, 0, 0, 0);
                    lct_vals[2 * slot + 1] = make_uint4(0, 0, 0, 0);
                }
            }
            __syncwarp();
            if (lane == 0) slice_off[(size_t)tile * kLctSlices + q] = g0;
        }
    }
}

struct LctView {
    const uint32_t* __restrict__ lc;         // [n_tiles * 768]
    const uint32_t* __restrict__ slice_off;  // [n_tiles * 24]
    const uint32_t* __restrict__ cols;       // [32 * n_groups]
    const uint4* __restrict__ vals;          // [2 * 32 * n_groups]
    uint32_t n_tiles;
};

// One term of a lane: column word -> class, witness element (gathered), coefficient (streamed).
struct LctTerm {
    uint32_t cls;
    uint32_t w[8], c[8];
};
// The three stages of a term's memory pipeline: its column word (coalesced stream, fetched three terms ahead), an L2 prefetch
// of its witness element and coefficient (two terms ahead: no registers held while DRAM answers), the loads proper (one term
// ahead: L2 hits by then).
__device__ __forceinline__ const uint4* lct_witness_addr(uint32_t col, const CsrView& m) {
    return ((col & kColAux) ? m.aux : m.inputs) + 2 * (size_t)(col & kColIdxMask);
}
template <int PF> __device__ __forceinline__ void lct_prefetch(uint32_t col, const LctView& v, const CsrView& m, size_t slot) {
    if (PF == 0) return;
    const uint32_t cls = (col >> kColClsShift) & 7u;
    if (cls == kClsZero) return;
    if (PF != 3) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(lct_witness_addr(col, m)));
    if (PF >= 2 && is_product_class(cls)) asm volatile("prefetch.global.L2 [%0];" ::"l"(v.vals + 2 * slot));
}
__device__ __forceinline__ void lct_fetch(LctTerm& t, uint32_t col, const LctView& v, const CsrView& m, size_t slot) {
    t.cls = (col >> kColClsShift) & 7u;
    if (t.cls == kClsZero) return;  // zero coefficient, or padding
    ld256_keep(t.w, lct_witness_addr(col, m));
    if (is_product_class(t.cls)) ld256_stream(t.c, v.vals + 2 * slot);
}
template <int F, int RIPPLE> __device__ __forceinline__ void lct_apply(uint32_t* acc, LctTerm& t, uint32_t& gen, uint32_t& mag) {
    const uint32_t cls = t.cls;
    if (cls == kClsZero) return;
    if (is_product_class(cls)) {
        mac_wide(acc, t.c, t.w);
        gen = 1;
        return;
    }
    if (cls == kClsM1 || cls == kClsM2) {
        uint32_t n[8];
        neg_mod<F>(n, t.w);
#pragma unroll
        for (int i = 0; i < 8; ++i) t.w[i] = n[i];
    }
    acc_add8<RIPPLE>(acc, t.w);
    mag += 1;
    if (cls == kClsP2 || cls == kClsM2) {
        acc_add8<RIPPLE>(acc, t.w);
        mag += 1;
    }
}

// PF: 0 = no L2 prefetch (column words three terms ahead, operands one term ahead in registers), 1 = L2 prefetch of the witness
// element two terms ahead, 2 = of the witness element and the coefficient, 3 = of the coefficient only (a prefetch always
// brings a whole 128-byte line: for the random witness element that doubles the DRAM bytes of the 64-byte-hinted gather),
// 4 = witness elements two terms ahead in REGISTERS + coefficient prefetch.
template <int F, bool EMIT, int PF>
__global__ void __launch_bounds__(kLctThreads, 2) check_lct(CsrView m, LctView v, CheckOut o, FieldConsts fc) {
    __shared__ uint32_t s_ab[2][8][kLctRows];   // A.w / B.w of the tile's rows (8 limbs, < 2^256), limb-major
    __shared__ uint32_t s_c[17][kLctRows];      // unreduced sum of the (negated) C terms
    __shared__ uint8_t s_flag[kLctRows];        // bit 0: C had a full product; bits 1..: plain magnitude of C (saturating at 8)
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    uint32_t my_bad = 0xffffffffu;
    for (uint32_t tile = blockIdx.x; tile < v.n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (uint32_t s3 = 0; s3 < 3u; ++s3) {
            // slices sorted by length: warp w takes q = w, 15 - w, 16 + w, so that every warp gets about the same number of terms
            const uint32_t q = s3 == 0 ? warp : (s3 == 1 ? 15u - warp : 16u + warp);
            const uint32_t word = __ldg(v.lc + (size_t)tile * kLctLcs + 32u * q + lane);
            const uint32_t len = word >> 16, id = word & 0xffffu;
            const uint32_t L = __shfl_sync(0xffffffffu, len, 0);
            if (L == 0) continue;  // (uniform) nothing but empty / skipped LCs: their shared-memory entries are written below
            const size_t slot0 = (size_t)__ldg(v.slice_off + (size_t)tile * kLctSlices + q) * 32u + lane;
            uint32_t acc[17];
            zeron<17>(acc);
            uint32_t gen = 0, mag = 0;
            const bool is_c = (id % 3u) == 2u;
            // software pipeline over the slice's L groups (padding groups carry the null column word)
            auto col_at = [&](uint32_t j) { return j < L ? ld32_stream(v.cols + slot0 + 32u * (size_t)j, pol) : kLctNullCol; };
            uint32_t c0 = col_at(0), c1 = col_at(1), c2 = col_at(2);
            if (PF == 4) {
                // witness elements TWO terms ahead in registers (a 64-byte DRAM fetch each), coefficients one term ahead behind an
                // L2 prefetch: the latency cover of PF = 2 without its 128-byte prefetch lines
                uint32_t wq[8];  // witness element of term j + 1, in flight
                LctTerm cur;
                lct_fetch(cur, c0, v, m, slot0);
                if (((c1 >> kColClsShift) & 7u) != kClsZero) ld256_keep(wq, lct_witness_addr(c1, m));
#pragma unroll 1
                for (uint32_t j = 0; j < L; ++j) {
                    const uint32_t c3 = col_at(j + 3u);
                    LctTerm nxt;
                    nxt.cls = (c1 >> kColClsShift) & 7u;
#pragma unroll
                    for (int i = 0; i < 8; ++i) nxt.w[i] = wq[i];
                    if (is_product_class(nxt.cls)) ld256_stream(nxt.c, v.vals + 2 * (slot0 + 32u * (size_t)(j + 1u)));
                    const uint32_t cls2 = (c2 >> kColClsShift) & 7u;
                    if (cls2 != kClsZero) ld256_keep(wq, lct_witness_addr(c2, m));
                    if (is_product_class(cls2)) asm volatile("prefetch.global.L2 [%0];" ::"l"(v.vals + 2 * (slot0 + 32u * (size_t)(j + 2u))));
                    lct_apply<F, 17>(acc, cur, gen, mag);
                    cur = nxt;
                    c1 = c2;
                    c2 = c3;
                }
            } else {
                lct_prefetch<PF>(c1, v, m, slot0 + 32u);
                LctTerm cur;
                lct_fetch(cur, c0, v, m, slot0);
#pragma unroll 1
                for (uint32_t j = 0; j < L; ++j) {
                    const uint32_t c3 = col_at(j + 3u);
                    lct_prefetch<PF>(c2, v, m, slot0 + 32u * (size_t)(j + 2u));
                    LctTerm nxt;
                    lct_fetch(nxt, c1, v, m, slot0 + 32u * (size_t)(j + 1u));
                    lct_apply<F, 17>(acc, cur, gen, mag);  // (one code path for A, B and C lanes: a slice mixes them)
                    cur = nxt;
                    c1 = c2;
                    c2 = c3;
                }
            }
            const uint32_t r = id / 3u;
            if (len) {
                if (is_c) {
#pragma unroll
                    for (int i = 0; i < 17; ++i) s_c[i][r] = acc[i];
                    s_flag[r] = (uint8_t)((gen ? 1u : 0u) | ((mag > 8u ? 8u : mag) << 1));
                } else {
                    uint32_t val[8];
                    finish_ab<F, true>(val, acc, gen, mag);
#pragma unroll
                    for (int i = 0; i < 8; ++i) s_ab[id % 3u][i][r] = val[i];
                }
            }
        }
        // LCs of length 0 were never visited by a lane with len > 0: their entries are zero
        // (three passes over the tile's LC words would cost more than writing the zeros first, so: a second, cheap sweep)
        __syncthreads();
        {
            const uint32_t r = threadIdx.x;
            const uint32_t row = tile * kLctRows + r;
            if (row < m.n_rows) {
                const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                               p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
                if (p3 - p0 <= m.fat_terms) {
                    uint32_t acc[17], az[8], bz[8];
                    uint32_t gen = 0, mag = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        az[i] = p1 > p0 ? s_ab[0][i][r] : 0u;
                        bz[i] = p2 > p1 ? s_ab[1][i][r] : 0u;
                    }
                    if (p3 > p2) {
#pragma unroll
                        for (int i = 0; i < 17; ++i) acc[i] = s_c[i][r];
                        gen = s_flag[r] & 1u;
                        mag = s_flag[r] >> 1;
                    } else {
                        zeron<17>(acc);
                    }
                    if (EMIT) {
                        uint32_t x[8];
                        if (o.cz) {
                            finish_c_canonical<F>(x, acc, fc);
                            st8(o.cz + 2 * (size_t)row, x);
                        }
                        if (o.az) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[i] = az[i];
                            canon_ab<F>(x);
                            st8(o.az + 2 * (size_t)row, x);
                        }
                        if (o.bz) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[i] = bz[i];
                            canon_ab<F>(x);
                            st8(o.bz + 2 * (size_t)row, x);
                        }
                    }
                    if (!row_satisfied<F, true>(acc, az, bz, gen, mag) && row < my_bad) my_bad = row;
                }
            }
        }
        __syncthreads();  // the tile's shared-memory results are free again
    }
    publish_first_bad(my_bad, m, o, 0u);
}

__global__ void init_result(long long* first_bad, unsigned int* err, uint32_t* n_deferred) {
    *first_bad = 0x7fffffffffffffffLL;
    *err = 0;
    if (n_deferred) { n_deferred[0] = 0; n_deferred[1] = 0; }  // thin rows deferred by check_small, fat rows left by check_fat_int
}

// ---- multi-GPU: the one-word MIN over the ranks, through peer-mapped mailboxes (no NCCL launch, no host) ------------------
// Every rank owns a mailbox of 2 x world slots in its own HBM; `boxes[r]` is rank r's mailbox as mapped into this process
// (CUDA IPC over NVLink; boxes[rank] is the local one).  Exchange number e (a device-side counter, the same on every rank
// because the call is collective) uses the slots of parity e & 1: lane j stores this rank's word into slot [e&1][rank] of
// rank j's mailbox and waits until slot [e&1][j] of its OWN mailbox carries exchange e, i.e. until rank j's word has arrived.
// A slot is two 8-byte packets {32 bits of the value, the 32-bit exchange number}: an aligned 8-byte store is one transaction,
// so a packet whose flag is current carries current data -- no fence, no second round trip (the flag-in-the-data idea of
// NCCL's LL protocol).  Parity is enough: a rank cannot start exchange e+2 before every other rank has finished reading
// exchange e (its e+1 needs their e+1, which they only start after their e).
struct GroupSlot {
    unsigned long long lo, hi;  // (exchange number << 32) | value bits 31..0 / 63..32
};
__global__ void group_exchange(long long* word, GroupSlot* const* __restrict__ boxes, GroupSlot* my_box, unsigned long long* epoch_ctr, int rank,
                               int world) {
    const int j = threadIdx.x;
    const unsigned long long e = *epoch_ctr + 1ull;
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    const int par = (int)(e & 1ull);
    const unsigned long long mine = (unsigned long long)*word;
    __syncwarp();
    long long got = 0x7fffffffffffffffLL;
    if (j < world) {
        GroupSlot* dst = boxes[j] + par * world + rank;
        asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(tag | (mine & 0xffffffffull)), "l"(tag | (mine >> 32)) : "memory");
        const GroupSlot* src = my_box + par * world + j;
        unsigned long long lo, hi;
        do {
            asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
        } while ((lo >> 32) != (tag >> 32) || (hi >> 32) != (tag >> 32));
        got = (long long)((hi << 32) | (lo & 0xffffffffull));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, got, d);
        got = o < got ? o : got;
    }
    if (j == 0) {
        *word = got;
        *epoch_ctr = e;
    }
}

// ---- witness program: the witness of a bit-logic gadget circuit generated on the device ----------------------
// Replaces, for the NEXT witness of an already synthesized circuit, the host closures of the gadgets (boolean.rs:68-272,
// 536-759; uint32.rs:306-406; sha256.rs:83-272) and the upload of their results.  The program (csrc/host/wtape.hpp builds it
// while the circuit is synthesized; layout in include/bp_r1cs.h) holds, per unit (e.g. one sha256 compression block), a tape:
// one entry per aux variable of the unit -- XOR / AND / AND_NOT / NOR / CH / MAJ of earlier values, or bit j of an integer sum
// of weighted bits (addmany) -- sorted into dependency levels.  One WARP evaluates one unit: the unit's values live in shared
// memory as a bit set (lanes OR their result bits in), a level's sums are reduced cooperatively (lanes stride the operands,
// shuffle-add), its entries are one per lane; operands outside the unit are bits of the message bytes or of the unit's
// chaining state (host-supplied: 8 words per unit).  At the end the warp writes the unit's values into the witness shadows.
constexpr uint32_t kWpMagic = 0x50575042u, kWpHeaderWords = 16;
constexpr uint32_t kWpFree = 0, kWpXor = 1, kWpAnd = 2, kWpAndNot = 3, kWpNor = 4, kWpCh = 5, kWpMaj = 6, kWpSumBit = 7;
constexpr int kWpWarps = 4;  // warps (= units) per CTA

__global__ void wprog_expand_msg(const uint8_t* __restrict__ msg, uint64_t n_bits, uint32_t msb_first, uint32_t* __restrict__ shadow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_bits; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = msg[i >> 3];
        shadow[i] = msb_first ? (b >> (7u - (uint32_t)(i & 7u))) & 1u : (b >> (uint32_t)(i & 7u)) & 1u;
    }
}

struct WpUnit {
    const uint32_t* bits;     // shared: the unit's values
    const uint8_t* msg;
    const uint32_t* state;    // 8 words
    uint32_t msg_bit_base, msb_first;
};
__device__ __forceinline__ uint32_t wp_value(uint32_t op, uint32_t payload_mask, const WpUnit& u) {
    const uint32_t kind = op >> 29, p = op & payload_mask;
    uint32_t v;
    if (kind < 2u) {
        v = kind;
    } else if (kind < 4u) {
        v = (u.bits[p >> 5] >> (p & 31u)) & 1u;
    } else if (kind < 6u) {
        const uint32_t g = u.msg_bit_base + p, b = u.msg[g >> 3];
        v = u.msb_first ? (b >> (7u - (g & 7u))) & 1u : (b >> (g & 7u)) & 1u;
    } else {
        v = (u.state[p >> 5] >> (p & 31u)) & 1u;
    }
    return (kind >= 2u && (kind & 1u)) ? v ^ 1u : v;
}

__global__ void __launch_bounds__(32 * kWpWarps) wprog_run(const uint32_t* __restrict__ prog, const uint8_t* __restrict__ msg,
                                                           const uint32_t* __restrict__ states, uint32_t* __restrict__ aux_shadow, uint32_t n_units,
                                                           uint32_t bit_words /*per warp*/, uint32_t sum_slots /*per warp*/) {
    extern __shared__ __align__(16) unsigned char wp_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned long long* sumv = reinterpret_cast<unsigned long long*>(wp_smem) + (size_t)warp * sum_slots;
    uint32_t* bits = reinterpret_cast<uint32_t*>(wp_smem + (size_t)kWpWarps * sum_slots * 8u) + (size_t)warp * bit_words;
    const uint32_t unit_off = prog[10], tape_off = prog[11], msb_first = prog[6];
    for (uint32_t unit = blockIdx.x * kWpWarps + warp; unit < n_units; unit += gridDim.x * kWpWarps) {
        const uint32_t* ur = prog + unit_off + 4u * unit;
        const uint32_t* tr = prog + tape_off + 8u * ur[0];
        const uint32_t n_vars = tr[0], n_levels = tr[2];
        const uint32_t* levels = prog + tr[1];
        const uint32_t* ents = prog + tr[3];
        const uint32_t* sums = prog + tr[5];
        const uint32_t* sumops = prog + tr[7];
        WpUnit u{bits, msg, states + 8u * ur[3], ur[2], msb_first};
        for (uint32_t i = lane; i < (n_vars + 31u) / 32u; i += 32u) bits[i] = 0u;
        __syncwarp();
        // A unit is a chain of ~10^3 short levels: what matters is the latency of ONE level, so nothing a level needs is
        // fetched when the level starts.  Level records are read 32 at a time (one per lane, handed round by shuffles); the
        // first 32 entries and the first sum record of level l+1 are loaded while level l is evaluated.
        const uint4* ents4 = reinterpret_cast<const uint4*>(ents);
        const uint4* sums4 = reinterpret_cast<const uint4*>(sums);
        uint4 ent_n = make_uint4(0, 0, 0, 0), sum_n = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (uint32_t lb = 0; lb < n_levels; lb += 32u) {
            const uint4 rec = lb + lane < n_levels ? __ldg(reinterpret_cast<const uint4*>(levels) + lb + lane) : make_uint4(0, 0, 0, 0);
            const uint32_t in_chunk = min(32u, n_levels - lb);
            if (lb == 0) {  // prime the pipeline with level 0
                const uint32_t e0 = __shfl_sync(0xffffffffu, rec.x, 0), e1 = __shfl_sync(0xffffffffu, rec.y, 0);
                const uint32_t s0 = __shfl_sync(0xffffffffu, rec.z, 0), s1 = __shfl_sync(0xffffffffu, rec.w, 0);
                if (e0 + lane < e1) ent_n = __ldg(ents4 + e0 + lane);
                if (s0 < s1) sum_n = __ldg(sums4 + s0);
            }
#pragma unroll 1
            for (uint32_t k = 0; k < in_chunk; ++k) {
                const uint32_t e0 = __shfl_sync(0xffffffffu, rec.x, k), e1 = __shfl_sync(0xffffffffu, rec.y, k);
                const uint32_t s0 = __shfl_sync(0xffffffffu, rec.z, k), s1 = __shfl_sync(0xffffffffu, rec.w, k);
                const uint4 ent_c = ent_n, sum_c = sum_n;
                if (k + 1u < in_chunk) {  // (the first level of the next chunk is fetched when its records are known: see below)
                    const uint32_t ne0 = __shfl_sync(0xffffffffu, rec.x, k + 1u), ne1 = __shfl_sync(0xffffffffu, rec.y, k + 1u);
                    const uint32_t ns0 = __shfl_sync(0xffffffffu, rec.z, k + 1u), ns1 = __shfl_sync(0xffffffffu, rec.w, k + 1u);
                    if (ne0 + lane < ne1) ent_n = __ldg(ents4 + ne0 + lane);
                    if (ns0 < ns1) sum_n = __ldg(sums4 + ns0);
                }
#pragma unroll 1
                for (uint32_t sidx = s0; sidx < s1; ++sidx) {  // integer sums of this level: lanes stride the weighted bits
                    const uint4 sr = sidx == s0 ? sum_c : __ldg(sums4 + sidx);
                    unsigned long long acc = 0;
                    for (uint32_t q = lane; q < sr.y; q += 32u) {
                        const uint32_t op = __ldg(sumops + sr.x + q);
                        acc += (unsigned long long)wp_value(op, 0x00ffffffu, u) << ((op >> 24) & 31u);
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                    if (lane == 0) sumv[sidx] = acc + (((unsigned long long)sr.w << 32) | sr.z);
                }
                if (s1 > s0) __syncwarp();
                for (uint32_t e = e0 + lane; e < e1; e += 32u) {
                    const uint4 r = e < e0 + 32u ? ent_c : __ldg(ents4 + e);
                    const uint32_t op = r.x >> 28, res = r.x & 0x0fffffffu;
                    uint32_t v;
                    if (op == kWpSumBit) {
                        v = (uint32_t)(sumv[r.y] >> r.z) & 1u;
                    } else {
                        const uint32_t a = wp_value(r.y, 0x1fffffffu, u), b = wp_value(r.z, 0x1fffffffu, u);
                        if (op == kWpXor) v = a ^ b;
                        else if (op == kWpAnd) v = a & b;
                        else if (op == kWpAndNot) v = a & (b ^ 1u);
                        else if (op == kWpNor) v = (a ^ 1u) & (b ^ 1u);
                        else {
                            const uint32_t c = wp_value(r.w, 0x1fffffffu, u);
                            v = op == kWpCh ? ((a & b) ^ ((a ^ 1u) & c)) : ((a & b) ^ (a & c) ^ (b & c));
                        }
                    }
                    if (v) atomicOr(bits + (res >> 5), 1u << (res & 31u));
                }
                __syncwarp();
            }
            if (lb + 32u < n_levels) {  // first level of the next chunk
                const uint4 nr = __ldg(reinterpret_cast<const uint4*>(levels) + lb + 32u);
                if (nr.x + lane < nr.y) ent_n = __ldg(ents4 + nr.x + lane);
                if (nr.z < nr.w) sum_n = __ldg(sums4 + nr.z);
            }
        }
        uint32_t* out = aux_shadow + ur[1];
        for (uint32_t i = lane; i < n_vars; i += 32u) out[i] = (bits[i >> 5] >> (i & 31u)) & 1u;
        __syncwarp();
    }
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
            if (cls == kClsP1 || cls == kClsM1) mag += 1;
            else if (cls == kClsP2 || cls == kClsM2) mag += 2;
            else if (cls != kClsZero) plain = 0;
            if (mag > 7) plain = 0;
        }
        kind[i] = (uint8_t)plain;
    }
}

// Pass 2, one thread per term: canonical -> internal form in place, class bits into the column word.
// err bit 1: coefficient >= p; bit 2: variable index does not fit 28 bits.
template <int F>
__global__ void convert_terms(uint4* vals, uint32_t* cols, uint16_t* __restrict__ kexp, const uint32_t* __restrict__ row_ptr,
                              const uint8_t* __restrict__ kind, uint32_t lc0, uint32_t n_lc, uint32_t k0, uint32_t n, FieldConsts fc, unsigned int* err) {
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
        if (type != 2u && !plain) {  // a general A/B LC is scaled as a whole: its +-1 / +-2 terms are +-2^0 / +-2^1
            if (cls == kClsP1 || cls == kClsP2) { s = (cls == kClsP2 ? 1u : 0u) | (kNoExp << 8); cls = kClsPow2P; }
            else if (cls == kClsM1 || cls == kClsM2) { s = (cls == kClsM2 ? 1u : 0u) | (kNoExp << 8); cls = kClsPow2M; }
        }
        if (cls == kClsGen || cls == kClsPow2P || cls == kClsPow2M) {
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
        }
        st8(vals + 2 * (size_t)k, r);
        kexp[k] = (uint16_t)s;
        cols[k] = (col & (kColAux | kColIdxMask)) | (cls << kColClsShift);
        // instance statistic for the launch heuristic: how many terms need a full product (err[1] is the counter)
        const unsigned int gen_mask = __ballot_sync(__activemask(), cls == kClsGen);
        if (cls == kClsGen && (threadIdx.x & 31u) == (unsigned)(__ffs(gen_mask) - 1)) atomicAdd(err + 1, (unsigned int)__popc(gen_mask));
    }
}

// bp_cs_export: terms [k0, k0+n) back from the internal form to CANONICAL coefficients (the inverse of convert_terms).
//   A/B terms: classes P1/M1/P2/M2 are +-1, +-2 of a plain LC (nothing stored); GEN / POW2* are stored as c * 2^288: one lazy
//   reduction undoes the scaling.   C terms: the class and the stored value describe the NEGATED coefficient.
template <int F>
__global__ void export_terms(const uint4* __restrict__ vals, const uint32_t* __restrict__ cols, const uint32_t* __restrict__ row_ptr, uint32_t n_lc,
                             uint32_t k0, uint32_t n, uint4* __restrict__ out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = k0 + i;
        uint32_t lo = 0, hi = n_lc;  // largest lc with row_ptr[lc] <= k
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (__ldg(row_ptr + mid) <= k) lo = mid; else hi = mid;
        }
        const bool is_c = lo % 3u == 2u;
        const uint32_t cls = (__ldg(cols + k) >> kColClsShift) & 7u;
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (cls == kClsP1 || cls == kClsM1) c[0] = 1u;
        else if (cls == kClsP2 || cls == kClsM2) c[0] = 2u;
        else if (cls != kClsZero) {
            ld8(c, vals + 2 * (size_t)k);
            if (!is_c) {
                uint32_t t[17];
#pragma unroll
                for (int j = 0; j < 17; ++j) t[j] = j < 8 ? c[j] : 0u;
                redc_acc<F>(c, t);  // stored * 2^-288, in [0, 2p)
                reduce_once<F>(c);
            }
        }
        // the sign: M classes are negative; for C everything describes -coefficient
        const bool stored_negative = cls == kClsM1 || cls == kClsM2;
        const bool flip = cls != kClsZero && (stored_negative != is_c);
        uint32_t nz = 0, ng[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) nz |= c[j];
        neg_mod<F>(ng, c);
        if (flip && nz) {
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = ng[j];
        }
        st8(out + 2 * (size_t)i, c);
    }
}

// Witness upload: value < p check, and the 4-byte shadow of every element.
template <int F> __global__ void validate_canonical(const uint4* __restrict__ v, uint64_t n, unsigned int* err, uint32_t* __restrict__ shadow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x[8];
        ld8(x, v + 2 * i);
        if (!is_canonical<F>(x)) atomicOr(err, 2u);
        shadow[i] = shadow_of(x);
    }
}

// K4 witness_patch: scatter n updated elements (test_cs.rs:270-282 `set`, batched): element idx[i] of one index space becomes
// vals[i] in both forms (32-byte element, shadow).  The host has validated the indices and values.
__global__ void patch_witness(const uint32_t* __restrict__ idx, const uint4* __restrict__ vals, uint32_t n, uint4* __restrict__ elems,
                              uint32_t* __restrict__ shadow) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t x[8];
        ld8(x, vals + 2 * (size_t)i);
        const uint32_t k = idx[i];
        st8(elems + 2 * (size_t)k, x);
        shadow[k] = shadow_of(x);
    }
}

// bp_cs_load: the offsets every kernel indexes with.  err bit 3: row_ptr[0] != 0, a decrease, or row_ptr[last] != nnz.
__global__ void validate_row_ptr(const uint32_t* __restrict__ row_ptr, uint64_t n /*entries*/, uint32_t nnz, unsigned int* err) {
    bool bad = false;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = row_ptr[i];
        if (i == 0 && v != 0u) bad = true;
        if (i + 1 == n && v != nnz) bad = true;
        if (i + 1 < n && row_ptr[i + 1] < v) bad = true;
    }
    if (bad) atomicOr(err, 8u);
}

// Packed witness upload (bp_cs_alloc_u8 / bp_cs_set_range_u8 / bp_cs_recheck_u8): element i = bytes[i].  Only the shadow is
// written (the handle's wide_valid flag goes to 0: see ld_witness); the 32-byte element is materialised when someone asks
// for it (bp_cs_witness / bp_cs_get).
__global__ void widen_u8(const uint8_t* __restrict__ bytes, uint64_t n, uint32_t* __restrict__ shadow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) shadow[i] = bytes[i];
}
// Same for the 1-bit-per-value form (bp_cs_set_range_bits / bp_cs_recheck_bits): element i = bit i of the byte string,
// least significant bit of a byte first.
__global__ void widen_bits(const uint8_t* __restrict__ bytes, uint64_t n, uint32_t* __restrict__ shadow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        shadow[i] = (bytes[i >> 3] >> (i & 7u)) & 1u;
}
__global__ void materialize_wide(const uint32_t* __restrict__ shadow, uint64_t n, uint4* __restrict__ out) {
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < 2 * n; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s = shadow[j >> 1];
        if (s != kShadowBig) out[j] = make_uint4((j & 1) ? 0u : s, 0, 0, 0);
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
        ld_witness(w, m, is_aux, idx);
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
__global__ void synth_witness(uint4* inputs, uint4* aux, uint32_t* inputs_s, uint32_t* aux_s, uint64_t seed, uint64_t n_vars, uint64_t n_inputs) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_vars; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t v[8];
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = j == 0 ? 1u : 0u;
        } else {
            sm_sample<F>(sm_key(seed, 4, i, 0), v);
        }
        st8((i < n_inputs ? inputs + 2 * i : aux + 2 * (i - n_inputs)), v);
        *(i < n_inputs ? inputs_s + i : aux_s + (i - n_inputs)) = shadow_of(v);
    }
}

}  // namespace bp
