/* TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * Plain-C CPU restatement of the reference's R1CS satisfaction check over a flat CSR, plus the
 * counter-based synthetic-instance recipe shared with the CUDA generator.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may load this.
 *
 * Restates (paths relative to /root/reference):
 *   eval_lc                 crates/bellpepper-core/src/util_cs/test_cs.rs:137-155   (always multiply, then add)
 *   which_is_unsatisfied    crates/bellpepper-core/src/util_cs/test_cs.rs:239-253   (sequential rows, first failure)
 *   LinearCombination::eval crates/bellpepper-core/src/lc.rs:245-267                (same dot product)
 * The field arithmetic lives in third-party crates that are NOT under /root/reference
 * (ff 0.13.0 trait, blstrs 0.7.0 Scalar: Cargo.toml:10,12).  blstrs keeps elements in Montgomery
 * form (R = 2^256) over 4 x u64 limbs; this file restates that published algorithm (CIOS) so the
 * per-term work -- one Montgomery multiplication and one modular addition -- is the reference's.
 *
 * Pinning: BLS12-381 Fr is pinned by the reference's known-answer tests via tests/test_oracle_kat.py
 * (this C code must agree with oracle/r1cs_py.py, which carries the KATs).  Pallas Fr / Vesta Fr:
 * PARITY UNPINNED (no reference test or dependency exists for them).
 *
 * Threading: the reference loop is single-threaded with early exit.  `threads > 1` runs the same
 * per-row code under OpenMP over contiguous row blocks ("stand-in for the rayon path north_star
 * names; the reference has none").  `early_exit = 0` makes every row do its work (for timing).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

typedef struct {
    uint64_t p[4];
    uint64_t inv;   /* -p^-1 mod 2^64 */
    uint64_t r2[4]; /* 2^512 mod p   */
} bpo_field;

/* Constants derived from the three moduli by oracle/fields.py (checked in tests/test_oracle_field.py). */
static const bpo_field FIELDS[3] = {
    /* 0: BLS12-381 Fr */
    {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
     0xfffffffeffffffffULL,
     {0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}},
    /* 1: Pallas Fr = pasta Fq */
    {{0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0000000000000000ULL, 0x4000000000000000ULL},
     0x8c46eb20ffffffffULL,
     {0xfc9678ff0000000fULL, 0x67bb433d891a16e3ULL, 0x7fae231004ccf590ULL, 0x096d41af7ccfdaa9ULL}},
    /* 2: Vesta Fr = pasta Fp */
    {{0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0000000000000000ULL, 0x4000000000000000ULL},
     0x992d30ecffffffffULL,
     {0x8c78ecb30000000fULL, 0xd7d30dbd8b0de0e7ULL, 0x7797a99bc3c95d18ULL, 0x096d41af7b9cb714ULL}},
};

int bpo_field_params(int field, uint64_t p[4], uint64_t *inv, uint64_t r2[4]) {
    if (field < 0 || field > 2) return -1;
    memcpy(p, FIELDS[field].p, 32);
    *inv = FIELDS[field].inv;
    memcpy(r2, FIELDS[field].r2, 32);
    return 0;
}

static inline int geq(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

static inline void sub_nb(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}

static inline void fadd(const bpo_field *f, uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; ++i) {
        c += (u128)a[i] + b[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    /* p < 2^255 so a + b < 2^256: no carry-out */
    if (geq(t, f->p)) sub_nb(r, t, f->p);
    else memcpy(r, t, 32);
}

/* CIOS Montgomery multiplication, R = 2^256: r = a * b / R mod p */
static inline void mmul(const bpo_field *f, uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * f->inv;
        c = ((u128)m * f->p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * f->p[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq(t, f->p)) sub_nb(r, t, f->p);
    else memcpy(r, t, 32);
}

static inline void to_mont(const bpo_field *f, uint64_t r[4], const uint64_t a[4]) { mmul(f, r, a, f->r2); }
static inline void from_mont(const bpo_field *f, uint64_t r[4], const uint64_t a[4]) {
    static const uint64_t one[4] = {1, 0, 0, 0};
    mmul(f, r, a, one);
}

/* ---- scalar entry points for unit tests against Python `%` ------------------------------- */
int bpo_mul(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
    if (field < 0 || field > 2) return -1;
    const bpo_field *f = &FIELDS[field];
    uint64_t am[4], bm[4], rm[4];
    to_mont(f, am, a);
    to_mont(f, bm, b);
    mmul(f, rm, am, bm);
    from_mont(f, r, rm);
    return 0;
}
int bpo_add(int field, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]) {
    if (field < 0 || field > 2) return -1;
    fadd(&FIELDS[field], r, a, b);
    return 0;
}
int bpo_is_canonical(int field, const uint64_t a[4]) {
    if (field < 0 || field > 2) return -1;
    return !geq(a, FIELDS[field].p);
}

/* ---- prepared instance: Montgomery-form coefficients and witness, as blstrs holds them ---- */
typedef struct {
    int field;
    uint64_t n_rows, nnz, n_inputs, n_aux;
    uint64_t *row_ptr;   /* 3*n_rows + 1 LC offsets */
    uint32_t *cols;      /* bit 31 = aux */
    uint64_t *coeffs_m;  /* nnz x 4, Montgomery */
    uint64_t *inputs_m;  /* Montgomery */
    uint64_t *aux_m;
} bpo_instance;

void bpo_free(bpo_instance *it) {
    if (!it) return;
    free(it->row_ptr); free(it->cols); free(it->coeffs_m); free(it->inputs_m); free(it->aux_m);
    free(it);
}

/* Returns NULL on bad arguments: non-canonical element, column out of range, OOM. */
bpo_instance *bpo_prepare(int field, uint64_t n_rows, const uint32_t *lens /*3 per row*/,
                          const uint32_t *cols, const uint64_t *coeffs /*nnz x 4 canonical*/,
                          const uint64_t *inputs, uint64_t n_inputs,
                          const uint64_t *aux, uint64_t n_aux) {
    if (field < 0 || field > 2) return NULL;
    const bpo_field *f = &FIELDS[field];
    bpo_instance *it = (bpo_instance *)calloc(1, sizeof(*it));
    if (!it) return NULL;
    it->field = field; it->n_rows = n_rows; it->n_inputs = n_inputs; it->n_aux = n_aux;
    it->row_ptr = (uint64_t *)malloc((3 * n_rows + 1) * sizeof(uint64_t));
    if (!it->row_ptr) { bpo_free(it); return NULL; }
    uint64_t nnz = 0;
    for (uint64_t i = 0; i < 3 * n_rows; ++i) { it->row_ptr[i] = nnz; nnz += lens[i]; }
    it->row_ptr[3 * n_rows] = nnz;
    it->nnz = nnz;
    it->cols = (uint32_t *)malloc((nnz ? nnz : 1) * sizeof(uint32_t));
    it->coeffs_m = (uint64_t *)malloc((nnz ? nnz : 1) * 32);
    it->inputs_m = (uint64_t *)malloc((n_inputs ? n_inputs : 1) * 32);
    it->aux_m = (uint64_t *)malloc((n_aux ? n_aux : 1) * 32);
    if (!it->cols || !it->coeffs_m || !it->inputs_m || !it->aux_m) { bpo_free(it); return NULL; }
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t k = 0; k < (int64_t)nnz; ++k) {
        uint32_t c = cols[k];
        uint64_t idx = c & 0x7fffffffu;
        if ((c >> 31) ? (idx >= n_aux) : (idx >= n_inputs)) bad |= 1;
        if (geq(coeffs + 4 * k, f->p)) bad |= 1;
        it->cols[k] = c;
        to_mont(f, it->coeffs_m + 4 * k, coeffs + 4 * k);
    }
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t k = 0; k < (int64_t)n_inputs; ++k) {
        if (geq(inputs + 4 * k, f->p)) bad |= 1;
        to_mont(f, it->inputs_m + 4 * k, inputs + 4 * k);
    }
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t k = 0; k < (int64_t)n_aux; ++k) {
        if (geq(aux + 4 * k, f->p)) bad |= 1;
        to_mont(f, it->aux_m + 4 * k, aux + 4 * k);
    }
    if (bad) { bpo_free(it); return NULL; }
    return it;
}

/* test_cs.rs:270-282 `set`: overwrite one witness element (canonical in). */
int bpo_set(bpo_instance *it, int is_aux, uint64_t idx, const uint64_t v[4]) {
    const bpo_field *f = &FIELDS[it->field];
    if (geq(v, f->p)) return -1;
    if (is_aux ? idx >= it->n_aux : idx >= it->n_inputs) return -1;
    to_mont(f, (is_aux ? it->aux_m : it->inputs_m) + 4 * idx, v);
    return 0;
}

/* test_cs.rs:137-155: acc = sum over terms (inputs-then-aux order as stored) of w[var] * coeff */
static inline void eval_lc(const bpo_instance *it, const bpo_field *f, uint64_t lo, uint64_t hi, uint64_t acc[4]) {
    acc[0] = acc[1] = acc[2] = acc[3] = 0;
    for (uint64_t k = lo; k < hi; ++k) {
        uint32_t c = it->cols[k];
        const uint64_t *w = ((c >> 31) ? it->aux_m : it->inputs_m) + 4 * (uint64_t)(c & 0x7fffffffu);
        uint64_t tmp[4];
        mmul(f, tmp, w, it->coeffs_m + 4 * k);
        fadd(f, acc, acc, tmp);
    }
}

/* test_cs.rs:239-253.  Returns the first unsatisfied row, or -1.
 * az/bz/cz (nullable, n_rows x 4 each) receive canonical A.w, B.w, C.w when given.
 * threads <= 1: the reference's sequential loop.  early_exit applies to the sequential loop only. */
int64_t bpo_check(const bpo_instance *it, int threads, int early_exit,
                  uint64_t *az, uint64_t *bz, uint64_t *cz) {
    const bpo_field *f = &FIELDS[it->field];
    int64_t first_bad = INT64_MAX;
    const int64_t n = (int64_t)it->n_rows;
    if (threads <= 1) {
        for (int64_t i = 0; i < n; ++i) {
            uint64_t a[4], b[4], c[4], ab[4];
            eval_lc(it, f, it->row_ptr[3 * i], it->row_ptr[3 * i + 1], a);
            eval_lc(it, f, it->row_ptr[3 * i + 1], it->row_ptr[3 * i + 2], b);
            eval_lc(it, f, it->row_ptr[3 * i + 2], it->row_ptr[3 * i + 3], c);
            if (az) from_mont(f, az + 4 * i, a);
            if (bz) from_mont(f, bz + 4 * i, b);
            if (cz) from_mont(f, cz + 4 * i, c);
            mmul(f, ab, a, b);
            if (memcmp(ab, c, 32) != 0) {
                if (i < first_bad) first_bad = i;
                if (early_exit) break;
            }
        }
    } else {
#ifdef _OPENMP
        omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(min : first_bad)
        for (int64_t i = 0; i < n; ++i) {
            uint64_t a[4], b[4], c[4], ab[4];
            eval_lc(it, f, it->row_ptr[3 * i], it->row_ptr[3 * i + 1], a);
            eval_lc(it, f, it->row_ptr[3 * i + 1], it->row_ptr[3 * i + 2], b);
            eval_lc(it, f, it->row_ptr[3 * i + 2], it->row_ptr[3 * i + 3], c);
            if (az) from_mont(f, az + 4 * i, a);
            if (bz) from_mont(f, bz + 4 * i, b);
            if (cz) from_mont(f, cz + 4 * i, c);
            mmul(f, ab, a, b);
            if (memcmp(ab, c, 32) != 0 && i < first_bad) first_bad = i;
        }
    }
    return first_bad == INT64_MAX ? -1 : first_bad;
}

int bpo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ==== synthetic instances: counter-based recipe, identical in oracle/synth.py and the CUDA generator ====
 *
 *   mix(x)            SplitMix64 finaliser
 *   g(seed,s,i,j)     mix(mix(mix(seed ^ (s << 56)) + i) + j)            s = stream id
 *   len(row,lc)       1 + g(seed,1,3*row+lc,0) mod (2t-1)                  in [1, 2t-1], mean t
 *   col(row,lc,k)     stratified: lo = k*n/len, hi = (k+1)*n/len, col = lo + g(seed,2,3*row+lc,k) mod (hi-lo)
 *                     -> strictly ascending, unique (the reference's LC invariant, lc.rs:74-113)
 *   unified col < n_inputs -> Input(col) else Aux(col - n_inputs)
 *   coeff(row,lc,k)   sample(g(seed,3,3*row+lc,k));  witness(i) = i == 0 ? 1 : sample(g(seed,4,i,0))
 *   sample(h)         for a = 0..63: limbs l_k = mix(h + 4a + k + 1), top limb masked to 63 bits; accept if < p.
 *                     (after 64 rejections: the last draw with its top limb >> 2; probability <= 2^-64)
 */
static inline uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t gkey(uint64_t seed, uint64_t s, uint64_t i, uint64_t j) {
    return mix(mix(mix(seed ^ (s << 56)) + i) + j);
}
static inline void sample(const bpo_field *f, uint64_t h, uint64_t out[4]) {
    for (int a = 0; a < 64; ++a) {
        for (int k = 0; k < 4; ++k) out[k] = mix(h + 4 * (uint64_t)a + (uint64_t)k + 1);
        out[3] &= 0x7fffffffffffffffULL;
        if (!geq(out, f->p)) return;
    }
    out[3] >>= 2;
}

uint32_t bpo_synth_len(uint64_t seed, uint32_t t, uint64_t row, int lc) {
    return 1 + (uint32_t)(gkey(seed, 1, 3 * row + (uint64_t)lc, 0) % (2 * (uint64_t)t - 1));
}

/* lens for rows [row0, row0+n_rows): 3 per row.  Returns total nnz. */
uint64_t bpo_synth_lens(uint64_t seed, uint32_t t, uint64_t row0, uint64_t n_rows, uint32_t *lens) {
    uint64_t nnz = 0;
    for (uint64_t r = 0; r < n_rows; ++r)
        for (int lc = 0; lc < 3; ++lc) {
            uint32_t l = bpo_synth_len(seed, t, row0 + r, lc);
            lens[3 * r + lc] = l;
            nnz += l;
        }
    return nnz;
}

/* Fill cols (tagged) and canonical coeffs for rows [row0, row0+n_rows); buffers sized from bpo_synth_lens. */
int bpo_synth_fill(int field, uint64_t seed, uint32_t t, uint64_t n_vars, uint64_t n_inputs,
                   uint64_t row0, uint64_t n_rows, uint32_t *cols, uint64_t *coeffs) {
    if (field < 0 || field > 2 || n_vars < 2 * (uint64_t)t || n_inputs > n_vars) return -1;
    const bpo_field *f = &FIELDS[field];
    uint64_t *off = (uint64_t *)malloc((3 * n_rows + 1) * sizeof(uint64_t));
    if (!off) return -1;
    uint64_t nnz = 0;
    for (uint64_t i = 0; i < 3 * n_rows; ++i) {
        off[i] = nnz;
        nnz += bpo_synth_len(seed, t, row0 + i / 3, (int)(i % 3));
    }
    off[3 * n_rows] = nnz;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)(3 * n_rows); ++i) {
        uint64_t lcid = 3 * row0 + (uint64_t)i;
        uint64_t len = off[i + 1] - off[i];
        for (uint64_t k = 0; k < len; ++k) {
            uint64_t lo = (k * n_vars) / len, hi = ((k + 1) * n_vars) / len;
            uint64_t col = lo + gkey(seed, 2, lcid, k) % (hi - lo);
            cols[off[i] + k] = col < n_inputs ? (uint32_t)col : ((uint32_t)(col - n_inputs) | 0x80000000u);
            sample(f, gkey(seed, 3, lcid, k), coeffs + 4 * (off[i] + k));
        }
    }
    free(off);
    return 0;
}

/* Witness elements with unified index in [i0, i0+n): canonical. */
int bpo_synth_witness(int field, uint64_t seed, uint64_t i0, uint64_t n, uint64_t *out) {
    if (field < 0 || field > 2) return -1;
    const bpo_field *f = &FIELDS[field];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t idx = i0 + (uint64_t)i;
        if (idx == 0) { out[0] = 1; out[1] = out[2] = out[3] = 0; continue; }
        sample(f, gkey(seed, 4, idx, 0), out + 4 * i);
    }
    return 0;
}

/* Witness elements at arbitrary unified indices idx[0..n): canonical.  (Sampled parity checks of instances whose whole
 * witness would not be worth materialising on the host: only the elements some sampled row reads are generated.) */
int bpo_synth_witness_at(int field, uint64_t seed, const uint64_t *idx, uint64_t n, uint64_t *out) {
    if (field < 0 || field > 2) return -1;
    const bpo_field *f = &FIELDS[field];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        if (idx[i] == 0) { out[4 * i] = 1; out[4 * i + 1] = out[4 * i + 2] = out[4 * i + 3] = 0; continue; }
        sample(f, gkey(seed, 4, idx[i], 0), out + 4 * i);
    }
    return 0;
}
