// CUDA kernels of the R1CS evaluation engine (sm_100a).  See DESIGN.md for the data layout.
//
//   K1 check_direct<F,false>   which_is_unsatisfied           test_cs.rs:239-253 (+ eval_lc :137-155)
//   K2 check_direct<F,true>    batched LinearCombination::eval lc.rs:245-267 (also yields the flag)
//   K3 to_internal<F>          canonical coefficient -> device-internal pre-scaled form (ff::PrimeField::from_repr)
//      validate_canonical<F>   value < p check for witness uploads
//   K5 synth_*                 synthetic instance generator (measurement fixture)
//      eval_lc_kernel<F>       one ad-hoc LinearCombination::eval
//
// Device layout (all per handle == per row shard):
//   row_ptr : u32[3N+1]   LC offsets; LC 3i, 3i+1, 3i+2 are A_i, B_i, C_i; their terms are contiguous
//   cols    : u32[nnz]    tagged column (bit 31 = aux index space)
//   vals    : uint4[2nnz] coefficient, 8 x u32 limbs, INTERNAL form:  A: c*2^288, B: c*2^576, C: -c*2^288  (mod p)
//   inputs  : uint4[2*n_inputs], aux : uint4[2*n_aux]   canonical witness
//
// Per row, check mode:  Az = redc(acc_A)            (= A.w,         in [0,2p))
//                       Bm = redc(acc_B)            (= B.w * 2^288, in [0,2p))
//                       Y  = redc(acc_C + Az*Bm)    (= (Az*Bz - Cz) mod p up to a multiple of p, in [0,2p))
//                       satisfied  <=>  Y in {0, p}
// i.e. T+1 wide multiply-accumulates and 3 lazy reductions per row; no per-term reduction.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "field.cuh"
#include "synth.cuh"

namespace bp {

struct FieldConsts {
    uint32_t kA[8];  // 2^544 mod p : mont_mul(c, kA) = c * 2^288
    uint32_t kB[8];  // 2^832 mod p : mont_mul(c, kB) = c * 2^576
    uint32_t kC[8];  // p - kA      : mont_mul(c, kC) = -c * 2^288
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
    unsigned long long row_base;
};

struct CheckOut {
    long long* first_bad;  // global row index, atomicMin; INT64_MAX = satisfied
    unsigned int* err;     // bit 0: a column was out of range
    uint4* az;             // emit mode only (nullable)
    uint4* bz;
    uint4* cz;
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

// acc += sum over terms [k0,k1) of vals[k] * w[cols[k]]
__device__ __forceinline__ void lc_accumulate(uint32_t* acc, uint32_t k0, uint32_t k1, const CsrView& m, unsigned int& err) {
#pragma unroll 1
    for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t col = __ldg(m.cols + k);
        const uint32_t idx = col & 0x7fffffffu;
        const bool is_aux = (col >> 31) != 0;
        if (idx >= (is_aux ? m.n_aux : m.n_inputs)) { err = 1; continue; }
        uint32_t c[8], w[8];
        ld8(w, (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx);
        ld8(c, m.vals + 2 * (size_t)k);
        mac_wide(acc, c, w);
    }
}

__device__ __forceinline__ void zero17(uint32_t* acc) {
#pragma unroll
    for (int i = 0; i < 17; ++i) acc[i] = 0;
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

// ---- K1 / K2: one thread per constraint, CSR read straight from global memory ---------------------------
template <int F, bool EMIT>
__global__ void __launch_bounds__(128, 4) check_direct(CsrView m, CheckOut o) {
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < m.n_rows; row += gridDim.x * blockDim.x) {
        const uint32_t p0 = __ldg(m.row_ptr + 3 * (size_t)row), p1 = __ldg(m.row_ptr + 3 * (size_t)row + 1),
                       p2 = __ldg(m.row_ptr + 3 * (size_t)row + 2), p3 = __ldg(m.row_ptr + 3 * (size_t)row + 3);
        uint32_t acc[17], az[8], bm[8], y[8];
        zero17(acc);
        lc_accumulate(acc, p0, p1, m, my_err);
        redc_acc<F>(az, acc);
        zero17(acc);
        lc_accumulate(acc, p1, p2, m, my_err);
        redc_acc<F>(bm, acc);
        zero17(acc);
        lc_accumulate(acc, p2, p3, m, my_err);
        if (EMIT) {
            // canonical outputs: Az, Bz = redc(Bm), Cz = -redc(acc_C)
            uint32_t t[17], v[8];
            if (o.cz) {
                redc_acc<F>(v, acc);
                reduce_once<F>(v);
                uint32_t pl[8], neg[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) pl[i] = PL<F>(i);
                (void)subn<8>(neg, pl, v);
                uint32_t nz = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) nz |= v[i];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = nz ? neg[i] : 0u;
                st8(o.cz + 2 * (size_t)row, v);
            }
            if (o.bz) {
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = bm[i];
#pragma unroll
                for (int i = 8; i < 17; ++i) t[i] = 0;
                redc_acc<F>(v, t);
                reduce_once<F>(v);
                st8(o.bz + 2 * (size_t)row, v);
            }
            if (o.az) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = az[i];
                reduce_once<F>(v);
                st8(o.az + 2 * (size_t)row, v);
            }
        }
        mac_wide(acc, az, bm);
        redc_acc<F>(y, acc);
        if (!is_zero_mod_p<F>(y) && row < my_bad) my_bad = row;
    }
    publish_first_bad(my_bad, m, o, my_err);
}

__global__ void init_result(long long* first_bad, unsigned int* err) {
    *first_bad = 0x7fffffffffffffffLL;
    *err = 0;
}

// ---- K3: canonical -> internal form, in place, for terms [k0, k0+n) of LCs [lc0, lc0+n_lc) -----------
// lc type = lc index mod 3 (A, B, C).  The term's LC is found by binary search in row_ptr.
template <int F>
__global__ void to_internal(uint4* vals, const uint32_t* __restrict__ row_ptr, uint32_t lc0, uint32_t n_lc, uint32_t k0,
                            uint32_t n, FieldConsts fc, unsigned int* err) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = k0 + i;
        // largest lc in [lc0, lc0+n_lc) with row_ptr[lc] <= k
        uint32_t lo = lc0, hi = lc0 + n_lc;
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (__ldg(row_ptr + mid) <= k) lo = mid; else hi = mid;
        }
        const uint32_t type = lo % 3u;
        uint32_t c[8], r[8];
        ld8(c, vals + 2 * (size_t)k);
        if (!is_canonical<F>(c)) { atomicOr(err, 2u); }
        const uint32_t* kk = type == 0 ? fc.kA : (type == 1 ? fc.kB : fc.kC);
        uint32_t kr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) kr[j] = kk[j];
        mont_mul<F>(r, c, kr);
        st8(vals + 2 * (size_t)k, r);
    }
}

template <int F> __global__ void validate_canonical(const uint4* __restrict__ v, uint64_t n, unsigned int* err) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t x[8];
        ld8(x, v + 2 * i);
        if (!is_canonical<F>(x)) atomicOr(err, 2u);
    }
}

// row_ptr[lc0 + i] += base for i in [0, n)   (after an exclusive scan of the chunk's lens)
__global__ void add_base(uint32_t* row_ptr, uint32_t n, uint32_t base) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) row_ptr[i] += base;
}

// ---- one ad-hoc LC (lc.rs:245-267): single warp, lanes stride the terms, then a serial fold -------------
template <int F>
__global__ void eval_lc_kernel(const uint32_t* __restrict__ cols, const uint4* __restrict__ vals_internal, uint32_t n,
                               CsrView m, uint4* out, unsigned int* err) {
    __shared__ uint32_t part[32][17];
    uint32_t acc[17];
    zero17(acc);
    unsigned int my_err = 0;
    for (uint32_t k = threadIdx.x; k < n; k += 32) {
        const uint32_t col = cols[k], idx = col & 0x7fffffffu;
        const bool is_aux = (col >> 31) != 0;
        if (idx >= (is_aux ? m.n_aux : m.n_inputs)) { my_err = 1; continue; }
        uint32_t c[8], w[8];
        ld8(w, (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx);
        ld8(c, vals_internal + 2 * (size_t)k);
        mac_wide(acc, c, w);
    }
#pragma unroll
    for (int i = 0; i < 17; ++i) part[threadIdx.x][i] = acc[i];
    if (my_err) atomicOr(err, 1u);
    __syncwarp();
    if (threadIdx.x == 0) {
        for (int l = 1; l < 32; ++l) {
            uint32_t t[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) t[i] = part[l][i];
            (void)addn<17>(acc, acc, t);
        }
        uint32_t v[8];
        redc_acc<F>(v, acc);
        reduce_once<F>(v);
        st8(out, v);
    }
}

// ---- K5: synthetic generator -------------------------------------------------------------------------------
__global__ void synth_lens(uint32_t* lens, uint64_t seed, uint32_t t, uint64_t lc0, uint32_t n_lc) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lc; i += gridDim.x * blockDim.x)
        lens[i] = sm_len(seed, t, lc0 + i);
}

// One thread per LC: writes its tagged columns and INTERNAL-form coefficients.
template <int F>
__global__ void synth_fill(uint32_t* cols, uint4* vals, const uint32_t* __restrict__ row_ptr, uint32_t lc_first /*index into row_ptr*/,
                           uint32_t n_lc, uint64_t seed, uint64_t lcid0 /*global LC id of lc_first*/, uint64_t n_vars,
                           uint64_t n_inputs, FieldConsts fc) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_lc; i += gridDim.x * blockDim.x) {
        const uint32_t k0 = row_ptr[lc_first + i], len = row_ptr[lc_first + i + 1] - k0;
        const uint64_t lcid = lcid0 + i;
        const uint32_t type = (uint32_t)(lcid % 3ull);
        uint32_t kr[8];
        const uint32_t* kk = type == 0 ? fc.kA : (type == 1 ? fc.kB : fc.kC);
#pragma unroll
        for (int j = 0; j < 8; ++j) kr[j] = kk[j];
        for (uint32_t k = 0; k < len; ++k) {
            cols[k0 + k] = sm_col(seed, lcid, k, len, n_vars, n_inputs);
            uint32_t c[8], r[8];
            sm_sample<F>(sm_key(seed, 3, lcid, k), c);
            mont_mul<F>(r, c, kr);
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
