// Warp-staged variants of the check kernels: every global access is either coalesced or explicitly asynchronous.
//
// The direct kernels (kernels.cuh) walk  row_ptr -> cols -> witness  as a chain of dependent loads per TERM and are
// latency-bound (IPC 0.25-0.49, DRAM <= 34 % in profiles/r1_ncu_full_check_rows_fat_sha256x512_v1.csv).  Here a warp
// stages the operands of a whole group of rows (or a 256-term chunk of a fat LC) in shared memory first:
//   * row_ptr and cols: coalesced loads by the 32 lanes,
//   * witness elements: one `cp.async` (LDGSTS, 2 x 16 B) per term, all in flight together,
//   * coefficients of a fat LC: 2 x 16 B `cp.async` per term, lane-contiguous (1 KB per warp instruction),
// so a group pays three memory round trips instead of three per term, and the arithmetic then runs out of shared memory.
// Same arithmetic, same verdicts as the direct kernels (the parity tests run both).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace bp {

constexpr uint32_t kStageCap = 256;   // terms staged per warp and round (thread-per-row kernel)
constexpr uint32_t kFatChunkIters = 8;  // fat kernel: terms per lane staged per round (8 x 32 = 256 terms)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

struct WarpStage {
    uint32_t rp[100];            // row_ptr slice of the warp's 32-row block (97 used)
    uint32_t cols[kStageCap];    // column words of the staged terms (class ZERO when the index was out of range)
    uint4 w[2 * kStageCap];      // gathered witness elements
};

// Fold terms [k0,k1) of the staged range (base kbase) into acc.
template <int F, int RIPPLE>
__device__ __forceinline__ void fold_staged(uint32_t* acc, uint32_t k0, uint32_t k1, uint32_t kbase, const WarpStage& st, const CsrView& m,
                                            uint32_t& gen, uint32_t& mag) {
#pragma unroll 1
    for (uint32_t k = k0; k < k1; ++k) {
        TermW t;
        const uint32_t col = st.cols[k - kbase];
        t.cls = (col >> kColClsShift) & 7u;
        if (t.cls == kClsZero) continue;
        const uint4 lo = st.w[2 * (k - kbase)], hi = st.w[2 * (k - kbase) + 1];
        t.w[0] = lo.x; t.w[1] = lo.y; t.w[2] = lo.z; t.w[3] = lo.w;
        t.w[4] = hi.x; t.w[5] = hi.y; t.w[6] = hi.z; t.w[7] = hi.w;
        apply_term<F, RIPPLE>(acc, t, k, m, gen, mag);
    }
}

// ---- K1, thin rows, staged: a warp owns 32 consecutive rows; one thread per row once the operands are in shared memory
template <int F, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_rows_staged(CsrView m, CheckOut o) {
    __shared__ WarpStage stage[4];
    WarpStage& st = stage[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t my_bad = 0xffffffffu;
    unsigned int my_err = 0;
    const uint32_t n_blocks = (m.n_rows + 31u) / 32u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < n_blocks; blk += n_warps) {
        const uint32_t row0 = blk * 32u;
        const uint32_t nr = min(32u, m.n_rows - row0);
        for (uint32_t i = lane; i <= 3u * nr; i += 32u) st.rp[i] = __ldg(m.row_ptr + 3 * (size_t)row0 + i);
        __syncwarp();
        uint32_t done = 0;
        while (done < nr) {
            const uint32_t kbase = st.rp[3u * done];
            // how many of the next rows fit the staging buffer (their term ranges are contiguous and increasing)
            const bool fits = (done + lane < nr) && (st.rp[3u * (done + lane + 1u)] - kbase <= kStageCap);
            const uint32_t r = __popc(__ballot_sync(0xffffffffu, fits));
            if (r == 0) {  // a single row larger than the buffer: it is a fat row (kStageCap >= fat_terms), not ours
                done += 1;
                continue;
            }
            const uint32_t nt = st.rp[3u * (done + r)] - kbase;
            // stage: coalesced column words, then the witness gathers, all asynchronous
#pragma unroll 2
            for (uint32_t t = lane; t < nt; t += 32u) {
                uint32_t col = __ldg(m.cols + kbase + t);
                const uint32_t cls = (col >> kColClsShift) & 7u;
                if (cls != kClsZero) {
                    const uint32_t idx = col & kColIdxMask;
                    const bool is_aux = (col & kColAux) != 0;
                    if (idx >= (is_aux ? m.n_aux : m.n_inputs)) {
                        my_err = 1;
                        col = (col & ~(7u << kColClsShift)) | (kClsZero << kColClsShift);
                    } else {
                        const uint4* src = (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx;
                        cp_async16(&st.w[2 * t], src);
                        cp_async16(&st.w[2 * t + 1], src + 1);
                    }
                }
                st.cols[t] = col;
            }
            cp_async_wait_all();
            __syncwarp();
            if (lane < r) {
                const uint32_t row = row0 + done + lane;
                const uint32_t p0 = st.rp[3u * (done + lane)], p1 = st.rp[3u * (done + lane) + 1u], p2 = st.rp[3u * (done + lane) + 2u],
                               p3 = st.rp[3u * (done + lane) + 3u];
                if (p3 - p0 <= m.fat_terms) {
                    uint32_t acc[17], az[8], bz[8];
                    uint32_t gen = 0, mag = 0;
                    zeron<17>(acc);
                    fold_staged<F, 9>(acc, p0, p1, kbase, st, m, gen, mag);
                    finish_ab<F, (V & kVMagSkip) != 0>(az, acc, gen, mag);
                    gen = 0; mag = 0;
                    zeron<17>(acc);
                    fold_staged<F, 9>(acc, p1, p2, kbase, st, m, gen, mag);
                    finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, gen, mag);
                    gen = 0; mag = 0;
                    zeron<17>(acc);
                    fold_staged<F, 17>(acc, p2, p3, kbase, st, m, gen, mag);
                    if (!row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, gen, mag) && row < my_bad) my_bad = row;
                }
            }
            __syncwarp();  // the buffer is reused by the next round
            done += r;
        }
        __syncwarp();  // rp is rewritten by the next block
    }
    publish_first_bad(my_bad, m, o, my_err);
}

// ---- K1, fat rows, staged: one warp per row; each lane's next kFatChunkIters terms (witness + coefficient) are
// brought into shared memory with cp.async before any of them is folded.
struct FatStage {
    uint4 w[kFatChunkIters][2][32];  // [iteration][half][lane]: conflict-free 16-byte accesses
    uint4 c[kFatChunkIters][2][32];
};

template <int F, int RIPPLE>
__device__ __forceinline__ void warp_fold_lc_staged(uint32_t* acc, uint32_t k0, uint32_t k1, const CsrView& m, FatStage& fs, unsigned int& err,
                                                    uint32_t& any_gen, uint32_t& mag) {
    const uint32_t lane = threadIdx.x & 31u;
    zeron<17>(acc);
    uint32_t g = 0, mg = 0;
    if (k1 - k0 <= 4) {
        if (lane == 0) fold_range<F, RIPPLE, 1, false>(acc, k0, k1, m, err, g, mg);
    } else {
#pragma unroll 1
        for (uint32_t kc = k0; kc < k1; kc += 32u * kFatChunkIters) {
            uint32_t clsmask = 0;  // 3 bits per staged iteration (kClsZero = 7 means "nothing to fold")
#pragma unroll
            for (uint32_t j = 0; j < kFatChunkIters; ++j) {
                const uint32_t k = kc + 32u * j + lane;
                uint32_t cj = kClsZero;
                if (k < k1) {
                    const uint32_t col = __ldg(m.cols + k);
                    const uint32_t c = (col >> kColClsShift) & 7u;
                    if (c != kClsZero) {
                        const uint32_t idx = col & kColIdxMask;
                        const bool is_aux = (col & kColAux) != 0;
                        if (idx >= (is_aux ? m.n_aux : m.n_inputs)) {
                            err = 1;
                        } else {
                            cj = c;
                            const uint4* src = (is_aux ? m.aux : m.inputs) + 2 * (size_t)idx;
                            cp_async16(&fs.w[j][0][lane], src);
                            cp_async16(&fs.w[j][1][lane], src + 1);
                            if (c == kClsGen) {
                                cp_async16(&fs.c[j][0][lane], m.vals + 2 * (size_t)k);
                                cp_async16(&fs.c[j][1][lane], m.vals + 2 * (size_t)k + 1);
                            } else if (c == kClsPS || c == kClsMS) {
                                cp_async16(&fs.c[j][0][lane], m.vals + 2 * (size_t)k);
                            }
                        }
                    }
                }
                clsmask |= cj << (3u * j);
            }
            cp_async_wait_all();  // each lane reads back only what it copied itself: no cross-lane hazard
#pragma unroll 1
            for (uint32_t j = 0; j < kFatChunkIters; ++j) {
                const uint32_t c = (clsmask >> (3u * j)) & 7u;
                if (c == kClsZero) continue;
                uint32_t w[8];
                {
                    const uint4 lo = fs.w[j][0][lane], hi = fs.w[j][1][lane];
                    w[0] = lo.x; w[1] = lo.y; w[2] = lo.z; w[3] = lo.w;
                    w[4] = hi.x; w[5] = hi.y; w[6] = hi.z; w[7] = hi.w;
                }
                if (c == kClsGen) {
                    uint32_t cf[8];
                    const uint4 lo = fs.c[j][0][lane], hi = fs.c[j][1][lane];
                    cf[0] = lo.x; cf[1] = lo.y; cf[2] = lo.z; cf[3] = lo.w;
                    cf[4] = hi.x; cf[5] = hi.y; cf[6] = hi.z; cf[7] = hi.w;
                    mac_wide(acc, cf, w);
                    g = 1;
                    continue;
                }
                if (c == kClsM1 || c == kClsM2 || c == kClsMS) {
                    uint32_t n[8];
                    neg_mod<F>(n, w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = n[i];
                }
                if (c == kClsP1 || c == kClsM1) {
                    acc_add8<17>(acc, w);
                    mg += 1;
                } else if (c == kClsP2 || c == kClsM2) {
                    acc_add8<17>(acc, w);
                    acc_add8<17>(acc, w);
                    mg += 2;
                } else {
                    const uint32_t sm = fs.c[j][0][lane].x;
                    acc_mad_small<17>(acc, w, sm);
                    mg += sm > 8u ? 8u : sm;
                }
            }
        }
    }
    warp_sum17(acc);
    any_gen = __any_sync(0xffffffffu, g != 0) ? 1u : 0u;
    mag = __reduce_add_sync(0xffffffffu, mg > 8u ? 8u : mg);
}

template <int F, int V, int MB>
__global__ void __launch_bounds__(128, MB) check_fat_rows_staged(CsrView m, CheckOut o, const uint32_t* __restrict__ fat_rows,
                                                                 const uint32_t* __restrict__ n_fat) {
    extern __shared__ __align__(16) unsigned char fat_smem[];
    FatStage& fs = reinterpret_cast<FatStage*>(fat_smem)[threadIdx.x >> 5];
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
        warp_fold_lc_staged<F, 9>(acc, p0, p1, m, fs, my_err, any_gen, mag);
        finish_ab<F, (V & kVMagSkip) != 0>(az, acc, any_gen, mag);
        warp_fold_lc_staged<F, 9>(acc, p1, p2, m, fs, my_err, any_gen, mag);
        finish_ab<F, (V & kVMagSkip) != 0>(bz, acc, any_gen, mag);
        warp_fold_lc_staged<F, 17>(acc, p2, p3, m, fs, my_err, any_gen, mag);
        if (!row_satisfied<F, (V & kVBitRow) != 0>(acc, az, bz, any_gen, mag) && row < my_bad) my_bad = row;
    }
    publish_first_bad(my_bad, m, o, my_err);
}

constexpr size_t kFatStageSmem = 4 * sizeof(FatStage);  // per 128-thread CTA (64 KB: needs the opt-in dynamic limit)

}  // namespace bp
