// libbp_r1cs.so -- C ABI (include/bp_r1cs.h) over the CUDA kernels in kernels.cuh.
//
// Host-side responsibilities: device buffer growth, pinned staging for H2D of enforce/alloc batches, ingest conversion
// launches, the plan (row kinds, term words, row lists, readiness of rows for pipelined uploads), the choice and order of
// the check kernels (two streams, fork/join), result read-back.  No evaluation ever happens on the CPU.
#include "../../include/bp_r1cs.h"

#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (dlopen), see NcclApi

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "kernels.cuh"
#include "host/pack.hpp"

namespace {

using namespace bp;

constexpr size_t kStageBytes = 32u << 20;  // per pinned staging buffer
constexpr int kNumStage = 2;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;  // bytes
};

}  // namespace

struct bp_cs {
    int field = 0;
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t side_stream = nullptr;            // the fat-row kernels of a check run beside the thin-row kernels
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_chunk[16] = {};                 // packed witness upload: copy of chunk i+1 overlaps the widening of chunk i
    cudaStream_t stream = nullptr;
    std::string err;

    uint64_t n_rows = 0, nnz = 0, n_inputs = 0, n_aux = 0, row_base = 0;
    uint64_t n_gen = 0;  // terms that need a full 256x256 product (drives the kernel-configuration heuristic)
    DevBuf row_ptr, cols, vals, kexp, inputs, aux;
    DevBuf shadow;             // witness shadows (u32 per element, kernels.cuh: shadow_of): inputs at [0, n_inputs), aux at
    uint64_t shadow_aux_off = 1u << 16;  // [shadow_aux_off, +n_aux): one array, so that a column word maps to one 32-bit index
    bool cols_in_range = true; // plan: every column index of every row exists (checked when the plan is built)
    bool fat_int_ok = false;   // plan: the fat rows have term words for the integer pass
    bool wide_valid = true;    // false after a packed upload: inputs/aux are current only where the shadow says "big"
    // plan of the pipelined re-check: after aux chunk i has arrived, rows [0, rows[i]) / fat list [0, fat[i]) are ready
    struct {
        bool valid = false, sparse = false;
        uint64_t n_aux = 0;
        int max_pieces = 16;
        int n_pieces = 0;                  // the aux witness arrives as up to 16 pieces [off, off+len), ascending
        uint64_t off[16], len[16];
        uint32_t rows[16], fat[16], gen[16];
    } chunk_plan;
    bool sparse_upload = false;  // recheck uploads only the aux chunks this handle's rows read (row-sharded use)
    DevBuf scan_tmp, scratch;  // CUB temp; ad-hoc LC scratch
    DevBuf u8_stage;           // packed witness uploads land here before widen_u8
    DevBuf row_meta;           // plan: offset + lengths + RowKind per row (kernels.cuh: meta_pack)
    DevBuf scols;              // plan: per term, the word check_small works from
    DevBuf fat_rows;           // plan: rows handled by check_fat_rows (+ one u32 counter at the end)
    DevBuf gen_rows;           // plan: rows of kind Generic when the instance also has plain rows (+ counter)
    DevBuf lct_lc, lct_slice_off, lct_cols, lct_vals, lct_tmp;  // plan: LC tiles of the streaming full-width kernel (check_lct)
    bool lct_ok = false;       // plan: the LC-tile layout exists and covers every non-fat row
    bool lct_enabled = true;   // bp_cs_set_option("stream_kernel", 0) keeps the thread-per-row kernel
    int lct_prefetch = -1;     // "stream_prefetch": L2 prefetch depth of check_lct (kernels.cuh); -1 = by witness size
    uint64_t lct_groups = 0;   // 32-term groups in the layout (padding included)
    DevBuf deferred;           // per check: plain rows check_small handed to check_rows
    DevBuf fat_undecided;      // per check: fat rows check_fat_int handed to check_fat_rows
    uint64_t n_fat_rows = 0, n_gen_rows = 0, n_plain_rows = 0;  // plan statistics (host copies)
    uint64_t n_terms_kind[3] = {0, 0, 0};                       // terms per RowKind
    uint64_t fat_terms = 96;   // rows with more terms than this go to the warp-per-row kernels (see eff_fat_terms)
    bool fat_terms_set = false;  // the caller chose the threshold (bp_cs_set_option): it is then used as it is
    int64_t fat_ctas_per_sm = 8;  // grid of check_fat_rows = sm_count * this
    int64_t fat_int_ctas_per_sm = 6;  // grid of check_fat_int = sm_count * this (it runs BESIDE check_small: see launch_check)
    int64_t small_ctas_per_sm = 5;    // grid of check_small = sm_count * this
    int64_t kernels_mask = 3;  // measurement aid: bit 0 = launch the thin-row kernels, bit 1 = launch check_fat_rows
    int64_t variant = -1;      // < 0: default; >= 0: bit 0 = no small-row kernel, bit 1 = no shadows in the fat kernel, bit 2 = park az/bz,
                               // bit 3 = no integer pass over the fat rows
    bool plan_valid = false;
    long long* d_result = nullptr;  // [0] first_bad
    unsigned int* d_err = nullptr;  // [0] err word, [1] GEN-term counter of the last ingest call
    uint32_t* d_ndef = nullptr;     // [0] number of rows in `deferred`, [1] in `fat_undecided`
    void* h_pinned_small = nullptr;  // 64 B: [0,32) element, [32,40) first_bad, [40,44) err word, [48,56) tiny row_ptr
    void* h_stage[kNumStage] = {nullptr, nullptr};
    cudaEvent_t stage_ev[kNumStage] = {nullptr, nullptr};
    int stage_next = 0;
    FieldConsts fc;
    int64_t launches = 0;
    // One check = init_result + up to four kernels on two streams (+ the group's exchange kernel): captured once into a CUDA
    // graph and replayed while nothing the kernels' arguments depend on has changed (see check_graphed).
    struct {
        cudaGraphExec_t exec = nullptr;
        bp::CsrView view;
        long long* out = nullptr;
        const void* group = nullptr;
        int64_t variant = 0, kernels_mask = 0;
        uint64_t plan_gen = 0;
        int64_t launches = 0;  // kernels one replay launches
    } graph;
    bool use_graph = true;      // bp_cs_set_option("graph", 0) turns it off
    uint64_t plan_gen = 0;      // bumped whenever the plan is rebuilt
    int64_t graph_replays = 0, graph_captures = 0;
    DevBuf wprog, wprog_in;       // witness program (kernels.cuh: wprog_run) and its per-run inputs (message bytes, chaining states)
    uint32_t wprog_hdr[16] = {};  // host copy of the program header (0 = none installed)
    void* h_pack = nullptr;       // pinned: bit-packed witness produced by bp_cs_recheck_scalars (inputs, then aux 64-byte aligned)
    size_t h_pack_cap = 0;
    void* h_patch = nullptr;      // pinned: exception list of the same call / batch of bp_cs_set_many (indices, then values)
    size_t h_patch_cap = 0;
    DevBuf patch_stage;           // device copy of h_patch
    bool caller_dma_pending = false;  // an H2D copy enqueued by upload() still reads the caller's own pinned memory
};

namespace {

int fail(bp_cs* h, int code, const char* fmt, ...) {
    if (h) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        h->err = buf;
    }
    return code;
}

#define CU(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(h, e__ == cudaErrorMemoryAllocation ? BP_E_OOM : BP_E_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                     \
    } while (0)

#define DISPATCH_FIELD(h, EXPR)                      \
    switch ((h)->field) {                            \
        case 0: { constexpr int F = 0; EXPR; } break; \
        case 1: { constexpr int F = 1; EXPR; } break; \
        case 2: { constexpr int F = 2; EXPR; } break; \
    }

// grow `b` to at least `need` bytes, preserving the first `keep` bytes
int ensure(bp_cs* h, DevBuf& b, size_t need, size_t keep) {
    if (need <= b.cap) return BP_OK;
    size_t cap = std::max(need, b.cap + b.cap / 2);
    cap = (cap + 255) & ~size_t(255);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, cap);
    if (e != cudaSuccess && cap > need) {  // retry with the exact size before giving up
        (void)cudaGetLastError();
        cap = (need + 255) & ~size_t(255);
        e = cudaMalloc(&np, cap);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, BP_E_OOM, "cudaMalloc(%zu bytes): %s", cap, cudaGetErrorString(e));
    }
    if (b.p && keep) CU(h, cudaMemcpyAsync(np, b.p, keep, cudaMemcpyDeviceToDevice, h->stream));
    if (b.p) {
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaFree(b.p));
    }
    b.p = np;
    b.cap = cap;
    return BP_OK;
}

// The row-length threshold in force.  Product-heavy instances (most terms a full product over a full-width witness) are
// gather-bound, and a thread per row keeps more gathers in flight than a warp per row until rows get really long
// (measured on 2^22 rows x ~96 terms over a 4 GiB witness: 22.5 ms at 96, 16.4 at 128, 14.1 at 160, 14.3 at 200+).
uint64_t eff_fat_terms(const bp_cs* h) {
    if (h->fat_terms_set || 2 * h->n_gen <= h->nnz) return h->fat_terms;
    return std::max<uint64_t>(h->fat_terms, 256);
}

uint32_t* shadow_ptr(bp_cs* h, int is_aux) { return (uint32_t*)h->shadow.p + (is_aux ? h->shadow_aux_off : 0); }

// Room for `need_in` input and `need_aux` aux shadows, keeping the current contents.  The aux region starts at a fixed
// offset; outgrowing the input region (rare: circuits have few public inputs) moves it.
int ensure_shadow(bp_cs* h, uint64_t need_in, uint64_t need_aux) {
    uint64_t off = h->shadow_aux_off;
    while (need_in + 1 > off) off *= 4;  // the last slot of the input region stays 0: the null term word points at it
    if (off + need_aux >= 0xffffffffull) return fail(h, BP_E_RANGE, "too many variables for one handle");
    const size_t have = h->shadow.cap / 4;
    if (off == h->shadow_aux_off && off + need_aux <= have) return BP_OK;
    size_t aux_cap = std::max<size_t>(need_aux, have > h->shadow_aux_off ? (have - h->shadow_aux_off) * 3 / 2 : 0);
    size_t bytes = ((off + aux_cap) * 4 + 255) & ~size_t(255);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        bytes = ((off + need_aux) * 4 + 255) & ~size_t(255);
        e = cudaMalloc(&np, bytes);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, BP_E_OOM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    if (h->shadow.p) {
        if (h->n_inputs) CU(h, cudaMemcpyAsync(np, h->shadow.p, h->n_inputs * 4, cudaMemcpyDeviceToDevice, h->stream));
        if (h->n_aux)
            CU(h, cudaMemcpyAsync((uint32_t*)np + off, (uint32_t*)h->shadow.p + h->shadow_aux_off, h->n_aux * 4, cudaMemcpyDeviceToDevice,
                                  h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaFree(h->shadow.p));
    }
    CU(h, cudaMemsetAsync((uint32_t*)np + off - 1, 0, 4, h->stream));
    h->shadow.p = np;
    h->shadow.cap = bytes;
    if (off != h->shadow_aux_off) h->plan_valid = false;  // the plan's term words hold shadow indices
    h->shadow_aux_off = off;
    return BP_OK;
}

// 2^k mod p on the host, with the same limb code the device uses.
template <int F> void pow2_mod_p(int k, uint32_t* out) {
    uint32_t x[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < k; ++i) {
        uint32_t y[8];
        (void)addn<8>(y, x, x);  // < 2p < 2^256
        reduce_once<F>(y);
        std::memcpy(x, y, 32);
    }
    std::memcpy(out, x, 32);
}

template <int F> void make_consts(FieldConsts& fc) {
    pow2_mod_p<F>(544, fc.k288m);
    pow2_mod_p<F>(576, fc.k576);
}

int grid_for(const bp_cs* h, uint64_t n, int block, int per_sm) {
    uint64_t g = (n + block - 1) / block;
    uint64_t cap = (uint64_t)h->sm_count * per_sm;
    return (int)std::max<uint64_t>(1, std::min(g, cap));
}

CsrView view(const bp_cs* h) {
    CsrView m;
    std::memset(&m, 0, sizeof m);  // (compared bytewise by the graph cache: no indeterminate padding)
    m.row_ptr = (const uint32_t*)h->row_ptr.p;
    m.cols = (const uint32_t*)h->cols.p;
    m.vals = (const uint4*)h->vals.p;
    m.kexp = (const uint16_t*)h->kexp.p;
    m.inputs = (const uint4*)h->inputs.p;
    m.aux = (const uint4*)h->aux.p;
    m.shadow = (const uint32_t*)h->shadow.p;
    m.aux_off = (uint32_t)h->shadow_aux_off;
    m.wide_valid = h->wide_valid ? 1u : 0u;
    m.row_meta = (const uint32_t*)h->row_meta.p;
    m.scols = (const uint32_t*)h->scols.p;
    m.n_rows = (uint32_t)h->n_rows;
    m.n_inputs = (uint32_t)h->n_inputs;
    m.n_aux = (uint32_t)h->n_aux;
    m.fat_terms = (uint32_t)eff_fat_terms(h);
    m.row_base = h->row_base;
    return m;
}

// Copy `bytes` from a caller buffer to device memory on the handle's stream.  Pinned/registered or
// device-accessible sources go straight through; pageable memory is bounced through the pinned ring so
// the DMA overlaps the next host memcpy.
int upload(bp_cs* h, void* dst, const void* src, size_t bytes) {
    if (!bytes) return BP_OK;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, src);
    if (e != cudaSuccess) (void)cudaGetLastError();
    if (e == cudaSuccess && at.type == cudaMemoryTypeDevice) {
        CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream));
        return BP_OK;
    }
    if (e == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged)) {
        CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
        h->caller_dma_pending = true;  // the copy reads the caller's buffer: see settle()
        return BP_OK;
    }
    size_t off = 0;
    while (off < bytes) {
        const int s = h->stage_next;
        h->stage_next = (s + 1) % kNumStage;
        const size_t n = std::min(kStageBytes, bytes - off);
        CU(h, cudaEventSynchronize(h->stage_ev[s]));
        std::memcpy(h->h_stage[s], (const char*)src + off, n);
        CU(h, cudaMemcpyAsync((char*)dst + off, h->h_stage[s], n, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaEventRecord(h->stage_ev[s], h->stream));
        off += n;
    }
    return BP_OK;
}

// "The caller owns every buffer it passes; the library copies before returning": the synchronous entry points call this
// before they return, so that a DMA straight out of pinned caller memory has finished (pageable memory was already copied
// into the staging ring).  The *_async / recheck entry points document that their buffers must outlive the stream's use.
int settle(bp_cs* h) {
    if (!h->caller_dma_pending) return BP_OK;
    h->caller_dma_pending = false;
    CU(h, cudaStreamSynchronize(h->stream));
    return BP_OK;
}

int read_flags(bp_cs* h, long long* first_bad, unsigned int* err) {
    char* hp = (char*)h->h_pinned_small;
    CU(h, cudaMemcpyAsync(hp + 32, h->d_result, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(hp + 40, h->d_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->caller_dma_pending = false;
    std::memcpy(first_bad, hp + 32, 8);
    std::memcpy(err, hp + 40, 4);
    return BP_OK;
}

int clear_err(bp_cs* h) {
    CU(h, cudaMemsetAsync(h->d_err, 0, 8, h->stream));  // err word + the per-call GEN-term counter
    return BP_OK;
}

int check_err_word(bp_cs* h, const char* what) {
    char* hp = (char*)h->h_pinned_small;
    CU(h, cudaMemcpyAsync(hp + 40, h->d_err, 8, cudaMemcpyDeviceToHost, h->stream));  // [40,44) err, [44,48) GEN counter
    CU(h, cudaStreamSynchronize(h->stream));
    h->caller_dma_pending = false;
    unsigned int e;
    std::memcpy(&e, hp + 40, 4);
    if (e & 2u) return fail(h, BP_E_RANGE, "%s: a field element is not canonical (>= p)", what);
    if (e & 4u) return fail(h, BP_E_RANGE, "%s: a variable index does not fit 28 bits", what);
    if (e & 1u) return fail(h, BP_E_RANGE, "%s: a column index is out of range", what);
    if (e & 8u) return fail(h, BP_E_STATE, "%s: row offsets are not a non-decreasing sequence from 0 to nnz", what);
    return BP_OK;
}

constexpr int kVDefault = kVMagSkip | kVBitRow;

struct RowKindIs {
    const uint32_t* meta;
    uint32_t kind;
    __host__ __device__ bool operator()(uint32_t row) const { return meta_kind(meta[row]) == kind; }
};

// Ascending list of the rows whose kind is `kind` (stable selection: warps that run at the same time then work on
// neighbouring rows, whose operands share cache lines).  list has room for `expect` rows + the counter word.
int select_rows(bp_cs* h, DevBuf& list, uint32_t kind, uint64_t expect) {
    int rc = ensure(h, list, ((size_t)expect + 1) * 4, 0);
    if (rc != BP_OK) return rc;
    uint32_t* cnt = (uint32_t*)list.p + expect;
    CU(h, cudaMemsetAsync(cnt, 0, 4, h->stream));
    if (!expect) return BP_OK;
    RowKindIs pred{(const uint32_t*)h->row_meta.p, kind};
    cub::CountingInputIterator<uint32_t> rows_begin(0);
    size_t tmp_bytes = 0;
    CU(h, cub::DeviceSelect::If(nullptr, tmp_bytes, rows_begin, (uint32_t*)list.p, cnt, (int)h->n_rows, pred, h->stream));
    if ((rc = ensure(h, h->scan_tmp, tmp_bytes, 0)) != BP_OK) return rc;
    CU(h, cub::DeviceSelect::If(h->scan_tmp.p, tmp_bytes, rows_begin, (uint32_t*)list.p, cnt, (int)h->n_rows, pred, h->stream));
    return BP_OK;
}

// LC tiles for the streaming full-width kernel (kernels.cuh: check_lct): only for product-heavy instances without plain rows
// (synthetic-style: every non-fat row would go to the thread-per-row kernel), every column in range.  A second copy of the
// coefficients in tile order (+ ~4 % padding); when it does not fit in memory the thread-per-row kernel simply stays in use.
int build_lct(bp_cs* h) {
    h->lct_ok = false;
    const uint64_t ft = eff_fat_terms(h);
    if (!h->lct_enabled || !(2 * h->n_gen > h->nnz) || h->n_plain_rows != 0 || !h->cols_in_range || ft > 65535 || !h->n_rows) return BP_OK;
    const uint32_t n = (uint32_t)h->n_rows;
    const uint32_t n_tiles = (n + kLctRows - 1) / kLctRows;
    int rc;
    if ((rc = ensure(h, h->lct_lc, (size_t)n_tiles * kLctLcs * 4, 0)) != BP_OK) return rc;
    if ((rc = ensure(h, h->lct_slice_off, (size_t)n_tiles * kLctSlices * 4, 0)) != BP_OK) return rc;
    if ((rc = ensure(h, h->lct_tmp, (size_t)n_tiles * 8 + 16, 0)) != BP_OK) return rc;
    uint32_t* groups = (uint32_t*)h->lct_tmp.p;       // per tile
    uint32_t* base = groups + n_tiles + 1;            // exclusive scan of it (+ total)
    lct_plan<<<std::min<uint32_t>(n_tiles, (uint32_t)h->sm_count * 8), 256, 0, h->stream>>>((const uint32_t*)h->row_ptr.p, n, (uint32_t)ft, n_tiles,
                                                                                         (uint32_t*)h->lct_lc.p, (uint32_t*)h->lct_slice_off.p, groups);
    h->launches++;
    CU(h, cudaMemsetAsync(groups + n_tiles, 0, 4, h->stream));
    size_t tmp_bytes = 0;
    CU(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, groups, base, (int)(n_tiles + 1), h->stream));
    if ((rc = ensure(h, h->scan_tmp, tmp_bytes, 0)) != BP_OK) return rc;
    CU(h, cub::DeviceScan::ExclusiveSum(h->scan_tmp.p, tmp_bytes, groups, base, (int)(n_tiles + 1), h->stream));
    uint32_t total = 0;
    CU(h, cudaMemcpyAsync(h->h_pinned_small, base + n_tiles, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    std::memcpy(&total, h->h_pinned_small, 4);
    if ((uint64_t)total * 32 > 2 * h->nnz + 4096) return BP_OK;  // pathological length mix: the padding would double the stream
    // the big arrays: give up quietly when they do not fit
    if (ensure(h, h->lct_cols, std::max<size_t>((size_t)total * 32 * 4, 256), 0) != BP_OK ||
        ensure(h, h->lct_vals, std::max<size_t>((size_t)total * 32 * 32, 256), 0) != BP_OK) {
        for (DevBuf* b : {&h->lct_cols, &h->lct_vals})
            if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
        h->err.clear();
        return BP_OK;
    }
    lct_fill<<<std::min<uint32_t>(n_tiles, (uint32_t)h->sm_count * 8), 256, 0, h->stream>>>(
        (const uint32_t*)h->row_ptr.p, (const uint32_t*)h->cols.p, (const uint4*)h->vals.p, n_tiles, (const uint32_t*)h->lct_lc.p, base,
        (uint32_t*)h->lct_slice_off.p, (uint32_t*)h->lct_cols.p, (uint4*)h->lct_vals.p);
    h->launches++;
    CU(h, cudaGetLastError());
    h->lct_groups = total;
    h->lct_ok = true;
    return BP_OK;
}

// (Re)build the plan when rows or the threshold changed: row_meta, the per-kind counts, the fat / generic row lists and the
// deferred-row buffer.
int ensure_plan(bp_cs* h) {
    if (h->plan_valid) return BP_OK;
    h->n_fat_rows = h->n_gen_rows = h->n_plain_rows = 0;
    h->chunk_plan.valid = false;
    if (h->n_rows) {
        const uint32_t n = (uint32_t)h->n_rows;
        int rc = ensure(h, h->row_meta, ((size_t)n + 1) * 4, 0);
        if (rc != BP_OK) return rc;
        if ((rc = ensure(h, h->scratch, 40, 0)) != BP_OK) return rc;
        const size_t n_scols = (((size_t)h->nnz + 3) & ~size_t(3)) + 4;  // whole 16-byte groups, and one past the end
        if ((rc = ensure(h, h->scols, n_scols * 4, 0)) != BP_OK) return rc;
        uint32_t* d_cnt = (uint32_t*)h->scratch.p;
        CU(h, cudaMemsetAsync(d_cnt, 0, 40, h->stream));
        fill_u32<<<grid_for(h, n_scols, 256, 8), 256, 0, h->stream>>>((uint32_t*)h->scols.p, n_scols, (uint32_t)h->shadow_aux_off - 1u);
        build_row_meta<<<grid_for(h, n, 256, 8), 256, 0, h->stream>>>((const uint32_t*)h->row_ptr.p, (const uint32_t*)h->cols.p, n,
                                                                       (uint32_t)eff_fat_terms(h), (uint32_t)h->n_inputs, (uint32_t)h->n_aux,
                                                                       (uint32_t)h->shadow_aux_off, (uint32_t*)h->row_meta.p,
                                                                       (uint32_t*)h->scols.p, d_cnt, (unsigned long long*)(d_cnt + 4));
        h->launches += 2;
        CU(h, cudaGetLastError());
        uint32_t* hc = (uint32_t*)((char*)h->h_pinned_small + 64);
        CU(h, cudaMemcpyAsync(hc, d_cnt, 40, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->n_gen_rows = hc[kRowGeneric];
        h->n_plain_rows = hc[kRowPlain];
        h->n_fat_rows = hc[kRowFat];
        h->cols_in_range = hc[3] == 0;  // (rows with a column that does not exist are generic: check_rows reports them)
        std::memcpy(h->n_terms_kind, hc + 4, 24);
        if ((rc = select_rows(h, h->fat_rows, kRowFat, h->n_fat_rows)) != BP_OK) return rc;
        h->fat_int_ok = false;
        if (h->n_fat_rows) {
            CU(h, cudaMemsetAsync(d_cnt, 0, 4, h->stream));
            build_fat_words<<<grid_for(h, h->n_fat_rows * 32, 128, 16), 128, 0, h->stream>>>(
                (const uint32_t*)h->fat_rows.p, (uint32_t)h->n_fat_rows, (const uint32_t*)h->row_ptr.p, (const uint32_t*)h->cols.p,
                (uint32_t)h->n_inputs, (uint32_t)h->n_aux, (uint32_t)h->shadow_aux_off, (uint32_t*)h->scols.p, (uint16_t*)h->kexp.p, d_cnt);
            h->launches++;
            CU(h, cudaGetLastError());
            CU(h, cudaMemcpyAsync(hc, d_cnt, 4, cudaMemcpyDeviceToHost, h->stream));
            CU(h, cudaStreamSynchronize(h->stream));
            h->fat_int_ok = hc[0] == 0;
        }
        // the generic list is only needed when the thin rows are split between check_small and check_rows
        if ((rc = select_rows(h, h->gen_rows, kRowGeneric, h->n_plain_rows ? h->n_gen_rows : 0)) != BP_OK) return rc;
        if ((rc = ensure(h, h->deferred, std::max<size_t>((size_t)h->n_plain_rows * 4, 4), 0)) != BP_OK) return rc;
        if ((rc = ensure(h, h->fat_undecided, std::max<size_t>((size_t)h->n_fat_rows * 4, 4), 0)) != BP_OK) return rc;
        if ((rc = build_lct(h)) != BP_OK) return rc;
    }
    h->plan_valid = true;
    h->plan_gen++;
    return BP_OK;
}

struct MaxU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// The pieces a new aux witness is uploaded in (all of it, or with "sparse_upload" only the 2^16-element chunks some row of
// this handle reads), and which rows / fat-list entries are ready after each piece has arrived.
int ensure_chunk_plan(bp_cs* h, int max_pieces = 16) {
    auto& cp = h->chunk_plan;
    max_pieces = std::max(1, std::min(16, max_pieces));
    if (cp.valid && cp.n_aux == h->n_aux && cp.sparse == h->sparse_upload && cp.max_pieces == max_pieces) return BP_OK;
    const uint64_t n_aux = h->n_aux;
    const uint32_t n = (uint32_t)h->n_rows;
    const size_t off_small = (((size_t)n * 4) + 255) & ~size_t(255);  // [row_max | bounds 16 | out 48]
    const size_t n_need = (size_t)((n_aux + (1u << kNeedChunkLog2) - 1) >> kNeedChunkLog2);
    int rc = ensure(h, h->scratch, std::max(off_small + 64 * 4, n_need + 256), 0);
    if (rc != BP_OK) return rc;
    // ---- pieces ----
    std::vector<std::pair<uint64_t, uint64_t>> ranges;  // [begin, end) in elements
    if (h->sparse_upload && n_need) {
        uint8_t* d_need = (uint8_t*)h->scratch.p;
        CU(h, cudaMemsetAsync(d_need, 0, n_need, h->stream));
        mark_needed_aux<<<grid_for(h, h->nnz, 256, 8), 256, 0, h->stream>>>((const uint32_t*)h->cols.p, (size_t)h->nnz, d_need);
        h->launches++;
        CU(h, cudaGetLastError());
        std::vector<uint8_t> need(n_need);
        CU(h, cudaMemcpyAsync(need.data(), d_need, n_need, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        for (size_t c = 0; c < n_need; ++c) {
            if (!need[c]) continue;
            const uint64_t b = (uint64_t)c << kNeedChunkLog2, e = std::min<uint64_t>(n_aux, b + (1ull << kNeedChunkLog2));
            if (!ranges.empty() && ranges.back().second == b) ranges.back().second = e;
            else ranges.push_back({b, e});
        }
        while ((int)ranges.size() > max_pieces) {  // too many islands: bridge the smallest gap (uploads a little that nobody reads)
            size_t best = 1;
            for (size_t i = 2; i < ranges.size(); ++i)
                if (ranges[i].first - ranges[i - 1].second < ranges[best].first - ranges[best - 1].second) best = i;
            ranges[best - 1].second = ranges[best].second;
            ranges.erase(ranges.begin() + best);
        }
    } else if (n_aux) {
        ranges.push_back({0, n_aux});
    }
    while (!ranges.empty() && (int)ranges.size() < max_pieces) {  // split the largest range: finer pipelining (pieces of >= 1 Mi elements)
        size_t big = 0;
        for (size_t i = 1; i < ranges.size(); ++i)
            if (ranges[i].second - ranges[i].first > ranges[big].second - ranges[big].first) big = i;
        const uint64_t b = ranges[big].first, e = ranges[big].second;
        if (e - b < (2u << 20)) break;
        const uint64_t mid = (b + (e - b) / 2 + 255) & ~uint64_t(255);
        ranges[big].second = mid;
        ranges.insert(ranges.begin() + big + 1, {mid, e});
    }
    cp.n_pieces = (int)ranges.size();
    for (int i = 0; i < cp.n_pieces; ++i) {
        cp.off[i] = ranges[i].first;
        cp.len[i] = ranges[i].second - ranges[i].first;
    }
    // ---- readiness ----
    if (cp.n_pieces && n) {
        uint32_t* row_max = (uint32_t*)h->scratch.p;
        uint32_t* d_bounds = (uint32_t*)((char*)h->scratch.p + off_small);
        uint32_t* d_out = d_bounds + 16;
        row_max_aux<<<grid_for(h, (uint64_t)n * 32, 256, 8), 256, 0, h->stream>>>((const uint32_t*)h->row_ptr.p, (const uint32_t*)h->cols.p, n,
                                                                                  row_max);
        size_t tmp_bytes = 0;
        CU(h, cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, row_max, row_max, MaxU32(), (int)n, h->stream));
        if ((rc = ensure(h, h->scan_tmp, tmp_bytes, 0)) != BP_OK) return rc;
        CU(h, cub::DeviceScan::InclusiveScan(h->scan_tmp.p, tmp_bytes, row_max, row_max, MaxU32(), (int)n, h->stream));
        uint32_t* hb = (uint32_t*)((char*)h->h_pinned_small + 64);  // 16 words of bounds out, 48 words back
        // rows whose largest aux index is below the end of piece i only read elements of pieces 0..i (or of chunks nobody
        // uploads because nobody reads them)
        for (int i = 0; i < cp.n_pieces; ++i) hb[i] = (uint32_t)(cp.off[i] + cp.len[i]);
        CU(h, cudaMemcpyAsync(d_bounds, hb, cp.n_pieces * 4, cudaMemcpyHostToDevice, h->stream));
        ready_rows<<<1, 32, 0, h->stream>>>(row_max, n, (const uint32_t*)h->fat_rows.p, (uint32_t)h->n_fat_rows, (const uint32_t*)h->gen_rows.p,
                                            (uint32_t)(h->n_plain_rows ? h->n_gen_rows : 0), d_bounds, (uint32_t)cp.n_pieces, d_out);
        h->launches += 2;
        CU(h, cudaGetLastError());
        CU(h, cudaMemcpyAsync(hb + 16, d_out, cp.n_pieces * 12, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        for (int i = 0; i < cp.n_pieces; ++i) {
            cp.rows[i] = hb[16 + 3 * i];
            cp.fat[i] = hb[16 + 3 * i + 1];
            cp.gen[i] = hb[16 + 3 * i + 2];
        }
        // after the LAST piece everything is ready
        cp.rows[cp.n_pieces - 1] = n;
        cp.fat[cp.n_pieces - 1] = (uint32_t)h->n_fat_rows;
        cp.gen[cp.n_pieces - 1] = (uint32_t)(h->n_plain_rows ? h->n_gen_rows : 0);
    }
    cp.valid = true;
    cp.n_aux = n_aux;
    cp.sparse = h->sparse_upload;
    cp.max_pieces = max_pieces;
    return BP_OK;
}

bool lct_usable(const bp_cs* h) { return h->lct_ok && h->lct_enabled && h->wide_valid && (h->kernels_mask & 1); }
LctView lct_view(const bp_cs* h) {
    LctView v;
    v.lc = (const uint32_t*)h->lct_lc.p;
    v.slice_off = (const uint32_t*)h->lct_slice_off.p;
    v.cols = (const uint32_t*)h->lct_cols.p;
    v.vals = (const uint4*)h->lct_vals.p;
    v.n_tiles = (uint32_t)((h->n_rows + kLctRows - 1) / kLctRows);
    return v;
}
// L2 prefetch two terms ahead pays while the prefetched lines survive until their load (measured: 7.46 -> 7.04 ms on 2^24 rows
// over a 512 MiB witness) and costs when they do not (10.2 -> 13.7 ms on 2^22 x 96-term rows over a 4 GiB witness).
int lct_prefetch_depth(const bp_cs* h) {
    if (h->lct_prefetch >= 0) return h->lct_prefetch;
    return (h->n_inputs + h->n_aux) * 32 <= (1ull << 30) ? 2 : 0;
}
int lct_grid(const bp_cs* h) {
    const uint64_t n_tiles = (h->n_rows + kLctRows - 1) / kLctRows;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)h->sm_count * 2));
}

// Launch K1/K2 on the handle's stream.  Emit mode (any of az/bz/cz set): the full-width kernels over every row.
// Check mode: result init; check_small over the plain rows (integer arithmetic on the shadows); check_rows over the generic
// rows and whatever check_small deferred; check_fat_rows (warp per row) over the fat rows.
int launch_check(bp_cs* h, long long* dev_first_bad, uint4* az, uint4* bz, uint4* cz) {
    init_result<<<1, 1, 0, h->stream>>>(dev_first_bad, h->d_err, h->d_ndef);
    h->launches++;
    if (h->n_rows == 0) {
        CU(h, cudaGetLastError());
        return BP_OK;
    }
    int rc = ensure_plan(h);
    if (rc != BP_OK) return rc;
    CsrView m = view(h);
    CheckOut o{dev_first_bad, h->d_err, az, bz, cz};
    const bool emit = az || bz || cz;
    const int block = 128;
    const int grid = grid_for(h, h->n_rows, block, 16);
    const uint32_t* fat = (const uint32_t*)h->fat_rows.p;
    const uint32_t* n_fat = fat + h->n_fat_rows;
    const int fat_grid = (int)std::min<uint64_t>((h->n_fat_rows + 3) / 4, (uint64_t)h->sm_count * (uint64_t)h->fat_ctas_per_sm);
    // (VT, MBT): feature bits / min blocks per SM of the thread-per-row kernel; (VF, MBF): of the warp-per-row kernel.
#define BP_LAUNCH(EMITF, VT, MBT, VF, MBF)                                                                                                     \
    do {                                                                                                                                       \
        if (h->kernels_mask & 1) {                                                                                                             \
            DISPATCH_FIELD(h, (check_rows<F, EMITF, VT, MBT><<<grid, block, 0, h->stream>>>(m, o, h->fc)));                                     \
            h->launches++;                                                                                                                     \
        }                                                                                                                                      \
        if ((h->kernels_mask & 2) && h->n_fat_rows) {                                                                                          \
            DISPATCH_FIELD(h, (check_fat_rows<F, EMITF, VF, MBF><<<fat_grid, block, 0, h->stream>>>(m, o, h->fc, fat, n_fat)));                 \
            h->launches++;                                                                                                                     \
        }                                                                                                                                      \
    } while (0)
    const int64_t ev = h->variant < 0 ? 0 : h->variant;
    if (emit && ev < 100 && !(ev & 1) && h->n_plain_rows > 0) {
        // emit mode on a gadget-shaped instance: the integer kernels write the rows they can decide, the full-width kernels
        // the generic rows and whatever was deferred / left undecided (same lists as in check mode)
        const bool fat_int = !(ev & 2) && !(ev & 8) && h->fat_int_ok && h->n_fat_rows;
        if (h->n_fat_rows) {
            CU(h, cudaEventRecord(h->ev_fork, h->stream));
            CU(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
            cudaStream_t fs = h->side_stream;
            if (fat_int) {
                const int igrid = (int)std::min<uint64_t>((h->n_fat_rows + 3) / 4, (uint64_t)h->sm_count * (uint64_t)h->fat_int_ctas_per_sm);
                DISPATCH_FIELD(h, (check_fat_int<F, true><<<igrid, block, 0, fs>>>(m, o, fat, (uint32_t)h->n_fat_rows,
                                                                                  (uint32_t*)h->fat_undecided.p, h->d_ndef + 1)));
                DISPATCH_FIELD(h, (check_fat_rows<F, true, 0, 4><<<std::min(fat_grid, h->sm_count), block, 0, fs>>>(
                                      m, o, h->fc, (const uint32_t*)h->fat_undecided.p, h->d_ndef + 1)));
                h->launches += 2;
            } else {
                DISPATCH_FIELD(h, (check_fat_rows<F, true, 0, 4><<<fat_grid, block, 0, fs>>>(m, o, h->fc, fat, n_fat)));
                h->launches++;
            }
            CU(h, cudaGetLastError());
            CU(h, cudaEventRecord(h->ev_join, fs));
        }
        const uint64_t sblocks = (h->n_rows + kSmallRows - 1) / kSmallRows;
        const int sgrid = (int)std::min<uint64_t>((sblocks + kSmallThreads / 32 - 1) / (kSmallThreads / 32), (uint64_t)h->sm_count * (uint64_t)h->small_ctas_per_sm);
        DISPATCH_FIELD(h, (check_small<F, true><<<sgrid, kSmallThreads, kSmallSmem, h->stream>>>(m, o, (uint32_t*)h->deferred.p, h->d_ndef, 0u,
                                                                                             0xffffffffu)));
        const int lgrid = grid_for(h, h->n_gen_rows + (uint64_t)h->sm_count * block, block, 16);
        DISPATCH_FIELD(h, (check_rows<F, true, 0, 4, true><<<lgrid, block, 0, h->stream>>>(
                              m, o, h->fc, (const uint32_t*)h->gen_rows.p, (uint32_t)h->n_gen_rows, (const uint32_t*)h->deferred.p, h->d_ndef)));
        h->launches += 2;
        if (h->n_fat_rows) CU(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    } else if (emit && h->variant < 0 && lct_usable(h)) {
        const LctView lv = lct_view(h);
        DISPATCH_FIELD(h, (check_lct<F, true, 0><<<lct_grid(h), kLctThreads, 0, h->stream>>>(m, lv, o, h->fc)));
        h->launches++;
        if ((h->kernels_mask & 2) && h->n_fat_rows) {
            DISPATCH_FIELD(h, (check_fat_rows<F, true, 0, 4><<<fat_grid, block, 0, h->stream>>>(m, o, h->fc, fat, n_fat)));
            h->launches++;
        }
    } else if (emit) {
        BP_LAUNCH(true, 0, 4, 0, 4);
    } else if (h->variant >= 100) {
#ifdef BP_EXPERIMENTAL_VARIANTS
        switch (h->variant - 100) {
            case 0: BP_LAUNCH(false, 0, 4, 0, 4); break;
            case 1: BP_LAUNCH(false, kVMagSkip, 4, kVMagSkip, 4); break;
            case 2: BP_LAUNCH(false, kVMagSkip | kVBitRow, 4, kVMagSkip | kVBitRow, 4); break;
            case 3: BP_LAUNCH(false, kVMagSkip | kVBitRow | kVPark, 6, kVMagSkip | kVBitRow | kVPark, 6); break;
            case 4: BP_LAUNCH(false, kVMagSkip | kVPrefetch, 4, kVMagSkip, 4); break;
            case 5: BP_LAUNCH(false, kVMagSkip | kVPipe, 5, kVMagSkip | kVPipe, 5); break;
            case 8: BP_LAUNCH(false, kVMagSkip | kVPark, 6, kVMagSkip | kVPark, 6); break;
            case 14: BP_LAUNCH(false, 0, 6, 0, 6); break;
            default: return fail(h, BP_E_ARG, "unknown experimental variant");
        }
#else
        return fail(h, BP_E_ARG, "experimental variants need a build with BP_EXPERIMENTAL_VARIANTS=1");
#endif
    } else {
        const int64_t v = h->variant < 0 ? 0 : h->variant;
        const bool use_small = !(v & 1) && h->n_plain_rows > 0;
        if ((h->kernels_mask & 2) && h->n_fat_rows) {  // fork: the fat-row kernels start after init_result, beside the thin ones
            CU(h, cudaEventRecord(h->ev_fork, h->stream));
            CU(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        }
        // product-heavy instances (synthetic: most terms a full product, witness values full-width): az/bz parked in shared
        // memory while C is folded; no shadow / integer passes over the fat rows (every gather would be done twice)
        const bool product_heavy = 2 * h->n_gen > h->nnz;
        const bool fat_shadow = !(v & 2) && !(h->variant < 0 && product_heavy);
        const bool park = (v & 4) || (h->variant < 0 && product_heavy);
        if ((h->kernels_mask & 2) && h->n_fat_rows) {
            cudaStream_t fs = h->side_stream;  // joined below: the check is complete on h->stream when the call returns
            const int fgrid_small = std::min(fat_grid, h->sm_count);
            if (fat_shadow && !(v & 8) && h->fat_int_ok) {  // integer pass first; the modular kernel takes what it could not decide
                const int igrid = (int)std::min<uint64_t>((h->n_fat_rows + 3) / 4, (uint64_t)h->sm_count * (uint64_t)h->fat_int_ctas_per_sm);
                DISPATCH_FIELD(h, (check_fat_int<F, false><<<igrid, block, 0, fs>>>(m, o, fat, (uint32_t)h->n_fat_rows,
                                                                                   (uint32_t*)h->fat_undecided.p, h->d_ndef + 1)));
                // (normally nothing is left: a small grid, grid-stride over whatever there is)
                DISPATCH_FIELD(h, (check_fat_rows<F, false, kVDefault | kVShadow | kVPark, 5><<<fgrid_small, block, 0, fs>>>(
                                      m, o, h->fc, (const uint32_t*)h->fat_undecided.p, h->d_ndef + 1)));
                h->launches += 2;
            } else if (fat_shadow) {
                DISPATCH_FIELD(h, (check_fat_rows<F, false, kVDefault | kVShadow | kVPark, 5><<<fat_grid, block, 0, fs>>>(m, o, h->fc, fat, n_fat)));
                h->launches++;
            } else {
                DISPATCH_FIELD(h, (check_fat_rows<F, false, kVDefault, 4><<<fat_grid, block, 0, fs>>>(m, o, h->fc, fat, n_fat)));
                h->launches++;
            }
            CU(h, cudaGetLastError());
            CU(h, cudaEventRecord(h->ev_join, fs));
        }
        if (h->kernels_mask & 1) {
            if (use_small) {
                const uint64_t sblocks = (h->n_rows + kSmallRows - 1) / kSmallRows;  // one warp per block of rows
                const int sgrid = (int)std::min<uint64_t>((sblocks + kSmallThreads / 32 - 1) / (kSmallThreads / 32), (uint64_t)h->sm_count * (uint64_t)h->small_ctas_per_sm);
                check_small<0, false><<<sgrid, kSmallThreads, kSmallSmem, h->stream>>>(m, o, (uint32_t*)h->deferred.p, h->d_ndef, 0u,
                                                                                       0xffffffffu);
                h->launches++;
                const uint32_t* gl = (const uint32_t*)h->gen_rows.p;
                // the plan's generic rows, plus whatever check_small deferred (normally nothing): grid-stride over both
                const int lgrid = grid_for(h, h->n_gen_rows + (uint64_t)h->sm_count * block, block, 16);
                if (park) {
                    DISPATCH_FIELD(h, (check_rows<F, false, kVDefault | kVPark, 6, true><<<lgrid, block, 0, h->stream>>>(
                                          m, o, h->fc, gl, (uint32_t)h->n_gen_rows, (const uint32_t*)h->deferred.p, h->d_ndef)));
                } else {
                    DISPATCH_FIELD(h, (check_rows<F, false, kVDefault, 6, true><<<lgrid, block, 0, h->stream>>>(
                                          m, o, h->fc, gl, (uint32_t)h->n_gen_rows, (const uint32_t*)h->deferred.p, h->d_ndef)));
                }
                h->launches++;
            } else if (h->variant < 0 && lct_usable(h)) {
                const LctView lv = lct_view(h);
                switch (lct_prefetch_depth(h)) {
                    case 0: DISPATCH_FIELD(h, (check_lct<F, false, 0><<<lct_grid(h), kLctThreads, 0, h->stream>>>(m, lv, o, h->fc))); break;
                    case 1: DISPATCH_FIELD(h, (check_lct<F, false, 1><<<lct_grid(h), kLctThreads, 0, h->stream>>>(m, lv, o, h->fc))); break;
                    case 2: DISPATCH_FIELD(h, (check_lct<F, false, 2><<<lct_grid(h), kLctThreads, 0, h->stream>>>(m, lv, o, h->fc))); break;
                    case 3: DISPATCH_FIELD(h, (check_lct<F, false, 3><<<lct_grid(h), kLctThreads, 0, h->stream>>>(m, lv, o, h->fc))); break;
                    default: DISPATCH_FIELD(h, (check_lct<F, false, 4><<<lct_grid(h), kLctThreads, 0, h->stream>>>(m, lv, o, h->fc))); break;
                }
                h->launches++;
            } else {
                if (park) { DISPATCH_FIELD(h, (check_rows<F, false, kVDefault | kVPark, 6><<<grid, block, 0, h->stream>>>(m, o, h->fc))); }
                else { DISPATCH_FIELD(h, (check_rows<F, false, kVDefault, 6><<<grid, block, 0, h->stream>>>(m, o, h->fc))); }
                h->launches++;
            }
        }
        if ((h->kernels_mask & 2) && h->n_fat_rows) CU(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));  // join
    }
#undef BP_LAUNCH
    CU(h, cudaGetLastError());
    return BP_OK;
}


// ---- multi-GPU group: row shards, replicated witness, one MIN over the ranks' first-unsatisfied rows -----------
// NCCL is bound at run time (dlopen): a single-GPU user needs no NCCL, and inside a process that already loaded a
// libnccl.so.2 (PyTorch bundles one) the same copy is used.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

bool nccl_bind(NcclApi& api) {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return false;
#define BP_NCCL_SYM(field, sym)                                         \
    *(void**)(&api.field) = dlsym(api.lib, sym);                        \
    if (!api.field) { api.lib = nullptr; return false; }
    BP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    BP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    BP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    BP_NCCL_SYM(Broadcast, "ncclBroadcast")
    BP_NCCL_SYM(AllGather, "ncclAllGather")
    BP_NCCL_SYM(AllReduce, "ncclAllReduce")
    BP_NCCL_SYM(GroupStart, "ncclGroupStart")
    BP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    BP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BP_NCCL_SYM
    return true;
}

const NcclApi* nccl_api() {  // bound once, whichever thread asks first (handles may live on different threads)
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] { (void)nccl_bind(api); });
    return api.lib ? &api : nullptr;
}

}  // namespace

constexpr int kMaxGroup = 16;

struct bp_group {
    bp_cs* cs = nullptr;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    // Peer-memory transport of the one-word reduction: every rank owns a mailbox of 2 x world slots (double-buffered by the
    // parity of the exchange count) in its own HBM, exported to the other ranks by CUDA IPC.  One exchange = every rank
    // stores its word into slot [parity][rank] of EVERY mailbox over NVLink (system-scope stores, then the epoch as the
    // release flag) and takes the minimum over its own mailbox once all epochs have arrived: no NCCL launch, no host.
    bp::GroupSlot* box = nullptr;             // my mailbox (device memory)
    bp::GroupSlot** d_peer_boxes = nullptr;   // device array [world]: every rank's mailbox as mapped here ([rank] = box)
    void* peer_mapped[kMaxGroup] = {};        // what cudaIpcOpenMemHandle returned (to close)
    unsigned long long* d_epoch = nullptr;    // exchanges done so far (device counter: a captured graph needs no new argument)
    bool mailbox = false;
    bool ready = false;                       // bp_group_init completed (teardown is collective only then)
    long long* h_result = nullptr;            // pinned host copy of the last synchronous result
};

namespace {

#define NC(h, api, call)                                                                                      \
    do {                                                                                                      \
        ncclResult_t r__ = (call);                                                                            \
        if (r__ != ncclSuccess) return fail(h, BP_E_CUDA, "%s: %s", #call, (api)->GetErrorString(r__));         \
    } while (0)

// MIN over the group of the int64 at `word` (device memory), in place, on the handle's stream.
int enqueue_reduce(bp_group* g, long long* word) {
    bp_cs* h = g->cs;
    if (g->world == 1) return BP_OK;
    if (g->mailbox) {
        group_exchange<<<1, 32, 0, h->stream>>>(word, g->d_peer_boxes, g->box, g->d_epoch, g->rank, g->world);
        h->launches++;
        CU(h, cudaGetLastError());
        return BP_OK;
    }
    const NcclApi* api = nccl_api();
    NC(h, api, api->AllReduce(word, word, 1, ncclInt64, ncclMin, g->comm, h->stream));
    return BP_OK;
}

void drop_graph(bp_cs* h) {
    if (h->graph.exec) {
        cudaGraphExecDestroy(h->graph.exec);
        h->graph.exec = nullptr;
    }
}

// One check (and, for a group, the exchange that follows it) as a single graph launch.  The graph is captured from the very
// same launch_check / enqueue_reduce code and replayed for as long as the kernels' arguments are what they were: the view
// (every pointer, count and flag the kernels read), the result word, the plan, the variant.  Anything else re-captures.
int check_graphed(bp_cs* h, long long* dev_first_bad, bp_group* g) {
    const bool mailbox_or_single = !g || g->world == 1 || g->mailbox;  // (an NCCL all-reduce stays outside the graph)
    if (!h->use_graph || h->n_rows == 0 || h->variant >= 100) {
        int rc = launch_check(h, dev_first_bad, nullptr, nullptr, nullptr);
        if (rc == BP_OK && g) rc = enqueue_reduce(g, dev_first_bad);
        return rc;
    }
    int rc = ensure_plan(h);  // (synchronises when it has to rebuild: must not happen inside a capture)
    if (rc != BP_OK) return rc;
    const CsrView m = view(h);
    auto& gc = h->graph;
    const void* gkey = (g && mailbox_or_single) ? (const void*)g : nullptr;
    if (gc.exec && std::memcmp(&gc.view, &m, sizeof m) == 0 && gc.out == dev_first_bad && gc.group == gkey && gc.variant == h->variant &&
        gc.kernels_mask == h->kernels_mask && gc.plan_gen == h->plan_gen) {
        CU(h, cudaGraphLaunch(gc.exec, h->stream));
        h->launches += gc.launches;
        h->graph_replays++;
        if (g && !mailbox_or_single) return enqueue_reduce(g, dev_first_bad);
        return BP_OK;
    }
    drop_graph(h);
    const int64_t l0 = h->launches;
    cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        rc = launch_check(h, dev_first_bad, nullptr, nullptr, nullptr);
        if (rc == BP_OK && gkey) rc = enqueue_reduce(g, dev_first_bad);
        cudaGraph_t graph = nullptr;
        e = cudaStreamEndCapture(h->stream, &graph);
        if (rc == BP_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&gc.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
    }
    const int64_t per_replay = h->launches - l0;
    h->launches = l0;
    if (rc != BP_OK || e != cudaSuccess || !gc.exec) {  // capture is an optimisation: fall back to plain launches for good
        (void)cudaGetLastError();
        gc.exec = nullptr;
        h->use_graph = false;
        rc = launch_check(h, dev_first_bad, nullptr, nullptr, nullptr);
        if (rc == BP_OK && g) rc = enqueue_reduce(g, dev_first_bad);
        return rc;
    }
    gc.view = m;
    gc.out = dev_first_bad;
    gc.group = gkey;
    gc.variant = h->variant;
    gc.kernels_mask = h->kernels_mask;
    gc.plan_gen = h->plan_gen;
    gc.launches = per_replay;
    h->graph_captures++;
    CU(h, cudaGraphLaunch(gc.exec, h->stream));
    h->launches += gc.launches;
    if (g && !mailbox_or_single) return enqueue_reduce(g, dev_first_bad);
    return BP_OK;
}

}  // namespace

extern "C" {

int bp_abi_version(void) { return BP_ABI_VERSION; }

const char* bp_cs_last_error(const bp_cs* cs) { return cs ? cs->err.c_str() : "null handle"; }

int bp_cs_new(int field, int device, uint64_t reserve_rows, uint64_t reserve_nnz, uint64_t reserve_vars, bp_cs** out) {
    if (!out || field < 0 || field > 2) return BP_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        (void)cudaGetLastError();
        return BP_E_CUDA;  // no CPU fallback
    }
    bp_cs* h = new (std::nothrow) bp_cs();
    if (!h) return BP_E_OOM;
    h->field = field;
    h->device = device;
    auto bail = [&](int code) {
        bp_cs_free(h);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(BP_E_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(BP_E_CUDA);
    h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(BP_E_CUDA);
    h->stream = h->own_stream;
    if (cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(BP_E_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail(BP_E_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) return bail(BP_E_CUDA);
    for (auto& e : h->ev_chunk)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(BP_E_CUDA);
    if (cudaMalloc(&h->d_result, 32) != cudaSuccess) return bail(BP_E_OOM);
    h->d_err = (unsigned int*)(h->d_result + 1);
    h->d_ndef = (uint32_t*)(h->d_result + 2);
    if (cudaMemset(h->d_result, 0, 32) != cudaSuccess) return bail(BP_E_CUDA);
    if (cudaMallocHost(&h->h_pinned_small, 512) != cudaSuccess) return bail(BP_E_OOM);
    for (int s = 0; s < kNumStage; ++s) {
        if (cudaMallocHost(&h->h_stage[s], kStageBytes) != cudaSuccess) return bail(BP_E_OOM);
        if (cudaEventCreateWithFlags(&h->stage_ev[s], cudaEventDisableTiming) != cudaSuccess) return bail(BP_E_CUDA);
    }
    DISPATCH_FIELD(h, make_consts<F>(h->fc));
    {  // check_small stages its term words in more than 48 KB of dynamic shared memory (set once: not inside captured regions)
        cudaError_t ae = cudaFuncSetAttribute(check_small<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem);
        DISPATCH_FIELD(h, { if (ae == cudaSuccess) ae = cudaFuncSetAttribute(check_small<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                              (int)kSmallSmem); });
        if (ae != cudaSuccess) return bail(BP_E_CUDA);
    }
    // inputs = [ONE]  (test_cs.rs:169, witness_cs.rs:95)
    const uint64_t one[4] = {1, 0, 0, 0};
    uint64_t idx = 0;
    size_t rv = (size_t)std::max<uint64_t>(reserve_vars, 1);
    if (ensure(h, h->inputs, 32 * 1024, 0) != BP_OK) return bail(BP_E_OOM);
    if (reserve_vars && ensure(h, h->aux, rv * 32, 0) != BP_OK) return bail(BP_E_OOM);
    if (ensure_shadow(h, 1, reserve_vars) != BP_OK) return bail(BP_E_OOM);
    if (reserve_rows && ensure(h, h->row_ptr, (3 * (size_t)reserve_rows + 1) * 4, 0) != BP_OK) return bail(BP_E_OOM);
    if (reserve_nnz) {
        if (ensure(h, h->cols, (size_t)reserve_nnz * 4, 0) != BP_OK) return bail(BP_E_OOM);
        if (ensure(h, h->vals, (size_t)reserve_nnz * 32, 0) != BP_OK) return bail(BP_E_OOM);
        if (ensure(h, h->kexp, (size_t)reserve_nnz * 2, 0) != BP_OK) return bail(BP_E_OOM);
    }
    if (ensure(h, h->row_ptr, 4, 0) != BP_OK) return bail(BP_E_OOM);
    if (cudaMemsetAsync(h->row_ptr.p, 0, 4, h->stream) != cudaSuccess) return bail(BP_E_CUDA);
    if (bp_cs_alloc(h, 0, one, 1, &idx) != BP_OK) return bail(BP_E_CUDA);
    *out = h;
    return BP_OK;
}

void bp_cs_free(bp_cs* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_graph(h);
    for (DevBuf* b : {&h->row_ptr, &h->cols, &h->vals, &h->kexp, &h->inputs, &h->aux, &h->shadow, &h->scan_tmp, &h->scratch, &h->u8_stage,
                      &h->row_meta, &h->scols, &h->fat_rows, &h->gen_rows, &h->deferred, &h->fat_undecided, &h->lct_lc, &h->lct_slice_off,
                      &h->lct_cols, &h->lct_vals, &h->lct_tmp})
        if (b->p) cudaFree(b->p);
    if (h->d_result) cudaFree(h->d_result);
    if (h->h_pinned_small) cudaFreeHost(h->h_pinned_small);
    if (h->h_pack) cudaFreeHost(h->h_pack);
    if (h->h_patch) cudaFreeHost(h->h_patch);
    if (h->patch_stage.p) cudaFree(h->patch_stage.p);
    if (h->wprog.p) cudaFree(h->wprog.p);
    if (h->wprog_in.p) cudaFree(h->wprog_in.p);
    for (int s = 0; s < kNumStage; ++s) {
        if (h->h_stage[s]) cudaFreeHost(h->h_stage[s]);
        if (h->stage_ev[s]) cudaEventDestroy(h->stage_ev[s]);
    }
    if (h->side_stream) {
        cudaStreamSynchronize(h->side_stream);
        cudaStreamDestroy(h->side_stream);
    }
    for (auto& e : h->ev_chunk)
        if (e) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    (void)cudaGetLastError();
    delete h;
}

int bp_cs_set_stream(bp_cs* h, void* s) {
    if (!h) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    h->stream = s ? (cudaStream_t)s : h->own_stream;  // cudaStreamLegacy (0x1) / cudaStreamPerThread (0x2) are valid values
    return BP_OK;
}

int bp_cs_set_row_base(bp_cs* h, uint64_t b) {
    if (!h) return BP_E_ARG;
    h->row_base = b;
    return BP_OK;
}

int bp_cs_sync(bp_cs* h) {
    if (!h) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    return BP_OK;
}

// ---- checkpoint / resume of an ingested system ---------------------------------------------------------------------------
// File = header (magic, ABI version, field, counts) followed by the device arrays as they are: row_ptr, cols (with class
// bits), vals (internal form), kexp, then the canonical witness (inputs, aux).  Loading is a plain copy plus the witness
// validation pass that also rebuilds the shadows; the plan is rebuilt lazily as after any structural change.
namespace {
struct FileHeader {
    char magic[8];  // "BPR1CS\0\2"
    uint32_t abi, field;
    uint32_t header_bytes, endian;  // sizeof(FileHeader); 0x01020304 as the writer's CPU stores it
    uint64_t n_rows, nnz, n_inputs, n_aux, n_gen, row_base;
};
const char kMagic[8] = {'B', 'P', 'R', '1', 'C', 'S', 0, 2};
constexpr uint32_t kEndianTag = 0x01020304u;

// Payload checksum (the file ends with it): four interleaved 64-bit FNV-1a style lanes over the 8-byte words of everything
// written after the header, folded at the end.  Every array's byte count is a multiple of 2; odd tails are zero-padded.
struct Checksum {
    uint64_t lane[4] = {0xcbf29ce484222325ull, 0x84222325cbf29ce4ull, 0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull};
    uint64_t words = 0;
    void add(const void* p, size_t bytes) {
        const unsigned char* b = (const unsigned char*)p;
        size_t i = 0;
        for (; i + 8 <= bytes; i += 8) {
            uint64_t w;
            std::memcpy(&w, b + i, 8);
            uint64_t& l = lane[words++ & 3];
            l = (l ^ w) * 0x100000001b3ull;
        }
        if (i < bytes) {
            uint64_t w = 0;
            std::memcpy(&w, b + i, bytes - i);
            uint64_t& l = lane[words++ & 3];
            l = (l ^ w) * 0x100000001b3ull;
        }
    }
    uint64_t value() const {
        uint64_t h = words;
        for (int i = 0; i < 4; ++i) h = (h ^ lane[i]) * 0x100000001b3ull;
        return h;
    }
};

int dump(bp_cs* h, FILE* f, const void* dev, size_t bytes, Checksum& ck) {
    size_t off = 0;
    while (off < bytes) {
        const size_t n = std::min(kStageBytes, bytes - off);
        CU(h, cudaMemcpyAsync(h->h_stage[0], (const char*)dev + off, n, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        ck.add(h->h_stage[0], n);
        if (fwrite(h->h_stage[0], 1, n, f) != n) return fail(h, BP_E_STATE, "short write");
        off += n;
    }
    return BP_OK;
}
int slurp(bp_cs* h, FILE* f, void* dev, size_t bytes, Checksum& ck) {
    size_t off = 0;
    while (off < bytes) {
        const int s = h->stage_next;
        h->stage_next = (s + 1) % kNumStage;
        const size_t n = std::min(kStageBytes, bytes - off);
        CU(h, cudaEventSynchronize(h->stage_ev[s]));
        if (fread(h->h_stage[s], 1, n, f) != n) return fail(h, BP_E_STATE, "short read: truncated file");
        ck.add(h->h_stage[s], n);
        CU(h, cudaMemcpyAsync((char*)dev + off, h->h_stage[s], n, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaEventRecord(h->stage_ev[s], h->stream));
        off += n;
    }
    return BP_OK;
}
}  // namespace

int bp_cs_save(bp_cs* h, const char* path) {
    if (!h || !path) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (!h->wide_valid) {  // small values live in the shadows only: write their 32-byte form first
        for (int k = 0; k < 2; ++k) {
            const uint64_t n = k ? h->n_aux : h->n_inputs;
            if (!n) continue;
            materialize_wide<<<grid_for(h, 2 * n, 256, 8), 256, 0, h->stream>>>(shadow_ptr(h, k), n, (uint4*)(k ? h->aux.p : h->inputs.p));
            h->launches++;
        }
        CU(h, cudaGetLastError());
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(h, BP_E_STATE, "cannot open %s for writing", path);
    FileHeader hd;
    std::memset(&hd, 0, sizeof hd);
    std::memcpy(hd.magic, kMagic, 8);
    hd.abi = BP_ABI_VERSION;
    hd.field = (uint32_t)h->field;
    hd.header_bytes = (uint32_t)sizeof hd;
    hd.endian = kEndianTag;
    hd.n_rows = h->n_rows; hd.nnz = h->nnz; hd.n_inputs = h->n_inputs; hd.n_aux = h->n_aux; hd.n_gen = h->n_gen; hd.row_base = h->row_base;
    Checksum ck;
    int rc = fwrite(&hd, sizeof hd, 1, f) == 1 ? BP_OK : fail(h, BP_E_STATE, "short write");
    if (rc == BP_OK) rc = dump(h, f, h->row_ptr.p, (3 * (size_t)h->n_rows + 1) * 4, ck);
    if (rc == BP_OK) rc = dump(h, f, h->cols.p, (size_t)h->nnz * 4, ck);
    if (rc == BP_OK) rc = dump(h, f, h->vals.p, (size_t)h->nnz * 32, ck);
    if (rc == BP_OK) rc = dump(h, f, h->kexp.p, (size_t)h->nnz * 2, ck);
    if (rc == BP_OK) rc = dump(h, f, h->inputs.p, (size_t)h->n_inputs * 32, ck);
    if (rc == BP_OK) rc = dump(h, f, h->aux.p, (size_t)h->n_aux * 32, ck);
    const uint64_t sum = ck.value();
    if (rc == BP_OK && fwrite(&sum, 8, 1, f) != 1) rc = fail(h, BP_E_STATE, "short write");
    if (fclose(f) != 0 && rc == BP_OK) rc = fail(h, BP_E_STATE, "close failed");
    return rc;
}

int bp_cs_load(const char* path, int device, bp_cs** out) {
    if (!path || !out) return BP_E_ARG;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return BP_E_STATE;
    FileHeader hd;
    // not one of ours / another layout version / written on a CPU of the other byte order / counts this build cannot hold
    if (fread(&hd, sizeof hd, 1, f) != 1 || std::memcmp(hd.magic, kMagic, 8) != 0 || hd.abi != BP_ABI_VERSION || hd.field > 2 ||
        hd.header_bytes != sizeof hd || hd.endian != kEndianTag || hd.n_inputs < 1 || hd.n_inputs > kColIdxMask || hd.n_aux > kColIdxMask ||
        hd.nnz >= 0xffffffffull || 3 * hd.n_rows + 1 >= 0xffffffffull || hd.n_gen > hd.nnz) {
        fclose(f);
        return BP_E_ARG;
    }
    bp_cs* h = nullptr;
    int rc = bp_cs_new((int)hd.field, device, hd.n_rows, hd.nnz, hd.n_aux, &h);
    if (rc != BP_OK) {
        fclose(f);
        return rc;
    }
    auto done = [&](int code) -> int {
        fclose(f);
        if (code != BP_OK) {
            bp_cs_free(h);
            return code;
        }
        *out = h;
        return BP_OK;
    };
    if ((rc = ensure(h, h->row_ptr, (3 * (size_t)hd.n_rows + 1) * 4, 0)) != BP_OK) return done(rc);
    if ((rc = ensure(h, h->cols, (size_t)hd.nnz * 4, 0)) != BP_OK) return done(rc);
    if ((rc = ensure(h, h->vals, (size_t)hd.nnz * 32, 0)) != BP_OK) return done(rc);
    if ((rc = ensure(h, h->kexp, (size_t)hd.nnz * 2, 0)) != BP_OK) return done(rc);
    if ((rc = ensure(h, h->inputs, (size_t)hd.n_inputs * 32, 0)) != BP_OK) return done(rc);
    if ((rc = ensure(h, h->aux, std::max<size_t>((size_t)hd.n_aux * 32, 32), 0)) != BP_OK) return done(rc);
    h->n_inputs = h->n_aux = 0;  // nothing to carry over when the shadows are sized
    if ((rc = ensure_shadow(h, hd.n_inputs, hd.n_aux)) != BP_OK) return done(rc);
    Checksum ck;
    if ((rc = slurp(h, f, h->row_ptr.p, (3 * (size_t)hd.n_rows + 1) * 4, ck)) != BP_OK) return done(rc);
    if ((rc = slurp(h, f, h->cols.p, (size_t)hd.nnz * 4, ck)) != BP_OK) return done(rc);
    if ((rc = slurp(h, f, h->vals.p, (size_t)hd.nnz * 32, ck)) != BP_OK) return done(rc);
    if ((rc = slurp(h, f, h->kexp.p, (size_t)hd.nnz * 2, ck)) != BP_OK) return done(rc);
    if ((rc = slurp(h, f, h->inputs.p, (size_t)hd.n_inputs * 32, ck)) != BP_OK) return done(rc);
    if ((rc = slurp(h, f, h->aux.p, (size_t)hd.n_aux * 32, ck)) != BP_OK) return done(rc);
    uint64_t sum = 0;
    if (fread(&sum, 8, 1, f) != 1) return done(BP_E_STATE);
    if (sum != ck.value()) return done(BP_E_STATE);  // corrupted payload: the kernels trust these offsets and class bits
    if ((rc = clear_err(h)) != BP_OK) return done(rc);
    // the structure the kernels index with: row_ptr starts at 0, never decreases and ends at nnz
    validate_row_ptr<<<grid_for(h, 3 * hd.n_rows + 1, 256, 8), 256, 0, h->stream>>>((const uint32_t*)h->row_ptr.p, 3 * hd.n_rows + 1,
                                                                                     (uint32_t)hd.nnz, h->d_err);
    h->launches++;
    for (int k = 0; k < 2; ++k) {
        const uint64_t n = k ? hd.n_aux : hd.n_inputs;
        if (!n) continue;
        DISPATCH_FIELD(h, (validate_canonical<F><<<grid_for(h, n, 256, 8), 256, 0, h->stream>>>((const uint4*)(k ? h->aux.p : h->inputs.p), n,
                                                                                              h->d_err, shadow_ptr(h, k))));
        h->launches++;
    }
    if (cudaGetLastError() != cudaSuccess) return done(BP_E_CUDA);
    h->n_rows = hd.n_rows; h->nnz = hd.nnz; h->n_inputs = hd.n_inputs; h->n_aux = hd.n_aux; h->n_gen = hd.n_gen; h->row_base = hd.row_base;
    h->wide_valid = true;
    h->plan_valid = false;
    if ((rc = check_err_word(h, "bp_cs_load")) != BP_OK) return done(rc);  // (also waits for the uploads)
    return done(BP_OK);
}

int bp_cs_set_option(bp_cs* h, const char* key, int64_t v) {
    if (!h || !key) return BP_E_ARG;
    if (!std::strcmp(key, "variant")) {
        h->variant = v;
        return BP_OK;
    }
    if (!std::strcmp(key, "fat_ctas_per_sm")) {
        if (v < 1 || v > 32) return fail(h, BP_E_ARG, "fat_ctas_per_sm out of range");
        h->fat_ctas_per_sm = v;
        drop_graph(h);  // (a launch dimension, not part of the graph key)
        return BP_OK;
    }
    if (!std::strcmp(key, "fat_int_ctas_per_sm") || !std::strcmp(key, "small_ctas_per_sm")) {
        if (v < 1 || v > 32) return fail(h, BP_E_ARG, "%s out of range", key);
        (key[0] == 'f' ? h->fat_int_ctas_per_sm : h->small_ctas_per_sm) = v;
        drop_graph(h);
        return BP_OK;
    }
    if (!std::strcmp(key, "sparse_upload")) {
        h->sparse_upload = v != 0;
        return BP_OK;
    }
    if (!std::strcmp(key, "stream_prefetch")) {
        if (v < -1 || v > 4) return fail(h, BP_E_ARG, "stream_prefetch out of range");
        h->lct_prefetch = (int)v;
        drop_graph(h);
        return BP_OK;
    }
    if (!std::strcmp(key, "stream_kernel")) {
        h->lct_enabled = v != 0;
        h->plan_valid = false;
        return BP_OK;
    }
    if (!std::strcmp(key, "graph")) {
        h->use_graph = v != 0;
        if (!h->use_graph) drop_graph(h);
        return BP_OK;
    }
    if (!std::strcmp(key, "kernels_mask")) {
        h->kernels_mask = v & 3;
        return BP_OK;
    }
    if (!std::strcmp(key, "fat_terms")) {
        if (v < 8 || v > (1 << 30)) return fail(h, BP_E_ARG, "fat_terms out of range");
        h->fat_terms = (uint64_t)v;
        h->fat_terms_set = true;
        h->plan_valid = false;
        return BP_OK;
    }
    return fail(h, BP_E_ARG, "unknown option '%s'", key);
}

int bp_cs_get_option(bp_cs* h, const char* key, int64_t* v) {
    if (!h || !key || !v) return BP_E_ARG;
    if (!std::strcmp(key, "fat_terms")) { *v = (int64_t)eff_fat_terms(h); return BP_OK; }
    if (!std::strcmp(key, "launches")) { *v = h->launches; return BP_OK; }
    if (!std::strcmp(key, "gen_terms")) { *v = (int64_t)h->n_gen; return BP_OK; }
    if (!std::strcmp(key, "sm_count")) { *v = h->sm_count; return BP_OK; }
    if (!std::strcmp(key, "graph")) { *v = h->use_graph ? 1 : 0; return BP_OK; }
    if (!std::strcmp(key, "stream_kernel") || !std::strcmp(key, "stream_groups")) {  // 1 when the LC-tile layout is in use; its size
        CU(h, cudaSetDevice(h->device));
        int rc = ensure_plan(h);
        if (rc != BP_OK) return rc;
        *v = key[7] == 'k' ? (lct_usable(h) ? 1 : 0) : (int64_t)(h->lct_ok ? h->lct_groups : 0);
        return BP_OK;
    }
    if (!std::strcmp(key, "graph_replays")) { *v = h->graph_replays; return BP_OK; }
    if (!std::strcmp(key, "graph_captures")) { *v = h->graph_captures; return BP_OK; }
    // plan statistics (build the plan if needed): rows per kernel
    if (!std::strcmp(key, "fat_rows") || !std::strcmp(key, "plain_rows") || !std::strcmp(key, "generic_rows")) {
        CU(h, cudaSetDevice(h->device));
        int rc = ensure_plan(h);
        if (rc != BP_OK) return rc;
        *v = (int64_t)(key[0] == 'f' ? h->n_fat_rows : (key[0] == 'p' ? h->n_plain_rows : h->n_gen_rows));
        return BP_OK;
    }
    if (!std::strcmp(key, "fat_row_terms") || !std::strcmp(key, "plain_row_terms") || !std::strcmp(key, "generic_row_terms")) {
        CU(h, cudaSetDevice(h->device));
        int rc = ensure_plan(h);
        if (rc != BP_OK) return rc;
        *v = (int64_t)h->n_terms_kind[key[0] == 'f' ? kRowFat : (key[0] == 'p' ? kRowPlain : kRowGeneric)];
        return BP_OK;
    }
    if (!std::strcmp(key, "recheck_upload_bytes")) {  // bytes one bp_cs_recheck_u8 copies to the device (inputs + aux pieces)
        CU(h, cudaSetDevice(h->device));
        int rc = ensure_plan(h);
        if (rc == BP_OK) rc = ensure_chunk_plan(h);
        if (rc != BP_OK) return rc;
        uint64_t b = h->n_inputs;
        for (int i = 0; i < h->chunk_plan.n_pieces; ++i) b += h->chunk_plan.len[i];
        *v = (int64_t)b;
        return BP_OK;
    }
    // plain rows / fat rows the last check handed on to the full-width kernels
    if (!std::strcmp(key, "deferred_rows") || !std::strcmp(key, "fat_undecided_rows")) {
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaMemcpyAsync(h->h_pinned_small, h->d_ndef, 8, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        uint32_t n[2];
        std::memcpy(n, h->h_pinned_small, 8);
        *v = n[key[0] == 'f' ? 1 : 0];
        return BP_OK;
    }
    return fail(h, BP_E_ARG, "unknown option '%s'", key);
}

int bp_cs_counts(bp_cs* h, uint64_t* n_inputs, uint64_t* n_aux, uint64_t* n_rows, uint64_t* nnz) {
    if (!h) return BP_E_ARG;
    if (n_inputs) *n_inputs = h->n_inputs;
    if (n_aux) *n_aux = h->n_aux;
    if (n_rows) *n_rows = h->n_rows;
    if (nnz) *nnz = h->nnz;
    return BP_OK;
}

int bp_cs_alloc(bp_cs* h, int is_aux, const uint64_t* vals, uint64_t n, uint64_t* first_index) {
    if (!h || (!vals && n)) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    DevBuf& b = is_aux ? h->aux : h->inputs;
    uint64_t& cnt = is_aux ? h->n_aux : h->n_inputs;
    if (cnt + n >= 0x80000000ull) return fail(h, BP_E_RANGE, "more than 2^31-1 variables in one index space");
    int rc = ensure(h, b, (size_t)(cnt + n) * 32, (size_t)cnt * 32);
    if (rc != BP_OK) return rc;
    if ((rc = ensure_shadow(h, h->n_inputs + (is_aux ? 0 : n), h->n_aux + (is_aux ? n : 0))) != BP_OK) return rc;
    if (n) {
        rc = clear_err(h);
        if (rc != BP_OK) return rc;
        rc = upload(h, (char*)b.p + cnt * 32, vals, (size_t)n * 32);
        if (rc != BP_OK) return rc;
        DISPATCH_FIELD(h, (validate_canonical<F><<<grid_for(h, n, 256, 8), 256, 0, h->stream>>>(
                              (const uint4*)((char*)b.p + cnt * 32), n, h->d_err, shadow_ptr(h, is_aux) + cnt)));
        h->launches++;
        CU(h, cudaGetLastError());
        rc = check_err_word(h, "bp_cs_alloc");
        if (rc != BP_OK) return rc;  // cnt not advanced: the rejected values are not part of the system
    }
    if (first_index) *first_index = cnt;
    cnt += n;
    return BP_OK;
}

int bp_cs_set_range(bp_cs* h, int is_aux, uint64_t first, uint64_t n, const uint64_t* vals) {
    if (!h || (!vals && n)) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    const uint64_t cnt = is_aux ? h->n_aux : h->n_inputs;
    if (first > cnt || n > cnt - first) return fail(h, BP_E_RANGE, "set_range [%llu,+%llu) exceeds %llu", (unsigned long long)first,
                                                   (unsigned long long)n, (unsigned long long)cnt);
    if (!n) return BP_OK;
    DevBuf& b = is_aux ? h->aux : h->inputs;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, vals);
    if (e != cudaSuccess) (void)cudaGetLastError();
    const bool host_readable = !(e == cudaSuccess && at.type == cudaMemoryTypeDevice);
    if (host_readable && n <= 4096) {
        // small updates (set / flip-and-recheck): validate on the host first, so a rejected value leaves the witness intact
        uint32_t pl[8];
        DISPATCH_FIELD(h, { for (int i = 0; i < 8; ++i) pl[i] = PL<F>(i); });
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t* x = (const uint32_t*)(vals + 4 * i);
            int lt = 0;
            for (int j = 7; j >= 0; --j) {
                if (x[j] < pl[j]) { lt = 1; break; }
                if (x[j] > pl[j]) break;
            }
            if (!lt) return fail(h, BP_E_RANGE, "bp_cs_set_range: element %llu is not canonical (>= p)", (unsigned long long)i);
        }
        int rc = clear_err(h);
        if (rc != BP_OK) return rc;
        if ((rc = upload(h, (char*)b.p + first * 32, vals, (size_t)n * 32)) != BP_OK) return rc;
        DISPATCH_FIELD(h, (validate_canonical<F><<<grid_for(h, n, 256, 8), 256, 0, h->stream>>>(  // refreshes the shadows
                              (const uint4*)((char*)b.p + first * 32), n, h->d_err, shadow_ptr(h, is_aux) + first)));
        h->launches++;
        CU(h, cudaGetLastError());
        return settle(h);
    }
    // bulk witness refresh: copy at link speed, validate on the device; a rejected batch leaves the range ZERO (canonical)
    int rc = clear_err(h);
    if (rc != BP_OK) return rc;
    if ((rc = upload(h, (char*)b.p + first * 32, vals, (size_t)n * 32)) != BP_OK) return rc;
    DISPATCH_FIELD(h, (validate_canonical<F><<<grid_for(h, n, 256, 8), 256, 0, h->stream>>>((const uint4*)((char*)b.p + first * 32), n,
                                                                                          h->d_err, shadow_ptr(h, is_aux) + first)));
    h->launches++;
    CU(h, cudaGetLastError());
    rc = check_err_word(h, "bp_cs_set_range");
    if (rc == BP_E_RANGE) {  // never leave values >= p behind: later checks assume canonical operands
        const std::string msg = h->err;
        CU(h, cudaMemsetAsync((char*)b.p + first * 32, 0, (size_t)n * 32, h->stream));
        CU(h, cudaMemsetAsync(shadow_ptr(h, is_aux) + first, 0, (size_t)n * 4, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->err = msg + "; the range was zeroed";
    }
    return rc;
}

int bp_cs_set(bp_cs* h, int is_aux, uint64_t idx, const uint64_t v[4]) { return bp_cs_set_range(h, is_aux, idx, 1, v); }

// bytes -> device staging -> widen_u8 into elements [first, first+n) of the index space (capacity already ensured)
// Launch the kernel that turns packed values (one byte each, or one bit each) of the staging buffer into shadows.
static void launch_widen(bp_cs* h, bool bits, uint64_t stage_elem_off, uint64_t len, uint32_t* shadow_dst) {
    if (bits) {
        widen_bits<<<grid_for(h, len, 256, 8), 256, 0, h->stream>>>((const uint8_t*)h->u8_stage.p + stage_elem_off / 8, len, shadow_dst);
    } else {
        widen_u8<<<grid_for(h, len, 256, 8), 256, 0, h->stream>>>((const uint8_t*)h->u8_stage.p + stage_elem_off, len, shadow_dst);
    }
    h->launches++;
}

// sync_after: the synchronous entry points wait for copies that read the caller's pinned memory (settle)
static int widen_into(bp_cs* h, int is_aux, uint64_t first, uint64_t n, const uint8_t* vals, bool bits = false, bool sync_after = true) {
    if (bits) {  // small path only: element offsets inside the staging buffer are those of `vals`
        const size_t nbytes = (size_t)((n + 7) / 8);
        int rc = ensure(h, h->u8_stage, nbytes, 0);
        if (rc != BP_OK) return rc;
        h->wide_valid = false;
        if ((rc = upload(h, h->u8_stage.p, vals, nbytes)) != BP_OK) return rc;
        launch_widen(h, true, 0, n, shadow_ptr(h, is_aux) + first);
        CU(h, cudaGetLastError());
        return sync_after ? settle(h) : BP_OK;
    }
    int rc = ensure(h, h->u8_stage, (size_t)n, 0);
    if (rc != BP_OK) return rc;
    h->wide_valid = false;  // only the shadows are written (kernels.cuh: ld_witness)
    auto widen = [&](uint64_t off, uint64_t len) { launch_widen(h, false, off, len, shadow_ptr(h, is_aux) + first + off); };
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, vals);
    if (e != cudaSuccess) (void)cudaGetLastError();
    const bool dma_able = e == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
    if (dma_able && n >= (8u << 20)) {
        // bulk refresh from pinned memory: the copies go down the side stream in up to 16 chunks, each widened on the
        // handle's stream as soon as it has landed (H2D of chunk i+1 overlaps the widening of chunk i)
        const uint64_t chunk = (((n + 15) / 16) + 255) & ~uint64_t(255);
        CU(h, cudaEventRecord(h->ev_fork, h->stream));  // earlier users of the staging buffer
        CU(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        int i = 0;
        for (uint64_t off = 0; off < n; off += chunk, ++i) {
            const uint64_t len = std::min(chunk, n - off);
            CU(h, cudaMemcpyAsync((char*)h->u8_stage.p + off, vals + off, len, cudaMemcpyHostToDevice, h->side_stream));
            CU(h, cudaEventRecord(h->ev_chunk[i], h->side_stream));
            CU(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[i], 0));
            widen(off, len);
        }
    } else {
        if ((rc = upload(h, h->u8_stage.p, vals, (size_t)n)) != BP_OK) return rc;
        widen(0, n);
    }
    CU(h, cudaGetLastError());
    if (sync_after) {
        if (dma_able && n >= (8u << 20)) CU(h, cudaStreamSynchronize(h->stream));  // (the chunked copies went down the side stream)
        return settle(h);
    }
    return BP_OK;
}

int bp_cs_alloc_u8(bp_cs* h, int is_aux, const uint8_t* vals, uint64_t n, uint64_t* first_index) {
    if (!h || (!vals && n)) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    DevBuf& b = is_aux ? h->aux : h->inputs;
    uint64_t& cnt = is_aux ? h->n_aux : h->n_inputs;
    if (cnt + n >= 0x80000000ull) return fail(h, BP_E_RANGE, "more than 2^31-1 variables in one index space");
    int rc = ensure(h, b, (size_t)(cnt + n) * 32, (size_t)cnt * 32);
    if (rc != BP_OK) return rc;
    if ((rc = ensure_shadow(h, h->n_inputs + (is_aux ? 0 : n), h->n_aux + (is_aux ? n : 0))) != BP_OK) return rc;
    if (n && (rc = widen_into(h, is_aux, cnt, n, vals)) != BP_OK) return rc;
    if (first_index) *first_index = cnt;
    cnt += n;
    return BP_OK;
}

// Packed overwrite of [first, first+n) (one byte or one bit per value); the recheck entry points skip the final wait.
static int set_range_packed(bp_cs* h, int is_aux, uint64_t first, uint64_t n, const uint8_t* vals, bool bits, bool sync_after) {
    CU(h, cudaSetDevice(h->device));
    const uint64_t cnt = is_aux ? h->n_aux : h->n_inputs;
    if (first > cnt || n > cnt - first) return fail(h, BP_E_RANGE, "set_range_%s [%llu,+%llu) exceeds %llu", bits ? "bits" : "u8",
                                                   (unsigned long long)first, (unsigned long long)n, (unsigned long long)cnt);
    if (!n) return BP_OK;
    return widen_into(h, is_aux, first, n, vals, bits, sync_after);
}

int bp_cs_set_range_u8(bp_cs* h, int is_aux, uint64_t first, uint64_t n, const uint8_t* vals) {
    if (!h || (!vals && n)) return BP_E_ARG;
    return set_range_packed(h, is_aux, first, n, vals, false, true);
}

int bp_cs_set_range_bits(bp_cs* h, int is_aux, uint64_t first, uint64_t n, const uint8_t* bits) {
    if (!h || (!bits && n)) return BP_E_ARG;
    return set_range_packed(h, is_aux, first, n, bits, true, true);
}

int bp_cs_witness(bp_cs* h, int is_aux, uint64_t first, uint64_t n, uint64_t* out) {
    if (!h || (!out && n)) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    const uint64_t cnt = is_aux ? h->n_aux : h->n_inputs;
    if (first > cnt || n > cnt - first) return fail(h, BP_E_RANGE, "witness range [%llu,+%llu) exceeds %llu", (unsigned long long)first,
                                                   (unsigned long long)n, (unsigned long long)cnt);
    if (!n) return BP_OK;
    const DevBuf& b = is_aux ? h->aux : h->inputs;
    if (!h->wide_valid) {  // small values live in the shadows only: write their 32-byte form before it is read
        materialize_wide<<<grid_for(h, 2 * n, 256, 8), 256, 0, h->stream>>>(shadow_ptr(h, is_aux) + first, n, (uint4*)((char*)b.p + first * 32));
        h->launches++;
        CU(h, cudaGetLastError());
    }
    CU(h, cudaMemcpyAsync(out, (const char*)b.p + first * 32, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return BP_OK;
}

int bp_cs_get(bp_cs* h, int is_aux, uint64_t idx, uint64_t v[4]) { return bp_cs_witness(h, is_aux, idx, 1, v); }

int bp_cs_enforce(bp_cs* h, uint64_t n_rows, const uint32_t* lens, const uint32_t* cols, const uint64_t* coeffs) {
    if (!h || (n_rows && !lens)) return BP_E_ARG;
    if (!n_rows) return BP_OK;
    CU(h, cudaSetDevice(h->device));
    uint64_t add = 0;
    for (uint64_t i = 0; i < 3 * n_rows; ++i) add += lens[i];
    if (add && (!cols || !coeffs)) return BP_E_ARG;
    if (h->nnz + add >= 0xffffffffull || 3 * (h->n_rows + n_rows) + 1 >= 0xffffffffull)
        return fail(h, BP_E_RANGE, "nnz or LC count would reach 2^32 in one handle; shard the rows");
    const size_t lc_old = 3 * (size_t)h->n_rows, lc_add = 3 * (size_t)n_rows;
    if (lc_add + 1 > 0x7fffffffull) return fail(h, BP_E_RANGE, "more than 2^31 LCs in one enforce batch; split the call");
    int rc;
    if ((rc = ensure(h, h->row_ptr, (lc_old + lc_add + 1) * 4, (lc_old + 1) * 4)) != BP_OK) return rc;
    if ((rc = ensure(h, h->cols, (size_t)(h->nnz + add) * 4, (size_t)h->nnz * 4)) != BP_OK) return rc;
    if ((rc = ensure(h, h->vals, (size_t)(h->nnz + add) * 32, (size_t)h->nnz * 32)) != BP_OK) return rc;
    if ((rc = ensure(h, h->kexp, (size_t)(h->nnz + add) * 2, (size_t)h->nnz * 2)) != BP_OK) return rc;
    uint32_t* rp = (uint32_t*)h->row_ptr.p;
    // lens -> row_ptr[lc_old .. lc_old+lc_add] : exclusive scan of (lens ++ [0]) then + nnz
    // stage the lens right where the scan output goes, shifted by one slot so in-place scan is safe
    if ((rc = ensure(h, h->scratch, (lc_add + 1) * 4, 0)) != BP_OK) return rc;
    if ((rc = upload(h, h->scratch.p, lens, lc_add * 4)) != BP_OK) return rc;
    CU(h, cudaMemsetAsync((char*)h->scratch.p + lc_add * 4, 0, 4, h->stream));
    size_t tmp_bytes = 0;
    CU(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t*)h->scratch.p, rp + lc_old, (int)(lc_add + 1), h->stream));
    if ((rc = ensure(h, h->scan_tmp, tmp_bytes, 0)) != BP_OK) return rc;
    CU(h, cub::DeviceScan::ExclusiveSum(h->scan_tmp.p, tmp_bytes, (const uint32_t*)h->scratch.p, rp + lc_old, (int)(lc_add + 1), h->stream));
    if (h->nnz) {
        add_base<<<grid_for(h, lc_add + 1, 256, 8), 256, 0, h->stream>>>(rp + lc_old, (uint32_t)(lc_add + 1), (uint32_t)h->nnz);
        h->launches++;
    }
    if (add) {
        if ((rc = upload(h, (uint32_t*)h->cols.p + h->nnz, cols, (size_t)add * 4)) != BP_OK) return rc;
        if ((rc = upload(h, (char*)h->vals.p + h->nnz * 32, coeffs, (size_t)add * 32)) != BP_OK) return rc;
        if ((rc = clear_err(h)) != BP_OK) return rc;
        // scratch was used for the lens; reuse it for the per-LC kind bytes once the scan has consumed them
        uint8_t* kind = (uint8_t*)h->scratch.p;
        DISPATCH_FIELD(h, (classify_lcs<F><<<grid_for(h, lc_add, 128, 16), 128, 0, h->stream>>>((const uint4*)h->vals.p, rp, (uint32_t)lc_old,
                                                                                          (uint32_t)lc_add, kind)));
        DISPATCH_FIELD(h, (convert_terms<F><<<grid_for(h, add, 128, 16), 128, 0, h->stream>>>(
                              (uint4*)h->vals.p, (uint32_t*)h->cols.p, (uint16_t*)h->kexp.p, rp, kind, (uint32_t)lc_old, (uint32_t)lc_add, (uint32_t)h->nnz,
                              (uint32_t)add, h->fc, h->d_err)));
        h->launches += 2;
        CU(h, cudaGetLastError());
        if ((rc = check_err_word(h, "bp_cs_enforce")) != BP_OK) return rc;  // rows not committed
        {
            unsigned int gen_chunk;
            std::memcpy(&gen_chunk, (char*)h->h_pinned_small + 44, 4);
            h->n_gen += gen_chunk;
        }
    } else {
        CU(h, cudaGetLastError());
        if ((rc = settle(h)) != BP_OK) return rc;  // (the lens may have been read straight from pinned caller memory)
    }
    h->n_rows += n_rows;
    h->nnz += add;
    h->plan_valid = false;
    return BP_OK;
}

int bp_cs_check_async(bp_cs* h, int64_t* dev_result) {
    if (!h || !dev_result) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    return check_graphed(h, (long long*)dev_result, nullptr);
}

int bp_cs_first_unsatisfied(bp_cs* h, int64_t* row) {
    if (!h || !row) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    int rc = check_graphed(h, h->d_result, nullptr);
    if (rc != BP_OK) return rc;
    long long fb;
    unsigned int e;
    if ((rc = read_flags(h, &fb, &e)) != BP_OK) return rc;
    if (e & 1u) return fail(h, BP_E_RANGE, "a term references a variable index that does not exist");
    *row = fb == 0x7fffffffffffffffLL ? -1 : (int64_t)(fb - (long long)h->row_base);
    return BP_OK;
}

// Core of bp_cs_recheck_u8[_async]: dev_first_bad receives the first failing GLOBAL row (INT64_MAX = satisfied).
static int recheck_u8(bp_cs* h, const uint8_t* inputs_u8, const uint8_t* aux_u8, long long* dev_first_bad, bool bits = false) {
    CU(h, cudaSetDevice(h->device));
    int rc;
    if (inputs_u8 && (rc = set_range_packed(h, 0, 0, h->n_inputs, inputs_u8, bits, false)) != BP_OK) return rc;
    const uint64_t n = h->n_aux;
    if ((rc = ensure_plan(h)) != BP_OK) return rc;
    cudaPointerAttributes at;
    cudaError_t pe = cudaPointerGetAttributes(&at, aux_u8);
    if (pe != cudaSuccess) (void)cudaGetLastError();
    const bool dma_able = pe == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged);
    // interleaving pays when the copy is long next to the check: from 4 MB in the 1-byte form, from 32 MB in the 1-bit form
    // (the 1-bit form of a 10^8-variable witness is a 0.25 ms copy: cheaper to finish it and run the check with its thin and
    // fat kernels side by side)
    const uint64_t xfer = bits ? n / 8 : n;
    const bool pipelined = dma_able && xfer >= (4u << 20) * (bits ? 8u : 1u) && h->n_plain_rows > 0 && h->variant < 0 && h->kernels_mask == 3 &&
                           h->n_rows > 0;
    if (!pipelined) {
        if (n && h->sparse_upload && h->n_rows) {  // only the chunks this handle's rows read, then the ordinary check
            if ((rc = ensure_chunk_plan(h, 16)) != BP_OK) return rc;
            if ((rc = ensure(h, h->u8_stage, (size_t)n, 0)) != BP_OK) return rc;
            h->wide_valid = false;
            for (int i = 0; i < h->chunk_plan.n_pieces; ++i) {
                const uint64_t off = h->chunk_plan.off[i], len = h->chunk_plan.len[i];
                const uint64_t boff = bits ? off / 8 : off, blen = bits ? (len + 7) / 8 : len;
                if ((rc = upload(h, (char*)h->u8_stage.p + boff, aux_u8 + boff, (size_t)blen)) != BP_OK) return rc;
                launch_widen(h, bits, off, len, shadow_ptr(h, 1) + off);
            }
            CU(h, cudaGetLastError());
        } else if (n && (rc = set_range_packed(h, 1, 0, n, aux_u8, bits, false)) != BP_OK) {
            return rc;
        }
        return launch_check(h, dev_first_bad, nullptr, nullptr, nullptr);
    }
    // Pipelined: the copy of aux chunk i+1 (side stream) overlaps the widening of chunk i and the check of the rows that
    // became ready with it (handle's stream); the full-width kernels take the generic / deferred / undecided rows at the end.
    // Pieces of >= 4 MB of transfer (~75 us of PCIe): finer pieces only add launches and tails once the copy is shorter than
    // the check (the 1-bit form of a 10^8-variable witness is 14 MB: three pieces; its 1-byte form 109 MB: sixteen).
    if ((rc = ensure_chunk_plan(h, (int)std::min<uint64_t>(16, std::max<uint64_t>(1, xfer >> 22)))) != BP_OK) return rc;
    const int n_chunks = h->chunk_plan.n_pieces;
    if ((rc = ensure(h, h->u8_stage, (size_t)n, 0)) != BP_OK) return rc;
    const auto& cp = h->chunk_plan;
    CsrView m = view(h);
    h->wide_valid = false;
    m.wide_valid = 0;
    CheckOut o{dev_first_bad, h->d_err, nullptr, nullptr, nullptr};
    init_result<<<1, 1, 0, h->stream>>>(dev_first_bad, h->d_err, h->d_ndef);
    h->launches++;
    CU(h, cudaEventRecord(h->ev_fork, h->stream));
    CU(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    const uint32_t n_blocks = (uint32_t)((h->n_rows + kSmallRows - 1) / kSmallRows);
    uint32_t blk_done = 0, fat_done = 0;
    for (int i = 0; i < n_chunks; ++i) {
        const uint64_t off = cp.off[i], len = cp.len[i];  // (piece offsets are multiples of 256 elements: whole bytes in either form)
        const uint64_t boff = bits ? off / 8 : off, blen = bits ? (len + 7) / 8 : len;
        CU(h, cudaMemcpyAsync((char*)h->u8_stage.p + boff, aux_u8 + boff, blen, cudaMemcpyHostToDevice, h->side_stream));
        CU(h, cudaEventRecord(h->ev_chunk[i], h->side_stream));
        CU(h, cudaStreamWaitEvent(h->stream, h->ev_chunk[i], 0));
        launch_widen(h, bits, off, len, shadow_ptr(h, 1) + off);
        const uint32_t blk_ready = i == n_chunks - 1 ? n_blocks : cp.rows[i] / kSmallRows;  // whole 64-row blocks only
        if (blk_ready > blk_done) {
            const uint32_t nb = blk_ready - blk_done;
            const int sgrid = (int)std::min<uint64_t>((nb + kSmallThreads / 32 - 1) / (kSmallThreads / 32), (uint64_t)h->sm_count * (uint64_t)h->small_ctas_per_sm);
            check_small<0, false><<<sgrid, kSmallThreads, kSmallSmem, h->stream>>>(m, o, (uint32_t*)h->deferred.p, h->d_ndef, blk_done,
                                                                                   blk_ready);
            h->launches++;
            blk_done = blk_ready;
        }
        // fat rows below the last whole block that is ready (their operands have arrived as well)
        uint32_t fat_ready = cp.fat[i];
        if (h->fat_int_ok && fat_ready > fat_done) {
            const uint32_t nf = fat_ready - fat_done;
            const int igrid = (int)std::min<uint64_t>((nf + 3) / 4, (uint64_t)h->sm_count * (uint64_t)h->fat_int_ctas_per_sm);
            DISPATCH_FIELD(h, (check_fat_int<F, false><<<igrid, 128, 0, h->stream>>>(m, o, (const uint32_t*)h->fat_rows.p + fat_done, nf,
                                                                                    (uint32_t*)h->fat_undecided.p, h->d_ndef + 1)));
            h->launches++;
            fat_done = fat_ready;
        }
    }
    CU(h, cudaGetLastError());
    // leftovers: the plan's generic rows + deferred plain rows; undecided fat rows (or all of them without an integer plan)
    {
        const int lgrid = grid_for(h, h->n_gen_rows + (uint64_t)h->sm_count * 128, 128, 16);
        DISPATCH_FIELD(h, (check_rows<F, false, kVDefault, 6, true><<<lgrid, 128, 0, h->stream>>>(
                              m, o, h->fc, (const uint32_t*)h->gen_rows.p, (uint32_t)h->n_gen_rows, (const uint32_t*)h->deferred.p, h->d_ndef)));
        h->launches++;
        if (h->n_fat_rows) {
            const int fat_grid = (int)std::min<uint64_t>((h->n_fat_rows + 3) / 4, (uint64_t)h->sm_count * (uint64_t)h->fat_ctas_per_sm);
            if (h->fat_int_ok) {
                DISPATCH_FIELD(h, (check_fat_rows<F, false, kVDefault | kVShadow | kVPark, 5><<<std::min(fat_grid, h->sm_count), 128, 0, h->stream>>>(
                                      m, o, h->fc, (const uint32_t*)h->fat_undecided.p, h->d_ndef + 1)));
            } else {
                const uint32_t* fat = (const uint32_t*)h->fat_rows.p;
                DISPATCH_FIELD(h, (check_fat_rows<F, false, kVDefault | kVShadow | kVPark, 5><<<fat_grid, 128, 0, h->stream>>>(m, o, h->fc, fat,
                                                                                                                    fat + h->n_fat_rows)));
            }
            h->launches++;
        }
    }
    CU(h, cudaGetLastError());
    return BP_OK;
}

int bp_cs_recheck_u8(bp_cs* h, const uint8_t* inputs_u8, const uint8_t* aux_u8, int64_t* row) {
    if (!h || !row || (!aux_u8 && h->n_aux)) return BP_E_ARG;
    int rc = recheck_u8(h, inputs_u8, aux_u8, h->d_result);
    if (rc != BP_OK) return rc;
    long long fb;
    unsigned int e;
    if ((rc = read_flags(h, &fb, &e)) != BP_OK) return rc;
    if (e & 1u) return fail(h, BP_E_RANGE, "a term references a variable index that does not exist");
    *row = fb == 0x7fffffffffffffffLL ? -1 : (int64_t)(fb - (long long)h->row_base);
    return BP_OK;
}

int bp_cs_recheck_u8_async(bp_cs* h, const uint8_t* inputs_u8, const uint8_t* aux_u8, int64_t* dev_result) {
    if (!h || !dev_result || (!aux_u8 && h->n_aux)) return BP_E_ARG;
    return recheck_u8(h, inputs_u8, aux_u8, (long long*)dev_result);
}

int bp_cs_recheck_bits(bp_cs* h, const uint8_t* inputs_bits, const uint8_t* aux_bits, int64_t* row) {
    if (!h || !row || (!aux_bits && h->n_aux)) return BP_E_ARG;
    int rc = recheck_u8(h, inputs_bits, aux_bits, h->d_result, true);
    if (rc != BP_OK) return rc;
    long long fb;
    unsigned int e;
    if ((rc = read_flags(h, &fb, &e)) != BP_OK) return rc;
    if (e & 1u) return fail(h, BP_E_RANGE, "a term references a variable index that does not exist");
    *row = fb == 0x7fffffffffffffffLL ? -1 : (int64_t)(fb - (long long)h->row_base);
    return BP_OK;
}

int bp_cs_recheck_bits_async(bp_cs* h, const uint8_t* inputs_bits, const uint8_t* aux_bits, int64_t* dev_result) {
    if (!h || !dev_result || (!aux_bits && h->n_aux)) return BP_E_ARG;
    return recheck_u8(h, inputs_bits, aux_bits, (long long*)dev_result, true);
}

int bp_cs_eval_async(bp_cs* h, uint64_t* az, uint64_t* bz, uint64_t* cz) {
    if (!h) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (!az && !bz && !cz) return BP_OK;
    return launch_check(h, h->d_result, (uint4*)az, (uint4*)bz, (uint4*)cz);
}

int bp_cs_eval(bp_cs* h, uint64_t* az, uint64_t* bz, uint64_t* cz) {
    if (!h) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if ((!az && !bz && !cz) || h->n_rows == 0) return BP_OK;
    const size_t bytes = (size_t)h->n_rows * 32;
    const int want = (az ? 1 : 0) + (bz ? 1 : 0) + (cz ? 1 : 0);
    int rc = ensure(h, h->scratch, bytes * want, 0);
    if (rc != BP_OK) return rc;
    char* p = (char*)h->scratch.p;
    uint4 *da = nullptr, *db = nullptr, *dc = nullptr;
    if (az) { da = (uint4*)p; p += bytes; }
    if (bz) { db = (uint4*)p; p += bytes; }
    if (cz) { dc = (uint4*)p; p += bytes; }
    if ((rc = launch_check(h, h->d_result, da, db, dc)) != BP_OK) return rc;
    if (az) CU(h, cudaMemcpyAsync(az, da, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (bz) CU(h, cudaMemcpyAsync(bz, db, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (cz) CU(h, cudaMemcpyAsync(cz, dc, bytes, cudaMemcpyDeviceToHost, h->stream));
    long long fb;
    unsigned int e;
    if ((rc = read_flags(h, &fb, &e)) != BP_OK) return rc;
    if (e & 1u) return fail(h, BP_E_RANGE, "a term references a variable index that does not exist");
    return BP_OK;
}

int bp_cs_eval_lc(bp_cs* h, const uint32_t* cols, const uint64_t* coeffs, uint32_t n, uint64_t out[4]) {
    if (!h || !out || (n && (!cols || !coeffs))) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    // scratch: [coeffs 32n][cols 4n (16-aligned)][out 32]
    const size_t off_cols = (size_t)n * 32, off_out = off_cols + (((size_t)n * 4 + 15) & ~size_t(15));
    int rc = ensure(h, h->scratch, off_out + 32, 0);
    if (rc != BP_OK) return rc;
    char* s = (char*)h->scratch.p;
    if ((rc = clear_err(h)) != BP_OK) return rc;
    if (n) {
        if ((rc = upload(h, s, coeffs, (size_t)n * 32)) != BP_OK) return rc;
        if ((rc = upload(h, s + off_cols, cols, (size_t)n * 4)) != BP_OK) return rc;
    }
    DISPATCH_FIELD(h, (eval_lc_kernel<F><<<1, 32, 0, h->stream>>>((const uint32_t*)(s + off_cols), (const uint4*)s, n, view(h), h->fc,
                                                                 (uint4*)(s + off_out), h->d_err)));
    h->launches++;
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(h->h_pinned_small, s + off_out, 32, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = check_err_word(h, "bp_cs_eval_lc")) != BP_OK) return rc;
    std::memcpy(out, h->h_pinned_small, 32);
    return BP_OK;
}

int bp_cs_synth_rows(bp_cs* h, uint64_t seed, uint32_t t, uint64_t n_vars, uint64_t n_inputs, uint64_t row0, uint64_t n_rows) {
    if (!h) return BP_E_ARG;
    if (t < 1 || n_vars < 2ull * t || n_inputs > n_vars || n_inputs < 1) return fail(h, BP_E_ARG, "bad synthetic parameters");
    if (!n_rows) return BP_OK;
    CU(h, cudaSetDevice(h->device));
    const size_t lc_old = 3 * (size_t)h->n_rows, lc_add = 3 * (size_t)n_rows;
    if (lc_old + lc_add + 1 >= 0xffffffffull || lc_add + 1 > 0x7fffffffull) return fail(h, BP_E_RANGE, "LC count too large for one handle/call");
    int rc;
    if ((rc = ensure(h, h->row_ptr, (lc_old + lc_add + 1) * 4, (lc_old + 1) * 4)) != BP_OK) return rc;
    if ((rc = ensure(h, h->scratch, (lc_add + 1) * 4, 0)) != BP_OK) return rc;
    uint32_t* rp = (uint32_t*)h->row_ptr.p;
    synth_lens<<<grid_for(h, lc_add, 256, 8), 256, 0, h->stream>>>((uint32_t*)h->scratch.p, seed, t, 3 * row0, (uint32_t)lc_add);
    h->launches++;
    CU(h, cudaMemsetAsync((char*)h->scratch.p + lc_add * 4, 0, 4, h->stream));
    size_t tmp_bytes = 0;
    CU(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t*)h->scratch.p, rp + lc_old, (int)(lc_add + 1), h->stream));
    if ((rc = ensure(h, h->scan_tmp, tmp_bytes, 0)) != BP_OK) return rc;
    CU(h, cub::DeviceScan::ExclusiveSum(h->scan_tmp.p, tmp_bytes, (const uint32_t*)h->scratch.p, rp + lc_old, (int)(lc_add + 1), h->stream));
    // the scan is 32-bit: bound the chunk's nnz with 64-bit host arithmetic first, then read the exact total back
    if ((uint64_t)lc_add * (2ull * t - 1) >= 0xffffffffull) return fail(h, BP_E_RANGE, "synthetic chunk too large; split the call");
    uint32_t add32 = 0;
    CU(h, cudaMemcpyAsync(h->h_pinned_small, rp + lc_old + lc_add, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    std::memcpy(&add32, h->h_pinned_small, 4);
    const uint64_t add = add32;
    if (h->nnz + add >= 0xffffffffull)
        return fail(h, BP_E_RANGE, "nnz would reach 2^32 in one handle; shard the rows");
    if (h->nnz) {
        add_base<<<grid_for(h, lc_add + 1, 256, 8), 256, 0, h->stream>>>(rp + lc_old, (uint32_t)(lc_add + 1), (uint32_t)h->nnz);
        h->launches++;
    }
    if ((rc = ensure(h, h->cols, (size_t)(h->nnz + add) * 4, (size_t)h->nnz * 4)) != BP_OK) return rc;
    if ((rc = ensure(h, h->vals, (size_t)(h->nnz + add) * 32, (size_t)h->nnz * 32)) != BP_OK) return rc;
    if ((rc = ensure(h, h->kexp, (size_t)(h->nnz + add) * 2, (size_t)h->nnz * 2)) != BP_OK) return rc;  // (unspecified: GEN terms)
    DISPATCH_FIELD(h, (synth_fill<F><<<grid_for(h, lc_add, 128, 16), 128, 0, h->stream>>>(
                          (uint32_t*)h->cols.p, (uint4*)h->vals.p, rp, (uint32_t)lc_old, (uint32_t)lc_add, seed, 3 * row0, n_vars, n_inputs,
                          h->fc)));
    h->launches++;
    CU(h, cudaGetLastError());
    h->n_rows += n_rows;
    h->nnz += add;
    h->n_gen += add;  // random coefficients: every term is a full product
    h->plan_valid = false;
    return BP_OK;
}

int bp_cs_synth_witness(bp_cs* h, uint64_t seed, uint64_t n_vars, uint64_t n_inputs) {
    if (!h) return BP_E_ARG;
    if (n_inputs < 1 || n_inputs > n_vars || n_vars >= 0x80000000ull) return fail(h, BP_E_ARG, "bad synthetic parameters");
    CU(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure(h, h->inputs, (size_t)n_inputs * 32, 0)) != BP_OK) return rc;
    if ((rc = ensure(h, h->aux, (size_t)std::max<uint64_t>(n_vars - n_inputs, 1) * 32, 0)) != BP_OK) return rc;
    h->n_inputs = h->n_aux = 0;  // the witness is REPLACED: nothing to carry over when the shadows move
    if ((rc = ensure_shadow(h, n_inputs, n_vars - n_inputs)) != BP_OK) return rc;
    DISPATCH_FIELD(h, (synth_witness<F><<<grid_for(h, n_vars, 256, 8), 256, 0, h->stream>>>(
                          (uint4*)h->inputs.p, (uint4*)h->aux.p, shadow_ptr(h, 0), shadow_ptr(h, 1), seed, n_vars, n_inputs)));
    h->launches++;
    CU(h, cudaGetLastError());
    h->n_inputs = n_inputs;
    h->n_aux = n_vars - n_inputs;
    h->wide_valid = true;   // the whole witness was just written in both forms
    h->plan_valid = false;  // the plan vouches for column ranges
    return BP_OK;
}


}  // extern "C"

// ---- prover hand-off: the system in a documented, implementation-independent layout ----------------------------------------
// Not a reference interface (LinearCombination is not serialisable there, lc.rs:34; the reference has no on-disk R1CS).  What
// a downstream prover (a Nova / Spartan style R1CSShape + witness) needs after the check: A, B, C with CANONICAL coefficients
// (no class bits, no internal scaling), the witness, and optionally A.w, B.w, C.w -- exactly the arrays bp_cs_enforce /
// bp_cs_alloc take, so another handle (or the oracle) can ingest the file as it is.  Little-endian:
//   char magic[8] = "BPR1CSX\1"; u32 version = 1; u32 field; u64 n_rows, n_inputs, n_aux, nnz, row_base; u32 flags (bit 0: A.w,
//   B.w, C.w follow the witness); u32 reserved;
//   u32 lens[3 n_rows] (|A_i|, |B_i|, |C_i|); u32 cols[nnz] (bit 31 = aux); u64 coeffs[nnz][4]; u64 inputs[n_inputs][4];
//   u64 aux[n_aux][4]; (flags & 1) u64 az[n_rows][4], bz[n_rows][4], cz[n_rows][4]; u64 checksum (as in bp_cs_save).
extern "C" int bp_cs_export(bp_cs* h, const char* path, int with_products) {
    if (!h || !path) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (!h->wide_valid) {
        for (int k = 0; k < 2; ++k) {
            const uint64_t n = k ? h->n_aux : h->n_inputs;
            if (!n) continue;
            materialize_wide<<<grid_for(h, 2 * n, 256, 8), 256, 0, h->stream>>>(shadow_ptr(h, k), n, (uint4*)(k ? h->aux.p : h->inputs.p));
            h->launches++;
        }
        CU(h, cudaGetLastError());
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(h, BP_E_STATE, "cannot open %s for writing", path);
    struct {
        char magic[8];
        uint32_t version, field;
        uint64_t n_rows, n_inputs, n_aux, nnz, row_base;
        uint32_t flags, reserved;
    } hd;
    std::memset(&hd, 0, sizeof hd);
    std::memcpy(hd.magic, "BPR1CSX\1", 8);
    hd.version = 1;
    hd.field = (uint32_t)h->field;
    hd.n_rows = h->n_rows; hd.n_inputs = h->n_inputs; hd.n_aux = h->n_aux; hd.nnz = h->nnz; hd.row_base = h->row_base;
    hd.flags = with_products ? 1u : 0u;
    Checksum ck;
    auto put = [&](const void* p, size_t n) {
        ck.add(p, n);
        return fwrite(p, 1, n, f) == n;
    };
    int rc = fwrite(&hd, sizeof hd, 1, f) == 1 ? BP_OK : fail(h, BP_E_STATE, "short write");
    const size_t n_lc = 3 * (size_t)h->n_rows;
    // lens: differences of the row offsets, chunk by chunk through the pinned staging buffer
    {
        uint32_t* st = (uint32_t*)h->h_stage[0];
        const size_t per = kStageBytes / 4 - 1;
        for (size_t i = 0; i < n_lc && rc == BP_OK; i += per) {
            const size_t n = std::min(per, n_lc - i);
            cudaError_t e = cudaMemcpyAsync(st, (const uint32_t*)h->row_ptr.p + i, (n + 1) * 4, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { rc = fail(h, BP_E_CUDA, "export: %s", cudaGetErrorString(e)); break; }
            for (size_t j = 0; j < n; ++j) st[j] = st[j + 1] - st[j];
            if (!put(st, n * 4)) rc = fail(h, BP_E_STATE, "short write");
        }
    }
    // columns without the class bits
    {
        uint32_t* st = (uint32_t*)h->h_stage[0];
        const size_t per = kStageBytes / 4;
        for (size_t i = 0; i < h->nnz && rc == BP_OK; i += per) {
            const size_t n = std::min(per, (size_t)h->nnz - i);
            cudaError_t e = cudaMemcpyAsync(st, (const uint32_t*)h->cols.p + i, n * 4, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { rc = fail(h, BP_E_CUDA, "export: %s", cudaGetErrorString(e)); break; }
            for (size_t j = 0; j < n; ++j) st[j] &= (kColAux | kColIdxMask);
            if (!put(st, n * 4)) rc = fail(h, BP_E_STATE, "short write");
        }
    }
    // canonical coefficients
    if (rc == BP_OK && h->nnz) {
        const size_t per = kStageBytes / 32;
        rc = ensure(h, h->scratch, per * 32, 0);
        for (size_t i = 0; i < h->nnz && rc == BP_OK; i += per) {
            const size_t n = std::min(per, (size_t)h->nnz - i);
            DISPATCH_FIELD(h, (export_terms<F><<<grid_for(h, n, 128, 16), 128, 0, h->stream>>>((const uint4*)h->vals.p, (const uint32_t*)h->cols.p,
                                                                                              (const uint32_t*)h->row_ptr.p, (uint32_t)n_lc + 1u,
                                                                                              (uint32_t)i, (uint32_t)n, (uint4*)h->scratch.p)));
            h->launches++;
            cudaError_t e = cudaMemcpyAsync(h->h_stage[0], h->scratch.p, n * 32, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { rc = fail(h, BP_E_CUDA, "export: %s", cudaGetErrorString(e)); break; }
            if (!put(h->h_stage[0], n * 32)) rc = fail(h, BP_E_STATE, "short write");
        }
    }
    auto put_dev = [&](const void* dev, size_t bytes) {
        for (size_t off = 0; off < bytes && rc == BP_OK; off += kStageBytes) {
            const size_t n = std::min(kStageBytes, bytes - off);
            cudaError_t e = cudaMemcpyAsync(h->h_stage[0], (const char*)dev + off, n, cudaMemcpyDeviceToHost, h->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
            if (e != cudaSuccess) { rc = fail(h, BP_E_CUDA, "export: %s", cudaGetErrorString(e)); break; }
            if (!put(h->h_stage[0], n)) rc = fail(h, BP_E_STATE, "short write");
        }
    };
    put_dev(h->inputs.p, (size_t)h->n_inputs * 32);
    put_dev(h->aux.p, (size_t)h->n_aux * 32);
    if (rc == BP_OK && with_products && h->n_rows) {
        const size_t bytes = (size_t)h->n_rows * 32;
        DevBuf out;  // (not h->scratch: launch_check's kernels use it through the plan)
        rc = ensure(h, out, 3 * bytes, 0);
        if (rc == BP_OK) {
            uint4* d = (uint4*)out.p;
            rc = launch_check(h, h->d_result, d, (uint4*)((char*)out.p + bytes), (uint4*)((char*)out.p + 2 * bytes));
            if (rc == BP_OK) put_dev(out.p, 3 * bytes);
        }
        if (out.p) {
            cudaStreamSynchronize(h->stream);
            cudaFree(out.p);
        }
    }
    const uint64_t sum = ck.value();
    if (rc == BP_OK && fwrite(&sum, 8, 1, f) != 1) rc = fail(h, BP_E_STATE, "short write");
    if (fclose(f) != 0 && rc == BP_OK) rc = fail(h, BP_E_STATE, "close failed");
    return rc;
}

// ---- witness program: generate the next witness on the device ---------------------------------------------------
namespace {
// Everything the kernel indexes with is checked here, once: a malformed program is refused, never run.
const char* wprog_validate(const uint32_t* w, uint64_t n_words, uint64_t n_aux) {
    if (n_words < kWpHeaderWords || w[0] != kWpMagic || w[1] != 1) return "not a witness program (magic / version)";
    const uint64_t n_units = w[2], n_tapes = w[3], msg_base = w[4], n_msg = w[5], total = w[7], unit_off = w[10], tape_off = w[11];
    if (w[12] != n_words || total != n_aux) return "program size or variable count does not match the system";
    if (unit_off != kWpHeaderWords || tape_off != unit_off + 4 * n_units || tape_off + 8 * n_tapes > n_words || !n_units || !n_tapes)
        return "bad section offsets";
    if (msg_base + n_msg > n_aux) return "message bits exceed the aux space";
    std::vector<uint32_t> tape_vars(n_tapes);
    for (uint64_t t = 0; t < n_tapes; ++t) {
        const uint32_t* tr = w + tape_off + 8 * t;
        const uint64_t n_vars = tr[0], lev = tr[1], n_lev = tr[2], ent = tr[3], n_ent = tr[4], sum = tr[5], n_sum = tr[6], sop = tr[7];
        tape_vars[t] = tr[0];
        if (lev % 4 || ent % 4 || sum % 4 || lev + 4 * n_lev > n_words || ent + 4 * n_ent > n_words || sum + 4 * n_sum > n_words || sop > n_words ||
            n_vars >= (1u << 24) || n_ent != n_vars || n_sum > w[9] || n_vars > w[8])
            return "bad tape record";
        uint64_t e_prev = 0, s_prev = 0;
        for (uint64_t l = 0; l < n_lev; ++l) {
            const uint32_t* lr = w + lev + 4 * l;
            if (lr[0] != e_prev || lr[1] < lr[0] || lr[1] > n_ent || lr[2] != s_prev || lr[3] < lr[2] || lr[3] > n_sum) return "bad level record";
            e_prev = lr[1];
            s_prev = lr[3];
        }
        if (e_prev != n_ent || s_prev != n_sum) return "levels do not cover the tape";
        uint64_t max_sop = 0;
        auto operand_ok = [&](uint32_t op, uint32_t mask) {
            const uint32_t kind = op >> 29, p = op & mask;
            if (kind < 2) return true;
            if (kind < 4) return p < n_vars;
            if (kind < 6) return true;  // message offsets depend on the unit: checked below per unit through max_msg
            return p < 256u;
        };
        for (uint64_t k = 0; k < n_sum; ++k) {
            const uint32_t* sr = w + sum + 4 * k;
            if ((uint64_t)sop + sr[0] + sr[1] > n_words) return "sum operands out of range";
            max_sop = std::max<uint64_t>(max_sop, (uint64_t)sr[0] + sr[1]);
            for (uint32_t i = 0; i < sr[1]; ++i)
                if (!operand_ok(w[sop + sr[0] + i], 0x00ffffffu)) return "bad sum operand";
        }
        std::vector<uint8_t> written(n_vars, 0);
        for (uint64_t e = 0; e < n_ent; ++e) {
            const uint32_t* er = w + ent + 4 * e;
            const uint32_t op = er[0] >> 28, res = er[0] & 0x0fffffffu;
            if (res >= n_vars || written[res] || op == kWpFree || op > kWpSumBit) return "bad tape entry";
            written[res] = 1;
            if (op == kWpSumBit) {
                if (er[1] >= n_sum || er[2] >= 64) return "bad sum bit";
            } else if (!operand_ok(er[1], 0x1fffffffu) || !operand_ok(er[2], 0x1fffffffu) ||
                       ((op == kWpCh || op == kWpMaj) && !operand_ok(er[3], 0x1fffffffu))) {
                return "bad operand";
            }
        }
    }
    // units: tapes exist, variables inside the aux space and disjoint from the message bits, message windows inside the message
    for (uint64_t u = 0; u < n_units; ++u) {
        const uint32_t* ur = w + unit_off + 4 * u;
        if (ur[0] >= n_tapes || (uint64_t)ur[1] + tape_vars[ur[0]] > n_aux || ur[3] >= n_units || ur[2] > n_msg) return "bad unit record";
    }
    // message operand offsets: at most 2^24 past the unit's base by construction; bound them by the message length here
    for (uint64_t u = 0; u < n_units; ++u) {
        const uint32_t* ur = w + unit_off + 4 * u;
        const uint32_t* tr = w + tape_off + 8 * ur[0];
        const uint64_t room = n_msg - ur[2];
        auto msg_ok = [&](uint32_t op, uint32_t mask) { const uint32_t k = op >> 29; return k < 4 || k >= 6 || (op & mask) < room; };
        for (uint64_t e = 0; e < tr[4]; ++e) {
            const uint32_t* er = w + tr[3] + 4 * e;
            if ((er[0] >> 28) != kWpSumBit && !(msg_ok(er[1], 0x1fffffffu) && msg_ok(er[2], 0x1fffffffu) && msg_ok(er[3], 0x1fffffffu)))
                return "message operand past the end of the message";
        }
        for (uint64_t k = 0; k < tr[6]; ++k) {
            const uint32_t* sr = w + tr[5] + 4 * k;
            for (uint32_t i = 0; i < sr[1]; ++i)
                if (!msg_ok(w[tr[7] + sr[0] + i], 0x00ffffffu)) return "message operand past the end of the message";
        }
    }
    return nullptr;
}
}  // namespace

extern "C" int bp_cs_set_witness_program(bp_cs* h, const uint32_t* words, uint64_t n_words) {
    if (!h || !words) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    if (const char* why = wprog_validate(words, n_words, h->n_aux)) return fail(h, BP_E_ARG, "bp_cs_set_witness_program: %s", why);
    int rc = ensure(h, h->wprog, (size_t)n_words * 4, 0);
    if (rc != BP_OK) return rc;
    CU(h, cudaMemcpyAsync(h->wprog.p, words, (size_t)n_words * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    std::memcpy(h->wprog_hdr, words, sizeof h->wprog_hdr);
    return BP_OK;
}

extern "C" int bp_cs_generate_witness_async(bp_cs* h, const uint8_t* msg, uint64_t msg_len, const uint32_t* states, uint64_t n_state_words) {
    if (!h || !msg || !states) return BP_E_ARG;
    if (h->wprog_hdr[0] != kWpMagic) return fail(h, BP_E_STATE, "no witness program installed (bp_cs_set_witness_program)");
    const uint32_t n_units = h->wprog_hdr[2], n_msg_bits = h->wprog_hdr[5];
    if (msg_len * 8 != n_msg_bits || n_state_words != 8ull * n_units || h->wprog_hdr[7] != h->n_aux)
        return fail(h, BP_E_ARG, "bp_cs_generate_witness: message of %llu bits / %llu state words, program wants %u / %llu",
                    (unsigned long long)(msg_len * 8), (unsigned long long)n_state_words, n_msg_bits, (unsigned long long)(8ull * n_units));
    CU(h, cudaSetDevice(h->device));
    const size_t state_off = ((size_t)msg_len + 15) & ~size_t(15);
    int rc = ensure(h, h->wprog_in, state_off + (size_t)n_state_words * 4, 0);
    if (rc != BP_OK) return rc;
    if ((rc = upload(h, h->wprog_in.p, msg, (size_t)msg_len)) != BP_OK) return rc;
    if ((rc = upload(h, (char*)h->wprog_in.p + state_off, states, (size_t)n_state_words * 4)) != BP_OK) return rc;
    uint32_t* aux_shadow = shadow_ptr(h, 1);
    h->wide_valid = false;  // only the shadows are written (kernels.cuh: ld_witness)
    wprog_expand_msg<<<grid_for(h, n_msg_bits, 256, 8), 256, 0, h->stream>>>((const uint8_t*)h->wprog_in.p, n_msg_bits, h->wprog_hdr[6],
                                                                            aux_shadow + h->wprog_hdr[4]);
    const uint32_t bit_words = (h->wprog_hdr[8] + 31) / 32 + 1, sum_slots = h->wprog_hdr[9] + 1;
    const size_t smem = (size_t)kWpWarps * ((size_t)sum_slots * 8 + (size_t)bit_words * 4);
    if (smem > 200 * 1024) return fail(h, BP_E_RANGE, "witness program unit too large for shared memory");
    if (smem > 48 * 1024) CU(h, cudaFuncSetAttribute(wprog_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n_units + kWpWarps - 1) / kWpWarps, (uint64_t)h->sm_count * 16));
    wprog_run<<<grid, 32 * kWpWarps, smem, h->stream>>>((const uint32_t*)h->wprog.p, (const uint8_t*)h->wprog_in.p,
                                                        (const uint32_t*)((char*)h->wprog_in.p + state_off), aux_shadow, n_units, bit_words, sum_slots);
    h->launches += 2;
    CU(h, cudaGetLastError());
    return settle(h);
}

// ---- K4: batched set (test_cs.rs:270-282) ----------------------------------------------------------------------------------
namespace {
int ensure_pinned(bp_cs* h, void*& p, size_t& cap, size_t need) {
    if (need <= cap) return BP_OK;
    if (p) {
        CU(h, cudaStreamSynchronize(h->stream));
        cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    size_t want = std::max(need + need / 4, (size_t)4096);
    if (cudaMallocHost(&p, want) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(h, BP_E_OOM, "cudaMallocHost(%zu bytes)", want);
    }
    cap = want;
    return BP_OK;
}

bool limbs_below_p(const bp_cs* h, const uint64_t* v) {
    uint32_t pl[8];
    DISPATCH_FIELD(h, { for (int i = 0; i < 8; ++i) pl[i] = PL<F>(i); });
    const uint32_t* x = (const uint32_t*)v;
    for (int j = 7; j >= 0; --j) {
        if (x[j] < pl[j]) return true;
        if (x[j] > pl[j]) return false;
    }
    return false;
}

// n (index, value) pairs already in h_patch ([u32 idx x n, padded to 32 bytes][32-byte values x n]) -> device -> scatter
int launch_patch(bp_cs* h, int is_aux, uint32_t n) {
    if (!n) return BP_OK;
    const size_t off_vals = ((size_t)n * 4 + 31) & ~size_t(31);
    int rc = ensure(h, h->patch_stage, off_vals + (size_t)n * 32, 0);
    if (rc != BP_OK) return rc;
    CU(h, cudaMemcpyAsync(h->patch_stage.p, h->h_patch, off_vals + (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
    patch_witness<<<grid_for(h, n, 256, 8), 256, 0, h->stream>>>((const uint32_t*)h->patch_stage.p, (const uint4*)((char*)h->patch_stage.p + off_vals),
                                                               n, (uint4*)(is_aux ? h->aux.p : h->inputs.p), shadow_ptr(h, is_aux));
    h->launches++;
    CU(h, cudaGetLastError());
    return BP_OK;
}
}  // namespace

extern "C" int bp_cs_set_many(bp_cs* h, int is_aux, uint64_t n, const uint64_t* idx, const uint64_t* vals) {
    if (!h || (n && (!idx || !vals))) return BP_E_ARG;
    if (!n) return BP_OK;
    if (n > 0x7fffffffull) return fail(h, BP_E_RANGE, "bp_cs_set_many: batch too large");
    CU(h, cudaSetDevice(h->device));
    const uint64_t cnt = is_aux ? h->n_aux : h->n_inputs;
    for (uint64_t i = 0; i < n; ++i) {  // validate everything first: a rejected batch changes nothing
        if (idx[i] >= cnt) return fail(h, BP_E_RANGE, "bp_cs_set_many: index %llu >= %llu", (unsigned long long)idx[i], (unsigned long long)cnt);
        if (!limbs_below_p(h, vals + 4 * i)) return fail(h, BP_E_RANGE, "bp_cs_set_many: element %llu is not canonical (>= p)", (unsigned long long)i);
    }
    const size_t off_vals = ((size_t)n * 4 + 31) & ~size_t(31);
    CU(h, cudaStreamSynchronize(h->stream));  // (an earlier batch may still be on its way out of h_patch)
    int rc = ensure_pinned(h, h->h_patch, h->h_patch_cap, off_vals + (size_t)n * 32);
    if (rc != BP_OK) return rc;
    uint32_t* pi = (uint32_t*)h->h_patch;
    for (uint64_t i = 0; i < n; ++i) pi[i] = (uint32_t)idx[i];
    std::memcpy((char*)h->h_patch + off_vals, vals, (size_t)n * 32);
    return launch_patch(h, is_aux, (uint32_t)n);
}

// ---- same circuit, next witness, given as the reference holds it: 32-byte scalars (WitnessCS::input_assignment /
// aux_assignment, witness_cs.rs:45-57) ---------------------------------------------------------------------------------------
// The host packs before it sends: every value that is 0 or 1 becomes one bit, everything else goes to an exception list
// (index + 32 bytes) that a scatter kernel applies after the bits have been widened.  A gadget witness of 10^8 bits is then
// 14 MB of PCIe traffic instead of 3.5 GB; the packing pass (all host threads, one sequential read of the scalars) is the
// cost of the call.  A witness with many non-bit values is sent as it is (bp_cs_set_range).
// `mont`: the scalars are in their in-memory Montgomery form (x * 2^256 mod p; bp_cs_recheck_scalars_mont) -- the pass looks
// for the limb patterns of 0 and 2^256 mod p instead of 0 and 1, and the exceptions are converted on the host.
static int recheck_scalars_impl(bp_cs* h, const uint64_t* inputs_le, const uint64_t* aux_le, int64_t* dev_result, bool mont) {
    if (!h || !dev_result || (!aux_le && h->n_aux)) return BP_E_ARG;
    CU(h, cudaSetDevice(h->device));
    const uint64_t n_in = inputs_le ? h->n_inputs : 0, n_aux = h->n_aux;
    const size_t in_bytes = (size_t)((n_in + 7) / 8), aux_bytes = (size_t)((n_aux + 7) / 8);
    const size_t aux_off = (in_bytes + 63) & ~size_t(63);
    CU(h, cudaStreamSynchronize(h->stream));  // the pinned buffers of the previous call are free again
    int rc = ensure_pinned(h, h->h_pack, h->h_pack_cap, aux_off + aux_bytes + 64);
    if (rc != BP_OK) return rc;
    const unsigned nt = pack_threads();
    using Exc = PackExc;  // host/pack.hpp: the packing kernel (AVX2 when the CPU has it) lives in its own host-only file
    uint8_t* bits_in = (uint8_t*)h->h_pack;
    uint8_t* bits_aux = bits_in + aux_off;
    std::vector<std::vector<Exc>> exc((size_t)nt * 2);
    // With "sparse_upload" (row shards) only the pieces of the aux witness that this handle's rows read are packed and sent.
    std::vector<std::pair<uint64_t, uint64_t>> aux_bytes_ranges;  // [first byte, end byte) of the aux bit string
    const bool sparse = h->sparse_upload && h->n_rows && n_aux;
    if (sparse) {
        if ((rc = ensure_plan(h)) != BP_OK || (rc = ensure_chunk_plan(h, 16)) != BP_OK) return rc;
        for (int i = 0; i < h->chunk_plan.n_pieces; ++i)  // (piece offsets are multiples of 256 elements)
            aux_bytes_ranges.push_back({h->chunk_plan.off[i] / 8, (h->chunk_plan.off[i] + h->chunk_plan.len[i] + 7) / 8});
    } else if (n_aux) {
        aux_bytes_ranges.push_back({0, aux_bytes});
    }
    uint64_t aux_range_bytes = 0;
    for (auto& r : aux_bytes_ranges) aux_range_bytes += r.second - r.first;
    // pack_bit_bytes[_mont]: bytes [b0, b1) of the bit string of `src` (elements [8*b0, min(8*b1, n)))
    uint64_t one_mont[4] = {1, 0, 0, 0};
    if (mont) mont_one(h->field, one_mont);
    auto pack_some = [&](const uint64_t* src, uint64_t n, uint8_t* dst, uint64_t b0, uint64_t b1, std::vector<Exc>& out) {
        if (mont) pack_bit_bytes_mont(src, n, one_mont, dst, b0, b1, out);
        else pack_bit_bytes(src, n, dst, b0, b1, out);
    };
    std::atomic<bool> pack_failed{false};
    auto work = [&](unsigned t) {
        try {
            if (n_in) pack_some(inputs_le, n_in, bits_in, in_bytes * t / nt, in_bytes * (t + 1) / nt, exc[2 * t]);
            // thread t takes the slice [s0, s1) of the concatenated aux byte ranges
            uint64_t s0 = aux_range_bytes * t / nt, s1 = aux_range_bytes * (t + 1) / nt, base = 0;
            for (auto& r : aux_bytes_ranges) {
                const uint64_t len = r.second - r.first, lo = std::max(s0, base), hi = std::min(s1, base + len);
                if (lo < hi) pack_some(aux_le, n_aux, bits_aux, r.first + (lo - base), r.first + (hi - base), exc[2 * t + 1]);
                base += len;
            }
        } catch (...) {  // (an exception list that cannot grow: nothing may leave a thread)
            pack_failed = true;
        }
    };
    try {
        run_on_threads(nt, work);
    } catch (...) {
        return fail(h, BP_E_OOM, "bp_cs_recheck_scalars: out of host memory while packing");
    }
    if (pack_failed) return fail(h, BP_E_OOM, "bp_cs_recheck_scalars: out of host memory while packing");
    size_t n_exc[2] = {0, 0};
    for (unsigned t = 0; t < nt; ++t) {
        n_exc[0] += exc[2 * t].size();
        n_exc[1] += exc[2 * t + 1].size();
    }
    if (n_exc[0] + n_exc[1] > (n_in + (sparse ? 8 * aux_range_bytes : n_aux)) / 16 + 1024) {  // not a bit witness: send it as it is
        if (mont) {  // ... in canonical form: one conversion pass on the host threads
            std::vector<uint64_t> canon;
            try {
                canon.resize(4 * (size_t)std::max(n_in, n_aux));
            } catch (...) {
                return fail(h, BP_E_OOM, "bp_cs_recheck_scalars_mont: no host memory for the canonical copy");
            }
            for (int k = 0; k < 2; ++k) {
                const uint64_t n = k ? n_aux : n_in;
                if (!n) continue;
                if (bp_scalars_from_mont(h->field, k ? aux_le : inputs_le, n, canon.data()) != BP_OK)
                    return fail(h, BP_E_RANGE, "bp_cs_recheck_scalars_mont: an %s element is >= p (not a Montgomery-form scalar)", k ? "aux" : "input");
                if ((rc = bp_cs_set_range(h, k, 0, n, canon.data())) != BP_OK) return rc;  // (synchronous: canon may be reused)
            }
            return check_graphed(h, (long long*)dev_result, nullptr);
        }
        if (n_in && (rc = bp_cs_set_range(h, 0, 0, n_in, inputs_le)) != BP_OK) return rc;
        if (n_aux && (rc = bp_cs_set_range(h, 1, 0, n_aux, aux_le)) != BP_OK) return rc;
        return check_graphed(h, (long long*)dev_result, nullptr);
    }
    for (int k = 0; k < 2; ++k)  // the exceptions must be field elements (the bits are by construction)
        for (unsigned t = 0; t < nt; ++t)
            for (Exc& e : exc[2 * t + k])
                if (mont ? !from_mont(h->field, e.v, e.v) : !limbs_below_p(h, e.v))
                    return fail(h, BP_E_RANGE, "bp_cs_recheck_scalars: %s element %llu is not canonical (>= p)", k ? "aux" : "input",
                                (unsigned long long)e.idx);
    if (n_in && (rc = set_range_packed(h, 0, 0, n_in, bits_in, true, false)) != BP_OK) return rc;
    if (sparse) {
        if ((rc = ensure(h, h->u8_stage, aux_bytes, 0)) != BP_OK) return rc;
        h->wide_valid = false;
        for (int i = 0; i < h->chunk_plan.n_pieces; ++i) {
            const uint64_t off = h->chunk_plan.off[i], len = h->chunk_plan.len[i];
            CU(h, cudaMemcpyAsync((char*)h->u8_stage.p + off / 8, bits_aux + off / 8, (size_t)((len + 7) / 8), cudaMemcpyHostToDevice, h->stream));
            launch_widen(h, true, off, len, shadow_ptr(h, 1) + off);
        }
        CU(h, cudaGetLastError());
    } else if (n_aux && (rc = set_range_packed(h, 1, 0, n_aux, bits_aux, true, false)) != BP_OK) {
        return rc;
    }
    for (int k = 0; k < 2; ++k) {
        const size_t n = n_exc[k];
        if (!n) continue;
        const size_t off_vals = (n * 4 + 31) & ~size_t(31);
        CU(h, cudaStreamSynchronize(h->stream));  // (h_patch is reused for the second index space)
        if ((rc = ensure_pinned(h, h->h_patch, h->h_patch_cap, off_vals + n * 32)) != BP_OK) return rc;
        uint32_t* pi = (uint32_t*)h->h_patch;
        uint64_t* pv = (uint64_t*)((char*)h->h_patch + off_vals);
        size_t j = 0;
        for (unsigned t = 0; t < nt; ++t)
            for (const Exc& e : exc[2 * t + k]) {
                pi[j] = (uint32_t)e.idx;
                std::memcpy(pv + 4 * j, e.v, 32);
                ++j;
            }
        if ((rc = launch_patch(h, k, (uint32_t)n)) != BP_OK) return rc;
    }
    return check_graphed(h, (long long*)dev_result, nullptr);
}

extern "C" int bp_cs_recheck_scalars_async(bp_cs* h, const uint64_t* inputs_le, const uint64_t* aux_le, int64_t* dev_result) {
    return recheck_scalars_impl(h, inputs_le, aux_le, dev_result, false);
}

extern "C" int bp_cs_recheck_scalars_mont_async(bp_cs* h, const uint64_t* inputs_mont, const uint64_t* aux_mont, int64_t* dev_result) {
    return recheck_scalars_impl(h, inputs_mont, aux_mont, dev_result, true);
}

static int recheck_scalars_sync(bp_cs* h, const uint64_t* inputs_le, const uint64_t* aux_le, int64_t* row, bool mont) {
    if (!h || !row) return BP_E_ARG;
    int rc = recheck_scalars_impl(h, inputs_le, aux_le, (int64_t*)h->d_result, mont);
    if (rc != BP_OK) return rc;
    long long fb;
    unsigned int e;
    if ((rc = read_flags(h, &fb, &e)) != BP_OK) return rc;
    if (e & 1u) return fail(h, BP_E_RANGE, "a term references a variable index that does not exist");
    *row = fb == 0x7fffffffffffffffLL ? -1 : (int64_t)(fb - (long long)h->row_base);
    return BP_OK;
}

extern "C" int bp_cs_recheck_scalars(bp_cs* h, const uint64_t* inputs_le, const uint64_t* aux_le, int64_t* row) {
    return recheck_scalars_sync(h, inputs_le, aux_le, row, false);
}

extern "C" int bp_cs_recheck_scalars_mont(bp_cs* h, const uint64_t* inputs_mont, const uint64_t* aux_mont, int64_t* row) {
    return recheck_scalars_sync(h, inputs_mont, aux_mont, row, true);
}

extern "C" {

// ---- multi-GPU group ----------------------------------------------------------------------------------------------------
int bp_group_unique_id(uint8_t id[BP_GROUP_ID_BYTES]) {
    if (!id) return BP_E_ARG;
    const NcclApi* api = nccl_api();
    if (!api) return BP_E_STATE;
    static_assert(sizeof(ncclUniqueId) == BP_GROUP_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    if (api->GetUniqueId(&u) != ncclSuccess) return BP_E_CUDA;
    std::memcpy(id, &u, sizeof u);
    return BP_OK;
}

void bp_group_free(bp_group* g) {
    if (!g) return;
    bp_cs* h = g->cs;
    if (h) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        if (h->graph.group == g) drop_graph(h);
    }
    for (int r = 0; r < g->world && r < kMaxGroup; ++r)
        if (g->peer_mapped[r]) cudaIpcCloseMemHandle(g->peer_mapped[r]);
    // Collective: nobody frees its mailbox while a peer still has it mapped (freeing exported memory before every importer
    // has closed it is undefined) -- one all-reduce after the closes is the barrier.
    if (g->ready && g->world > 1 && g->comm && g->d_epoch && h) {
        const NcclApi* api = nccl_api();
        if (api && api->AllReduce(g->d_epoch, g->d_epoch, 1, ncclUint64, ncclMin, g->comm, h->stream) == ncclSuccess)
            cudaStreamSynchronize(h->stream);
    }
    if (g->d_peer_boxes) cudaFree(g->d_peer_boxes);
    if (g->box) cudaFree(g->box);
    if (g->d_epoch) cudaFree(g->d_epoch);
    if (g->h_result) cudaFreeHost(g->h_result);
    if (g->comm) {
        const NcclApi* api = nccl_api();
        if (api) api->CommDestroy(g->comm);
    }
    (void)cudaGetLastError();
    delete g;
}

int bp_group_init(bp_cs* h, const uint8_t id[BP_GROUP_ID_BYTES], int rank, int world, bp_group** out) {
    if (!h || !out || world < 1 || world > kMaxGroup || rank < 0 || rank >= world || (world > 1 && !id)) return BP_E_ARG;
    *out = nullptr;
    CU(h, cudaSetDevice(h->device));
    bp_group* g = new (std::nothrow) bp_group();
    if (!g) return BP_E_OOM;
    g->cs = h;
    g->rank = rank;
    g->world = world;
    auto bail = [&](int code) {
        bp_group_free(g);
        return code;
    };
    if (cudaMallocHost((void**)&g->h_result, 64) != cudaSuccess) return bail(fail(h, BP_E_OOM, "cudaMallocHost"));
    if (world == 1) {
        *out = g;
        return BP_OK;
    }
    const NcclApi* api = nccl_api();
    if (!api) return bail(fail(h, BP_E_STATE, "libnccl.so.2 not found: a group of more than one rank needs NCCL"));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    {
        ncclResult_t r = api->CommInitRank(&g->comm, world, u, rank);
        if (r != ncclSuccess) return bail(fail(h, BP_E_CUDA, "ncclCommInitRank: %s", api->GetErrorString(r)));
    }
    // mailbox: allocate + zero, export, all-gather the IPC handles (the all-gather also orders every rank's memset before any
    // peer's first store), map the peers.  Any failure on any rank -> everybody uses the NCCL all-reduce instead.
    const size_t box_bytes = sizeof(GroupSlot) * 2 * world;
    unsigned char* d_handles = nullptr;  // world x (64-byte handle + 8-byte ok flag), 128 bytes per rank
    constexpr size_t kRec = 128;
    bool ok = cudaMalloc((void**)&g->box, box_bytes) == cudaSuccess && cudaMalloc((void**)&g->d_epoch, 8) == cudaSuccess &&
              cudaMalloc((void**)&g->d_peer_boxes, sizeof(GroupSlot*) * world) == cudaSuccess &&
              cudaMalloc((void**)&d_handles, kRec * world) == cudaSuccess;
    if (!ok) {
        (void)cudaGetLastError();
        if (d_handles) cudaFree(d_handles);
        return bail(fail(h, BP_E_OOM, "group mailbox allocation"));
    }
    CU(h, cudaMemsetAsync(g->box, 0, box_bytes, h->stream));
    CU(h, cudaMemsetAsync(g->d_epoch, 0, 8, h->stream));
    unsigned char rec[kRec];
    std::memset(rec, 0, sizeof rec);
    cudaIpcMemHandle_t mh;
    static_assert(sizeof mh <= 64, "cudaIpcMemHandle_t size");
    uint64_t my_ok = cudaIpcGetMemHandle(&mh, g->box) == cudaSuccess ? 1 : 0;
    if (!my_ok) (void)cudaGetLastError();
    std::memcpy(rec, &mh, sizeof mh);
    std::memcpy(rec + 64, &my_ok, 8);
    CU(h, cudaMemcpyAsync(d_handles + kRec * rank, rec, kRec, cudaMemcpyHostToDevice, h->stream));
    {
        ncclResult_t r = api->AllGather(d_handles + kRec * rank, d_handles, kRec, ncclUint8, g->comm, h->stream);
        if (r != ncclSuccess) {
            cudaFree(d_handles);
            return bail(fail(h, BP_E_CUDA, "ncclAllGather: %s", api->GetErrorString(r)));
        }
    }
    std::vector<unsigned char> all(kRec * world);
    CU(h, cudaMemcpyAsync(all.data(), d_handles, kRec * world, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    std::vector<GroupSlot*> peers(world, nullptr);
    uint64_t mapped = 1;
    for (int r = 0; r < world; ++r) {
        uint64_t rok;
        std::memcpy(&rok, all.data() + kRec * r + 64, 8);
        if (!rok) { mapped = 0; continue; }
        if (r == rank) { peers[r] = g->box; continue; }
        cudaIpcMemHandle_t ph;
        std::memcpy(&ph, all.data() + kRec * r, sizeof ph);
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            (void)cudaGetLastError();
            mapped = 0;
            continue;
        }
        g->peer_mapped[r] = p;
        peers[r] = (GroupSlot*)p;
    }
    // everybody must agree on the transport: MIN over the ranks of "I mapped every peer" (reusing the handle buffer)
    CU(h, cudaMemcpyAsync(d_handles, &mapped, 8, cudaMemcpyHostToDevice, h->stream));
    {
        ncclResult_t r = api->AllReduce(d_handles, d_handles, 1, ncclUint64, ncclMin, g->comm, h->stream);
        if (r != ncclSuccess) {
            cudaFree(d_handles);
            return bail(fail(h, BP_E_CUDA, "ncclAllReduce: %s", api->GetErrorString(r)));
        }
    }
    CU(h, cudaMemcpyAsync(&mapped, d_handles, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    cudaFree(d_handles);
    g->mailbox = mapped != 0 && !getenv("BP_GROUP_NO_MAILBOX");
    if (g->mailbox) CU(h, cudaMemcpyAsync(g->d_peer_boxes, peers.data(), sizeof(GroupSlot*) * world, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    g->ready = true;
    *out = g;
    return BP_OK;
}

int bp_group_info(bp_group* g, int* rank, int* world, int* transport) {
    if (!g) return BP_E_ARG;
    if (rank) *rank = g->rank;
    if (world) *world = g->world;
    if (transport) *transport = g->world == 1 ? 0 : (g->mailbox ? 2 : 1);
    return BP_OK;
}

int bp_group_reduce_async(bp_group* g, int64_t* dev_result) {
    if (!g || !dev_result) return BP_E_ARG;
    CU(g->cs, cudaSetDevice(g->cs->device));
    return enqueue_reduce(g, (long long*)dev_result);
}

int bp_group_check_async(bp_group* g, int64_t* dev_result) {
    if (!g || !dev_result) return BP_E_ARG;
    CU(g->cs, cudaSetDevice(g->cs->device));
    return check_graphed(g->cs, (long long*)dev_result, g);
}

int bp_group_check(bp_group* g, int64_t* row) {
    if (!g || !row) return BP_E_ARG;
    bp_cs* h = g->cs;
    CU(h, cudaSetDevice(h->device));
    int rc = check_graphed(h, h->d_result, g);
    if (rc != BP_OK) return rc;
    long long fb;
    unsigned int e;
    if ((rc = read_flags(h, &fb, &e)) != BP_OK) return rc;
    if (e & 1u) return fail(h, BP_E_RANGE, "a term references a variable index that does not exist");
    *row = fb == 0x7fffffffffffffffLL ? -1 : (int64_t)fb;  // GLOBAL row (row bases included), the same on every rank
    return BP_OK;
}

// The witness of `root` replaces everybody's: inputs, aux and their shadows travel once over NVLink (ncclBroadcast); the
// counts must agree (the witness is replicated).
int bp_group_broadcast_witness(bp_group* g, int root) {
    if (!g || root < 0 || root >= g->world) return BP_E_ARG;
    bp_cs* h = g->cs;
    CU(h, cudaSetDevice(h->device));
    if (g->world == 1) return BP_OK;
    const NcclApi* api = nccl_api();
    // header: counts + wide_valid of the root, so that a mismatch is an error instead of a hang or a corrupted witness
    unsigned long long hdr[4] = {h->n_inputs, h->n_aux, h->wide_valid ? 1ull : 0ull, 0ull};
    int rc = ensure(h, h->scratch, 64, 0);
    if (rc != BP_OK) return rc;
    CU(h, cudaMemcpyAsync(h->scratch.p, hdr, 32, cudaMemcpyHostToDevice, h->stream));
    NC(h, api, api->Broadcast(h->scratch.p, h->scratch.p, 32, ncclUint8, root, g->comm, h->stream));
    unsigned long long got[4];
    CU(h, cudaMemcpyAsync(got, h->scratch.p, 32, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    // (every rank learns the verdict from the same header, so either all proceed or none does)
    unsigned long long mine_ok = (got[0] == h->n_inputs && got[1] == h->n_aux) ? 1ull : 0ull;
    CU(h, cudaMemcpyAsync(h->scratch.p, &mine_ok, 8, cudaMemcpyHostToDevice, h->stream));
    NC(h, api, api->AllReduce(h->scratch.p, h->scratch.p, 1, ncclUint64, ncclMin, g->comm, h->stream));
    CU(h, cudaMemcpyAsync(&mine_ok, h->scratch.p, 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (!mine_ok) return fail(h, BP_E_STATE, "bp_group_broadcast_witness: the ranks hold different numbers of variables");
    NC(h, api, api->GroupStart());
    if (got[2]) {  // the 32-byte form is current on the root
        NC(h, api, api->Broadcast(h->inputs.p, h->inputs.p, (size_t)h->n_inputs * 32, ncclUint8, root, g->comm, h->stream));
        if (h->n_aux) NC(h, api, api->Broadcast(h->aux.p, h->aux.p, (size_t)h->n_aux * 32, ncclUint8, root, g->comm, h->stream));
    }
    NC(h, api, api->Broadcast(shadow_ptr(h, 0), shadow_ptr(h, 0), (size_t)h->n_inputs * 4, ncclUint8, root, g->comm, h->stream));
    if (h->n_aux) NC(h, api, api->Broadcast(shadow_ptr(h, 1), shadow_ptr(h, 1), (size_t)h->n_aux * 4, ncclUint8, root, g->comm, h->stream));
    NC(h, api, api->GroupEnd());
    if (!got[2] && h->wide_valid) {
        // the root's small values live in its shadows only; ours would now be stale where the shadow says "small"
        h->wide_valid = false;
    } else if (got[2]) {
        h->wide_valid = true;
    }
    return BP_OK;
}

// Every rank uploads ONE SLICE of a new witness from its own host memory (rank r: elements [r*cnt/world, (r+1)*cnt/world) of
// the index space; `slice` points at the first of them), the slices are all-gathered over NVLink, and the whole witness is
// validated (canonical, shadows) everywhere: world PCIe links carry 1/world of the bytes each instead of every rank pulling
// the whole witness through its own link.
int bp_group_set_witness_sharded(bp_group* g, int is_aux, const uint64_t* slice) {
    if (!g) return BP_E_ARG;
    bp_cs* h = g->cs;
    CU(h, cudaSetDevice(h->device));
    const uint64_t cnt = is_aux ? h->n_aux : h->n_inputs;
    if (!cnt) return BP_OK;
    DevBuf& b = is_aux ? h->aux : h->inputs;
    const int W = g->world;
    auto lo = [&](int r) { return (uint64_t)r * cnt / (uint64_t)W; };
    const uint64_t my0 = lo(g->rank), my1 = lo(g->rank + 1);
    if (my1 > my0 && !slice) return BP_E_ARG;
    int rc = clear_err(h);
    if (rc != BP_OK) return rc;
    if (my1 > my0 && (rc = upload(h, (char*)b.p + my0 * 32, slice, (size_t)(my1 - my0) * 32)) != BP_OK) return rc;
    if (W > 1) {
        const NcclApi* api = nccl_api();
        NC(h, api, api->GroupStart());
        for (int r = 0; r < W; ++r) {
            const uint64_t r0 = lo(r), r1 = lo(r + 1);
            if (r1 > r0) NC(h, api, api->Broadcast((char*)b.p + r0 * 32, (char*)b.p + r0 * 32, (size_t)(r1 - r0) * 32, ncclUint8, r, g->comm, h->stream));
        }
        NC(h, api, api->GroupEnd());
    }
    DISPATCH_FIELD(h, (validate_canonical<F><<<grid_for(h, cnt, 256, 8), 256, 0, h->stream>>>((const uint4*)b.p, cnt, h->d_err,
                                                                                            shadow_ptr(h, is_aux))));
    h->launches++;
    CU(h, cudaGetLastError());
    // (the other index space may still be shadow-only: wide_valid describes both, so it only goes up when that one is current)
    rc = check_err_word(h, "bp_group_set_witness_sharded");
    if (rc == BP_E_RANGE) {
        const std::string msg = h->err;
        CU(h, cudaMemsetAsync(b.p, 0, (size_t)cnt * 32, h->stream));
        CU(h, cudaMemsetAsync(shadow_ptr(h, is_aux), 0, (size_t)cnt * 4, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->err = msg + "; the index space was zeroed";
    }
    return rc;
}

// Contiguous row shards balanced by TERMS: bounds[r] = first row of rank r, bounds[world] = n_rows.
int bp_split_rows_by_nnz(const uint32_t* lens, uint64_t n_rows, int world, uint64_t* bounds) {
    if (world < 1 || !bounds || (n_rows && !lens)) return BP_E_ARG;
    uint64_t total = 0;
    for (uint64_t i = 0; i < 3 * n_rows; ++i) total += lens[i];
    total += n_rows;  // every row costs at least its offsets: keeps shards of empty rows apart
    bounds[0] = 0;
    uint64_t acc = 0, row = 0;
    for (int r = 1; r < world; ++r) {
        const uint64_t target = (uint64_t)((unsigned __int128)total * (unsigned)r / (unsigned)world);
        while (row < n_rows && acc < target) {
            acc += (uint64_t)lens[3 * row] + lens[3 * row + 1] + lens[3 * row + 2] + 1;
            ++row;
        }
        bounds[r] = row;
    }
    bounds[world] = n_rows;
    return BP_OK;
}

}  // extern "C"
