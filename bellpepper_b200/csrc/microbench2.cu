// Memory-side micro-benchmarks behind the design of the full-width check kernel (DESIGN.md "Full-width kernel: what bounds it").
// Question: how fast can random 32-byte witness gathers go on a B200 when the witness does not fit the 126 MB L2, alone and
// next to the coalesced coefficient stream, and which knobs move it (256-bit loads, L2 eviction hints, the persisting-L2
// carve-out, the L2 fetch granularity, the number of SMs issuing)?  Prints one JSON object per line.  Not on the product path.
//   usage: microbench2 [l2_fetch_granularity_bytes]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "synth.cuh"

using namespace bp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct U8 { uint32_t v[8]; };

// MODE 0: two LDG.128 (.nc)   1: one LDG.256 (.nc)   2: LDG.256 + L2::evict_last   3: LDG.256 + L2::evict_first
template <int MODE> __device__ __forceinline__ U8 ld32(const uint4* p) {
    U8 r;
    if (MODE == 0) {
        const uint4 a = __ldg(p), b = __ldg(p + 1);
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    } else if (MODE == 1) {
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    } else if (MODE == 2) {
        asm volatile("ld.global.nc.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    } else {
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    }
    return r;
}

// Random 32-byte gathers, UNROLL independent ones in flight per thread.
template <int MODE, int UNROLL> __global__ void gather(const uint4* __restrict__ w, uint32_t mask, int iters, uint32_t* out) {
    uint64_t h = sm_mix(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    uint32_t s = 0;
    for (int it = 0; it < iters; ++it) {
        U8 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            h = h * 6364136223846793005ULL + 1442695040888963407ULL;
            v[u] = ld32<MODE>(w + 2 * (size_t)((uint32_t)(h >> 33) & mask));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) s ^= v[u].v[0] ^ v[u].v[7];
    }
    if (s == 0x12345678u) out[0] = s;
}

// The memory side of the full-width kernel without its arithmetic: per "term" a warp reads 32 x 32 B of coefficients and
// 32 x 4 B of columns COALESCED (streamed once: L1 no-allocate, L2 evict-first when SHINT) and every lane gathers one random
// 32-byte witness element (GMODE as above).  DEPTH terms in flight per thread.
template <int GMODE, bool SHINT, int DEPTH>
__global__ void stream_gather(const uint4* __restrict__ vals, const uint32_t* __restrict__ cols, size_t n_terms, const uint4* __restrict__ w,
                              uint32_t mask, uint32_t* out) {
    uint32_t s = 0;
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k + (DEPTH - 1) * stride < n_terms; k += DEPTH * stride) {
        uint32_t c[DEPTH];
        U8 cv[DEPTH], wv[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            if (SHINT) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(c[d]) : "l"(cols + k + d * stride), "l"(pol));
            else c[d] = __ldg(cols + k + d * stride);
        }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) cv[d] = ld32<SHINT ? 3 : 1>(vals + 2 * (k + d * stride));
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) wv[d] = ld32<GMODE>(w + 2 * (size_t)(c[d] & mask));
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) s ^= cv[d].v[0] ^ cv[d].v[7] ^ wv[d].v[0] ^ wv[d].v[7];
    }
    if (s == 0x12345678u) out[0] = s;
}

__global__ void fill_cols(uint32_t* cols, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) cols[i] = (uint32_t)(sm_mix(i) >> 20);
}

template <typename Fn> float time_ms(Fn fn, int reps = 4) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fn();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        fn();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    size_t gran = 0;
    if (argc > 1) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        if (e != cudaSuccess) { fprintf(stderr, "set granularity: %s\n", cudaGetErrorString(e)); (void)cudaGetLastError(); }
    }
    CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
    printf("{\"bench\":\"device\",\"name\":\"%s\",\"sms\":%d,\"l2_bytes\":%d,\"persisting_l2_max\":%d,\"l2_fetch_granularity\":%zu}\n", prop.name, sms,
           prop.l2CacheSize, prop.persistingL2CacheMaxSize, gran);
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, 64));
    const size_t wmax = 4ull << 30;
    char* d_w;
    CK(cudaMalloc(&d_w, wmax));
    CK(cudaMemset(d_w, 1, wmax));
    const uint4* w = (const uint4*)d_w;

#define RUN_GATHER(MODE, UNROLL, GRID, LABEL)                                                                                             \
    do {                                                                                                                                   \
        const int iters = 64 / UNROLL, threads = 256;                                                                                      \
        float ms = time_ms([&] { gather<MODE, UNROLL><<<GRID, threads>>>(w, mask, iters, d_out); });                                     \
        const double n = (double)(GRID) * threads * iters * UNROLL;                                                                        \
        printf("{\"bench\":\"gather\",\"load\":\"%s\",\"unroll\":%d,\"grid\":%d,\"witness_MiB\":%zu,\"ms\":%.4f,\"gathers_per_s\":%.4e}\n", LABEL, \
               UNROLL, (int)(GRID), wbytes >> 20, ms, n / (ms * 1e-3));                                                                    \
        fflush(stdout);                                                                                                                    \
    } while (0)

    for (size_t wbytes : {size_t(64) << 20, size_t(128) << 20, size_t(512) << 20, size_t(4) << 30}) {
        const uint32_t mask = (uint32_t)(wbytes / 32 - 1);
        RUN_GATHER(0, 1, sms * 8, "2xLDG128");
        RUN_GATHER(1, 1, sms * 8, "LDG256");
        RUN_GATHER(1, 4, sms * 8, "LDG256");
        RUN_GATHER(2, 1, sms * 8, "LDG256.evict_last");
        RUN_GATHER(1, 1, sms * 4, "LDG256");   // half the threads per SM
        RUN_GATHER(1, 2, sms * 1, "LDG256");   // 256 threads per SM
        RUN_GATHER(1, 4, (sms / 2) * 1, "LDG256");  // about half the SMs
        RUN_GATHER(1, 4, (sms / 2) * 8, "LDG256");
    }
    // persisting-L2 carve-out over a 512 MiB witness: a fixed fraction of the lines stays resident
    {
        const size_t wbytes = size_t(512) << 20;
        const uint32_t mask = (uint32_t)(wbytes / 32 - 1);
        cudaStream_t st;
        CK(cudaStreamCreate(&st));
        size_t carve = (size_t)prop.persistingL2CacheMaxSize;
        cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        if (e != cudaSuccess) { fprintf(stderr, "persisting limit: %s\n", cudaGetErrorString(e)); (void)cudaGetLastError(); }
        for (float ratio : {0.1f, 0.15f, 0.2f, 0.25f}) {
            cudaStreamAttrValue at = {};
            at.accessPolicyWindow.base_ptr = d_w;
            at.accessPolicyWindow.num_bytes = std::min<size_t>(wbytes, (size_t)prop.accessPolicyMaxWindowSize);
            at.accessPolicyWindow.hitRatio = ratio;
            at.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            at.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            e = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &at);
            if (e != cudaSuccess) { fprintf(stderr, "policy window: %s\n", cudaGetErrorString(e)); (void)cudaGetLastError(); break; }
            const int iters = 64, threads = 256, grid = sms * 8;
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float best = 1e30f;
            for (int r = 0; r < 5; ++r) {
                CK(cudaEventRecord(e0, st));
                gather<1, 1><<<grid, threads, 0, st>>>(w, mask, iters, d_out);
                CK(cudaEventRecord(e1, st));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r && ms < best) best = ms;
            }
            printf("{\"bench\":\"gather_persisting\",\"window_MiB\":%zu,\"carve_MiB\":%zu,\"hit_ratio\":%.2f,\"ms\":%.4f,\"gathers_per_s\":%.4e}\n",
                   at.accessPolicyWindow.num_bytes >> 20, carve >> 20, ratio, best, (double)grid * threads * iters / (best * 1e-3));
            fflush(stdout);
        }
        cudaStreamAttrValue at = {};
        at.accessPolicyWindow.num_bytes = 0;
        (void)cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &at);
        (void)cudaCtxResetPersistingL2Cache();
        (void)cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        CK(cudaStreamDestroy(st));
    }
    // coefficient stream + gathers together (config-3 proportions: 3.02e8 terms of 36 B, 512 MiB witness)
    {
        const size_t n_terms = 302008621;
        uint4* vals;
        uint32_t* cols;
        CK(cudaMalloc(&vals, n_terms * 32));
        CK(cudaMalloc(&cols, n_terms * 4));
        CK(cudaMemset(vals, 3, n_terms * 32));
        fill_cols<<<sms * 8, 256>>>(cols, n_terms);
        CK(cudaDeviceSynchronize());
#define RUN_SG(GMODE, SHINT, DEPTH, CTAS, LABEL)                                                                                                 \
    do {                                                                                                                                          \
        float ms = time_ms([&] { stream_gather<GMODE, SHINT, DEPTH><<<sms * CTAS, 256>>>(vals, cols, n_terms, w, mask, d_out); }, 3);             \
        printf("{\"bench\":\"stream_gather\",\"variant\":\"%s\",\"depth\":%d,\"ctas_per_sm\":%d,\"witness_MiB\":%zu,\"terms\":%zu,\"ms\":%.4f,"       \
               "\"terms_per_s\":%.4e,\"stream_GBps\":%.1f}\n", LABEL, DEPTH, CTAS, wbytes >> 20, n_terms, ms, n_terms / (ms * 1e-3),                 \
               n_terms * 36.0 / (ms * 1e-3) / 1e9);                                                                                               \
        fflush(stdout);                                                                                                                           \
    } while (0)
        for (size_t wbytes : {size_t(64) << 20, size_t(512) << 20, size_t(4) << 30}) {
            const uint32_t mask = (uint32_t)(wbytes / 32 - 1);
            RUN_SG(1, false, 1, 8, "plain");
            RUN_SG(1, false, 2, 4, "plain");
            RUN_SG(2, true, 1, 8, "stream evict_first, witness evict_last");
            RUN_SG(2, true, 2, 4, "stream evict_first, witness evict_last");
            RUN_SG(2, true, 4, 2, "stream evict_first, witness evict_last");
            RUN_SG(1, true, 2, 4, "stream evict_first");
        }
        CK(cudaFree(vals));
        CK(cudaFree(cols));
    }
    CK(cudaFree(d_w));
    return 0;
}
