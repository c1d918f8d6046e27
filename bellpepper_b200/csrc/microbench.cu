// Micro-benchmarks that size the R1CS kernels' two ceilings on this B200: the integer (fma-pipe) rate of the
// wide multiply-accumulate and the HBM rates of the access patterns the kernel uses (streaming 128-bit loads,
// bulk async copies, random 32-byte gathers).  Prints one JSON object per line.  Not part of the product path.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "field.cuh"
#include "synth.cuh"

using namespace bp;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// ---- raw instruction rates ---------------------------------------------------------------------------------
template <int MODE> __global__ void imad_rate(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // 16 independent IMAD.LO
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] * a + b;
        } else if (MODE == 1) {  // 8 independent IMAD.WIDE (64-bit accumulators)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint64_t v = (uint64_t)a * (b + i) + (((uint64_t)x[2 * i + 1] << 32) | x[2 * i]);
                x[2 * i] = (uint32_t)v; x[2 * i + 1] = (uint32_t)(v >> 32);
            }
        } else if (MODE == 2) {  // two 4-lane carry chains (8 IMAD.WIDE.X) + carry words
            x[8] += mad4_0(x, a, b, a ^ 5, b ^ 9, a + it);
            x[8] += mad4_0(x, b, a, b ^ 3, a ^ 7, b + it);
        } else if (MODE == 3) {  // 16 independent IADD3
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = x[i] + a + (b ^ i);
        } else if (MODE == 4) {  // 8 IMAD.WIDE + 8 IADD3 interleaved (dual-pipe issue)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint64_t v = (uint64_t)a * (b + i) + (((uint64_t)x[2 * i + 1] << 32) | x[2 * i]);
                x[2 * i] = (uint32_t)v; x[2 * i + 1] = (uint32_t)(v >> 32);
            }
#pragma unroll
            for (int i = 8; i < 16; ++i) x[i] = x[i] + a + (b ^ i);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint64_t v = (uint64_t)b * (a + i) + (((uint64_t)x[2 * i + 1] << 32) | x[2 * i]);
                x[2 * i] = (uint32_t)v; x[2 * i + 1] = (uint32_t)(v >> 32);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}

// ---- field-level rates: mac_wide per term, redc_acc per LC ---------------------------------------------------
template <int F, int MODE> __global__ void field_rate(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8], acc[17];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed ^ (0x9e3779b9u * (i + 1)); }
    a[7] &= 0x3fffffffu; b[7] &= 0x3fffffffu;
#pragma unroll
    for (int i = 0; i < 17; ++i) acc[i] = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            mac_wide(acc, a, b);
            a[0] += acc[3];
        } else if (MODE == 1) {
            uint32_t u[8];
            acc[16] &= 0xffu;
            redc_acc<F>(u, acc);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] ^= u[i];
            acc[9] += it;
        } else {
            uint32_t r[8];
            mont_mul<F>(r, a, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = r[i];
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 17; ++i) s ^= acc[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;
}

// ---- memory patterns -------------------------------------------------------------------------------------------
__global__ void stream_ldg(const uint4* __restrict__ p, size_t n16, uint32_t* out) {
    uint32_t s = 0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 v0 = __ldg(p + i), v1 = __ldg(p + i + stride), v2 = __ldg(p + i + 2 * stride), v3 = __ldg(p + i + 3 * stride);
        s ^= v0.x ^ v1.y ^ v2.z ^ v3.w;
    }
    if (s == 0x12345678u) out[0] = s;
}

// one warp-elected thread streams `bytes_per_copy` chunks with cp.async.bulk into a smem ring (depth STAGES)
template <int STAGES> __global__ void stream_bulk(const char* __restrict__ p, size_t total, uint32_t chunk, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    const size_t n_chunks = total / chunk;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    uint32_t acc = 0;
    size_t c = blockIdx.x;
    // prologue
    if (threadIdx.x == 0) {
        size_t cc = c;
        for (int s = 0; s < STAGES && cc < n_chunks; ++s, cc += gridDim.x) {
            unsigned ba = (unsigned)__cvta_generic_to_shared(&bar[s]);
            unsigned da = (unsigned)__cvta_generic_to_shared(smem + (size_t)s * chunk);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(chunk));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(da),
                         "l"(p + cc * chunk), "r"(chunk), "r"(ba)
                         : "memory");
        }
    }
    int stage = 0;
    unsigned phase = 0;
    for (; c < n_chunks; c += gridDim.x) {
        unsigned ba = (unsigned)__cvta_generic_to_shared(&bar[stage]);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(ba), "r"(phase)
                         : "memory");
        }
        // touch the data lightly (one word per thread) so the copy cannot be elided
        acc ^= ((const uint32_t*)(smem + (size_t)stage * chunk))[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            size_t nc = c + (size_t)STAGES * gridDim.x;
            if (nc < n_chunks) {
                unsigned da = (unsigned)__cvta_generic_to_shared(smem + (size_t)stage * chunk);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(chunk));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(da),
                             "l"(p + nc * chunk), "r"(chunk), "r"(ba)
                             : "memory");
            }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// random 32-byte gathers: each thread issues UNROLL independent gathers per iteration
template <int UNROLL> __global__ void gather32(const uint4* __restrict__ w, uint32_t n_elems_mask, int iters, uint32_t* out) {
    uint64_t h = sm_mix(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    uint32_t s = 0;
    for (int it = 0; it < iters; ++it) {
        uint4 v[UNROLL][2];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            h = h * 6364136223846793005ULL + 1442695040888963407ULL;
            const uint32_t idx = (uint32_t)(h >> 33) & n_elems_mask;
            v[u][0] = __ldg(w + 2 * (size_t)idx);
            v[u][1] = __ldg(w + 2 * (size_t)idx + 1);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) s ^= v[u][0].x ^ v[u][1].w;
    }
    if (s == 0x12345678u) out[0] = s;
}

template <typename Fn> float time_ms(Fn fn, int reps = 5) {
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

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\"bench\":\"device\",\"name\":\"%s\",\"sms\":%d,\"clock_mhz\":%d}\n", prop.name, sms, clk_khz / 1000);
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, 64));

    // instruction rates: 1024 threads/SM (8 warps per SMSP)
    {
        const int blocks = sms * 4, threads = 256, iters = 4096;
        const double lanes = (double)blocks * threads * iters;
        struct { const char* name; int mode; double ops; } m[] = {
            {"imad_lo", 0, 16}, {"imad_wide", 1, 8}, {"imad_wide_chain", 2, 8}, {"iadd3", 3, 16}, {"imad_wide8+iadd3x8", 4, 16}};
        for (auto& e : m) {
            float ms = 0;
            switch (e.mode) {
                case 0: ms = time_ms([&] { imad_rate<0><<<blocks, threads>>>(d_out, iters, 7); }); break;
                case 1: ms = time_ms([&] { imad_rate<1><<<blocks, threads>>>(d_out, iters, 7); }); break;
                case 2: ms = time_ms([&] { imad_rate<2><<<blocks, threads>>>(d_out, iters, 7); }); break;
                case 3: ms = time_ms([&] { imad_rate<3><<<blocks, threads>>>(d_out, iters, 7); }); break;
                case 4: ms = time_ms([&] { imad_rate<4><<<blocks, threads>>>(d_out, iters, 7); }); break;
            }
            const double ops_per_s = lanes * e.ops / (ms * 1e-3);
            printf("{\"bench\":\"instr_rate\",\"op\":\"%s\",\"ms\":%.4f,\"lane_ops_per_s\":%.4e,\"per_sm_per_clk_at_%dMHz\":%.2f}\n", e.name, ms,
                   ops_per_s, clk_khz / 1000, ops_per_s / sms / (clk_khz * 1e3));
        }
    }
    // field rates
    for (int threads : {128, 256}) {
        const int blocks = sms * (1024 / threads), iters = 512;
        const double n = (double)blocks * threads * iters;
        float ms;
        ms = time_ms([&] { field_rate<0, 0><<<blocks, threads>>>(d_out, iters, 7); });
        printf("{\"bench\":\"field_rate\",\"op\":\"mac_wide\",\"threads\":%d,\"ms\":%.4f,\"per_s\":%.4e}\n", threads, ms, n / (ms * 1e-3));
        ms = time_ms([&] { field_rate<0, 1><<<blocks, threads>>>(d_out, iters, 7); });
        printf("{\"bench\":\"field_rate\",\"op\":\"redc_acc_bls\",\"threads\":%d,\"ms\":%.4f,\"per_s\":%.4e}\n", threads, ms, n / (ms * 1e-3));
        ms = time_ms([&] { field_rate<1, 1><<<blocks, threads>>>(d_out, iters, 7); });
        printf("{\"bench\":\"field_rate\",\"op\":\"redc_acc_pallas\",\"threads\":%d,\"ms\":%.4f,\"per_s\":%.4e}\n", threads, ms, n / (ms * 1e-3));
        ms = time_ms([&] { field_rate<0, 2><<<blocks, threads>>>(d_out, iters, 7); });
        printf("{\"bench\":\"field_rate\",\"op\":\"mont_mul_bls\",\"threads\":%d,\"ms\":%.4f,\"per_s\":%.4e}\n", threads, ms, n / (ms * 1e-3));
    }
    // memory: 4 GiB buffer
    {
        const size_t bytes = 4ull << 30;
        char* d;
        CK(cudaMalloc(&d, bytes));
        CK(cudaMemset(d, 1, bytes));
        for (int per_sm : {4, 8, 16}) {
            float ms = time_ms([&] { stream_ldg<<<sms * per_sm, 256>>>((const uint4*)d, bytes / 16, d_out); });
            printf("{\"bench\":\"stream_ldg128\",\"blocks_per_sm\":%d,\"ms\":%.4f,\"GBps\":%.1f}\n", per_sm, ms, bytes / (ms * 1e-3) / 1e9);
        }
        for (uint32_t chunk : {8192u, 16384u, 32768u}) {
            const int smem = 4 * chunk;
            CK(cudaFuncSetAttribute(stream_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            for (int per_sm : {1, 2}) {
                if ((size_t)per_sm * smem > 200 * 1024) continue;
                float ms = time_ms([&] { stream_bulk<4><<<sms * per_sm, 128, smem>>>(d, bytes, chunk, d_out); });
                printf("{\"bench\":\"stream_bulk\",\"chunk\":%u,\"stages\":4,\"ctas_per_sm\":%d,\"ms\":%.4f,\"GBps\":%.1f}\n", chunk, per_sm, ms,
                       bytes / (ms * 1e-3) / 1e9);
            }
        }
        // random 32B gathers over 32 MiB (L2 resident), 512 MiB (cfg 4 witness) and 4 GiB (cfg 5 witness)
        for (size_t wbytes : {size_t(32) << 20, size_t(512) << 20, size_t(4) << 30}) {
            const uint32_t mask = (uint32_t)(wbytes / 32 - 1);
            const int iters = 64;
            {
                const int blocks = sms * 8, threads = 256;
                float ms = time_ms([&] { gather32<1><<<blocks, threads>>>((const uint4*)d, mask, iters, d_out); });
                const double n = (double)blocks * threads * iters * 1;
                printf("{\"bench\":\"gather32\",\"unroll\":1,\"witness_MiB\":%zu,\"ms\":%.4f,\"gathers_per_s\":%.4e,\"GBps\":%.1f}\n", wbytes >> 20, ms,
                       n / (ms * 1e-3), n * 32 / (ms * 1e-3) / 1e9);
            }
            {
                const int blocks = sms * 8, threads = 256;
                float ms = time_ms([&] { gather32<4><<<blocks, threads>>>((const uint4*)d, mask, iters, d_out); });
                const double n = (double)blocks * threads * iters * 4;
                printf("{\"bench\":\"gather32\",\"unroll\":4,\"witness_MiB\":%zu,\"ms\":%.4f,\"gathers_per_s\":%.4e,\"GBps\":%.1f}\n", wbytes >> 20, ms,
                       n / (ms * 1e-3), n * 32 / (ms * 1e-3) / 1e9);
            }
        }
        CK(cudaFree(d));
    }
    return 0;
}
