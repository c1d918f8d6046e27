// How many DRAM bytes does one random 32-byte gather cost on a B200, and does the flavour of the load change it?
// (DESIGN.md "Full-width kernel: what bounds it".)  Same random-gather loop as microbench2 with eight load flavours; run it
// plainly for the rates and under `ncu --metrics dram__bytes_read.sum` for the bytes.  Not on the product path.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "synth.cuh"
using namespace bp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
struct U8 { uint32_t v[8]; };
#define LD(NAME, INSTR)                                                                                                           \
    __device__ __forceinline__ U8 NAME(const uint4* p) {                                                                          \
        U8 r;                                                                                                                     \
        asm volatile(INSTR " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                                    \
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) \
                     : "l"(p));                                                                                                   \
        return r;                                                                                                                 \
    }
LD(ld0, "ld.global.nc.v8.u32")
LD(ld1, "ld.global.v8.u32")
LD(ld2, "ld.global.cg.v8.u32")
LD(ld3, "ld.global.nc.L2::64B.v8.u32")
LD(ld4, "ld.global.L1::no_allocate.v8.u32")
LD(ld5, "ld.global.cv.v8.u32")
LD(ld6, "ld.global.nc.L2::128B.v8.u32")
LD(ld7, "ld.global.nc.L1::no_allocate.L2::64B.v8.u32")
static const char* kNames[8] = {"nc", "default(ca)", "cg", "nc.L2::64B", "L1::no_allocate", "cv", "nc.L2::128B", "nc.L1::no_allocate.L2::64B"};

template <int FL> __global__ void gatherx(const uint4* __restrict__ w, uint32_t mask, int iters, uint32_t* out) {
    uint64_t h = sm_mix(blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    uint32_t s = 0;
    for (int it = 0; it < iters; ++it) {
        h = h * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint4* p = w + 2 * (size_t)((uint32_t)(h >> 33) & mask);
        U8 v;
        if (FL == 0) v = ld0(p); else if (FL == 1) v = ld1(p); else if (FL == 2) v = ld2(p); else if (FL == 3) v = ld3(p);
        else if (FL == 4) v = ld4(p); else if (FL == 5) v = ld5(p); else if (FL == 6) v = ld6(p); else v = ld7(p);
        s ^= v.v[0] ^ v.v[7];
    }
    if (s == 0x12345678u) out[0] = s;
}

template <int FL> void run(const uint4* w, size_t wbytes, int sms, uint32_t* d_out) {
    const uint32_t mask = (uint32_t)(wbytes / 32 - 1);
    const int grid = sms * 8, threads = 256, iters = 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        gatherx<FL><<<grid, threads>>>(w, mask, iters, d_out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r && ms < best) best = ms;
    }
    printf("{\"bench\":\"gather_flavour\",\"load\":\"%s\",\"witness_MiB\":%zu,\"gathers\":%lld,\"ms\":%.4f,\"gathers_per_s\":%.4e}\n", kNames[FL],
           wbytes >> 20, (long long)grid * threads * iters, best, (double)grid * threads * iters / (best * 1e-3));
    fflush(stdout);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, 64));
    char* d_w;
    CK(cudaMalloc(&d_w, 4ull << 30));
    CK(cudaMemset(d_w, 1, 4ull << 30));
    for (size_t wbytes : {size_t(512) << 20, size_t(4) << 30}) {
        const uint4* w = (const uint4*)d_w;
        run<0>(w, wbytes, sms, d_out); run<1>(w, wbytes, sms, d_out); run<2>(w, wbytes, sms, d_out); run<3>(w, wbytes, sms, d_out);
        run<4>(w, wbytes, sms, d_out); run<5>(w, wbytes, sms, d_out); run<6>(w, wbytes, sms, d_out); run<7>(w, wbytes, sms, d_out);
    }
    return 0;
}
