// Host-side packing of a witness given as the reference holds it -- 32-byte canonical little-endian scalars, the Vec<Scalar>s of
// WitnessCS (witness_cs.rs:45-57) -- into what crosses PCIe: one bit per 0/1 value plus an exception list for everything else.
// Pure host code (g++; no CUDA): bp_cs_recheck_scalars[_async] and the exported bp_pack_scalars use it, and the CPU tests
// drive it without a device.
#pragma once

#include <cstdint>
#include <thread>
#include <vector>

namespace bp {

struct PackExc {  // an element whose value is neither 0 nor 1
    uint64_t idx;
    uint64_t v[4];
};

// Bytes [b0, b1) of the bit string of src[0 .. n): element i is bit (i & 7) of dst[i >> 3] when its value is 0 or 1; any other
// element leaves a 0 bit and is appended to `out` (ascending index).  Bits past element n - 1 in the last byte are 0.
// Dispatches once to an AVX2 kernel (8 elements = one output byte per step, software prefetch ahead of the stream) when the CPU
// has it; environment variable BP_PACK_SIMD=0 forces the portable loop.
void pack_bit_bytes(const uint64_t* src, uint64_t n, uint8_t* dst, uint64_t b0, uint64_t b1, std::vector<PackExc>& out);

// The same for scalars as they sit IN MEMORY in the reference's Vec<Scalar> (witness_cs.rs:45-57): blstrs::Scalar and
// pasta_curves::{Fp, Fq} are 4 x u64 little-endian limbs of the MONTGOMERY form x * 2^256 mod p, so the bit 1 is the limb
// pattern `one` = 2^256 mod p (mont_one) and 0 is all-zero.  An element that is neither goes to `out` with its limbs as they
// are (still Montgomery: from_mont converts the few there are).  No pass over the witness to call to_repr() is needed.
void pack_bit_bytes_mont(const uint64_t* src, uint64_t n, const uint64_t one[4], uint8_t* dst, uint64_t b0, uint64_t b1,
                         std::vector<PackExc>& out);

// Montgomery constants / conversion on the host (field ids of include/bp_r1cs.h): one = 2^256 mod p; from_mont: x * 2^-256
// mod p, the canonical residue (false when x >= p: not a value a Scalar can hold).
void mont_one(int field, uint64_t one[4]);
bool from_mont(int field, const uint64_t x[4], uint64_t out[4]);

// work(0 .. nt - 1), each on its own thread (work(0) on the caller's).  A thread that cannot be created is not an error: its
// share runs on the caller's thread instead.  `work` must not throw.
template <class W> void run_on_threads(unsigned nt, W&& work) {
    std::vector<std::thread> th;
    th.reserve(nt);
    unsigned started = 1;
    try {
        for (; started < nt; ++started) th.emplace_back(work, started);
    } catch (...) {  // std::system_error: no more threads
    }
    work(0u);
    for (unsigned t = started; t < nt; ++t) work(t);
    for (auto& x : th) x.join();
}

// Threads of one packing pass: BP_PACK_THREADS, else every hardware thread; 1 .. 64.
unsigned pack_threads();

// "avx2" or "portable": what pack_bit_bytes runs on this CPU (bp_pack_kernel reports it).
const char* pack_kernel_name();

}  // namespace bp
