// libbp_frontend.so -- the C++ gadget front-end (cs.hpp / lc.hpp / gadgets.hpp behind include/bp_fixtures.h) on its own, for
// HOST RECORDING only: it synthesizes a circuit into host CSR arrays + witness (bp_tcs_new with device < 0) and never touches
// CUDA.  Built with g++, no nvcc, no libcudart.  Users: the CPU legs that only need a circuit's structure -- bench.py's
// `--impl reference` arm and cpu_baseline sample, the CPU-side tests -- so that those processes never map libbp_r1cs.so.
//
// The front-end's device sink calls the C ABI; here those entry points are local stubs that report "no device" (a device
// handle cannot be created through this library: bp_tcs_new(device >= 0) fails with BP_E_CUDA).  Only bp_tcs_* / bp_wcs_*
// are exported (frontend.map).
#include "../../../include/bp_r1cs.h"

extern "C" {
int bp_cs_new(int, int, uint64_t, uint64_t, uint64_t, bp_cs** out) {
    if (out) *out = nullptr;
    return BP_E_CUDA;
}
void bp_cs_free(bp_cs*) {}
const char* bp_cs_last_error(const bp_cs*) { return "libbp_frontend.so records on the host only; link libbp_r1cs.so for a device"; }
int bp_cs_alloc(bp_cs*, int, const uint64_t*, uint64_t, uint64_t*) { return BP_E_CUDA; }
int bp_cs_alloc_u8(bp_cs*, int, const uint8_t*, uint64_t, uint64_t*) { return BP_E_CUDA; }
int bp_cs_set(bp_cs*, int, uint64_t, const uint64_t*) { return BP_E_CUDA; }
int bp_cs_get(bp_cs*, int, uint64_t, uint64_t*) { return BP_E_CUDA; }
int bp_cs_set_range(bp_cs*, int, uint64_t, uint64_t, const uint64_t*) { return BP_E_CUDA; }
int bp_cs_witness(bp_cs*, int, uint64_t, uint64_t, uint64_t*) { return BP_E_CUDA; }
int bp_cs_enforce(bp_cs*, uint64_t, const uint32_t*, const uint32_t*, const uint64_t*) { return BP_E_CUDA; }
int bp_cs_first_unsatisfied(bp_cs*, int64_t*) { return BP_E_CUDA; }
}

#include "fixtures.cpp"
