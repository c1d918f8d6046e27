// TMA-staged K1 (placeholder until the pipeline kernel lands): reports "not eligible" so the direct kernel runs.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "kernels.cuh"

namespace bp {

struct StagedPlan {
    uint32_t tile_terms = 1024;
    uint32_t n_tiles = 0;
    uint32_t* d_tile_row = nullptr;  // n_tiles+1 first rows
    bool last_used = false;
};

inline void staged_plan_free(StagedPlan& p) {
    if (p.d_tile_row) cudaFree(p.d_tile_row);
    p.d_tile_row = nullptr;
}

// returns 0 = launched, 1 = allocation failure, 2 = not eligible (caller uses the direct kernel)
template <int F>
int staged_launch(StagedPlan& plan, bool& plan_valid, const CsrView& m, const CheckOut& o, int sm_count, cudaStream_t stream,
                  int64_t& launches) {
    (void)plan_valid; (void)m; (void)o; (void)sm_count; (void)stream; (void)launches;
    plan.last_used = false;
    return 2;
}

}  // namespace bp
