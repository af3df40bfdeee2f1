// Host entry points of the temporally blocked forward (wt_tile.cu).
#pragma once
#include "wt_common.cuh"

namespace wt {

bool tile_eligible(const wt_problem* p);
size_t tile_extra_ws_bytes(const wt_problem* p);
int tile_forward(const wt_problem* p, const float* a1, const float* a3, const float* x, const int32_t* src_ij,
                 const int32_t* prb_ij, const int32_t* prb_sq, float* u1, float* u2, float* probe_out, float* probe_raw,
                 float* tape, float* extra_ws, cudaStream_t st, int* launches, const wt_slab* slab = nullptr);

size_t tile_extra_ws_bwd_bytes(const wt_problem* p);
int tile_launches_fwd(const wt_problem* p);
int tile_launches_bwd(const wt_problem* p);
int tile_backward(const wt_problem* p, const float* a1, const float* a3, const float* c, const int32_t* src_ij,
                  const int32_t* prb_ij, const int32_t* prb_sq, const float* grad_probe, const float* probe_raw,
                  const float* tape, float* state1, float* state2, float* spare1, float* spare2, float* G, float* grad_c,
                  float* grad_x, bool chained, cudaStream_t st, const wt_slab* slab = nullptr);

}  // namespace wt
