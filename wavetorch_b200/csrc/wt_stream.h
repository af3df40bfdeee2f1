// Host entry points of the HBM-streaming path (wt_stream.cu).
#pragma once
#include "wt_common.cuh"

namespace wt {

size_t stream_tape_bytes(const wt_problem* p);
size_t stream_ws_fwd_bytes(const wt_problem* p);
size_t stream_ws_bwd_bytes(const wt_problem* p);

int stream_forward(const wt_problem* p, const float* c, const float* b, const float* rho, const float* x,
                   const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_sq, float* u1, float* u2,
                   float* probe_out, float* probe_raw, float* fields_out, void* history, void* workspace,
                   cudaStream_t st, const wt_slab* slab = nullptr);

int stream_backward(const wt_problem* p, const float* c, const float* b, const float* rho, const int32_t* src_ij,
                    const int32_t* prb_ij, const int32_t* prb_sq, const float* grad_probe, const float* probe_raw,
                    const float* grad_fields, const void* history, float* adj1, float* adj2, float* grad_c,
                    float* grad_b, float* grad_rho, float* grad_x, void* workspace, cudaStream_t st,
                    const wt_slab* slab = nullptr);

int step_forward(const wt_problem* p, const float* b, int bb, const float* c, int cb, const float* y1,
                 const float* y2, float* y, cudaStream_t st);

int step_backward(const wt_problem* p, const float* b, int bb, const float* c, int cb, const float* y1,
                  const float* y2, const float* g, float* gb, float* gc, float* gy1, float* gy2, cudaStream_t st);

// shared small kernels used by the resident path as well
__global__ void k_coeff(const float* __restrict__ b, const float* __restrict__ c, int n, double dt, double kappa,
                        float* __restrict__ a1, float* __restrict__ a3, float* __restrict__ gscale);
__global__ void k_finish_grad(const float* __restrict__ G, const float* __restrict__ gscale, int n_part, size_t stride,
                              size_t plane, float* __restrict__ grad_c);

}  // namespace wt
