// Host entry points of the on-chip path (wt_resident.cu).
#pragma once
#include "wt_common.cuh"

namespace wt {

// Fills `plan` and returns true when the problem can run on the on-chip path on this device.
bool resident_plan(const wt_problem* p, const cudaDeviceProp& prop, bool need_adjoint, wt_plan* plan);

int resident_forward(const wt_problem* p, const wt_plan& plan, const float* c, const float* b, const float* rho,
                     const float* x,
                     const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_sq, float* u1, float* u2,
                     float* probe_out, float* probe_raw, float* fields_out, void* history, void* workspace,
                     cudaStream_t st);

int resident_backward(const wt_problem* p, const wt_plan& plan, const float* c, const float* b, const float* rho,
                      const int32_t* src_ij,
                      const int32_t* prb_ij, const int32_t* prb_sq, const float* grad_probe, const float* probe_raw,
                      const void* history, float* grad_c, float* grad_b, float* grad_rho, float* grad_x,
                      void* workspace, cudaStream_t st);

}  // namespace wt
