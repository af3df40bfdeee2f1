// Halo exchange of the row-slab domain decomposition (wt_slab.cu).
#pragma once
#include "wt_common.cuh"

namespace wt {

// Validates the descriptor for a [B,Nx,Ny] slab (host side, launches nothing).
int slab_check(const wt_slab* s, int B, int Nx, int Ny);

// One exchange of both fields (no-op when s == nullptr or the slab has no neighbour).
int slab_exchange(const wt_slab* s, int B, int Nx, int Ny, float* f1, float* f2, cudaStream_t st);

}  // namespace wt
