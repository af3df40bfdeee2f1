// Geometry parameterisation on the device: rho -> blur^N -> tanh projection -> c, and its
// reverse.  Reference: wavetorch/geom.py:207-233 (_apply_blur, _apply_projection, c) and what autograd derives from it.
// Evaluated once per forward; in PyTorch it is ~50 small launches per training iteration, here it is N+1 (N blur
// passes fused with the projection on the last one).  Scalars (eta, beta, c0, c1) are read from the module's 0-dim
// device buffers so that no host synchronisation is needed (beta changes during optimize_lens.py's schedule).
#include "wt_common.cuh"

namespace wt {

constexpr int GEOM_MAX_TAPS = 121;   // blur radius <= 5

struct GeomArgs {
  int Nx, Ny, radius;
  const float* taps;     // [(2r+1)^2] normalised disk stencil (geom.py:149-152), device memory
  const float* eta;      // 0-dim device buffers
  const float* beta;
  const float* c0;
  const float* c1;
};

__device__ __forceinline__ float blur_at(const GeomArgs& g, const float* __restrict__ src, int i, int j) {
  const int r = g.radius, n = 2 * r + 1;
  float acc = 0.f;
  for (int di = -r; di <= r; ++di) {
    const int ii = i + di;
    if (ii < 0 || ii >= g.Nx) continue;
    for (int dj = -r; dj <= r; ++dj) {
      const int jj = j + dj;
      if (jj < 0 || jj >= g.Ny) continue;
      const float w = g.taps[(di + r) * n + (dj + r)];
      if (w != 0.f) acc = fmaf(w, src[(size_t)ii * g.Ny + jj], acc);
    }
  }
  return acc;
}

// one blur pass; when `project` is set also writes c = c0 + (c1-c0) * proj(blurred)
__global__ void k_geom_blur(GeomArgs g, const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ c_out,
                            int project) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.Nx || j >= g.Ny) return;
  const float v = blur_at(g, src, i, j);
  dst[(size_t)i * g.Ny + j] = v;
  if (project) {
    const float eta = *g.eta, beta = *g.beta, c0 = *g.c0, c1 = *g.c1;
    const float lo = tanhf(beta * eta);
    const float den = lo + tanhf(beta * (1.f - eta));
    c_out[(size_t)i * g.Ny + j] = c0 + (c1 - c0) * ((lo + tanhf(beta * (v - eta))) / den);
  }
}

// g_blurred = grad_c * (c1-c0) * beta * (1 - tanh^2(beta*(blurred-eta))) / den
__global__ void k_geom_proj_bwd(GeomArgs g, const float* __restrict__ blurred, const float* __restrict__ grad_c,
                                float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.Nx || j >= g.Ny) return;
  const size_t o = (size_t)i * g.Ny + j;
  const float eta = *g.eta, beta = *g.beta, c0 = *g.c0, c1 = *g.c1;
  const float den = tanhf(beta * eta) + tanhf(beta * (1.f - eta));
  const float th = tanhf(beta * (blurred[o] - eta));
  out[o] = grad_c[o] * (c1 - c0) * beta * (1.f - th * th) / den;
}

// adjoint of one zero-padded correlation pass = correlation with the flipped stencil
__global__ void k_geom_blur_bwd(GeomArgs g, const float* __restrict__ src, float* __restrict__ dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.Nx || j >= g.Ny) return;
  const int r = g.radius, n = 2 * r + 1;
  float acc = 0.f;
  for (int di = -r; di <= r; ++di) {
    const int ii = i - di;
    if (ii < 0 || ii >= g.Nx) continue;
    for (int dj = -r; dj <= r; ++dj) {
      const int jj = j - dj;
      if (jj < 0 || jj >= g.Ny) continue;
      const float w = g.taps[(di + r) * n + (dj + r)];
      if (w != 0.f) acc = fmaf(w, src[(size_t)ii * g.Ny + jj], acc);
    }
  }
  dst[(size_t)i * g.Ny + j] = acc;
}

}  // namespace wt

using namespace wt;

extern "C" {

// blurred: [passes, Nx, Ny] every intermediate blurred field (the last one feeds the projection and the backward).
int wt_geom_forward(int Nx, int Ny, int radius, int passes, const float* rho, const float* taps, const float* eta,
                    const float* beta, const float* c0, const float* c1, float* blurred, float* c_out, int device,
                    void* stream) {
  WT_REQUIRE(Nx > 0 && Ny > 0 && passes >= 1 && radius >= 0 && (2 * radius + 1) * (2 * radius + 1) <= GEOM_MAX_TAPS,
             "wt_geom_forward: bad shape (Nx=%d Ny=%d radius=%d passes=%d)", Nx, Ny, radius, passes);
  WT_REQUIRE(rho && taps && eta && beta && c0 && c1 && blurred && c_out, "wt_geom_forward: NULL argument");
  WT_CUDA(cudaSetDevice(device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GeomArgs g = {Nx, Ny, radius, taps, eta, beta, c0, c1};
  dim3 block(32, 8), grid((Ny + 31) / 32, (Nx + 7) / 8);
  const size_t plane = (size_t)Nx * Ny;
  const float* src = rho;
  for (int p = 0; p < passes; ++p) {
    float* dst = blurred + (size_t)p * plane;
    k_geom_blur<<<grid, block, 0, st>>>(g, src, dst, c_out, p == passes - 1);
    src = dst;
  }
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

// grad_rho = blur^T^passes( proj'(blurred_last) * (c1-c0) * grad_c ); scratch: [2, Nx, Ny]
int wt_geom_backward(int Nx, int Ny, int radius, int passes, const float* blurred_last, const float* grad_c,
                     const float* taps, const float* eta, const float* beta, const float* c0, const float* c1,
                     float* grad_rho, float* scratch, int device, void* stream) {
  WT_REQUIRE(Nx > 0 && Ny > 0 && passes >= 1 && radius >= 0 && (2 * radius + 1) * (2 * radius + 1) <= GEOM_MAX_TAPS,
             "wt_geom_backward: bad shape");
  WT_REQUIRE(blurred_last && grad_c && taps && eta && beta && c0 && c1 && grad_rho && scratch, "wt_geom_backward: NULL argument");
  WT_CUDA(cudaSetDevice(device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GeomArgs g = {Nx, Ny, radius, taps, eta, beta, c0, c1};
  dim3 block(32, 8), grid((Ny + 31) / 32, (Nx + 7) / 8);
  const size_t plane = (size_t)Nx * Ny;
  float* a = scratch;
  float* b = scratch + plane;
  k_geom_proj_bwd<<<grid, block, 0, st>>>(g, blurred_last, grad_c, a);
  for (int p = 0; p < passes; ++p) {
    float* dst = (p == passes - 1) ? grad_rho : b;
    k_geom_blur_bwd<<<grid, block, 0, st>>>(g, a, dst);
    float* t = a; a = b; b = t;
  }
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

}  // extern "C"
