// Shared declarations for the wavetorch_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/wavetorch_b200.h"

namespace wt {

void set_error(const char* fmt, ...);

#define WT_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      wt::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return WT_ECUDA;                                                                       \
    }                                                                                        \
  } while (0)

#define WT_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      wt::set_error(__VA_ARGS__);      \
      return WT_EINVAL;                \
    }                                  \
  } while (0)

#define WT_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != WT_OK) return _s; \
  } while (0)

// Scalars every kernel needs, derived once on the host in double precision.
struct Scalars {
  float dt;      // time step
  float kappa;   // dt^2 / h^2  (the reference multiplies c^2 * h^-2 * stencil, then divides by dt^-2 + b/dt)
  float b0;      // saturable damping strength (0 = off)
  float inv_uth; // 1 / uth
  float c_nl;    // Kerr coefficient (0 = off)
};

inline Scalars make_scalars(const wt_problem* p) {
  Scalars s;
  s.dt = (float)p->dt;
  s.kappa = (float)((p->dt * p->dt) / (p->h * p->h));
  s.b0 = (float)p->b0;
  s.inv_uth = (p->b0 > 0) ? (float)(1.0 / p->uth) : 0.f;
  s.c_nl = (float)p->c_nl;
  return s;
}

inline int nonlinear_mask(const wt_problem* p) { return (p->b0 > 0 ? 1 : 0) | (p->c_nl != 0 ? 2 : 0); }

// ---------------------------------------------------------------------------------------------
// The cell: per-point arithmetic shared by every kernel (device side).
//
// Reference: y = (dt^-2 + b/dt)^-1 * (2/dt^2*y1 - (dt^-2 - b/dt)*y2 + c^2*L_h(y1))   (cell.py:12-17)
// With beta = b*dt, q = 1/(1+beta), kappa = dt^2/h^2 and L the unscaled 5-point stencil this is
//     y = 2q*y1 + (1-2q)*y2 + q*kappa*c^2*L(y1)  =  y2 + 2q*(y1 - y2) + (q*kappa*c^2)*L(y1).
// The last form keeps the y1 and y2 weights summing to exactly one (as the reference's evaluation does
// in the undamped interior, see oracle/wave_oracle.py:time_step) which matters for float32 drift.
// ---------------------------------------------------------------------------------------------
struct CellCoef {
  float a1;  // 2q
  float a3;  // q*kappa*c^2
};

__device__ __forceinline__ float wt_update(float a1, float a3, float u1, float u2, float lap) {
  return fmaf(a3, lap, fmaf(a1, u1 - u2, u2));
}

// b(u), c(u) of WaveCell.forward (cell.py:94-102)
template <bool SAT, bool KERR>
__device__ __forceinline__ void wt_nl_bc(const Scalars& s, float bpml, float clin, float rho, float u1, float& b,
                                         float& c, float& d) {
  d = 1.f;
  b = bpml;
  c = clin;
  if (SAT) {
    float r = u1 * s.inv_uth;
    d = fmaf(r, r, 1.f);
    b = fmaf(rho * s.b0, __frcp_rn(d), bpml);   // correctly rounded reciprocal: a few instructions, no division routine
  }
  if (KERR) c = fmaf(rho * s.c_nl, u1 * u1, clin);
}

__device__ __forceinline__ CellCoef wt_coef(const Scalars& s, float b, float c) {
  float q = __frcp_rn(fmaf(b, s.dt, 1.f));
  CellCoef k;
  k.a1 = 2.f * q;
  k.a3 = q * s.kappa * c * c;
  return k;
}

// ---------------------------------------------------------------------------------------------
// Rare-cell updates of a thread's R x 4 register patch (sources, probe seeds).  They are executed by one or two threads per
// sample but sit inside the time-step body of every thread: selecting the ROW with R compares and touching its four
// registers (three of them with a zero addend) keeps that code to a few instructions; selecting the single register
// would take a 4R-way branch tree in every unrolled copy of the step and puts ~100 instructions on the critical
// path of the owning warp, which every other warp then waits for at the step's barrier.
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void patch_add_cell(float (&P)[R][4], int cell, float v) {   // P[cell] += v
  const int prow = cell >> 2, pcol = cell & 3;
  const float s0 = pcol == 0 ? v : 0.f, s1 = pcol == 1 ? v : 0.f, s2 = pcol == 2 ? v : 0.f, s3 = pcol == 3 ? v : 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (prow == r) { P[r][0] += s0; P[r][1] += s1; P[r][2] += s2; P[r][3] += s3; }
}
template <int R>
__device__ __forceinline__ void patch_fma_cell(float (&P)[R][4], const float (&K)[R][4], int cell, float v) {   // P[cell] += K[cell]*v
  const int prow = cell >> 2, pcol = cell & 3;
  const float s0 = pcol == 0 ? v : 0.f, s1 = pcol == 1 ? v : 0.f, s2 = pcol == 2 ? v : 0.f, s3 = pcol == 3 ? v : 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (prow == r) {
      P[r][0] = fmaf(K[r][0], s0, P[r][0]);
      P[r][1] = fmaf(K[r][1], s1, P[r][1]);
      P[r][2] = fmaf(K[r][2], s2, P[r][2]);
      P[r][3] = fmaf(K[r][3], s3, P[r][3]);
    }
}
// P += x at the cells listed in m1, once more at those in m2 and m3 (bit r*4+k = cell (r,k)); source.py:19-22
template <int R>
__device__ __forceinline__ void patch_inject(float (&P)[R][4], unsigned m1, unsigned m2, unsigned m3, float x) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const unsigned a = (m1 >> (4 * r)) & 0xFu;
    if (a) {
      const unsigned b = (m2 >> (4 * r)) & 0xFu, c = (m3 >> (4 * r)) & 0xFu;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        P[r][k] += (a >> k & 1u) ? x : 0.f;
        P[r][k] += (b >> k & 1u) ? x : 0.f;
        P[r][k] += (c >> k & 1u) ? x : 0.f;
      }
    }
  }
}

// P += x at the cells listed in m (bit r*4+k = cell (r,k)) as 4R predicated adds without a branch.  For the on-chip kernels,
// whose step is paced by the slowest warp of the CTA: the caller tests "does any lane of my warp own a source" (warp-uniform)
// and only that warp issues these ~2*4R instructions; the row-wise select of patch_inject costs it a divergent branch
// region per row, a switch over the cell index (tried) a compare tree with 13-cycle predicate-to-branch latencies.
// inline PTX: written as C++ the 4R bit tests are loop invariant and nvcc keeps 4R extracted bits in registers
template <int BIT>
__device__ __forceinline__ void add_if_bit(float& p, unsigned m, float x) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t;\n"
      "and.b32 t, %1, %2;\n"
      "setp.ne.b32 p, t, 0;\n"
      "@p add.f32 %0, %0, %3;\n"
      "}\n"
      : "+f"(p)
      : "r"(m), "n"(1u << BIT), "f"(x));
}
template <int R, int I = 0>
__device__ __forceinline__ void patch_inject_pred(float (&P)[R][4], unsigned m, float x) {
  if constexpr (I < 4 * R) {
    add_if_bit<I>(P[I / 4][I % 4], m, x);
    patch_inject_pred<R, I + 1>(P, m, x);
  }
}

// P[cell] += K[cell] * v as 4R compare-and-predicated-FMA pairs without a branch (cell < 0: nothing); see patch_inject_pred
template <int I>
__device__ __forceinline__ void fma_if_cell(float& p, float k, int cell, float v) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.s32 p, %1, %2;\n"
      "@p fma.rn.f32 %0, %3, %4, %0;\n"
      "}\n"
      : "+f"(p)
      : "r"(cell), "n"(I), "f"(k), "f"(v));
}
template <int R, int I = 0>
__device__ __forceinline__ void patch_fma_pred(float (&P)[R][4], const float (&K)[R][4], int cell, float v) {
  if constexpr (I < 4 * R) {
    fma_if_cell<I>(P[I / 4][I % 4], K[I / 4][I % 4], cell, v);
    patch_fma_pred<R, I + 1>(P, K, cell, v);
  }
}

}  // namespace wt
