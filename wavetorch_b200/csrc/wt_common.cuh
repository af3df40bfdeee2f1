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

// P[cell] += v through a switch over the 4R registers: ptxas turns it into a short compare tree plus jump tables, ~10
// instructions on the owning warp's path where the row-wise select above costs a branch region per row.  Used by the
// on-chip kernels, whose step is paced by the slowest warp of the CTA.
template <int R>
__device__ __forceinline__ void patch_add_switch(float (&P)[R][4], int cell, float v) {
#define WT_CASE(i) case i: if (i < 4 * R) P[(i) / 4 < R ? (i) / 4 : 0][(i) % 4] += v; break;
  switch (cell) {
    WT_CASE(0) WT_CASE(1) WT_CASE(2) WT_CASE(3) WT_CASE(4) WT_CASE(5) WT_CASE(6) WT_CASE(7)
    WT_CASE(8) WT_CASE(9) WT_CASE(10) WT_CASE(11) WT_CASE(12) WT_CASE(13) WT_CASE(14) WT_CASE(15)
    WT_CASE(16) WT_CASE(17) WT_CASE(18) WT_CASE(19) WT_CASE(20) WT_CASE(21) WT_CASE(22) WT_CASE(23)
    WT_CASE(24) WT_CASE(25) WT_CASE(26) WT_CASE(27) WT_CASE(28) WT_CASE(29) WT_CASE(30) WT_CASE(31)
    default: break;
  }
#undef WT_CASE
}
// P += x at the cells listed in m1, once more at those also in m2 (source.py:19-22 adds x once per listing, in sequence)
template <int R>
__device__ __forceinline__ void patch_inject_sw(float (&P)[R][4], unsigned m1, unsigned m2, float x) {
  for (unsigned mm = m1; mm; mm &= mm - 1u) {
    const int bit = __ffs(mm) - 1;
    for (int n = (m2 >> bit & 1u) ? 2 : 1; n; --n) patch_add_switch<R>(P, bit, x);
  }
}

}  // namespace wt
