// HBM-streaming kernels: one launch per time step, fields in natural [B,Nx,Ny] layout.
//
// This is the path for grids that do not fit on-chip (BASELINE config 5) and the general fallback
// (nonlinear terms, grad w.r.t. damping, dLoss/dfields, adjoint-state chaining).  Each thread owns VEC
// consecutive cells of one row and loops over the samples of its batch chunk, so the coefficient fields
// are read once per chunk instead of once per sample.
//
// Reference semantics: wavetorch/rnn.py:50-67 (loop body), cell.py:12-17 / 27-44 / 94-102,
// operators.py:5-11, source.py:15-22, probe.py:14-27.
#include "wt_common.cuh"
#include "wt_slab.h"
#include "wt_stream.h"
#include "wt_tile.h"

namespace wt {

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld(const float* __restrict__ p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void st(float* __restrict__ p, const float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    *p = v[0];
  }
}
template <int VEC>
__device__ __forceinline__ void zero(float (&v)[VEC]) {
#pragma unroll
  for (int k = 0; k < VEC; ++k) v[k] = 0.f;
}

// 5-point neighbourhood of VEC cells at (i, j0..j0+VEC-1) of one [Nx,Ny] plane; zero outside the domain
// (operators.py:11, conv2d padding=1).
template <int VEC>
struct Hood {
  float ce[VEC], up[VEC], dn[VEC], lf, rt;
  __device__ __forceinline__ void load(const float* __restrict__ plane, int Nx, int Ny, int i, int j0) {
    const float* row = plane + (size_t)i * Ny + j0;
    ld<VEC>(row, ce);
    if (i > 0) ld<VEC>(row - Ny, up); else zero<VEC>(up);
    if (i + 1 < Nx) ld<VEC>(row + Ny, dn); else zero<VEC>(dn);
    lf = (j0 > 0) ? row[-1] : 0.f;
    rt = (j0 + VEC < Ny) ? row[VEC] : 0.f;
  }
  __device__ __forceinline__ float west(int k) const { return k == 0 ? lf : ce[k - 1]; }
  __device__ __forceinline__ float east(int k) const { return k == VEC - 1 ? rt : ce[k + 1]; }
  // unscaled Laplacian; the association order is the same in every kernel of this library
  __device__ __forceinline__ float lap(int k) const {
    return fmaf(-4.f, ce[k], (up[k] + dn[k]) + (west(k) + east(k)));
  }
};

// ------------------------------------------------------------------------------------------------
// coefficient fields for the linear case, evaluated in double and rounded once
// ------------------------------------------------------------------------------------------------
__global__ void k_coeff(const float* __restrict__ b, const float* __restrict__ c, int n, double dt, double kappa,
                        float* __restrict__ a1, float* __restrict__ a3, float* __restrict__ gscale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double q = 1.0 / (1.0 + (double)b[i] * dt);
  double cc = (double)c[i];
  a1[i] = (float)(2.0 * q);
  a3[i] = (float)(q * kappa * cc * cc);
  gscale[i] = (float)(2.0 * q * kappa * cc);  // d a3 / d c : grad_c = gscale * sum_t,b L(u_{t-1}) * lambda_t  (cell.py:36)
}

// ------------------------------------------------------------------------------------------------
// forward step
// ------------------------------------------------------------------------------------------------
struct FwdArgs {
  int Nx, Ny, B, bchunk;
  size_t plane;
  const float* u1;   // [B,plane] field at t-1
  float* u2;         // [B,plane] field at t-2, overwritten with the new field
  const float* a1;   // linear: precomputed coefficient fields
  const float* a3;
  const float* bpml; // nonlinear: raw fields
  const float* clin;
  const float* rho;
  float* tape_lap;   // nullable: [B,plane] slot for L(u1)
  float* tape_u1;    // nullable: [B,plane] slot receiving u1
  float* tape_u2;    // nullable: [B,plane] slot receiving u2 (first step only)
  float* fields;     // nullable: fields_out + t*plane, sample stride fields_bstride
  size_t fields_bstride;
  Scalars s;
};

template <int VEC, bool LINEAR, bool SAT, bool KERR>
__global__ void __launch_bounds__(128) k_stream_fwd(FwdArgs a) {
  const int j0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int i = blockIdx.y * 4 + threadIdx.y;
  if (j0 >= a.Ny || i >= a.Nx) return;
  const size_t cell = (size_t)i * a.Ny + j0;
  float k1[VEC], k3[VEC], bp[VEC], cl[VEC], rh[VEC];
  if constexpr (LINEAR) {
    ld<VEC>(a.a1 + cell, k1);
    ld<VEC>(a.a3 + cell, k3);
  } else {
    ld<VEC>(a.bpml + cell, bp);
    ld<VEC>(a.clin + cell, cl);
    if (SAT || KERR) ld<VEC>(a.rho + cell, rh); else zero<VEC>(rh);
  }
  const int b0 = blockIdx.z * a.bchunk;
  const int b1 = min(a.B, b0 + a.bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * a.plane;
    Hood<VEC> h;
    h.load(a.u1 + off, a.Nx, a.Ny, i, j0);
    float w[VEC], y[VEC], l[VEC];
    ld<VEC>(a.u2 + off + cell, w);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      l[k] = h.lap(k);
      if constexpr (LINEAR) {
        y[k] = wt_update(k1[k], k3[k], h.ce[k], w[k], l[k]);
      } else {
        float bb, cc, d;
        wt_nl_bc<SAT, KERR>(a.s, bp[k], cl[k], rh[k], h.ce[k], bb, cc, d);
        CellCoef kc = wt_coef(a.s, bb, cc);
        y[k] = wt_update(kc.a1, kc.a3, h.ce[k], w[k], l[k]);
      }
    }
    if (a.tape_u2) st<VEC>(a.tape_u2 + off + cell, w);
    st<VEC>(a.u2 + off + cell, y);
    if (a.tape_lap) st<VEC>(a.tape_lap + off + cell, l);
    if (a.tape_u1) st<VEC>(a.tape_u1 + off + cell, h.ce);
    if (a.fields) st<VEC>(a.fields + (size_t)b * a.fields_bstride + cell, y);
  }
}

// Source injection (source.py:19-22, called with its dt=1.0 default from rnn.py:57) and probe readout
// (probe.py:15,27) for one time step.  One block per sample.
__global__ void k_src_prb(float* __restrict__ U, size_t plane, float* fields, size_t fields_bstride,
                          const float* __restrict__ x, int t, int T, const int32_t* __restrict__ src_off, int n_src,
                          const int32_t* __restrict__ prb_off, const int32_t* __restrict__ prb_sq, int n_prb,
                          float* __restrict__ probe_out, float* __restrict__ probe_raw) {
  const int b = blockIdx.x;
  float* u = U + (size_t)b * plane;
  const float xv = x[(size_t)b * T + t];
  // all addends are equal, so the result does not depend on the order of the atomics
  for (int s = threadIdx.x; s < n_src; s += blockDim.x) atomicAdd(u + src_off[s], xv);
  __syncthreads();
  if (fields) {
    float* f = fields + (size_t)b * fields_bstride;
    for (int s = threadIdx.x; s < n_src; s += blockDim.x) f[src_off[s]] = u[src_off[s]];
  }
  for (int p = threadIdx.x; p < n_prb; p += blockDim.x) {
    float v = u[prb_off[p]];
    size_t o = ((size_t)b * T + t) * n_prb + p;
    if (probe_raw) probe_raw[o] = v;
    if (probe_out) probe_out[o] = prb_sq[p] ? v * v : v;
  }
}

__global__ void k_swap(float* __restrict__ a, float* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float t = a[i]; a[i] = b[i]; b[i] = t;
  }
}

__global__ void k_off(const int32_t* __restrict__ ij, int n, int Nx, int Ny, int32_t* __restrict__ off, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int r = ij[2 * i], c = ij[2 * i + 1];
  if (r < 0 || r >= Nx || c < 0 || c >= Ny) { atomicExch(bad, 1); r = 0; c = 0; }
  off[i] = r * Ny + c;
}

// ------------------------------------------------------------------------------------------------
// adjoint, linear:  lambda_{t-1} = a1*lambda_t + L(a3*lambda_t) + (1-a1)*lambda_{t+1}   (cell.py:39-42)
// ------------------------------------------------------------------------------------------------
struct AdjLinArgs {
  int Nx, Ny, B, bchunk;
  size_t plane;
  const float* lam1;  // [B,plane] lambda_t (complete, seeded)
  float* lam2;        // [B,plane] in: lambda_{t+1} (or an already weighted carry when `premul`); out: lambda_{t-1} w/o seed
  const float* a1;
  const float* a3;
  const float* tape_lap;  // [B,plane] L(u_{t-1}) recorded by the forward step t
  const float* gfields;   // nullable: dLoss/dfields[:, t-1] (sample stride gf_bstride)
  size_t gf_bstride;
  float* G;           // [plane] accumulates sum_b L(u_{t-1})*lambda_t
  int premul;
  int atomic_G;
};

template <int VEC>
__global__ void __launch_bounds__(128) k_stream_adj_lin(AdjLinArgs a) {
  const int j0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  const int i = blockIdx.y * 4 + threadIdx.y;
  if (j0 >= a.Ny || i >= a.Nx) return;
  const size_t cell = (size_t)i * a.Ny + j0;
  float k1[VEC];
  ld<VEC>(a.a1 + cell, k1);
  Hood<VEC> k3;
  k3.load(a.a3, a.Nx, a.Ny, i, j0);
  float g[VEC];
  zero<VEC>(g);
  const int b0 = blockIdx.z * a.bchunk;
  const int b1 = min(a.B, b0 + a.bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * a.plane;
    Hood<VEC> h;
    h.load(a.lam1 + off, a.Nx, a.Ny, i, j0);
    float w[VEC], l[VEC], y[VEC], gf[VEC];
    ld<VEC>(a.lam2 + off + cell, w);
    ld<VEC>(a.tape_lap + off + cell, l);
    if (a.gfields) ld<VEC>(a.gfields + (size_t)b * a.gf_bstride + cell, gf); else zero<VEC>(gf);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float pw = k3.west(k) * h.west(k), pe = k3.east(k) * h.east(k);
      float lapP = fmaf(-4.f, k3.ce[k] * h.ce[k], (k3.up[k] * h.up[k] + k3.dn[k] * h.dn[k]) + (pw + pe));
      float carry = a.premul ? w[k] : (1.f - k1[k]) * w[k];
      y[k] = fmaf(k1[k], h.ce[k], lapP) + carry + gf[k];
      g[k] = fmaf(l[k], h.ce[k], g[k]);
    }
    st<VEC>(a.lam2 + off + cell, y);
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    if (a.atomic_G) atomicAdd(a.G + cell + k, g[k]);
    else a.G[cell + k] += g[k];
  }
}

// Adds the probe seeds of step t to lambda_t and gathers dLoss/dx[:, t].  One block (128 threads) per sample.
__global__ void k_adj_seed(float* __restrict__ lam, size_t plane, const float* __restrict__ grad_probe,
                           const float* __restrict__ probe_raw, int t, int T, const int32_t* __restrict__ prb_off,
                           const int32_t* __restrict__ prb_sq, int n_prb, const int32_t* __restrict__ src_off,
                           int n_src, float* __restrict__ grad_x) {
  const int b = blockIdx.x;
  float* u = lam + (size_t)b * plane;
  for (int p = threadIdx.x; p < n_prb; p += blockDim.x) {
    size_t o = ((size_t)b * T + t) * n_prb + p;
    float g = grad_probe[o];
    if (prb_sq[p]) g *= 2.f * probe_raw[o];   // d(u^2) = 2u  (probe.py:27)
    atomicAdd(u + prb_off[p], g);
  }
  if (!grad_x) return;
  __syncthreads();
  __shared__ float red[128];
  float s = 0.f;
  for (int k = threadIdx.x; k < n_src; k += blockDim.x) s += u[src_off[k]];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) grad_x[(size_t)b * T + t] = red[0];
}

__global__ void k_add_fields(float* __restrict__ lam, size_t plane, const float* __restrict__ gf, size_t gf_bstride) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < plane) lam[(size_t)blockIdx.y * plane + i] += gf[(size_t)blockIdx.y * gf_bstride + i];
}

__global__ void k_scale_carry(float* __restrict__ lam, const float* __restrict__ a1, size_t plane, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    lam[i] = (1.f - a1[i % plane]) * lam[i];
}

__global__ void k_finish_grad(const float* __restrict__ G, const float* __restrict__ gscale, int n_part, size_t stride,
                              size_t plane, float* __restrict__ grad_c) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= plane) return;
  float s = 0.f;
  for (int k = 0; k < n_part; ++k) s += G[(size_t)k * stride + i];
  grad_c[i] = gscale ? gscale[i] * s : s;
}

// ------------------------------------------------------------------------------------------------
// adjoint, general (nonlinear b(u), c(u); grad w.r.t. damping): two passes per step
//   pass A (own cell): coefficients of step t from u_{t-1}; gradient accumulation; P = kappa*c^2*q*lambda
//   pass B (stencil) : carry1' += L(P)
// SURVEY appendix A.2/A.3; reference: cell.py:27-44 + autograd through cell.py:94-100.
// ------------------------------------------------------------------------------------------------
struct AdjNlArgs {
  int Nx, Ny, B, bchunk;
  size_t plane;
  float* c1;          // [B,plane] in: lambda_t (seeded); out: carry2' = (beta-1)*q*lambda_t
  float* c2;          // [B,plane] in: carry2; out: carry1' without the stencil term
  float* P;           // [B,plane] out
  const float* tu1;   // [B,plane] u_{t-1}
  const float* tu2;   // [B,plane] u_{t-2}
  const float* bpml;
  const float* clin;
  const float* rho;
  float* Gc;          // [plane]
  float* Gb;          // [plane] nullable
  float* Grho;        // [plane] nullable
  int atomic_G;
  Scalars s;
};

template <bool SAT, bool KERR>
__global__ void __launch_bounds__(128) k_stream_adjA(AdjNlArgs a) {
  const int j0 = blockIdx.x * 32 + threadIdx.x;
  const int i = blockIdx.y * 4 + threadIdx.y;
  if (j0 >= a.Ny || i >= a.Nx) return;
  const size_t cell = (size_t)i * a.Ny + j0;
  const float bp = a.bpml[cell], cl = a.clin[cell];
  const float rh = (SAT || KERR) ? a.rho[cell] : 0.f;
  float gc = 0.f, gb = 0.f, gr = 0.f;
  const int b0 = blockIdx.z * a.bchunk;
  const int b1 = min(a.B, b0 + a.bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * a.plane;
    Hood<1> h;
    h.load(a.tu1 + off, a.Nx, a.Ny, i, j0);
    const float u1 = h.ce[0], u2 = a.tu2[off + cell], lap = h.lap(0);
    const float lam = a.c1[off + cell], car2 = a.c2[off + cell];
    float bb, cc, d;
    wt_nl_bc<SAT, KERR>(a.s, bp, cl, rh, u1, bb, cc, d);
    const float beta = bb * a.s.dt;
    const float q = __frcp_rn(1.f + beta);
    const float ql = q * lam;
    const float kl = a.s.kappa * lap;
    const float S = fmaf(cc * cc, kl, 2.f * (u1 - u2));
    const float g_b = -a.s.dt * q * S * ql;          // cell.py:33-34
    const float g_c = 2.f * cc * kl * ql;             // cell.py:36
    float gu1 = 2.f * ql;                             // own-cell part of cell.py:39-40
    if (SAT) {
      const float iu = a.s.inv_uth;
      const float rd = __frcp_rn(d);
      gr = fmaf(g_b, a.s.b0 * rd, gr);
      gu1 = fmaf(g_b, rh * a.s.b0 * (-2.f * u1 * iu * iu) * (rd * rd), gu1);
    }
    if (KERR) {
      gr = fmaf(g_c, a.s.c_nl * u1 * u1, gr);
      gu1 = fmaf(g_c, 2.f * rh * a.s.c_nl * u1, gu1);
    }
    gc += g_c;
    gb += g_b;
    a.P[off + cell] = a.s.kappa * cc * cc * ql;
    a.c2[off + cell] = car2 + gu1;
    a.c1[off + cell] = (beta - 1.f) * ql;             // cell.py:42
  }
  if (a.atomic_G) {
    atomicAdd(a.Gc + cell, gc);
    if (a.Gb) atomicAdd(a.Gb + cell, gb);
    if (a.Grho) atomicAdd(a.Grho + cell, gr);
  } else {
    a.Gc[cell] += gc;
    if (a.Gb) a.Gb[cell] += gb;
    if (a.Grho) a.Grho[cell] += gr;
  }
}

__global__ void __launch_bounds__(128) k_stream_adjB(int Nx, int Ny, int B, int bchunk, size_t plane,
                                                     const float* __restrict__ P, float* __restrict__ c2,
                                                     const float* __restrict__ gfields, size_t gf_bstride) {
  const int j0 = blockIdx.x * 32 + threadIdx.x;
  const int i = blockIdx.y * 4 + threadIdx.y;
  if (j0 >= Ny || i >= Nx) return;
  const size_t cell = (size_t)i * Ny + j0;
  const int b0 = blockIdx.z * bchunk;
  const int b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * plane;
    Hood<1> h;
    h.load(P + off, Nx, Ny, i, j0);
    float v = c2[off + cell] + h.lap(0);
    if (gfields) v += gfields[(size_t)b * gf_bstride + cell];
    c2[off + cell] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// single step (TimeStep.forward / TimeStep.backward), general per-sample or shared b and c
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_step_fwd(int Nx, int Ny, int B, size_t plane, const float* __restrict__ b,
                                                  size_t bs, const float* __restrict__ c, size_t cs,
                                                  const float* __restrict__ y1, const float* __restrict__ y2,
                                                  float* __restrict__ y, Scalars s) {
  const int j0 = blockIdx.x * 32 + threadIdx.x;
  const int i = blockIdx.y * 4 + threadIdx.y;
  const int n = blockIdx.z;
  if (j0 >= Ny || i >= Nx) return;
  const size_t cell = (size_t)i * Ny + j0, off = (size_t)n * plane;
  Hood<1> h;
  h.load(y1 + off, Nx, Ny, i, j0);
  CellCoef k = wt_coef(s, b[n * bs + cell], c[n * cs + cell]);
  y[off + cell] = wt_update(k.a1, k.a3, h.ce[0], y2[off + cell], h.lap(0));
}

__global__ void __launch_bounds__(128) k_step_bwd(int Nx, int Ny, int B, size_t plane, const float* __restrict__ b,
                                                  size_t bs, const float* __restrict__ c, size_t cs,
                                                  const float* __restrict__ y1, const float* __restrict__ y2,
                                                  const float* __restrict__ g, float* __restrict__ gb,
                                                  float* __restrict__ gc, float* __restrict__ gy1,
                                                  float* __restrict__ gy2, Scalars s) {
  const int j0 = blockIdx.x * 32 + threadIdx.x;
  const int i = blockIdx.y * 4 + threadIdx.y;
  const int n = blockIdx.z;
  if (j0 >= Ny || i >= Nx) return;
  const size_t cell = (size_t)i * Ny + j0, off = (size_t)n * plane;
  const float* bn = b + n * bs;
  const float* cn = c + n * cs;
  const float* gn = g + off;
  auto pval = [&](int ii, int jj) -> float {   // kappa*c^2*q*g at a neighbour, zero outside (cell.py:39)
    if (ii < 0 || ii >= Nx || jj < 0 || jj >= Ny) return 0.f;
    size_t o = (size_t)ii * Ny + jj;
    float cc = cn[o];
    return s.kappa * cc * cc * gn[o] / fmaf(bn[o], s.dt, 1.f);
  };
  const float bb = bn[cell], cc = cn[cell], gg = gn[cell];
  const float beta = bb * s.dt, q = 1.f / (1.f + beta), ql = q * gg;
  if (gb || gc) {
    Hood<1> h;
    h.load(y1 + off, Nx, Ny, i, j0);
    const float kl = s.kappa * h.lap(0);
    if (gb) gb[off + cell] = -s.dt * q * fmaf(cc * cc, kl, 2.f * (h.ce[0] - y2[off + cell])) * ql;
    if (gc) gc[off + cell] = 2.f * cc * kl * ql;
  }
  if (gy1) {
    float lapP = fmaf(-4.f, pval(i, j0), (pval(i - 1, j0) + pval(i + 1, j0)) + (pval(i, j0 - 1) + pval(i, j0 + 1)));
    gy1[off + cell] = fmaf(2.f, ql, lapP);
  }
  if (gy2) gy2[off + cell] = (beta - 1.f) * ql;
}

// ------------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------------
static inline dim3 stream_grid(int Nx, int Ny, int vec, int nbz) {
  int groups = (Ny + vec - 1) / vec;
  return dim3((groups + 31) / 32, (Nx + 3) / 4, nbz);
}

static int pick_batch_chunks(const wt_problem* p, int vec) {
  // enough threads to fill the chip (148 SMs x 2048 threads), but never more chunks than samples
  long per_plane = (long)p->Nx * ((p->Ny + vec - 1) / vec);
  long want = (148L * 2048 * 2 + per_plane - 1) / per_plane;
  if (want < 1) want = 1;
  if (want > p->B) want = p->B;
  return (int)want;
}

static bool vec4_ok(const wt_problem* p, std::initializer_list<const void*> ptrs) {
  if (p->Ny % 4) return false;
  for (const void* q : ptrs)
    if (q && ((uintptr_t)q & 15)) return false;
  return true;
}

size_t stream_tape_bytes(const wt_problem* p) {
  size_t field = (size_t)p->B * p->Nx * p->Ny * sizeof(float);
  bool general = nonlinear_mask(p) || (p->flags & WT_F_NEED_GRAD_B);
  return general ? field * ((size_t)p->T + 1) : field * (size_t)p->T;
}

static size_t stream_ws_fwd_base(const wt_problem* p) {
  size_t plane = (size_t)p->Nx * p->Ny;
  size_t n = 3 * plane * sizeof(float) + (size_t)(p->n_src + p->n_prb + 4) * sizeof(int32_t) + 256;
  return (n + 255) & ~(size_t)255;
}

size_t stream_ws_fwd_bytes(const wt_problem* p) { return stream_ws_fwd_base(p) + tile_extra_ws_bytes(p); }

size_t stream_ws_bwd_bytes(const wt_problem* p) {
  size_t plane = (size_t)p->Nx * p->Ny;
  size_t field = (size_t)p->B * plane;
  // coefficients (3) + accumulators (3) + two adjoint state fields + P + offsets
  size_t n = (6 * plane + 3 * field) * sizeof(float) + (size_t)(p->n_src + p->n_prb + 4) * sizeof(int32_t) + 512;
  return ((n + 255) & ~(size_t)255) + tile_extra_ws_bwd_bytes(p);
}

struct Offsets {
  int32_t* src;
  int32_t* prb;
  int* bad;
};

static int make_offsets(const wt_problem* p, const int32_t* src_ij, const int32_t* prb_ij, char* base, Offsets* o,
                        cudaStream_t st) {
  o->src = reinterpret_cast<int32_t*>(base);
  o->prb = o->src + p->n_src;
  o->bad = reinterpret_cast<int*>(o->prb + p->n_prb);
  WT_CUDA(cudaMemsetAsync(o->bad, 0, sizeof(int), st));
  if (p->n_src) k_off<<<(p->n_src + 127) / 128, 128, 0, st>>>(src_ij, p->n_src, p->Nx, p->Ny, o->src, o->bad);
  if (p->n_prb) k_off<<<(p->n_prb + 127) / 128, 128, 0, st>>>(prb_ij, p->n_prb, p->Nx, p->Ny, o->prb, o->bad);
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

int stream_forward(const wt_problem* p, const float* c, const float* b, const float* rho, const float* x,
                   const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_sq, float* u1, float* u2,
                   float* probe_out, float* probe_raw, float* fields_out, void* history, void* workspace,
                   cudaStream_t st, const wt_slab* slab) {
  const size_t plane = (size_t)p->Nx * p->Ny, field = plane * p->B;
  const int nl = nonlinear_mask(p);
  const bool general = nl || (p->flags & WT_F_NEED_GRAD_B);
  const Scalars s = make_scalars(p);
  float* a1 = reinterpret_cast<float*>(workspace);
  float* a3 = a1 + plane;
  float* gs = a3 + plane;
  Offsets off;
  WT_TRY(make_offsets(p, src_ij, prb_ij, reinterpret_cast<char*>(gs + plane), &off, st));
  if (!nl) {
    k_coeff<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(b, c, (int)plane, p->dt, (p->dt * p->dt) / (p->h * p->h),
                                                              a1, a3, gs);
  }
  if (p->flags & WT_F_ZERO_INIT) {
    WT_CUDA(cudaMemsetAsync(u1, 0, field * sizeof(float), st));
    WT_CUDA(cudaMemsetAsync(u2, 0, field * sizeof(float), st));
  }
  float* tape = reinterpret_cast<float*>(history);
  if (!fields_out && tile_eligible(p) && vec4_ok(p, {u1, u2, history, workspace})) {
    // large grid: K time steps per HBM round trip (wt_tile.cu)
    float* extra = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + stream_ws_fwd_base(p));
    return tile_forward(p, a1, a3, x, src_ij, prb_ij, prb_sq, u1, u2, probe_out, probe_raw, tape, extra, st, nullptr, slab);
  }
  const bool v4 = vec4_ok(p, {u1, u2, history, fields_out, workspace});
  const int vec = v4 ? 4 : 1;
  const int nbz = pick_batch_chunks(p, vec);
  float* cur1 = u1;
  float* cur2 = u2;
  for (int t = 0; t < p->T; ++t) {
    FwdArgs a;
    a.Nx = p->Nx; a.Ny = p->Ny; a.B = p->B; a.bchunk = (p->B + nbz - 1) / nbz; a.plane = plane;
    a.u1 = cur1; a.u2 = cur2; a.a1 = a1; a.a3 = a3; a.bpml = b; a.clin = c; a.rho = rho;
    a.tape_lap = (tape && !general) ? tape + (size_t)t * field : nullptr;
    a.tape_u1 = (tape && general) ? tape + (size_t)(t + 1) * field : nullptr;
    a.tape_u2 = (tape && general && t == 0) ? tape : nullptr;
    const int fe = p->field_every > 1 ? p->field_every : 1;      // time-decimated snapshots: step t -> slot t / fe
    a.fields = (fields_out && (t + 1) % fe == 0) ? fields_out + (size_t)(t / fe) * plane : nullptr;
    a.fields_bstride = (size_t)(p->T / fe) * plane;
    a.s = s;
    dim3 grid = stream_grid(p->Nx, p->Ny, vec, nbz), block(32, 4);
#define WT_LAUNCH_FWD(V)                                                          \
  switch (nl) {                                                                   \
    case 0: k_stream_fwd<V, true, false, false><<<grid, block, 0, st>>>(a); break;  \
    case 1: k_stream_fwd<V, false, true, false><<<grid, block, 0, st>>>(a); break;  \
    case 2: k_stream_fwd<V, false, false, true><<<grid, block, 0, st>>>(a); break;  \
    default: k_stream_fwd<V, false, true, true><<<grid, block, 0, st>>>(a); break;  \
  }
    if (v4) { WT_LAUNCH_FWD(4) } else { WT_LAUNCH_FWD(1) }
#undef WT_LAUNCH_FWD
    // the new field now lives in cur2
    k_src_prb<<<p->B, 128, 0, st>>>(cur2, plane, a.fields, a.fields_bstride, x, t, p->T, off.src, p->n_src, off.prb,
                                    prb_sq, p->n_prb, probe_out, probe_raw);
    float* tmp = cur1; cur1 = cur2; cur2 = tmp;
    // slab decomposition: refresh the ghost rows every halo steps (halo is even: cur1 == u1 here)
    if (slab && t + 1 < p->T && (t + 1) % slab->halo == 0) WT_TRY(slab_exchange(slab, p->B, p->Nx, p->Ny, cur1, cur2, st));
  }
  WT_CUDA(cudaGetLastError());
  if (cur1 != u1) k_swap<<<592, 256, 0, st>>>(u1, u2, field);   // odd T: put the latest field back into u1
  WT_CUDA(cudaGetLastError());
  WT_TRY(slab_exchange(slab, p->B, p->Nx, p->Ny, u1, u2, st));
  return WT_OK;
}

int stream_backward(const wt_problem* p, const float* c, const float* b, const float* rho, const int32_t* src_ij,
                    const int32_t* prb_ij, const int32_t* prb_sq, const float* grad_probe, const float* probe_raw,
                    const float* grad_fields, const void* history, float* adj1, float* adj2, float* grad_c,
                    float* grad_b, float* grad_rho, float* grad_x, void* workspace, cudaStream_t st,
                    const wt_slab* slab) {
  const size_t plane = (size_t)p->Nx * p->Ny, field = plane * p->B;
  const int nl = nonlinear_mask(p);
  const bool general = nl || (p->flags & WT_F_NEED_GRAD_B);
  const Scalars s = make_scalars(p);
  float* a1 = reinterpret_cast<float*>(workspace);
  float* a3 = a1 + plane;
  float* gs = a3 + plane;
  float* Gc = gs + plane;
  float* Gb = Gc + plane;
  float* Gr = Gb + plane;
  float* w1 = Gr + plane;        // adjoint state buffers when the caller does not chain
  float* w2 = w1 + field;
  float* P = w2 + field;
  Offsets off;
  WT_TRY(make_offsets(p, src_ij, prb_ij, reinterpret_cast<char*>(P + field), &off, st));
  WT_CUDA(cudaMemsetAsync(Gc, 0, 3 * plane * sizeof(float), st));
  const bool chained = adj1 && adj2;
  float* l1 = chained ? adj1 : w1;   // lambda_t / carry1
  float* l2 = chained ? adj2 : w2;   // lambda_{t+1} / carry2
  if (!chained) WT_CUDA(cudaMemsetAsync(w1, 0, 2 * field * sizeof(float), st));
  const float* tape = reinterpret_cast<const float*>(history);
  const size_t gfb = (size_t)p->T * plane;

  if (!general) {
    k_coeff<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(b, c, (int)plane, p->dt, (p->dt * p->dt) / (p->h * p->h),
                                                              a1, a3, gs);
    if (!grad_fields && tile_eligible(p) && vec4_ok(p, {l1, l2, history, workspace})) {
      // large grid: K reverse steps per HBM round trip (wt_tile.cu), state kept as P = a3*lambda
      size_t base = (6 * plane + 3 * field) * sizeof(float) + (size_t)(p->n_src + p->n_prb + 4) * sizeof(int32_t) + 512;
      base = (base + 255) & ~(size_t)255;
      float* extra = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + base);
      float* spare1 = chained ? w1 : P;
      float* spare2 = chained ? w2 : extra;
      WT_TRY(tile_backward(p, a1, a3, c, src_ij, prb_ij, prb_sq, grad_probe, probe_raw, tape, l1, l2, spare1, spare2, Gc, grad_c,
                           grad_x, chained, st, slab));
      if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));
      if (grad_rho) WT_CUDA(cudaMemsetAsync(grad_rho, 0, plane * sizeof(float), st));
      return WT_OK;
    }
    const bool v4 = vec4_ok(p, {l1, l2, history, grad_fields, workspace});
    const int vec = v4 ? 4 : 1;
    const int nbz = pick_batch_chunks(p, vec);
    // lambda_{T-1} = adj1_in + dLoss/dfields[T-1] (+ seed_{T-1}, added inside the loop)
    if (grad_fields)
      k_add_fields<<<dim3((unsigned)((plane + 255) / 256), p->B), 256, 0, st>>>(l1, plane,
                                                                               grad_fields + (size_t)(p->T - 1) * plane, gfb);
    for (int t = p->T - 1; t >= 0; --t) {
      k_adj_seed<<<p->B, 128, 0, st>>>(l1, plane, grad_probe, probe_raw, t, p->T, off.prb, prb_sq, p->n_prb, off.src,
                                       p->n_src, grad_x);
      AdjLinArgs a;
      a.Nx = p->Nx; a.Ny = p->Ny; a.B = p->B; a.bchunk = (p->B + nbz - 1) / nbz; a.plane = plane;
      a.lam1 = l1; a.lam2 = l2; a.a1 = a1; a.a3 = a3; a.tape_lap = tape + (size_t)t * field;
      a.gfields = (grad_fields && t > 0) ? grad_fields + (size_t)(t - 1) * plane : nullptr;
      a.gf_bstride = gfb;
      a.G = Gc; a.premul = (t == p->T - 1) ? 1 : 0; a.atomic_G = nbz > 1;
      dim3 grid = stream_grid(p->Nx, p->Ny, vec, nbz), block(32, 4);
      if (v4) k_stream_adj_lin<4><<<grid, block, 0, st>>>(a);
      else k_stream_adj_lin<1><<<grid, block, 0, st>>>(a);
      float* tmp = l1; l1 = l2; l2 = tmp;   // l1 = lambda_{t-1} (unseeded), l2 = lambda_t
      const int done = p->T - t;
      if (slab && t > 0 && done % slab->halo == 0) WT_TRY(slab_exchange(slab, p->B, p->Nx, p->Ny, l1, l2, st));
    }
    WT_CUDA(cudaGetLastError());
    k_finish_grad<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(Gc, gs, 1, plane, plane, grad_c);
    if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));
    if (grad_rho) WT_CUDA(cudaMemsetAsync(grad_rho, 0, plane * sizeof(float), st));
    if (chained) {
      // l1 = dLoss/du1_in, l2 = lambda_0 -> weight it to get dLoss/du2_in; then restore the caller's order
      k_scale_carry<<<592, 256, 0, st>>>(l2, a1, plane, field);
      if (l1 != adj1) k_swap<<<592, 256, 0, st>>>(adj1, adj2, field);
      WT_TRY(slab_exchange(slab, p->B, p->Nx, p->Ny, adj1, adj2, st));
    }
    WT_CUDA(cudaGetLastError());
    return WT_OK;
  }
  if (slab) {
    set_error("wt_slab_backward: saturable damping / Kerr terms / grad_b are not supported under domain decomposition");
    return WT_EUNSUPPORTED;
  }

  // general path
  const int nbz = pick_batch_chunks(p, 1);
  if (grad_fields)
    k_add_fields<<<dim3((unsigned)((plane + 255) / 256), p->B), 256, 0, st>>>(l1, plane,
                                                                             grad_fields + (size_t)(p->T - 1) * plane, gfb);
  for (int t = p->T - 1; t >= 0; --t) {
    k_adj_seed<<<p->B, 128, 0, st>>>(l1, plane, grad_probe, probe_raw, t, p->T, off.prb, prb_sq, p->n_prb, off.src,
                                     p->n_src, grad_x);
    AdjNlArgs a;
    a.Nx = p->Nx; a.Ny = p->Ny; a.B = p->B; a.bchunk = (p->B + nbz - 1) / nbz; a.plane = plane;
    a.c1 = l1; a.c2 = l2; a.P = P; a.tu1 = tape + (size_t)(t + 1) * field; a.tu2 = tape + (size_t)t * field;
    a.bpml = b; a.clin = c; a.rho = rho; a.Gc = Gc; a.Gb = Gb; a.Grho = Gr; a.atomic_G = nbz > 1; a.s = s;
    dim3 grid = stream_grid(p->Nx, p->Ny, 1, nbz), block(32, 4);
    switch (nl) {
      case 0: k_stream_adjA<false, false><<<grid, block, 0, st>>>(a); break;
      case 1: k_stream_adjA<true, false><<<grid, block, 0, st>>>(a); break;
      case 2: k_stream_adjA<false, true><<<grid, block, 0, st>>>(a); break;
      default: k_stream_adjA<true, true><<<grid, block, 0, st>>>(a); break;
    }
    k_stream_adjB<<<grid, block, 0, st>>>(p->Nx, p->Ny, p->B, a.bchunk, plane, P, l2,
                                          (grad_fields && t > 0) ? grad_fields + (size_t)(t - 1) * plane : nullptr, gfb);
    float* tmp = l1; l1 = l2; l2 = tmp;     // l1 = carry1 for step t-1, l2 = carry2
  }
  WT_CUDA(cudaGetLastError());
  k_finish_grad<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(Gc, nullptr, 1, plane, plane, grad_c);
  if (grad_b) k_finish_grad<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(Gb, nullptr, 1, plane, plane, grad_b);
  if (grad_rho) k_finish_grad<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(Gr, nullptr, 1, plane, plane, grad_rho);
  if (chained && l1 != adj1) k_swap<<<592, 256, 0, st>>>(adj1, adj2, field);
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

int step_forward(const wt_problem* p, const float* b, int bb, const float* c, int cb, const float* y1,
                 const float* y2, float* y, cudaStream_t st) {
  const size_t plane = (size_t)p->Nx * p->Ny;
  k_step_fwd<<<stream_grid(p->Nx, p->Ny, 1, p->B), dim3(32, 4), 0, st>>>(p->Nx, p->Ny, p->B, plane, b, bb ? plane : 0,
                                                                        c, cb ? plane : 0, y1, y2, y, make_scalars(p));
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

int step_backward(const wt_problem* p, const float* b, int bb, const float* c, int cb, const float* y1,
                  const float* y2, const float* g, float* gb, float* gc, float* gy1, float* gy2, cudaStream_t st) {
  const size_t plane = (size_t)p->Nx * p->Ny;
  k_step_bwd<<<stream_grid(p->Nx, p->Ny, 1, p->B), dim3(32, 4), 0, st>>>(p->Nx, p->Ny, p->B, plane, b, bb ? plane : 0,
                                                                        c, cb ? plane : 0, y1, y2, g, gb, gc, gy1,
                                                                        gy2, make_scalars(p));
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

}  // namespace wt
