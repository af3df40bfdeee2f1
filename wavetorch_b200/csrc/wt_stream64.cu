// float64 time loop and adjoint: the reference's `utils.set_dtype('float64')` mode (utils.py:14-20; study/example.yml:5).
//
// The float32 paths are the product (BASELINE); this file exists so that a model built in float64 is integrated in float64
// instead of being cast -- and so that the float32 kernels can be cross-checked on the GPU against a double-precision run of
// the same problem.  One launch per time step, one thread per cell looping over the samples of its batch chunk, fields in
// natural [B,Nx,Ny] layout: the structure of the general float32 streaming path (wt_stream.cu), nothing blocked or staged --
// a 5-point stencil in double is HBM-bound at half the float32 cell rate and B200's FP64 rate is ample for it.
//
//   forward  y = u2 + 2q (u1 - u2) + q kappa c^2 L(u1),  q = 1/(1 + dt b),  b = b_pml + rho b0/(1 + (u1/uth)^2),
//            c = c_lin + rho c_nl u1^2                                                   (cell.py:12-17, 94-102)
//   adjoint  two passes per step (own-cell pass A: coefficients from u_{t-1}, gradients, P = kappa c^2 q lambda; stencil
//            pass B: lambda_{t-1} += L(P)), tape = every field                          (cell.py:27-44 + autograd of 94-100)
#include "wt_common.cuh"

namespace wt {

struct S64 {
  double dt, kappa, b0, inv_uth, c_nl;
  int sat, kerr;
};

static S64 make_s64(const wt_problem* p) {
  S64 s;
  s.dt = p->dt;
  s.kappa = (p->dt * p->dt) / (p->h * p->h);
  s.b0 = p->b0;
  s.inv_uth = p->b0 > 0 ? 1.0 / p->uth : 0.0;
  s.c_nl = p->c_nl;
  s.sat = p->b0 > 0;
  s.kerr = p->c_nl != 0;
  return s;
}

__device__ __forceinline__ double lap64(const double* __restrict__ u, int Nx, int Ny, int i, int j) {
  const size_t o = (size_t)i * Ny + j;
  const double n = i > 0 ? u[o - Ny] : 0.0, s = i + 1 < Nx ? u[o + Ny] : 0.0;
  const double w = j > 0 ? u[o - 1] : 0.0, e = j + 1 < Ny ? u[o + 1] : 0.0;
  return fma(-4.0, u[o], (n + s) + (w + e));
}

__device__ __forceinline__ void coef64(const S64& s, double bp, double cl, double rh, double u1, double& beta, double& cc,
                                       double& d) {
  d = 1.0;
  double b = bp;
  cc = cl;
  if (s.sat) {
    const double r = u1 * s.inv_uth;
    d = fma(r, r, 1.0);
    b = bp + rh * s.b0 / d;
  }
  if (s.kerr) cc = cl + rh * s.c_nl * u1 * u1;
  beta = b * s.dt;
}

__global__ void __launch_bounds__(128) k64_fwd(int Nx, int Ny, int B, int bchunk, size_t plane, const double* __restrict__ u1,
                                               double* __restrict__ u2, const double* __restrict__ bpml,
                                               const double* __restrict__ clin, const double* __restrict__ rho,
                                               double* __restrict__ tape_u1, double* __restrict__ tape_u2,
                                               double* __restrict__ fields, size_t fields_bstride, S64 s) {
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 4 + threadIdx.y;
  if (j >= Ny || i >= Nx) return;
  const size_t cell = (size_t)i * Ny + j;
  const double bp = bpml[cell], cl = clin[cell], rh = (s.sat || s.kerr) ? rho[cell] : 0.0;
  const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * plane;
    const double c = u1[off + cell], w = u2[off + cell];
    const double l = lap64(u1 + off, Nx, Ny, i, j);
    double beta, cc, d;
    coef64(s, bp, cl, rh, c, beta, cc, d);
    const double q = 1.0 / (1.0 + beta);
    const double y = fma(q * s.kappa * cc * cc, l, fma(2.0 * q, c - w, w));
    if (tape_u2) tape_u2[off + cell] = w;
    if (tape_u1) tape_u1[off + cell] = c;
    u2[off + cell] = y;
    if (fields) fields[(size_t)b * fields_bstride + cell] = y;
  }
}

// source injection (source.py:19-22) and probe readout (probe.py:15,27) of one step; one block per sample
__global__ void k64_src_prb(double* __restrict__ U, size_t plane, double* fields, size_t fields_bstride,
                            const double* __restrict__ x, int t, int T, const int32_t* __restrict__ src_ij, int n_src,
                            const int32_t* __restrict__ prb_ij, const int32_t* __restrict__ prb_sq, int n_prb, int Ny,
                            double* __restrict__ probe_out, double* __restrict__ probe_raw) {
  const int b = blockIdx.x;
  double* u = U + (size_t)b * plane;
  const double xv = x[(size_t)b * T + t];
  for (int k = threadIdx.x; k < n_src; k += blockDim.x) atomicAdd(u + (size_t)src_ij[2 * k] * Ny + src_ij[2 * k + 1], xv);
  __syncthreads();
  if (fields) {
    double* f = fields + (size_t)b * fields_bstride;
    for (int k = threadIdx.x; k < n_src; k += blockDim.x) {
      const size_t o = (size_t)src_ij[2 * k] * Ny + src_ij[2 * k + 1];
      f[o] = u[o];
    }
  }
  for (int p = threadIdx.x; p < n_prb; p += blockDim.x) {
    const double v = u[(size_t)prb_ij[2 * p] * Ny + prb_ij[2 * p + 1]];
    const size_t o = ((size_t)b * T + t) * n_prb + p;
    if (probe_raw) probe_raw[o] = v;
    if (probe_out) probe_out[o] = prb_sq[p] ? v * v : v;
  }
}

// lambda_t += seeds of step t; dLoss/dx[:, t] = sum over source listings of lambda_t
__global__ void k64_seed(double* __restrict__ lam, size_t plane, const double* __restrict__ grad_probe,
                         const double* __restrict__ probe_raw, int t, int T, const int32_t* __restrict__ prb_ij,
                         const int32_t* __restrict__ prb_sq, int n_prb, const int32_t* __restrict__ src_ij, int n_src, int Ny,
                         double* __restrict__ grad_x) {
  const int b = blockIdx.x;
  double* u = lam + (size_t)b * plane;
  for (int p = threadIdx.x; p < n_prb; p += blockDim.x) {
    const size_t o = ((size_t)b * T + t) * n_prb + p;
    double g = grad_probe[o];
    if (prb_sq[p]) g *= 2.0 * probe_raw[o];
    atomicAdd(u + (size_t)prb_ij[2 * p] * Ny + prb_ij[2 * p + 1], g);
  }
  if (!grad_x) return;
  __syncthreads();
  __shared__ double red[128];
  double s = 0.0;
  for (int k = threadIdx.x; k < n_src; k += blockDim.x) s += u[(size_t)src_ij[2 * k] * Ny + src_ij[2 * k + 1]];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 64; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) grad_x[(size_t)b * T + t] = red[0];
}

// pass A: c1 in lambda_t (seeded) / out carry2' = (beta-1) q lambda_t; c2 in carry2 / out lambda_{t-1} without the stencil term
__global__ void __launch_bounds__(128) k64_adjA(int Nx, int Ny, int B, int bchunk, size_t plane, double* __restrict__ c1,
                                                double* __restrict__ c2, double* __restrict__ P,
                                                const double* __restrict__ tu1, const double* __restrict__ tu2,
                                                const double* __restrict__ bpml, const double* __restrict__ clin,
                                                const double* __restrict__ rho, double* __restrict__ Gc,
                                                double* __restrict__ Gb, double* __restrict__ Gr, int atomic_G, S64 s) {
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 4 + threadIdx.y;
  if (j >= Ny || i >= Nx) return;
  const size_t cell = (size_t)i * Ny + j;
  const double bp = bpml[cell], cl = clin[cell], rh = (s.sat || s.kerr) ? rho[cell] : 0.0;
  double gc = 0.0, gb = 0.0, gr = 0.0;
  const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * plane;
    const double u1 = tu1[off + cell], u2 = tu2[off + cell];
    const double lap = lap64(tu1 + off, Nx, Ny, i, j);
    const double lam = c1[off + cell], car2 = c2[off + cell];
    double beta, cc, d;
    coef64(s, bp, cl, rh, u1, beta, cc, d);
    const double q = 1.0 / (1.0 + beta), ql = q * lam, kl = s.kappa * lap;
    const double S = fma(cc * cc, kl, 2.0 * (u1 - u2));
    const double g_b = -s.dt * q * S * ql;            // cell.py:33-34
    const double g_c = 2.0 * cc * kl * ql;            // cell.py:36
    double gu1 = 2.0 * ql;                            // own-cell part of cell.py:39-40
    if (s.sat) {
      gr = fma(g_b, s.b0 / d, gr);
      gu1 = fma(g_b, rh * s.b0 * (-2.0 * u1 * s.inv_uth * s.inv_uth) / (d * d), gu1);
    }
    if (s.kerr) {
      gr = fma(g_c, s.c_nl * u1 * u1, gr);
      gu1 = fma(g_c, 2.0 * rh * s.c_nl * u1, gu1);
    }
    gc += g_c;
    gb += g_b;
    P[off + cell] = s.kappa * cc * cc * ql;
    c2[off + cell] = car2 + gu1;
    c1[off + cell] = (beta - 1.0) * ql;               // cell.py:42
  }
  if (atomic_G) {
    atomicAdd(Gc + cell, gc);
    if (Gb) atomicAdd(Gb + cell, gb);
    if (Gr) atomicAdd(Gr + cell, gr);
  } else {
    Gc[cell] += gc;
    if (Gb) Gb[cell] += gb;
    if (Gr) Gr[cell] += gr;
  }
}

__global__ void __launch_bounds__(128) k64_adjB(int Nx, int Ny, int B, int bchunk, size_t plane, const double* __restrict__ P,
                                                double* __restrict__ c2) {
  const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 4 + threadIdx.y;
  if (j >= Ny || i >= Nx) return;
  const size_t cell = (size_t)i * Ny + j;
  const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    const size_t off = (size_t)b * plane;
    c2[off + cell] += lap64(P + off, Nx, Ny, i, j);
  }
}

__global__ void k64_swap(double* __restrict__ a, double* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double t = a[i]; a[i] = b[i]; b[i] = t;
  }
}

static int chunks64(const wt_problem* p) {
  long per_plane = (long)p->Nx * p->Ny;
  long want = (148L * 2048 * 2 + per_plane - 1) / per_plane;
  if (want < 1) want = 1;
  if (want > p->B) want = p->B;
  return (int)want;
}

static int check64(const wt_problem* p) {
  WT_REQUIRE(p != nullptr, "wt_problem is NULL");
  WT_REQUIRE(p->Nx >= 1 && p->Ny >= 1 && p->B >= 1 && p->T >= 0, "bad problem %dx%d B=%d T=%d", p->Nx, p->Ny, p->B, p->T);
  WT_REQUIRE(p->dt > 0 && p->h > 0, "dt and h must be positive");
  WT_REQUIRE(!(p->b0 > 0) || p->uth != 0, "saturable damping needs uth != 0");
  return WT_OK;
}

}  // namespace wt

using namespace wt;

extern "C" {

int wt_query_plan_f64(const wt_problem* p, wt_plan* plan) {
  WT_TRY(check64(p));
  WT_REQUIRE(plan != nullptr, "plan is NULL");
  memset(plan, 0, sizeof(*plan));
  const size_t plane = (size_t)p->Nx * p->Ny, field = plane * p->B;
  plan->path = WT_PATH_STREAM;
  plan->cluster = 1;
  plan->threads = 128;
  plan->nonlinear = nonlinear_mask(p);
  plan->history_bytes = field * ((size_t)p->T + 1) * sizeof(double);       // every field (u_{-2}, u_{-1} .. u_{T-2})
  plan->workspace_fwd_bytes = 256;
  plan->workspace_bwd_bytes = (3 * plane + 3 * field) * sizeof(double) + 256;  // Gc, Gb, Grho; lambda pair and P
  plan->launches_fwd = 2 * p->T + 2;
  plan->launches_bwd = 3 * p->T + 4;
  return WT_OK;
}

int wt_forward_f64(const wt_problem* p, const double* c, const double* b, const double* rho, const double* x,
                   const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, double* u1, double* u2,
                   double* probe_out, double* probe_raw, double* fields_out, void* history, size_t history_bytes,
                   void* workspace, size_t workspace_bytes, void* stream) {
  wt_plan plan;
  WT_TRY(wt_query_plan_f64(p, &plan));
  WT_REQUIRE(c && b && x && u1 && u2, "wt_forward_f64: c, b, x, u1, u2 must not be NULL");
  WT_REQUIRE(!plan.nonlinear || rho, "wt_forward_f64: rho is required when b0 > 0 or c_nl != 0");
  WT_REQUIRE(p->n_src == 0 || src_ij, "wt_forward_f64: src_ij is NULL");
  WT_REQUIRE(p->n_prb == 0 || (prb_ij && prb_square), "wt_forward_f64: prb_ij / prb_square is NULL");
  if (history && history_bytes < plan.history_bytes) {
    set_error("wt_forward_f64: history %zu < %llu bytes", history_bytes, (unsigned long long)plan.history_bytes);
    return WT_ENOSPACE;
  }
  (void)workspace; (void)workspace_bytes;
  WT_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t plane = (size_t)p->Nx * p->Ny, field = plane * p->B;
  const S64 s = make_s64(p);
  if (p->flags & WT_F_ZERO_INIT) {
    WT_CUDA(cudaMemsetAsync(u1, 0, field * sizeof(double), st));
    WT_CUDA(cudaMemsetAsync(u2, 0, field * sizeof(double), st));
  }
  double* tape = reinterpret_cast<double*>(history);
  const int nbz = chunks64(p), bchunk = (p->B + nbz - 1) / nbz;
  const dim3 grid((p->Ny + 31) / 32, (p->Nx + 3) / 4, nbz), block(32, 4);
  double* cur1 = u1;
  double* cur2 = u2;
  const int fe = p->field_every > 1 ? p->field_every : 1;
  for (int t = 0; t < p->T; ++t) {
    double* f = (fields_out && (t + 1) % fe == 0) ? fields_out + (size_t)(t / fe) * plane : nullptr;
    const size_t fbs = (size_t)(p->T / fe) * plane;
    k64_fwd<<<grid, block, 0, st>>>(p->Nx, p->Ny, p->B, bchunk, plane, cur1, cur2, b, c, rho,
                                    tape ? tape + (size_t)(t + 1) * field : nullptr, (tape && t == 0) ? tape : nullptr, f, fbs, s);
    k64_src_prb<<<p->B, 128, 0, st>>>(cur2, plane, f, fbs, x, t, p->T, src_ij, p->n_src, prb_ij, prb_square, p->n_prb, p->Ny,
                                      probe_out, probe_raw);
    double* tmp = cur1; cur1 = cur2; cur2 = tmp;
  }
  WT_CUDA(cudaGetLastError());
  if (cur1 != u1) {   // odd T: the latest field sits in the caller's u2
    k64_swap<<<592, 256, 0, st>>>(u1, u2, field);
    WT_CUDA(cudaGetLastError());
  }
  return WT_OK;
}

int wt_backward_f64(const wt_problem* p, const double* c, const double* b, const double* rho, const int32_t* src_ij,
                    const int32_t* prb_ij, const int32_t* prb_square, const double* grad_probe, const double* probe_raw,
                    const void* history, size_t history_bytes, double* grad_c, double* grad_b, double* grad_rho, double* grad_x,
                    void* workspace, size_t workspace_bytes, void* stream) {
  wt_plan plan;
  WT_TRY(wt_query_plan_f64(p, &plan));
  WT_REQUIRE(c && b && history && grad_c && workspace, "wt_backward_f64: c, b, history, grad_c, workspace must not be NULL");
  WT_REQUIRE(!plan.nonlinear || rho, "wt_backward_f64: rho is required when b0 > 0 or c_nl != 0");
  WT_REQUIRE(p->n_prb == 0 || (prb_ij && prb_square && grad_probe && probe_raw), "wt_backward_f64: probe arrays missing");
  if (workspace_bytes < plan.workspace_bwd_bytes || history_bytes < plan.history_bytes) {
    set_error("wt_backward_f64: workspace %zu / history %zu too small (%llu / %llu)", workspace_bytes, history_bytes,
              (unsigned long long)plan.workspace_bwd_bytes, (unsigned long long)plan.history_bytes);
    return WT_ENOSPACE;
  }
  WT_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t plane = (size_t)p->Nx * p->Ny, field = plane * p->B;
  const S64 s = make_s64(p);
  double* Gc = reinterpret_cast<double*>(workspace);
  double* Gb = Gc + plane;
  double* Gr = Gb + plane;
  double* l1 = Gr + plane;
  double* l2 = l1 + field;
  double* P = l2 + field;
  WT_CUDA(cudaMemsetAsync(Gc, 0, (3 * plane + 2 * field) * sizeof(double), st));
  const double* tape = reinterpret_cast<const double*>(history);
  const int nbz = chunks64(p), bchunk = (p->B + nbz - 1) / nbz;
  const dim3 grid((p->Ny + 31) / 32, (p->Nx + 3) / 4, nbz), block(32, 4);
  for (int t = p->T - 1; t >= 0; --t) {
    k64_seed<<<p->B, 128, 0, st>>>(l1, plane, grad_probe, probe_raw, t, p->T, prb_ij, prb_square, p->n_prb, src_ij, p->n_src,
                                   p->Ny, grad_x);
    k64_adjA<<<grid, block, 0, st>>>(p->Nx, p->Ny, p->B, bchunk, plane, l1, l2, P, tape + (size_t)(t + 1) * field,
                                     tape + (size_t)t * field, b, c, rho, Gc, grad_b ? Gb : nullptr,
                                     (grad_rho && plan.nonlinear) ? Gr : nullptr, nbz > 1, s);
    k64_adjB<<<grid, block, 0, st>>>(p->Nx, p->Ny, p->B, bchunk, plane, P, l2);
    double* tmp = l1; l1 = l2; l2 = tmp;
  }
  WT_CUDA(cudaGetLastError());
  WT_CUDA(cudaMemcpyAsync(grad_c, Gc, plane * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (grad_b) WT_CUDA(cudaMemcpyAsync(grad_b, Gb, plane * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (grad_rho) {
    if (plan.nonlinear) WT_CUDA(cudaMemcpyAsync(grad_rho, Gr, plane * sizeof(double), cudaMemcpyDeviceToDevice, st));
    else WT_CUDA(cudaMemsetAsync(grad_rho, 0, plane * sizeof(double), st));
  }
  return WT_OK;
}

}  // extern "C"
