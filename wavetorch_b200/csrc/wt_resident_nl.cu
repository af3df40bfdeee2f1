// On-chip kernels for the nonlinear cell (BASELINE config 4): saturable damping b(u) = b_pml + rho*b0/(1+(u/uth)^2)
// and Kerr-like wave speed c(u) = c_lin + rho*c_nl*u^2, evaluated from u_{t-1} every step (cell.py:94-102).
//
// Same decomposition, ghost-row exchange and staging as the linear kernels (wt_resident.cu); differences:
//   * per-cell registers hold b_pml, c_lin, rho instead of the precomputed a1, a3; the coefficients are rebuilt per
//     step with the same device functions as the streaming path (wt_common.cuh), so both paths agree bitwise
//   * the tape holds u_{t-1} AND L(u_{t-1}) per step (8 B/cell); u_{t-2} is read from the next ring stage
//   * the adjoint accumulates dLoss/dc_lin and the direct dLoss/drho per cluster
#include <type_traits>

#include "wt_resident.h"
#include "wt_resident_dev.cuh"

namespace wt {

template <int R>
constexpr int res_nl_max_threads() {
  return R <= 2 ? 512 : R == 3 ? 384 : 384;   // register budget: the nonlinear adjoint keeps ~9 values per cell live
}

// Per-cell constants of the nonlinear coefficients, kept in registers for the whole loop:
//   e  = 1 + dt*b_pml          f = dt*rho*b0          cl = c_lin          g = rho*c_nl
// so that, with d = 1 + (u/uth)^2,
//   q = 1/(1 + dt*b(u)) = 1/(e + f/d) = d/(d*e + f)          c(u) = cl + g*u^2          (cell.py:94-102, :12-17)
// One MUFU reciprocal per cell and step in the forward kernel, two in the adjoint (which also needs 1/d).  rcp.approx is
// within 1 ulp; q only scales the damping, the u_{t-1}/u_{t-2} weights a1 and 1-a1 still sum to exactly one.
template <int R>
__device__ __forceinline__ void load_nl_consts(const ResArgs& a, bool active, int gi0, int j0, float (&e)[R][4],
                                               float (&f)[R][4], float (&cl)[R][4], float (&g)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int gi = gi0 + r, j = j0 + k;
      bool ok = active && gi < a.Nx && j < a.Ny;
      size_t o = (size_t)gi * a.Ny + j;
      // cells outside the domain: c = 0 makes a3 = 0, so they stay exactly zero like the linear kernels' padding
      const float bp = ok ? a.bpml[o] : 0.f;
      const float rh = ok ? a.rho[o] : 0.f;
      e[r][k] = fmaf(bp, a.s.dt, 1.f);
      f[r][k] = a.s.dt * (rh * a.s.b0);
      cl[r][k] = ok ? a.clin[o] : 0.f;
      g[r][k] = rh * a.s.c_nl;
    }
}

__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// =================================================================================================
// forward
// =================================================================================================
// CKPT: checkpoint-and-recompute instantiation (see k_res_fwd in wt_resident.cu): steps [t_off, t_off + T) of longer
// sequences, optional start from a register-patch snapshot, snapshots every snap_every steps
template <int R, bool SAT, bool KERR, bool FIELDS = false, int PITCH = 0, int NTC = 0, bool CKPT = false>
__global__ void __launch_bounds__(NTC ? NTC : res_nl_max_threads<R>()) k_res_fwd_nl(ResArgs a) {
  extern __shared__ float4 smem4[];
  const int pitch = PITCH ? PITCH : a.pitch;
  const int slab_f = slab_words(R, a.Hc, pitch);
  float* fld = reinterpret_cast<float*>(smem4);
  float* xs = fld + 2 * slab_f;
  float* ps = xs + 2 * TB;
  int* poff = reinterpret_cast<int*>(ps + 2 * TB * a.n_prb);
  uint64_t* bars = reinterpret_cast<uint64_t*>(poff + a.n_prb + (a.n_prb & 1));

  Lane<R> L;
  L.init(a, fld, bars);
  const int tid = L.tid, NT = NTC ? NTC : blockDim.x;
  float ce[R][4], cf[R][4], cl[R][4], cg[R][4];
  load_nl_consts<R>(a, L.active, L.gi0, L.j0, ce, cf, cl, cg);
  if (!KERR) {   // constant wave speed: keep kappa*c^2
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) cl[r][k] = a.s.kappa * (cl[r][k] * cl[r][k]);
  }
  if (!SAT) {    // constant damping: keep q
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) ce[r][k] = __frcp_rn(ce[r][k]);
  }
  unsigned m1, m2;
  source_masks<R>(a, L.active, L.gi0, L.j0, m1, m2);
  for (int p = tid; p < a.n_prb; p += NT) {
    int li = a.prb_ij[2 * p] - L.rank * a.Hc, pj = a.prb_ij[2 * p + 1];
    poff[p] = (li >= 0 && li < a.Hc) ? slab_cell(R, pitch, li, pj) : -1;
  }
  L.pub_all = L.active && probe_in_interior<R>(a, L.rank, L.lt);
  for (int i = tid; i < 2 * slab_f; i += NT) fld[i] = 0.f;
  if (a.C > 1) cg::this_cluster().sync(); else __syncthreads();
  const int plane_lane = NT - 1 - tid;   // probe p is sampled by lane NT-1-p: the highest warp has issue priority
  const int my_poff = (plane_lane < a.n_prb) ? poff[plane_lane] : -1;
  const int own = (L.lr0 + 1) * pitch + (L.run + 1) * slab_skew(R, pitch) + L.g;
  const size_t tape_step = (size_t)a.C * 2 * R * NT;
  const size_t plane = (size_t)a.Nx * a.Ny;
  const Scalars s = a.s;

  for (int b = L.cid; b < a.B; b += a.n_clusters) {
    float v[R][4], w[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int gi = L.gi0 + r, j = L.j0 + k;
        bool ok = !CKPT && L.active && gi < a.Nx && j < a.Ny && !(a.flags & WT_F_ZERO_INIT);
        size_t o = ((size_t)b * a.Nx + gi) * a.Ny + j;
        v[r][k] = ok ? a.u1[o] : 0.f;
        w[r][k] = ok ? a.u2[o] : 0.f;
      }
    if (CKPT && a.snap_in) {   // resume from a snapshot: my own registers, as I stored them
      const float4* sp = a.snap_in + ((size_t)b * a.C + L.rank) * 2 * R * NT + tid;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p = sp[r * NT], q = sp[(R + r) * NT];
        v[r][0] = p.x; v[r][1] = p.y; v[r][2] = p.z; v[r][3] = p.w;
        w[r][0] = q.x; w[r][1] = q.y; w[r][2] = q.z; w[r][3] = q.w;
      }
    }
    if (L.active) L.publish(pitch, fld, 0, v);
    ++L.npub;
    const int Tst = CKPT ? a.Tstride : a.T;
    const int toff = CKPT ? a.t_off : 0;
    const float* xb = a.x + (size_t)b * Tst + toff;
    for (int i = tid; i < TB && i < a.T; i += NT) xs[i] = xb[i];
    __syncthreads();

    float4* tape = a.tape ? a.tape + (((size_t)b * a.T) * a.C + L.rank) * 2 * R * NT + tid : nullptr;
    float* fout = FIELDS ? a.fields + ((size_t)b * (a.T / a.field_every)) * plane + (size_t)L.gi0 * a.Ny + L.j0 : nullptr;

    auto flush = [&](int blk) {
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      const float* src = ps + (blk & 1) * TB;      // ps: [n_prb][2*TB], a ring of 2*TB samples per probe
      for (int i = tid; i < n * a.n_prb; i += NT) {
        int p = i % a.n_prb;
        if (poff[p] >= 0) {
          float val = src[p * (2 * TB) + i / a.n_prb];
          size_t o = ((size_t)b * Tst + toff + t0 + i / a.n_prb) * a.n_prb + p;
          if (a.probe_raw) a.probe_raw[o] = val;
          if (a.probe_out) a.probe_out[o] = a.prb_sq[p] ? val * val : val;
        }
      }
    };
    auto step = [&](float (&cu)[R][4], float (&pr)[R][4], int t, int blk, int tt) {
      const float* cur = fld + (t & 1) * L.slab;
      L.acquire_ghosts();
      if (t > 0 && my_poff >= 0) ps[plane_lane * (2 * TB) + ((t - 1) & (2 * TB - 1))] = cur[my_poff];
      if (L.active) {
        float lap[R][4];
        patch_laplacian<R>(pitch, cur + own, cu, lap);
        if (tape) {   // the adjoint needs u_{t-1} itself (coefficients) and L(u_{t-1})
#pragma unroll
          for (int r = 0; r < R; ++r) {
            st_stream(tape + (size_t)r * NT, make_float4(cu[r][0], cu[r][1], cu[r][2], cu[r][3]));
            st_stream(tape + (size_t)(R + r) * NT, make_float4(lap[r][0], lap[r][1], lap[r][2], lap[r][3]));
          }
          tape += tape_step;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float u = cu[r][k];
            float q = ce[r][k];              // !SAT: 1/(1 + dt*b_pml)
            if (SAT) {
              const float rr = u * s.inv_uth;
              const float d = fmaf(rr, rr, 1.f);
              q = d * rcp_fast(fmaf(d, ce[r][k], cf[r][k]));
            }
            float kc2 = cl[r][k];            // !KERR: kappa*c_lin^2
            if (KERR) {
              const float cc = fmaf(cg[r][k], u * u, cl[r][k]);
              kc2 = s.kappa * (cc * cc);
            }
            pr[r][k] = wt_update(q + q, q * kc2, u, pr[r][k], lap[r][k]);
          }
        if (m1) patch_inject<R>(pr, m1, m2, 0u, xs[(blk & 1) * TB + tt]);
        L.publish(pitch, fld, (t + 1) & 1, pr);
        if (FIELDS && (t + 1) % a.field_every == 0) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (L.gi0 + r < a.Nx) {
              float* f = fout + (size_t)r * a.Ny;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (L.j0 + k < a.Ny) f[k] = pr[r][k];
            }
          }
          fout += plane;
        }
      }
      ++L.npub;
      __syncthreads();
    };

    const int nblk = (a.T + TB - 1) / TB;
    for (int blk = 0; blk < nblk; ++blk) {
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      if ((blk + 1) * TB < a.T) {
        float* dst = xs + ((blk + 1) & 1) * TB;
        const int t1 = (blk + 1) * TB;
        for (int i = tid; i < TB && t1 + i < a.T; i += NT) dst[i] = xb[t1 + i];
      }
      if (blk >= 2) flush(blk - 2);
      if (CKPT && a.snap_every && t0 > 0 && (toff + t0) % a.snap_every == 0) {   // v = u_{t-1}, w = u_{t-2}: blocks are even
        float4* sp = a.snap + ((((size_t)((toff + t0) / a.snap_every - 1) * a.B + b) * a.C + L.rank) * 2 * R) * NT + tid;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          sp[r * NT] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
          sp[(R + r) * NT] = make_float4(w[r][0], w[r][1], w[r][2], w[r][3]);
        }
      }
      int tt = 0;
      for (; tt + 1 < n; tt += 2) {
        step(v, w, t0 + tt, blk, tt);
        step(w, v, t0 + tt + 1, blk, tt + 1);
      }
      if (tt < n) {
        step(v, w, t0 + tt, blk, tt);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) { float tmp = v[r][k]; v[r][k] = w[r][k]; w[r][k] = tmp; }
      }
    }
    L.acquire_ghosts();
    if (my_poff >= 0) ps[plane_lane * (2 * TB) + ((a.T - 1) & (2 * TB - 1))] = fld[(a.T & 1) * L.slab + my_poff];
    __syncthreads();
    for (int blk = max(0, nblk - 2); blk < nblk; ++blk) flush(blk);
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int gi = L.gi0 + r, j = L.j0 + k;
        if (L.active && gi < a.Nx && j < a.Ny && (!CKPT || a.u1)) {
          size_t o = ((size_t)b * a.Nx + gi) * a.Ny + j;
          a.u1[o] = v[r][k];
          a.u2[o] = w[r][k];
        }
      }
    __syncthreads();
  }
  if (a.C > 1) cg::this_cluster().sync();
}

// =================================================================================================
// adjoint
// =================================================================================================
// PITCH / NTC: row pitch and threads per CTA as compile-time constants (0 = from the launch), see wt_resident.cu
// GRADX = 0: dLoss/dx code compiled out (shape-specialised instances only); 1: decided at run time
// CHAIN: checkpoint-and-recompute instantiation: reverse steps of the segment [t_off, t_off + T); the carried pair
// (lambda_{t-1} without its seeds, carry into lambda_{t-2}) passes from launch to launch through a.chain, u_{t-2} of the
// segment's first step comes from the snapshot the segment started from (a.snap_in), gradient partials accumulate.
template <int R, bool SAT, bool KERR, int PITCH = 0, int NTC = 0, int GRADX = 1, bool CHAIN = false>
__global__ void __launch_bounds__(NTC ? NTC : res_nl_max_threads<R>()) k_res_adj_nl(ResArgs a) {
  const int NT = NTC ? NTC : blockDim.x;
  const int pitch = PITCH ? PITCH : a.pitch;
  const int slab_f = slab_words(R, a.Hc, pitch);
  const int RG = a.ring;                        // 2 or 4 (resident_plan)
  const unsigned rg_mask = (unsigned)RG - 1u, rg_shift = RG == 4 ? 2u : 1u;
  const int stage_f4 = 2 * R * NT;
  const unsigned stage_bytes = (unsigned)(stage_f4 * sizeof(float4));

  extern __shared__ float4 smem4[];
  float4* ring = smem4;                                          // [RG][2*R*NT]: u_{t-1} rows, then L(u_{t-1}) rows
  float* fld = reinterpret_cast<float*>(ring + RG * stage_f4);    // [2][slab]   P = kappa*c^2*q*lambda
  float* ss = fld + 2 * slab_f;
  float* gxs = ss + 2 * TB * a.n_prb;
  int* pown = reinterpret_cast<int*>(gxs + 2 * TB);
  int* pcell = pown + a.n_prb;
  uint64_t* full = reinterpret_cast<uint64_t*>(pcell + a.n_prb);  // [RING] (max depth allocated)
  uint64_t* bars = full + RING;

  Lane<R> L;
  L.init(a, fld, bars);
  const int tid = L.tid;
  float ce[R][4], cf[R][4], cl[R][4], cg[R][4];
  load_nl_consts<R>(a, L.active, L.gi0, L.j0, ce, cf, cl, cg);
  unsigned m1, m2;
  source_masks<R>(a, L.active, L.gi0, L.j0, m1, m2);
  for (int p = tid; p < a.n_prb; p += NT) {
    int li = a.prb_ij[2 * p] - L.rank * a.Hc, pj = a.prb_ij[2 * p + 1];
    bool mine = li >= 0 && li < a.Hc;
    pown[p] = mine ? (li / R) * a.P4 + pj / 4 : -1;
    pcell[p] = mine ? (li % R) * 4 + (pj & 3) : 0;
  }
  for (int i = tid; i < 2 * slab_f; i += NT) fld[i] = 0.f;
  for (int i = tid; i < 2 * TB; i += NT) gxs[i] = 0.f;
  if (tid == 0) {
    for (int q = 0; q < RING; ++q) mbar_init(full + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (a.C > 1) cg::this_cluster().sync(); else __syncthreads();
  int pc0 = -1, pi0 = 0;
  bool more_probes = false;
  for (int p = 0; p < a.n_prb; ++p)
    if (pown[p] == L.lt) {
      if (pc0 < 0) { pc0 = pcell[p]; pi0 = p; } else more_probes = true;
    }
  const int own = (L.lr0 + 1) * pitch + (L.run + 1) * slab_skew(R, pitch) + L.g;
  const Scalars s = a.s;
  const float ndtb0 = -s.dt * s.b0, two_iu2 = 2.f * s.inv_uth * s.inv_uth;

  float Gc[R][4], Gr[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) { Gc[r][k] = 0.f; Gr[r][k] = 0.f; }

  unsigned it_global = 0;
  for (int b = L.cid; b < a.B; b += a.n_clusters) {
    float lam[R][4], c2[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) { lam[r][k] = 0.f; c2[r][k] = 0.f; }
    if (CHAIN && a.chain_in) {
      const float4* cp = a.chain + ((size_t)b * a.C + L.rank) * 2 * R * NT + tid;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p = cp[r * NT], q = cp[(R + r) * NT];
        lam[r][0] = p.x; lam[r][1] = p.y; lam[r][2] = p.z; lam[r][3] = p.w;
        c2[r][0] = q.x; c2[r][1] = q.y; c2[r][2] = q.z; c2[r][3] = q.w;
      }
    }
    const int Tst = CHAIN ? a.Tstride : a.T;
    const int toff = CHAIN ? a.t_off : 0;

    auto tape_ptr = [&](int t) { return a.tape + ((((size_t)b * a.T + t) * a.C + L.rank) * 2 * R) * NT; };
    auto stage_seeds = [&](int blk) {
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      float* dst = ss + (blk & 1) * TB * a.n_prb;
      for (int i = tid; i < n * a.n_prb; i += NT) {
        int p = i % a.n_prb;
        size_t o = ((size_t)b * Tst + toff + t0 + i / a.n_prb) * a.n_prb + p;
        float g = a.grad_probe[o];
        if (a.prb_sq[p]) g *= 2.f * a.probe_raw[o];
        dst[i] = g;
      }
    };
    auto flush_gx = [&](int blk) {
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      float* src = gxs + (blk & 1) * TB;
      for (int i = tid; i < n; i += NT) {
        float sv = src[i];
        src[i] = 0.f;
        if (sv != 0.f) atomicAdd(a.grad_x + (size_t)b * Tst + toff + t0 + i, sv);
      }
    };
    if (tid == 0) {
      for (int q = 0; q < RG && q < a.T; ++q) {
        unsigned slot = (it_global + q) & rg_mask;
        mbar_expect_tx(full + slot, stage_bytes);
        bulk_g2s(ring + slot * stage_f4, tape_ptr(a.T - 1 - q), stage_bytes, full + slot);
      }
    }
    stage_seeds((a.T - 1) / TB);
    if ((a.T - 1) / TB > 0) stage_seeds((a.T - 1) / TB - 1);
    __syncthreads();

    // One reverse step; PAR = it & 1 is a compile-time constant of each of the two unrolled copies.
    auto step = [&](auto par, int t, int it) {
      constexpr int PAR = decltype(par)::value;
      const int blk = t / TB, tt = t - blk * TB;
      float* cur = fld + PAR * L.slab;
      const unsigned gi = it_global + it;
      const unsigned slot = gi & rg_mask, parity = (gi >> rg_shift) & 1u;
      const unsigned slot2 = (gi + 1) & rg_mask, parity2 = ((gi + 1) >> rg_shift) & 1u;   // stage of step t-1: its u is my u_{t-2}
      float pv[R][4];   // across the barrier: P (stencil centre); lam holds the old carry + own-cell part, c2 the new carry
      if (L.active) {
        if (pc0 >= 0) {   // lambda_t += dLoss/du_t through the probes (first probe of my patch: fast path)
          const float* srow = ss + (blk & 1) * TB * a.n_prb + tt * a.n_prb;
          patch_add_cell<R>(lam, pc0, srow[pi0]);
          if (more_probes) {
            for (int p = pi0 + 1; p < a.n_prb; ++p)
              if (pown[p] == L.lt) patch_add_cell<R>(lam, pcell[p], srow[p]);
          }
        }
        if (GRADX && a.grad_x && m1) {
          float sv = 0.f;
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (m1 >> (r * 4 + k) & 1u) sv += lam[r][k];
              if (m2 >> (r * 4 + k) & 1u) sv += lam[r][k];
            }
          atomicAdd(gxs + (blk & 1) * TB + tt, sv);
        }
        mbar_wait(full + slot, parity);
        if (t > 0) mbar_wait(full + slot2, parity2);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 u1v = ring[slot * stage_f4 + r * NT + tid];
          const float4 lpv = ring[slot * stage_f4 + (R + r) * NT + tid];
          float4 u2v = make_float4(0.f, 0.f, 0.f, 0.f);      // t = 0: u_{-2} is the zero initial field (WT_F_ZERO_INIT)
          if (t > 0) u2v = ring[slot2 * stage_f4 + r * NT + tid];
          else if (CHAIN && a.snap_in) u2v = a.snap_in[(((size_t)b * a.C + L.rank) * 2 * R + R + r) * NT + tid];   // ... or the snapshot's u_{t-2}
          const float u1a[4] = {u1v.x, u1v.y, u1v.z, u1v.w};
          const float lpa[4] = {lpv.x, lpv.y, lpv.z, lpv.w};
          const float u2a[4] = {u2v.x, u2v.y, u2v.z, u2v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float u1 = u1a[k], u2 = u2a[k];
            // coefficients at u_{t-1}: q = 1/(1+dt*b), c, and rd = 1/d of the saturable term
            float rd = 1.f, den = ce[r][k], cc = cl[r][k];
            if (SAT) {
              const float rr = u1 * s.inv_uth;
              rd = rcp_fast(fmaf(rr, rr, 1.f));
              den = fmaf(cf[r][k], rd, ce[r][k]);               // 1 + dt*b(u)
            }
            const float uu = u1 * u1;
            if (KERR) cc = fmaf(cg[r][k], uu, cl[r][k]);
            const float q = rcp_fast(den);
            const float ql = q * lam[r][k];
            const float kl = s.kappa * lpa[k];
            const float cc2 = cc * cc;
            const float S = fmaf(cc2, kl, 2.f * (u1 - u2));
            const float g_c = (cc + cc) * (kl * ql);            // cell.py:36
            float gu1 = ql + ql;                                // own-cell part of cell.py:39-40
            if (SAT) {   // dLoss/db = -dt*q*S*ql (cell.py:33-34) through b = b_pml + rho*b0/d
              const float X = (q * S) * (ql * rd);              // -dLoss/db / (dt*d)
              Gr[r][k] = fmaf(X, ndtb0, Gr[r][k]);              // d b/d rho = b0/d
              gu1 = fmaf((X * rd) * cf[r][k], two_iu2 * u1, gu1);   // d b/d u = -2*rho*b0*u/(uth^2 d^2)
            }
            if (KERR) {
              Gr[r][k] = fmaf(g_c, s.c_nl * uu, Gr[r][k]);
              gu1 = fmaf(g_c, (cg[r][k] + cg[r][k]) * u1, gu1);
            }
            Gc[r][k] += g_c;
            pv[r][k] = (s.kappa * cc2) * ql;
            lam[r][k] = c2[r][k] + gu1;                       // everything of lambda_{t-1} except the stencil term
            c2[r][k] = (den - 2.f) * ql;                        // new carry (beta-1)*q*lambda, cell.py:42
          }
        }
        L.publish(pitch, fld, PAR, pv);
      }
      ++L.npub;
      __syncthreads();
      if (tid == 0 && it + RG < a.T) {
        mbar_expect_tx(full + slot, stage_bytes);
        bulk_g2s(ring + slot * stage_f4, tape_ptr(t - RG), stage_bytes, full + slot);
      }
      L.acquire_ghosts();
      if (L.active) {
        float lapP[R][4];
        patch_laplacian<R>(pitch, cur + own, pv, lapP);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) lam[r][k] += lapP[r][k];
      }
    };
    {
      using P0 = std::integral_constant<int, 0>;
      using P1 = std::integral_constant<int, 1>;
      int t = a.T - 1, it = 0;
      for (; t >= 1; t -= 2, it += 2) {
        step(P0{}, t, it);
        step(P1{}, t - 1, it + 1);
        // Per-block staging outside the step bodies.  A block's seeds are read from its first (highest) step on, so the
        // seeds of the NEXT block are staged after the pair that contains that first step: the half they overwrite
        // belonged to the block that has just ended.
        const int tb = ((t & (TB - 1)) == TB - 1) ? t : ((((t - 1) & (TB - 1)) == TB - 1) ? t - 1 : -1);
        if (tb >= TB && tb != a.T - 1) stage_seeds(tb / TB - 1);
        if (GRADX && a.grad_x) {
          if ((t & (TB - 1)) == 0 && t >= TB) flush_gx(t / TB);
          if (((t - 1) & (TB - 1)) == 0 && t - 1 >= TB) flush_gx((t - 1) / TB);
        }
      }
      if (t == 0) step(P0{}, 0, it);
    }
    it_global += (unsigned)a.T;
    if (CHAIN && a.chain_out) {
      float4* cp = a.chain + ((size_t)b * a.C + L.rank) * 2 * R * NT + tid;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        cp[r * NT] = make_float4(lam[r][0], lam[r][1], lam[r][2], lam[r][3]);
        cp[(R + r) * NT] = make_float4(c2[r][0], c2[r][1], c2[r][2], c2[r][3]);
      }
    }
    __syncthreads();
    if (GRADX && a.grad_x) flush_gx(0);
    __syncthreads();
  }
  const size_t plane = (size_t)a.Nx * a.Ny;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int gi = L.gi0 + r, j = L.j0 + k;
      if (L.active && gi < a.Nx && j < a.Ny) {
        size_t o = ((size_t)L.cid * 2) * plane + (size_t)gi * a.Ny + j;
        const bool acc = CHAIN && a.accumulate;
        a.Gpart[o] = acc ? a.Gpart[o] + Gc[r][k] : Gc[r][k];
        a.Gpart[o + plane] = acc ? a.Gpart[o + plane] + Gr[r][k] : Gr[r][k];
      }
    }
  if (a.C > 1) cg::this_cluster().sync();
}

// =================================================================================================
// host side
// =================================================================================================
int res_nl_max_threads_rt(int R) {
  switch (R) {
    case 1: return 512;
    case 2: return 512;
    case 3: return 384;
    case 4: return 384;
    default: return 0;
  }
}

size_t res_nl_smem_fwd(int Hc, int pitch, int n_prb, int R) {
  return (size_t)2 * slab_words(R, Hc, pitch) * 4 + 2 * TB * 4 + (size_t)2 * TB * n_prb * 4 + (size_t)(n_prb + 1) * 4 + 4 * 8 + 16;
}
size_t res_nl_smem_adj(int Hc, int pitch, int n_prb, int R, int threads, int ring) {
  return (size_t)ring * 2 * R * threads * 16 + (size_t)2 * slab_words(R, Hc, pitch) * 4 + (size_t)2 * TB * n_prb * 4 + 2 * TB * 4 +
         (size_t)2 * n_prb * 4 + 8 + RING * 8 + 4 * 8 + 16;
}

template <typename K>
static int nl_active_clusters(K kernel, int C, int threads, size_t smem) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (C > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C * 1024);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

#define WT_NL_CASE(R_, NL_, CALLF)                                                          \
  if (R == R_ && nl == NL_) { constexpr int RR = R_; constexpr bool SAT = (NL_ & 1) != 0, KERR = (NL_ & 2) != 0; CALLF; }
#define WT_NL_ALL(CALLF)                                                                    \
  WT_NL_CASE(1, 1, CALLF) WT_NL_CASE(1, 2, CALLF) WT_NL_CASE(1, 3, CALLF)                  \
  WT_NL_CASE(2, 1, CALLF) WT_NL_CASE(2, 2, CALLF) WT_NL_CASE(2, 3, CALLF)                  \
  WT_NL_CASE(3, 1, CALLF) WT_NL_CASE(3, 2, CALLF) WT_NL_CASE(3, 3, CALLF)                  \
  WT_NL_CASE(4, 1, CALLF) WT_NL_CASE(4, 2, CALLF) WT_NL_CASE(4, 3, CALLF)

int res_nl_clusters(int R, int nl, int C, int threads, size_t smem_fwd, size_t smem_bwd) {
  int nf = 0, nb = 0;
  WT_NL_ALL((nf = nl_active_clusters(k_res_fwd_nl<RR, SAT, KERR>, C, threads, smem_fwd),
             nb = nl_active_clusters(k_res_adj_nl<RR, SAT, KERR>, C, threads, smem_bwd)))
  return nf < nb ? nf : nb;
}

template <typename K>
static int nl_launch(K kernel, const wt_plan& plan, size_t smem, const ResArgs& a, cudaStream_t st) {
  WT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (plan.cluster > 8) WT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.n_clusters * plan.cluster);
  cfg.blockDim = dim3(plan.threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  WT_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  return WT_OK;
}

int res_nl_launch_fwd(const wt_plan& plan, const ResArgs& a, cudaStream_t st, bool ckpt) {
  const int R = plan.rows_per_thread, nl = plan.nonlinear;
  int rc = WT_EINVAL;
  if (ckpt) {   // checkpoint-and-recompute instantiations (generic shapes)
    WT_NL_ALL((rc = nl_launch(k_res_fwd_nl<RR, SAT, KERR, false, 0, 0, true>, plan, plan.smem_fwd, a, st)))
    if (rc == WT_EINVAL) set_error("nonlinear on-chip kernel R=%d nl=%d not instantiated", R, nl);
    return rc;
  }
  const bool spec = R == 2 && a.pitch == 104 && plan.threads == 480 && !(a.flags & WT_F_NO_SPECIALIZE);   // BASELINE config 4
  if (a.fields) {   // output_fields=True: separate instantiation, keeps the field stores out of the common step body
    WT_NL_ALL((rc = nl_launch(k_res_fwd_nl<RR, SAT, KERR, true>, plan, plan.smem_fwd, a, st)))
  } else if (spec) {
    if (nl == 1) rc = nl_launch(k_res_fwd_nl<2, true, false, false, 104, 480>, plan, plan.smem_fwd, a, st);
    if (nl == 2) rc = nl_launch(k_res_fwd_nl<2, false, true, false, 104, 480>, plan, plan.smem_fwd, a, st);
    if (nl == 3) rc = nl_launch(k_res_fwd_nl<2, true, true, false, 104, 480>, plan, plan.smem_fwd, a, st);
  } else {
    WT_NL_ALL((rc = nl_launch(k_res_fwd_nl<RR, SAT, KERR>, plan, plan.smem_fwd, a, st)))
  }
  if (rc == WT_EINVAL) set_error("nonlinear on-chip kernel R=%d nl=%d not instantiated", R, nl);
  return rc;
}

int res_nl_launch_adj(const wt_plan& plan, const ResArgs& a, cudaStream_t st, bool chain) {
  const int R = plan.rows_per_thread, nl = plan.nonlinear;
  int rc = WT_EINVAL;
  if (chain) {
    WT_NL_ALL((rc = nl_launch(k_res_adj_nl<RR, SAT, KERR, 0, 0, 1, true>, plan, plan.smem_bwd, a, st)))
    if (rc == WT_EINVAL) set_error("nonlinear on-chip kernel R=%d nl=%d not instantiated", R, nl);
    return rc;
  }
  if (R == 2 && a.pitch == 104 && plan.threads == 480 && !(a.flags & WT_F_NO_SPECIALIZE)) {   // BASELINE config 4
    if (a.grad_x) {
      if (nl == 1) rc = nl_launch(k_res_adj_nl<2, true, false, 104, 480, 1>, plan, plan.smem_bwd, a, st);
      if (nl == 2) rc = nl_launch(k_res_adj_nl<2, false, true, 104, 480, 1>, plan, plan.smem_bwd, a, st);
      if (nl == 3) rc = nl_launch(k_res_adj_nl<2, true, true, 104, 480, 1>, plan, plan.smem_bwd, a, st);
    } else {
      if (nl == 1) rc = nl_launch(k_res_adj_nl<2, true, false, 104, 480, 0>, plan, plan.smem_bwd, a, st);
      if (nl == 2) rc = nl_launch(k_res_adj_nl<2, false, true, 104, 480, 0>, plan, plan.smem_bwd, a, st);
      if (nl == 3) rc = nl_launch(k_res_adj_nl<2, true, true, 104, 480, 0>, plan, plan.smem_bwd, a, st);
    }
  } else {
    WT_NL_ALL((rc = nl_launch(k_res_adj_nl<RR, SAT, KERR>, plan, plan.smem_bwd, a, st)))
  }
  if (rc == WT_EINVAL) set_error("nonlinear on-chip kernel R=%d nl=%d not instantiated", R, nl);
  return rc;
}

}  // namespace wt
