// On-chip ("resident") kernels: the whole time loop of a sample runs in ONE launch.
//
// A sample's [Nx,Ny] field is split by rows over a thread-block cluster of C CTAs.  Inside a CTA each thread
// owns a patch of R rows x 4 columns and keeps, in registers for the whole loop, the two time levels of its
// cells (u_t, u_{t-1}) and their coefficients.  Shared memory only carries what neighbours need: the current
// field of the slab (double buffered, with one ghost row per side).  Ghost rows are written straight into
// the neighbouring CTA's shared memory (st.async over DSMEM, completion counted on an mbarrier there, see
// wt_resident_dev.cuh); inside a CTA there is one __syncthreads() per step and no cluster-wide barrier.
// HBM is touched only for x[b,t], the probe samples, and -- when a gradient is wanted -- the adjoint tape.
//
// Forward  : u_{t}   = u_{t-2} + a1*(u_{t-1}-u_{t-2}) + a3*L(u_{t-1}) ; += x[b,t] at sources ; probes read
// Adjoint  : lam_{t-1} = a1*lam_t + L(a3*lam_t) + (1-a1)*lam_{t+1} + seed_{t-1};  G += L(u_{t-1})*lam_t, carried as
//            P = a3*lam, which obeys the forward update (see k_res_adj)
//            (the tape holds L(u_{t-1}) per step in the thread-major order the adjoint reads it back in,
//             fetched by cp.async.bulk into a shared-memory ring ahead of use)
// The step bodies are kept small on purpose: they are fetched 2T times per sample and everything inlined into them --
// also code that a branch skips -- costs instruction-cache bandwidth.
// Each kernel carries TWO instantiations of its step: the general one, and a PLAIN one for warps in which no lane waits for
// or pushes ghost rows, owns a source or a probe, refills the tape ring or is inactive -- most warps of a CTA.  The choice is
// made once per sample by a warp-uniform branch around the whole time loop; all warps meet at the same barrier 0 every step.
// Slab buffers: plane layout with a per-run skew, see wt_resident_dev.cuh (every access is one conflict-free wavefront).
//
// Reference semantics: wavetorch/rnn.py:36-70, cell.py:12-17, cell.py:27-44, operators.py:5-11,
// source.py:15-22, probe.py:14-27.
#include <mutex>
#include <type_traits>
#include <vector>

#include "wt_resident.h"
#include "wt_resident_dev.cuh"
#include "wt_stream.h"

namespace wt {

// =================================================================================================
// forward
// =================================================================================================
// PITCH / NTC: shared-memory row pitch and threads per CTA as compile-time constants (0 = take them from the launch).  The
// BASELINE config-3 shape is instantiated with constants: every shared-memory and tape offset of the step becomes an immediate.
// FIELDS: also write every field to HBM (output_fields=True); a separate instantiation keeps that code out of the step body
// of the common kernels, which is fetched from the instruction cache 2T times per sample.
// CKPT: the checkpoint-and-recompute instantiation -- the launch covers steps [t_off, t_off + T) of longer sequences, can
// start from a stored snapshot of the register patches and stores snapshots every snap_every steps (a multiple of TB) on
// its way; the common kernels carry none of that code.
template <int R, bool TAPE, int PITCH = 0, int NTC = 0, bool FIELDS = false, bool CKPT = false, bool NOPLAIN = false, bool TMAST = false>
__global__ void __launch_bounds__(NTC ? NTC : res_max_threads<R>(), res_min_blocks<R>(NTC)) k_res_fwd(ResArgs a) {
  extern __shared__ float4 smem4[];
#ifdef WT_DEBUG_CLOCK
  __shared__ long long wt_dbg[16][2][8];
#endif
  const int pitch = PITCH ? PITCH : a.pitch;
  const int slab_f = slab_words(R, a.Hc, pitch);
  float* fld = reinterpret_cast<float*>(smem4);       // [2][slab]
  float* xs = fld + 2 * slab_f;                        // [2][TB]
  float* ps = xs + 2 * TB;                             // [n_prb][2*TB]: a ring of 2*TB samples per probe
  int* poff = reinterpret_cast<int*>(ps + 2 * TB * a.n_prb);  // [n_prb] offset into a slab buffer, or -1
  uint64_t* bars = reinterpret_cast<uint64_t*>(poff + a.n_prb + (a.n_prb & 1));
  // TMAST: the tape rows of a step are staged in shared memory (three buffers) and leave through ONE TMA bulk store per step
  float4* tst = TMAST ? reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(bars + 4) + 127) & ~(uintptr_t)127) : nullptr;

  Lane<R> L;
  L.init(a, fld, bars);
  const int tid = L.tid, NT = NTC ? NTC : blockDim.x;
  const int store_tid = ((NT / 32) / 2) * 32;   // TMAST: the lane that issues the bulk stores (its warp runs the general step)
  float k1[R][4], k3[R][4];
  load_coef<R>(a, L.active, L.gi0, L.j0, k1, k3);
  unsigned m1, m2;
  source_masks<R>(a, L.active, L.gi0, L.j0, m1, m2);
  const bool src_warp = __any_sync(0xffffffffu, m1 != 0u), src2_warp = __any_sync(0xffffffffu, m2 != 0u);
  for (int p = tid; p < a.n_prb; p += NT) {
    int li = a.prb_ij[2 * p] - L.rank * a.Hc, pj = a.prb_ij[2 * p + 1];
    poff[p] = (li >= 0 && li < a.Hc) ? slab_cell(R, pitch, li, pj) : -1;
  }
  L.pub_all = L.active && probe_in_interior<R>(a, L.rank, L.lt);
  for (int i = tid; i < 2 * slab_f; i += NT) fld[i] = 0.f;
  if (a.C > 1) cg::this_cluster().sync(); else __syncthreads();
  const int plane_lane = NT - 1 - tid;   // probe p is sampled by lane NT-1-p: the highest warp has issue priority
  const int my_poff = (plane_lane < a.n_prb) ? poff[plane_lane] : -1;
  // a warp without special duties: all lanes own cells, none borders another CTA, owns a source, samples a probe or has to
  // publish interior cells for a probe lane; and the FIELDS / CKPT instantiations keep to the general step
  const bool plain_warp = !FIELDS && !NOPLAIN && !__any_sync(0xffffffffu, !L.active || L.edge_up || L.edge_dn || L.pub_all || m1 != 0u || my_poff >= 0 || (TMAST && tid == store_tid));
  const int own = (L.lr0 + 1) * pitch + (L.run + 1) * slab_skew(R, pitch) + L.g;       // my first row inside a slab buffer
  const size_t tape_step = (size_t)a.C * R * NT;        // float4 per time step of one sample
  const size_t plane = (size_t)a.Nx * a.Ny;

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
    const int Tst = CKPT ? a.Tstride : a.T;         // length of the sequences x / probe_out are laid out for
    const int toff = CKPT ? a.t_off : 0;
    const float* xb = a.x + (size_t)b * Tst + toff;
    for (int i = tid; i < TB && i < a.T; i += NT) xs[i] = xb[i];
    __syncthreads();

    float4* tape = TAPE ? a.tape + (((size_t)b * a.T) * a.C + L.rank) * R * NT + tid : nullptr;
    float* fout = FIELDS ? a.fields + ((size_t)b * (a.T / a.field_every)) * plane + (size_t)L.gi0 * a.Ny + L.j0 : nullptr;

    auto flush = [&](int blk) {   // probe samples of time block blk -> HBM
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      const float* src = ps + (blk & 1) * TB;
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
    // One time step: `cu` = u_t (kept), `pr` = u_{t-1} on entry and u_{t+1} on exit.  t is the index of the new field.
    // PAR = t & 1 is a compile-time constant of each of the two unrolled copies, so every shared-memory address below
    // is loop invariant.  x and the probe samples live in rings of 2*TB steps.
    const float* rd0 = fld + own;                 // my patch in slab buffer 0 / 1
    const float* rd1 = fld + L.slab + own;
    float* psw = ps + plane_lane * (2 * TB);      // sample ring of this lane's probe (the last n_prb lanes only)
    int tslot = 0;                                // TMAST: staging buffer of the current step
    // PLAIN: the instantiation for warps without special duties (plain_warp below) carries none of the flag tests and
    // branch regions of the general step -- ghost-row waits and pushes, probe sampling, source injection, inactive lanes.
    // Those cost a warp ~30 instructions and six divergence regions per step even when every one of them is skipped, and
    // the step is paced by the sum of what the three warps of a scheduler issue.
    auto step = [&](auto par, auto plain_t, float (&cu)[R][4], float (&pr)[R][4], int t) {
      constexpr int PAR = decltype(par)::value;
      constexpr bool PLAIN = decltype(plain_t)::value;
      const float* cur = PAR ? rd1 : rd0;
#ifdef WT_DEBUG_CLOCK
      long long c0 = clock64(), c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#endif
      if (!PLAIN) {
        L.acquire_ghosts();
#ifdef WT_DEBUG_CLOCK
        c1 = clock64();
#endif
      }
#ifdef WT_DEBUG_CLOCK
      c2 = clock64();
#endif
      if (PLAIN || L.active) {
        float xv = 0.f;
        if (!PLAIN) xv = src_warp ? xs[t & (2 * TB - 1)] : 0.f;   // fetched ahead of the stencil: off the source warp's path
        float lap[R][4];
        patch_laplacian<R>(pitch, cur, cu, lap);
        if (TAPE && TMAST) {
          float4* dst = tst + tslot * (R * NT) + tid;
#pragma unroll
          for (int r = 0; r < R; ++r) dst[r * NT] = make_float4(lap[r][0], lap[r][1], lap[r][2], lap[r][3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my rows, visible to the TMA engine after the barrier
        }
        if (TAPE && !TMAST && PLAIN) {
          // The tape rows go out BEFORE the update and the publish.  The LSU queue is in order: issued last, the 128-bit
          // tape stores of the warps that finish early sit in front of the rim stores and ghost-row pushes of the warps
          // that finish late -- the ones the step barrier waits for (measured: 700-1200 cycles for an edge warp's publish).
#pragma unroll
          for (int r = 0; r < R; ++r) st_stream(tape + (size_t)r * NT, make_float4(lap[r][0], lap[r][1], lap[r][2], lap[r][3]));
          tape += tape_step;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) pr[r][k] = wt_update(k1[r][k], k3[r][k], cu[r][k], pr[r][k], lap[r][k]);
#ifdef WT_DEBUG_CLOCK
        c3 = clock64();
#endif
        if (!PLAIN && src_warp) {   // source.py:19-22 (dt = 1.0 there): every listed pixel receives x[b,t], once per listing
          patch_inject_pred<R>(pr, m1, xv);
          if (src2_warp) patch_inject_pred<R>(pr, m2, xv);
        }
#ifdef WT_DEBUG_CLOCK
        c4 = clock64();
#endif
        L.template publish<PLAIN>(pitch, fld, PAR ^ 1, pr);
#ifdef WT_DEBUG_CLOCK
        c5 = clock64();
#endif
        if (TAPE && !TMAST && !PLAIN) {   // a warp with special duties: its rim stores and pushes are what others wait for, the tape comes last
#pragma unroll
          for (int r = 0; r < R; ++r) st_stream(tape + (size_t)r * NT, make_float4(lap[r][0], lap[r][1], lap[r][2], lap[r][3]));
          tape += tape_step;
        }
        if (FIELDS && (t + 1) % a.field_every == 0) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (L.gi0 + r < a.Nx) {
              float* f = fout + (size_t)r * a.Ny;
              if (a.vec_fields) {
                *reinterpret_cast<float4*>(f) = make_float4(pr[r][0], pr[r][1], pr[r][2], pr[r][3]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (L.j0 + k < a.Ny) f[k] = pr[r][k];
              }
            }
          }
          fout += plane;
        }
      }
      if (!PLAIN) {
        ++L.npub;
        // The probe lanes sample the field of the previous step at the END of this step (its buffer is not written before
        // the barrier): at the top the load would queue behind the stencil loads of the whole CTA and the dependent store
        // would hold this warp back for ~400 cycles (measured with the WT_DEBUG_CLOCK build).
        if (my_poff >= 0 && t > 0) psw[(t - 1) & (2 * TB - 1)] = (PAR ? fld + L.slab : fld)[my_poff];
      }
#ifdef WT_DEBUG_CLOCK
      if ((tid & 31) == 0 && (int)blockIdx.x < 2 && b == (int)blockIdx.x / a.C && (t == 500 || t == 501)) {
        long long* d = wt_dbg[tid >> 5][t - 500];
        d[0] = c0; d[1] = c1; d[2] = c2; d[3] = c3; d[4] = c4; d[5] = c5; d[6] = clock64(); d[7] = PLAIN;
      }
#endif
      step_barrier(NT);
      if (TAPE && TMAST) {
        if (!PLAIN && tid == store_tid) {   // every row of this step is staged: one bulk store; the buffer written two steps
          // from now was read by the store before the previous one, which wait_group.read 1 has seen finish before the NEXT barrier
          const float4* gdst = a.tape + (((size_t)b * a.T + t) * a.C + L.rank) * R * NT;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(tst + tslot * (R * NT))),
                       "r"((unsigned)(R * NT * sizeof(float4)))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        tslot = tslot == 2 ? 0 : tslot + 1;
      }
    };
    using P0 = std::integral_constant<int, 0>;
    using P1 = std::integral_constant<int, 1>;

    const int nblk = (a.T + TB - 1) / TB;
    const unsigned npub0 = L.npub;
    auto run = [&](auto plain_t) {
      for (int blk = 0; blk < nblk; ++blk) {
        const int t0 = blk * TB, n = min(TB, a.T - t0);
        if ((blk + 1) * TB < a.T) {   // stage the next block of x
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
        for (; tt + 1 < n; tt += 2) {     // two steps per iteration: the two time levels swap roles, no moves
          step(P0{}, plain_t, v, w, t0 + tt);
          step(P1{}, plain_t, w, v, t0 + tt + 1);
        }
        if (tt < n) {                      // odd tail (last block only): keep "v = latest" by swapping once
          step(P0{}, plain_t, v, w, t0 + tt);
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) { float tmp = v[r][k]; v[r][k] = w[r][k]; w[r][k] = tmp; }
        }
      }
    };
    // Every warp executes the same sequence of __syncthreads(); only the instruction stream between them differs.
    if (plain_warp) run(std::true_type{}); else run(std::false_type{});
    L.npub = npub0 + (unsigned)a.T;
    if (TAPE && TMAST && tid == store_tid) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    L.acquire_ghosts();   // consume the last publish so that no st.async is in flight past this point
    if (my_poff >= 0) psw[(a.T - 1) & (2 * TB - 1)] = fld[(a.T & 1) * L.slab + my_poff];
    __syncthreads();
    for (int blk = max(0, nblk - 2); blk < nblk; ++blk) flush(blk);
    // final state back to HBM: u1 = latest field, u2 = the one before (cell.py:107)
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
#ifdef WT_DEBUG_CLOCK
  __syncthreads();
  if (tid == 0 && blockIdx.x < 2 && a.T > 502)
    for (int w = 0; w < NT / 32; ++w)
      for (int k = 0; k < 2; ++k) {
        long long* d = wt_dbg[w][k];
        printf("F cta=%d step=%d w=%d plain=%lld c0=%lld acq=%lld prb=%lld upd=%lld inj=%lld pub=%lld arrive=%lld\n", (int)blockIdx.x, k, w, d[7],
               d[0] - wt_dbg[0][0][0], d[1] ? d[1] - d[0] : 0, d[2] - d[0], d[3] - d[0], d[4] - d[0], d[5] - d[0], d[6] - d[0]);
      }
#endif
  if (a.C > 1) cg::this_cluster().sync();   // nobody exits while a neighbour could still address its shared memory
}

// =================================================================================================
// adjoint
// =================================================================================================
// Written in the variable P_t = a3 * lambda_t.  Multiplying the adjoint recursion
//     lambda_{t-1} = a1*lambda_t + L(a3*lambda_t) + (1-a1)*lambda_{t+1} + seed_{t-1}            (cell.py:39-42)
// by a3 (all factors are own-cell) gives
//     P_{t-1} = P_{t+1} + a1*(P_t - P_{t+1}) + a3*L(P_t) + a3*seed_{t-1},
// i.e. exactly the forward update (wt_update) run backwards in time, with the probe seeds playing the role of the
// sources.  So this kernel is the forward kernel plus the tape: sum_t L(u_{t-1})*P_t accumulates per cell and
// dLoss/dc = gscale * sum / a3 = (2/c) * sum  (cell.py:36).  dLoss/dx[b,t] = sum over source pixels of P_t/a3.
// GRADX = 0: dLoss/dx is not wanted, its code is compiled out (shape-specialised instances only); 1: decided at run time.
// RINGC: tape prefetch depth as a compile-time constant (0 = a.ring, a power of two chosen by resident_plan: small patches
// run a step in a few hundred nanoseconds and need a deeper ring to cover the HBM latency of the bulk copies).
// CHAIN: the checkpoint-and-recompute instantiation -- the launch covers the reverse steps of the segment [t_off, t_off+T)
// of longer sequences; the pair (P_{t-1}, P_t) at the segment boundary is handed from launch to launch through a.chain
// in the register layout (no lambda <-> P conversion, no division), and the per-cluster gradient partials accumulate.
template <int R, int PITCH = 0, int NTC = 0, int GRADX = 1, int RINGC = 0, bool CHAIN = false, bool NOPLAIN = false>
__global__ void __launch_bounds__(NTC ? NTC : res_max_threads<R>(), res_min_blocks<R>(NTC)) k_res_adj(ResArgs a) {
  constexpr bool EARLY = R <= 2;   // see the step body
  constexpr bool SEED_PRED = R >= 4;   // big patches: probe seeds fetched ahead and added with predicated FMAs (R = 5: adjoint -3 %; slower at R = 2)
  const int NT = NTC ? NTC : blockDim.x;
  const int RG = RINGC ? RINGC : a.ring;
  const int RG_LOG = RINGC ? (RINGC == 2 ? 1 : RINGC == 4 ? 2 : RINGC == 8 ? 3 : 4) : (31 - __clz(a.ring));
  const int pitch = PITCH ? PITCH : a.pitch;
  const int slab_f = slab_words(R, a.Hc, pitch);
  const unsigned stage_bytes = (unsigned)(R * NT * sizeof(float4));

  extern __shared__ float4 smem4[];
#ifdef WT_DEBUG_CLOCK
  __shared__ long long wt_dbg[16][2][8];
#endif
  float4* ring = smem4;                                        // [RG][R*NT] tape stages
  float* fld = reinterpret_cast<float*>(ring + RG * R * NT);    // [2][slab]   P
  float* ss = fld + 2 * slab_f;                                 // [2][TB][n_prb] probe seeds
  float* gxs = ss + 2 * TB * a.n_prb;                           // [2][TB]     dLoss/dx staging
  int* pown = reinterpret_cast<int*>(gxs + 2 * TB);             // [n_prb] owning thread, or -1
  int* pcell = pown + a.n_prb;                                  // [n_prb] cell index inside the owner's patch
  uint64_t* full = reinterpret_cast<uint64_t*>(pcell + a.n_prb);  // [MAX_RING]; 8-byte aligned (even word count before)
  uint64_t* bars = full + MAX_RING;                             // [4] ghost rows

  Lane<R> L;
  L.init(a, fld, bars);
  const int tid = L.tid;
  float k1[R][4], k3[R][4];
  load_coef<R>(a, L.active, L.gi0, L.j0, k1, k3);
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
    for (int s = 0; s < RG; ++s) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (a.C > 1) cg::this_cluster().sync(); else __syncthreads();
  int pc0 = -1, pi0 = 0;       // first probe inside my patch: cell index and probe index
  bool more_probes = false;
  for (int p = 0; p < a.n_prb; ++p)
    if (pown[p] == L.lt) {
      if (pc0 < 0) { pc0 = pcell[p]; pi0 = p; } else more_probes = true;
    }
  const int own = (L.lr0 + 1) * pitch + (L.run + 1) * slab_skew(R, pitch) + L.g;
  const size_t tape_step = (size_t)a.C * R * NT;
  // the lane that re-issues tape copies sits in a middle warp: the first and last warps already wait for ghost rows
  const int refill_tid = ((NT / 32) / 2) * 32;
  const bool seed_warp = __any_sync(0xffffffffu, pc0 >= 0);
  const bool plain_warp = !NOPLAIN && !__any_sync(0xffffffffu, !L.active || L.edge_up || L.edge_dn || pc0 >= 0 || tid == refill_tid ||
                                                       (GRADX && a.grad_x && m1 != 0u));

  float G[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) G[r][k] = 0.f;

  unsigned it_global = 0;   // tape stages consumed so far (ring slot / parity bookkeeping across samples)
  for (int b = L.cid; b < a.B; b += a.n_clusters) {
    const float4* tape_b = a.tape + (((size_t)b * a.T) * a.C + L.rank) * R * NT;   // stage of step t at + t*tape_step
    const int Tst = CHAIN ? a.Tstride : a.T;        // length of the sequences grad_probe / probe_raw / grad_x are laid out for
    const int toff = CHAIN ? a.t_off : 0;
    auto stage_seeds = [&](int blk) {   // seeds of time block blk: dLoss/d(raw probe value)
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      float* dst = ss + (blk & 1) * TB;               // ss: [n_prb][2*TB], a ring of 2*TB seeds per probe
      for (int i = tid; i < n * a.n_prb; i += NT) {
        int p = i % a.n_prb;
        size_t o = ((size_t)b * Tst + toff + t0 + i / a.n_prb) * a.n_prb + p;
        float g = a.grad_probe[o];
        if (a.prb_sq[p]) g *= 2.f * a.probe_raw[o];   // probe.py:27
        dst[p * (2 * TB) + i / a.n_prb] = g;
      }
    };
    auto flush_gx = [&](int blk) {
      const int t0 = blk * TB, n = min(TB, a.T - t0);
      float* src = gxs + (blk & 1) * TB;
      for (int i = tid; i < n; i += NT) {
        float s = src[i];
        src[i] = 0.f;
        if (s != 0.f) atomicAdd(a.grad_x + (size_t)b * Tst + toff + t0 + i, s);
      }
    };
    // P += a3 * seed_t at the probe cells of my patch.  The common case (at most one probe per thread) needs one shared
    // load and a compile-time unrolled select; further probes of the same thread go through the general loop.
    auto add_seeds = [&](float (&P)[R][4], int t) {
      if (pc0 >= 0) {
        const float* srow = ss + (t & (2 * TB - 1));
        patch_fma_cell<R>(P, k3, pc0, srow[pi0 * (2 * TB)]);
        if (more_probes) {
          for (int p = pi0 + 1; p < a.n_prb; ++p)
            if (pown[p] == L.lt) patch_fma_cell<R>(P, k3, pcell[p], srow[p * (2 * TB)]);
        }
      }
    };
    // In the step: the seed of my first probe is fetched at the top of the step and added with predicated FMAs (no branch
    // tree on the owning warp's path); further probes of the same thread take the general route.
    auto add_more_seeds = [&](float (&P)[R][4], int t) {
      if (more_probes) {
        const float* srow = ss + (t & (2 * TB - 1));
        for (int p = pi0 + 1; p < a.n_prb; ++p)
          if (pown[p] == L.lt) patch_fma_cell<R>(P, k3, pcell[p], srow[p * (2 * TB)]);
      }
    };
    const float* seed0 = ss + pi0 * (2 * TB);
    if (tid == 0) {   // prime the tape ring
      for (int s = 0; s < RG && s < a.T; ++s) {
        unsigned slot = (it_global + s) & (RG - 1);
        mbar_expect_tx(full + slot, stage_bytes);
        bulk_g2s(ring + slot * R * NT, tape_b + (size_t)(a.T - 1 - s) * tape_step, stage_bytes, full + slot);
      }
    }
    const int last_blk = (a.T - 1) / TB;
    stage_seeds(last_blk);
    if (last_blk > 0) stage_seeds(last_blk - 1);
    __syncthreads();

    float v[R][4], w[R][4];   // v = P_t, w = P_{t+1}
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) { v[r][k] = 0.f; w[r][k] = 0.f; }
    if (CHAIN && a.chain_in) {   // (P_{T-1} before its seeds, P_T) as the later segment left them
      const float4* cp = a.chain + ((size_t)b * a.C + L.rank) * 2 * R * NT + tid;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p = cp[r * NT], q = cp[(R + r) * NT];
        v[r][0] = p.x; v[r][1] = p.y; v[r][2] = p.z; v[r][3] = p.w;
        w[r][0] = q.x; w[r][1] = q.y; w[r][2] = q.z; w[r][3] = q.w;
      }
    }
    if (L.active) {
      add_seeds(v, a.T - 1);
      L.publish(pitch, fld, 0, v);
    }
    ++L.npub;
    __syncthreads();

    // One reverse step: `cu` = P_t (kept), `pr` = P_{t+1} on entry and P_{t-1} on exit.  PAR = it & 1 is a compile-time
    // constant of each of the two unrolled copies (shared-memory addresses are loop invariant).
    const float* rd0 = fld + own;
    const float* rd1 = fld + L.slab + own;
    const float4* ring_me = ring + tid;
    // PLAIN: instantiation for warps without special duties (plain_warp), as in k_res_fwd: no ghost-row wait or push, no
    // probe seeds, no dLoss/dx gather, no tape refill, all lanes active
    auto step = [&](auto par, auto plain_t, float (&cu)[R][4], float (&pr)[R][4], int t, int it) {
      constexpr int PAR = decltype(par)::value;
      constexpr bool PLAIN = decltype(plain_t)::value;
      const float* cur = PAR ? rd1 : rd0;
      const unsigned gi = it_global + it;
      const unsigned slot = gi & (RG - 1), parity = (gi >> RG_LOG) & 1u;
#ifdef WT_DEBUG_CLOCK
      long long c0 = clock64(), c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#endif
      if (!PLAIN) L.acquire_ghosts();
      float sv0 = 0.f;
      if (!PLAIN && SEED_PRED && seed_warp && t > 0) sv0 = pc0 >= 0 ? seed0[(t - 1) & (2 * TB - 1)] : 0.f;
#ifdef WT_DEBUG_CLOCK
      c1 = clock64();
#endif
      // EARLY (small patches): the stencil update and the ghost-row push come first, the tape stage and the gradient
      // accumulation -- which need neither the ghost rows nor the new field -- after.  With a handful of rows per CTA the
      // step time is the ring  push -> DSMEM flight -> neighbour's wait -> its update -> its push;  everything an edge
      // warp does between its wait and its push sits on that ring.  Big patches (config 3: R = 5) keep the old order:
      // there the step is issue-bound and consuming the tape stage first frees its registers before the stencil.
      auto stencil = [&]() {
        if (t > 0 || (CHAIN && a.chain_out)) {   // chained: P_{-1} of this segment is P_{T-1} of the one before it
          float lap[R][4];
          patch_laplacian<R>(pitch, cur, cu, lap);
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) pr[r][k] = wt_update(k1[r][k], k3[r][k], cu[r][k], pr[r][k], lap[r][k]);
          if (!CHAIN || t > 0) {
            if (!PLAIN && SEED_PRED && seed_warp) {
              patch_fma_pred<R>(pr, k3, pc0, sv0);
              add_more_seeds(pr, t - 1);
            }
            if (!PLAIN && !SEED_PRED) add_seeds(pr, t - 1);
            L.template publish<PLAIN>(pitch, fld, PAR ^ 1, pr);
          }
        }
      };
      if (EARLY && (PLAIN || L.active)) stencil();
#ifdef WT_DEBUG_CLOCK
      c2 = clock64();
#endif
      if (PLAIN || L.active) {
        if (!PLAIN && GRADX && a.grad_x && m1) {   // source.py:22: dLoss/dx[b,t] = sum over listed pixels of lambda_t = P_t / a3
          // a loop over the (few) source cells of this thread with one division each: 4R unrolled divisions would
          // triple the size of the step body for code that one thread per sample executes
          float s = 0.f;
          for (unsigned mm = m1; mm; mm &= mm - 1u) {
            const int bit = __ffs(mm) - 1;
            float cv = 0.f, kv = 1.f;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (bit == r * 4 + k) { cv = cu[r][k]; kv = k3[r][k]; }
            const float q = kv != 0.f ? cv / kv : 0.f;   // a cell with c == 0 carries no P
            s += q;
            if (m2 >> bit & 1u) s += q;
          }
          atomicAdd(gxs + (t & (2 * TB - 1)), s);
        }
        mbar_wait(full + slot, parity);
#ifdef WT_DEBUG_CLOCK
        c3 = clock64();
#endif
        const float4* rs = ring_me + slot * R * NT;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 l = rs[r * NT];
          G[r][0] = fmaf(l.x, cu[r][0], G[r][0]);     // cell.py:36 up to the factor 2/c applied at the end
          G[r][1] = fmaf(l.y, cu[r][1], G[r][1]);
          G[r][2] = fmaf(l.z, cu[r][2], G[r][2]);
          G[r][3] = fmaf(l.w, cu[r][3], G[r][3]);
        }
#ifdef WT_DEBUG_CLOCK
        c4 = clock64();
#endif
        if (!EARLY) stencil();
      }
      if (!PLAIN && t > 0) ++L.npub;
#ifdef WT_DEBUG_CLOCK
      if ((tid & 31) == 0 && (int)blockIdx.x < 2 && b == (int)blockIdx.x / a.C && (t == 500 || t == 501)) {
        long long* d = wt_dbg[tid >> 5][501 - t];
        d[0] = c0; d[1] = c1; d[2] = c2; d[3] = c3; d[4] = c4; d[5] = 0; d[6] = clock64(); d[7] = PLAIN;
      }
#endif
      step_barrier(NT);
      if (!PLAIN && tid == refill_tid && it + RG < a.T) {   // every thread has read this slot: refill it RG steps ahead
        mbar_expect_tx(full + slot, stage_bytes);
        bulk_g2s(ring + slot * R * NT, tape_b + (size_t)(t - RG) * tape_step, stage_bytes, full + slot);
      }
    };
    using P0 = std::integral_constant<int, 0>;
    using P1 = std::integral_constant<int, 1>;
    const unsigned npub0 = L.npub;
    auto run = [&](auto plain_t) {
      int t = a.T - 1, it = 0;
      for (; t >= 1; t -= 2, it += 2) {
        // Staging bookkeeping, once per block of TB steps and outside the step bodies (their code size is what the
        // instruction cache sees 2T times per sample).  When a block starts at the second step of the pair its seeds are
        // staged one step early: the half they go to held the seeds of the block before the previous one, all consumed.
        const int tb = ((t & (TB - 1)) == TB - 1) ? t : ((((t - 1) & (TB - 1)) == TB - 1) ? t - 1 : -1);
        if (tb >= 0 && tb != a.T - 1 && tb >= TB) stage_seeds(tb / TB - 1);
        step(P0{}, plain_t, v, w, t, it);
        step(P1{}, plain_t, w, v, t - 1, it + 1);
        if (GRADX && a.grad_x) {   // dLoss/dx of a block goes out right after its last (lowest) step
          if ((t & (TB - 1)) == 0 && t >= TB) flush_gx(t / TB);
          if (((t - 1) & (TB - 1)) == 0 && t - 1 >= TB) flush_gx((t - 1) / TB);
        }
      }
      if (t == 0) step(P0{}, plain_t, v, w, 0, it);
    };
    // every warp executes the same sequence of __syncthreads(); only the instruction stream between them differs
    if (plain_warp) run(std::true_type{}); else run(std::false_type{});
    L.npub = npub0 + (unsigned)(a.T - 1);
    it_global += (unsigned)a.T;
    if (CHAIN && a.chain_out) {   // (P_{-1}, P_0): after an even number of steps they sit in (v, w), else in (w, v)
      float4* cp = a.chain + ((size_t)b * a.C + L.rank) * 2 * R * NT + tid;
      const bool even = (a.T & 1) == 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 p = even ? make_float4(v[r][0], v[r][1], v[r][2], v[r][3]) : make_float4(w[r][0], w[r][1], w[r][2], w[r][3]);
        const float4 q = even ? make_float4(w[r][0], w[r][1], w[r][2], w[r][3]) : make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
        cp[r * NT] = p;
        cp[(R + r) * NT] = q;
      }
    }
    __syncthreads();
    if (GRADX && a.grad_x) flush_gx(0);
    __syncthreads();
  }
  // per-cluster partial of sum_{b,t} L(u_{t-1})*P_t ; reduced and scaled by 2/c in k_finish_grad_p
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int gi = L.gi0 + r, j = L.j0 + k;
      if (L.active && gi < a.Nx && j < a.Ny) {
        float* gp = a.Gpart + ((size_t)L.cid * a.Nx + gi) * a.Ny + j;
        *gp = (CHAIN && a.accumulate) ? *gp + G[r][k] : G[r][k];
      }
    }
#ifdef WT_DEBUG_CLOCK
  __syncthreads();
  if (tid == 0 && blockIdx.x < 2 && a.T > 502)
    for (int w = 0; w < NT / 32; ++w)
      for (int k = 0; k < 2; ++k) {
        long long* d = wt_dbg[w][k];
        printf("A cta=%d step=%d w=%d plain=%lld c0=%lld acq=%lld sten=%lld tapew=%lld G=%lld arrive=%lld\n", (int)blockIdx.x, k, w, d[7],
               d[0] - wt_dbg[0][0][0], d[1] - d[0], d[2] - d[0], d[3] ? d[3] - d[0] : 0, d[4] ? d[4] - d[0] : 0, d[6] - d[0]);
      }
#endif
  if (a.C > 1) cg::this_cluster().sync();
}

__global__ void k_finish_grad_p(const float* __restrict__ G, const float* __restrict__ c, int n_part, size_t stride,
                                size_t plane, float* __restrict__ grad_c) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= plane) return;
  float s = 0.f;
  for (int k = 0; k < n_part; ++k) s += G[(size_t)k * stride + i];
  const float ci = c[i];
  grad_c[i] = ci != 0.f ? 2.f * s / ci : 0.f;   // cell.py:36 is proportional to c: exactly zero where c == 0
}

// =================================================================================================
// host side
// =================================================================================================
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

static size_t smem_fwd_bytes(int Hc, int pitch, int n_prb, int R) {
  return (size_t)2 * slab_words(R, Hc, pitch) * 4 + 2 * TB * 4 + (size_t)2 * TB * n_prb * 4 + (size_t)(n_prb + 1) * 4 + 4 * 8 + 16;
}
static size_t smem_adj_bytes(int Hc, int pitch, int n_prb, int R, int threads, int ring) {
  return (size_t)ring * R * threads * 16 + (size_t)2 * slab_words(R, Hc, pitch) * 4 + (size_t)2 * TB * n_prb * 4 + 2 * TB * 4 +
         (size_t)2 * n_prb * 4 + 8 + MAX_RING * 8 + 4 * 8 + 16;
}

static int max_threads_for(int R) {
  switch (R) {
    case 1: return 512;
    case 2: return 512;
    case 3: return 384;
    case 4: return 384;
    case 5: return 384;
    case 6: return 320;
    case 8: return 256;
    default: return 0;
  }
}

// ---- on-chip checkpoint-and-recompute (linear kernels) ---------------------------------------------------------------
// wt_problem.checkpoint_every = S: the forward keeps no tape; every S steps (rounded up to a multiple of the TB = 64 step
// staging block) each thread stores its register patch (u_{t-1}, u_{t-2}).  The backward walks the segments in reverse:
// re-run the segment from its snapshot WITH a tape (which lives for one segment only), then run the adjoint over it, the
// pair (P_{t-1}, P_t) passing from launch to launch.  Memory: B*T*4 (x) + (T/S - 1) snapshots of 2 fields + S tape steps,
// instead of T tape steps.
static int ckpt_interval(const wt_problem* p) {
  if (p->checkpoint_every <= 0 || !(p->flags & WT_F_ZERO_INIT)) return 0;
  const int S = round_up(p->checkpoint_every, TB);
  return S < p->T ? S : 0;
}
struct CkptLayout { int n_seg; size_t x_bytes, patch, snaps, tape, total; };
static CkptLayout ckpt_layout(const wt_problem* p, int C, int R, int threads, int S) {
  CkptLayout l;
  l.n_seg = (p->T + S - 1) / S;
  l.x_bytes = ((size_t)p->B * p->T * 4 + 255) & ~(size_t)255;
  l.patch = (size_t)p->B * C * 2 * R * threads * 16;          // one snapshot (or the chain pair) of the whole batch
  l.snaps = l.patch * (l.n_seg - 1);
  l.tape = (size_t)p->B * S * C * R * threads * 16 * (nonlinear_mask(p) ? 2 : 1);   // nonlinear: u_{t-1} and L(u_{t-1})
  l.total = l.x_bytes + l.snaps + l.tape;
  return l;
}

template <typename K>
static int active_clusters(K kernel, int C, int threads, size_t smem) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (C > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
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

// Shape-specialised instantiations: (rows per thread, row pitch, threads per CTA, tape ring) as compile-time constants, so
// that every shared-memory and tape offset of the step body is an immediate (5-14 % on the step).  One entry per plan the
// planner picks for the grids and batch sizes of the reference's study configs; anything else runs the generic kernels
// (bitwise the same results: tests/test_gpu_parity.py::test_shape_specialised_kernels_match_generic_ones).
//   X(R, PITCH, THREADS, RING, TMAST)
// TMAST = 1: the tape-writing forward of this shape stages the tape rows of a step in shared memory (three buffers) and
// stores them with ONE TMA bulk copy per step instead of five 128-bit stores per thread.  Measured for the config-3 shape
// (same-box A/B, forward with tape: 0.859 -> 0.748 ms); WT_FWD_TMAST=0 / 1 overrides the table for A/B measurements.
#define WT_SPEC_SHAPES(X)                                                                                              \
  X(5, 104, 384, 4, 1)  /* 150x100, C=2: study/example.yml geometry at B >= 64 (BASELINE config 3, bench.py)          */ \
  X(2, 104, 256, 16, 0) /* 150x100, C=8: example.yml at its own batch_size 6; config 3 sharded 8 per GPU             */ \
  X(3, 104, 352, 4, 0)  /* 150x100, C=4: config 3 sharded 32 per GPU (64 waveforms over 2 GPUs)                       */ \
  X(2, 104, 352, 8, 0)  /* 150x100, C=6, R=2: config 3 sharded 16 per GPU (64 waveforms over 4 GPUs)                   */ \
  X(2, 144, 320, 8, 0)  /* 140x140, C=8: study/linear/linear.yml (batch_size 9)                                      */ \
  X(2, 156, 384, 8, 0)  /* 151x151, C=8: study/propagate.py, study/optimize_lens.py (BASELINE configs 1-2)           */

static size_t tape_stage_bytes(int R, int threads) { return (size_t)3 * R * threads * 16 + 256; }
static int fwd_tmast_override() {   // -1: follow the table
  static const int v = [] { const char* e = getenv("WT_FWD_TMAST"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  return v;
}

// Clusters that can be co-resident for both kernels of a decomposition (0 = cannot launch).  Cached per device.
static int resident_clusters(int device, int R, int C, int threads, size_t smem_fwd, size_t smem_bwd, int pitch, int ring,
                             bool specialize) {
  struct Key { int dev, R, C, threads; size_t sf, sb; int spec, n; };
  static std::mutex mu;
  static std::vector<Key> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (const Key& k : cache)
    if (k.dev == device && k.R == R && k.C == C && k.threads == threads && k.sf == smem_fwd && k.sb == smem_bwd && k.spec == (int)specialize) return k.n;
  int nf = 0, nb = 0;
  switch (R) {
#define WT_OCC(R_) case R_: { int n0 = active_clusters(k_res_fwd<R_, false>, C, threads, smem_fwd); int n1 = active_clusters(k_res_fwd<R_, true>, C, threads, smem_fwd); nf = n0 < n1 ? n0 : n1; nb = active_clusters(k_res_adj<R_>, C, threads, smem_bwd); } break;
    WT_OCC(1) WT_OCC(2) WT_OCC(3) WT_OCC(4) WT_OCC(5) WT_OCC(6) WT_OCC(8)
#undef WT_OCC
    default: break;
  }
  // a shape-specialised instantiation is what will be launched: its register budget (two CTAs per SM for the small ones) counts
  if (specialize) {
#define WT_SPEC_OCC(R_, P_, N_, G_, T_)                                                                    \
    if (R == R_ && pitch == P_ && threads == N_) {                                                           \
      const bool tm = fwd_tmast_override() < 0 ? (T_ != 0) : fwd_tmast_override() != 0;                       \
      int n0 = active_clusters(k_res_fwd<R_, false, P_, N_>, C, threads, smem_fwd);                          \
      int n1 = tm ? active_clusters(k_res_fwd<R_, true, P_, N_, false, false, false, true>, C, threads,      \
                                    smem_fwd + tape_stage_bytes(R_, N_))                                      \
                  : active_clusters(k_res_fwd<R_, true, P_, N_>, C, threads, smem_fwd);                       \
      nf = n0 < n1 ? n0 : n1;                                                                                \
      if (ring == G_) {                                                                                      \
        int b0 = active_clusters(k_res_adj<R_, P_, N_, 0, G_>, C, threads, smem_bwd);                        \
        int b1 = active_clusters(k_res_adj<R_, P_, N_, 1, G_>, C, threads, smem_bwd);                        \
        nb = b0 < b1 ? b0 : b1;                                                                              \
      }                                                                                                      \
    }
    WT_SPEC_SHAPES(WT_SPEC_OCC)
#undef WT_SPEC_OCC
  }
  int n = nf < nb ? nf : nb;
  cache.push_back(Key{device, R, C, threads, smem_fwd, smem_bwd, (int)specialize, n});
  return n;
}

static int res_nl_clusters_cached(int device, int R, int nl, int C, int threads, size_t sf, size_t sb) {
  struct Key { int dev, R, nl, C, threads; size_t sf, sb; int n; };
  static std::mutex mu;
  static std::vector<Key> cache;
  std::lock_guard<std::mutex> lock(mu);
  for (const Key& k : cache)
    if (k.dev == device && k.R == R && k.nl == nl && k.C == C && k.threads == threads && k.sf == sf && k.sb == sb) return k.n;
  int n = res_nl_clusters(R, nl, C, threads, sf, sb);
  cache.push_back(Key{device, R, nl, C, threads, sf, sb, n});
  return n;
}

bool resident_plan(const wt_problem* p, const cudaDeviceProp& prop, bool need_adjoint, wt_plan* plan) {
  const int nl = nonlinear_mask(p);
  if (p->flags & WT_F_NEED_GRAD_B) return false;
  if (nl && !(p->flags & WT_F_ZERO_INIT)) return false;   // the on-chip nonlinear adjoint assumes zero initial fields
  if (p->n_prb > MAX_PRB || p->T < 1) return false;
  const int P4 = (p->Ny + 3) / 4, pitch = 4 * P4 + 4;
  const int smem_cap = (int)prop.sharedMemPerBlockOptin;
  static const int Rs_lin[] = {8, 6, 5, 4, 3, 2, 1};
  static const int Rs_nl[] = {4, 3, 2, 1};
  static const int Cs[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 16};   // any cluster size works; > 8 is non-portable (opt-in)
  const int* Rs = nl ? Rs_nl : Rs_lin;
  const int nR = nl ? 4 : 7;
  auto ring_for = [&](int Hc, int R, int threads) {   // deepest tape ring that fits (nonlinear stages are twice as big)
    for (int ring = RING; ring >= 2; ring /= 2)   // powers of two: slot and parity of a stage are a mask and a shift
      if ((int)res_nl_smem_adj(Hc, pitch, p->n_prb, R, threads, ring) <= smem_cap) return ring;
    return 0;
  };
  // linear adjoint: deepest power-of-two ring (<= MAX_RING stages, <= 128 KB) that fits next to the slab buffers.  Big patches
  // (config 3: 30 KB per stage) get 4 stages as before; small ones (R = 1..2, latency-bound small batches) get 8..16
  auto ring_lin = [&](int Hc, int R, int threads) {
    for (int ring = MAX_RING; ring >= 2; ring /= 2)
      if ((size_t)ring * R * threads * 16 <= (size_t)128 * 1024 &&
          (int)smem_adj_bytes(Hc, pitch, p->n_prb, R, threads, ring) <= smem_cap) return ring;
    return 0;
  };
  int bestC = 0, bestR = 0;
  double best_score = -1;
  for (int C : Cs) {
    if (p->cluster && C != p->cluster) continue;
    if (C > 8 && !p->cluster) continue;   // non-portable cluster sizes only on request
    for (int ri = 0; ri < nR; ++ri) {
      const int R = Rs[ri];
      if (p->rows_per_thread && R != p->rows_per_thread) continue;
      const int Hc = round_up((p->Nx + C - 1) / C, R);
      if ((C - 1) * Hc >= p->Nx) continue;            // every CTA must own at least one real row
      const int runs = Hc / R, nact = runs * P4, threads = round_up(nact, 32);
      if (threads > (nl ? res_nl_max_threads_rt(R) : max_threads_for(R))) continue;
      if (threads < p->n_prb) continue;                // the forward kernels sample the probes with one lane per probe
      if (nl) {
        if (!ring_for(Hc, R, threads)) continue;
      } else {
        if (need_adjoint ? !ring_lin(Hc, R, threads) : (int)smem_fwd_bytes(Hc, pitch, p->n_prb, R) > smem_cap) continue;
      }
      // estimated time per step ~ (waves of clusters) x (rows per CTA) / (per-thread efficiency)
      size_t sf, sb;
      int ncl;
      if (nl) {
        sf = res_nl_smem_fwd(Hc, pitch, p->n_prb, R);
        sb = res_nl_smem_adj(Hc, pitch, p->n_prb, R, threads, ring_for(Hc, R, threads));
        ncl = res_nl_clusters_cached(p->device, R, nl, C, threads, sf, sb);
      } else {
        sf = smem_fwd_bytes(Hc, pitch, p->n_prb, R);
        sb = smem_adj_bytes(Hc, pitch, p->n_prb, R, threads, ring_lin(Hc, R, threads));
        ncl = resident_clusters(p->device, R, C, threads, sf, sb, pitch, ring_lin(Hc, R, threads), !(p->flags & WT_F_NO_SPECIALIZE));
      }
      if (ncl < 1) continue;
      const int waves = (p->B + ncl - 1) / ncl;
      const int used = p->B < ncl ? p->B : ncl;
      const int per_sm = (used * C + prop.multiProcessorCount - 1) / prop.multiProcessorCount;   // CTAs sharing an SM
      const double rim = (double)(4 * R) / (4 * R + 2 * R + 8);   // own cells / (own + rim loads)
      const double lane = (double)nact / threads;
      const double sync_cost = C > 1 ? 0.85 : 1.0;
      const double par = threads >= 384 ? 1.0 : threads / 384.0;
      // measured: linear R = 4..5 beats 6..8; the nonlinear adjoint keeps ~10 values per cell live and is fastest at R = 1
      const double regs = nl ? (R == 1 ? 1.6 : R == 2 ? 1.0 : 0.6) : (R >= 6 ? 0.8 : 1.0);
      double score = (rim * lane * sync_cost * par * regs) / ((double)waves * Hc * per_sm);
      if (!nl) {
        // Linear kernels: a cost model fitted to measured sweeps (profiles/r2_sweeps.md; fwd-with-tape + adjoint, us per time
        // step of one CTA with Hc rows of 100 cells):  t = a_R + b_R * Hc.  The fixed part a_R is the latency chain of one
        // thread's patch plus the step barrier (small patches win when there are SMs to spread over), the slope b_R the
        // issue-bound rate (big patches win when every SM is busy).  score = 1 / (waves * t).
        // (refitted at the end of round 2: R = 3, 4 are compiled for 384 threads = 168 registers -- at 640 / 512 threads they
        //  spilled inside the loop -- and the R = 5 step barely depends on the rows of a CTA any more)
        static const double aR[9] = {0, 0.57, 0.50, 0.925, 0.983, 1.50, 1.30, 0, 1.275};
        static const double bR[9] = {0, 0.031, 0.0216, 0.0082, 0.0086, 0.0015, 0.0098, 0, 0.0104};
        const double t = aR[R] + bR[R] * Hc * (P4 / 25.0) * (per_sm > 1 ? per_sm : 1);
        score = 1.0 / ((double)waves * t);
        score *= 1.0 - 1e-3 * C;      // ties: the smaller cluster
      }
      if (score > best_score) { best_score = score; bestC = C; bestR = R; }
    }
  }
  if (!bestC) return false;
  const int Hc = round_up((p->Nx + bestC - 1) / bestC, bestR);
  const int runs = Hc / bestR, nact = runs * P4, threads = round_up(nact, 32);
  plan->path = WT_PATH_RESIDENT;
  plan->cluster = bestC;
  plan->rows_per_thread = bestR;
  plan->threads = threads;
  plan->rows_per_cta = Hc;
  plan->nonlinear = nl;
  const size_t plane = (size_t)p->Nx * p->Ny;
  int ncl;
  if (nl) {
    const int ring = ring_for(Hc, bestR, threads);
    plan->reserved[0] = ring;
    plan->smem_fwd = (int)res_nl_smem_fwd(Hc, pitch, p->n_prb, bestR);
    plan->smem_bwd = (int)res_nl_smem_adj(Hc, pitch, p->n_prb, bestR, threads, ring);
    ncl = res_nl_clusters_cached(p->device, bestR, nl, bestC, threads, plan->smem_fwd, plan->smem_bwd);
  } else {
    const int ring = ring_lin(Hc, bestR, threads);
    plan->reserved[0] = ring;
    plan->smem_fwd = (int)smem_fwd_bytes(Hc, pitch, p->n_prb, bestR);
    plan->smem_bwd = (int)smem_adj_bytes(Hc, pitch, p->n_prb, bestR, threads, ring);
    ncl = resident_clusters(p->device, bestR, bestC, threads, plan->smem_fwd, plan->smem_bwd, pitch, ring, !(p->flags & WT_F_NO_SPECIALIZE));
  }
  if (ncl < 1) return false;
  plan->n_clusters = p->B < ncl ? p->B : ncl;
  plan->history_bytes = (uint64_t)p->B * p->T * bestC * (nl ? 2 : 1) * bestR * threads * 16;
  plan->workspace_fwd_bytes = 3 * plane * 4 + 64;
  plan->workspace_bwd_bytes = (3 + 2 * (size_t)plan->n_clusters) * plane * 4 + 64;
  plan->reserved[2] = 0;
  const int S = ckpt_interval(p);
  if (S) {
    // checkpoint-and-recompute on chip: history = [x copy][register-patch snapshots every S steps][tape of ONE segment]
    const CkptLayout lay = ckpt_layout(p, bestC, bestR, threads, S);
    plan->reserved[2] = S;
    plan->history_bytes = lay.total;
    plan->workspace_bwd_bytes += lay.patch + 256;       // the (P_{t-1}, P_t) pair handed from segment to segment
    plan->launches_fwd = 3;
    plan->launches_bwd = 2 * lay.n_seg + 2;
    return true;
  }
  plan->launches_fwd = nl ? 1 : 2;
  plan->launches_bwd = 3;
  return true;
}

static void fill_args(const wt_problem* p, const wt_plan& plan, ResArgs* a) {
  a->Nx = p->Nx; a->Ny = p->Ny; a->B = p->B; a->T = p->T;
  a->C = plan.cluster; a->Hc = plan.rows_per_cta; a->P4 = (p->Ny + 3) / 4; a->pitch = 4 * a->P4 + 4;
  a->runs = plan.rows_per_cta / plan.rows_per_thread; a->nact = a->runs * a->P4;
  a->n_src = p->n_src; a->n_prb = p->n_prb; a->n_clusters = plan.n_clusters; a->flags = p->flags;
  a->field_every = p->field_every > 1 ? p->field_every : 1;
}

template <typename K>
static int launch_cluster(K kernel, const wt_plan& plan, size_t smem, const ResArgs& a, cudaStream_t st) {
  WT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (plan.cluster > 8) WT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.n_clusters * plan.cluster);
  cfg.blockDim = dim3(plan.threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int sched = [] {   // WT_CLUSTER_SCHED=spread|lb: cluster scheduling policy preference (A/B measurements)
    const char* e = getenv("WT_CLUSTER_SCHED");
    return !e ? 0 : (e[0] == 's' ? 1 : 2);
  }();
  if (sched) {
    attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
    attr[1].val.clusterSchedulingPolicyPreference = sched == 1 ? cudaClusterSchedulingPolicySpread : cudaClusterSchedulingPolicyLoadBalancing;
    cfg.numAttrs = 2;
  }
  WT_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  return WT_OK;
}

#define WT_DISPATCH_R(R_, CALL)                  \
  switch (R_) {                                  \
    case 1: { constexpr int R = 1; CALL; } break; \
    case 2: { constexpr int R = 2; CALL; } break; \
    case 3: { constexpr int R = 3; CALL; } break; \
    case 4: { constexpr int R = 4; CALL; } break; \
    case 5: { constexpr int R = 5; CALL; } break; \
    case 6: { constexpr int R = 6; CALL; } break; \
    case 8: { constexpr int R = 8; CALL; } break; \
    default: wt::set_error("rows_per_thread=%d not instantiated", R_); return WT_EINVAL; \
  }

int resident_forward(const wt_problem* p, const wt_plan& plan, const float* c, const float* b, const float* rho,
                     const float* x,
                     const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_sq, float* u1, float* u2,
                     float* probe_out, float* probe_raw, float* fields_out, void* history, void* workspace,
                     cudaStream_t st) {
  const size_t plane = (size_t)p->Nx * p->Ny;
  float* a1 = reinterpret_cast<float*>(workspace);
  float* a3 = a1 + plane;
  float* gs = a3 + plane;
  int* status = reinterpret_cast<int*>(gs + plane);
  if (!plan.nonlinear)
    k_coeff<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(b, c, (int)plane, p->dt, (p->dt * p->dt) / (p->h * p->h),
                                                              a1, a3, gs);
  WT_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  ResArgs a = {};
  fill_args(p, plan, &a);
  a.bpml = b; a.clin = c; a.rho = rho; a.s = make_scalars(p); a.ring = plan.reserved[0];
  a.a1 = a1; a.a3 = a3; a.x = x; a.src_ij = src_ij; a.prb_ij = prb_ij; a.prb_sq = prb_sq;
  a.u1 = u1; a.u2 = u2; a.probe_out = probe_out; a.probe_raw = probe_raw; a.fields = fields_out;
  a.vec_fields = (p->Ny % 4 == 0) && (((uintptr_t)fields_out & 15) == 0);
  a.tape = reinterpret_cast<float4*>(history);
  a.status = status;
  if (plan.nonlinear && !(plan.reserved[2] && history)) return res_nl_launch_fwd(plan, a, st);
  if (plan.reserved[2] && history) {   // checkpoint-and-recompute: no tape, snapshots every S steps, x kept for the recompute
    if (a.fields) { wt::set_error("wt_forward: fields_out together with history needs WT_F_FORCE_STREAM"); return WT_EUNSUPPORTED; }
    const CkptLayout lay = ckpt_layout(p, plan.cluster, plan.rows_per_thread, plan.threads, plan.reserved[2]);
    char* hb = reinterpret_cast<char*>(history);
    WT_CUDA(cudaMemcpyAsync(hb, x, (size_t)p->B * p->T * 4, cudaMemcpyDeviceToDevice, st));
    a.tape = nullptr;
    a.Tstride = p->T; a.t_off = 0; a.snap_every = plan.reserved[2];
    a.snap = reinterpret_cast<float4*>(hb + lay.x_bytes);
    a.snap_in = nullptr;
    if (plan.nonlinear) return res_nl_launch_fwd(plan, a, st, true);
    WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, false, 0, 0, false, true>, plan, plan.smem_fwd, a, st)));
    return WT_OK;
  }
  if (!a.fields && (a.flags & WT_F_NO_PLAIN_WARPS)) {   // A/B and test switch: every warp through the general step (generic kernels)
    if (a.tape) { WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, true, 0, 0, false, false, true>, plan, plan.smem_fwd, a, st))); }
    else { WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, false, 0, 0, false, false, true>, plan, plan.smem_fwd, a, st))); }
    return WT_OK;
  }
  if (!a.fields && !(a.flags & WT_F_NO_SPECIALIZE)) {
#define WT_SPEC_F(R_, P_, N_, G_, T_)                                                                   \
    if (plan.rows_per_thread == R_ && a.pitch == P_ && plan.threads == N_) {                              \
      const bool tm = fwd_tmast_override() < 0 ? (T_ != 0) : fwd_tmast_override() != 0;                    \
      if (a.tape && tm)                                                                                   \
        WT_TRY(launch_cluster(k_res_fwd<R_, true, P_, N_, false, false, false, true>, plan,               \
                              plan.smem_fwd + tape_stage_bytes(R_, N_), a, st));                           \
      else if (a.tape) WT_TRY(launch_cluster(k_res_fwd<R_, true, P_, N_>, plan, plan.smem_fwd, a, st));   \
      else WT_TRY(launch_cluster(k_res_fwd<R_, false, P_, N_>, plan, plan.smem_fwd, a, st));              \
      return WT_OK;                                                                                       \
    }
    WT_SPEC_SHAPES(WT_SPEC_F)
#undef WT_SPEC_F
  }
  if (a.fields) {
    if (a.tape) { wt::set_error("wt_forward: fields_out together with history needs WT_F_FORCE_STREAM"); return WT_EUNSUPPORTED; }
    WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, false, 0, 0, true>, plan, plan.smem_fwd, a, st)));
    return WT_OK;
  }
  if (a.tape) { WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, true>, plan, plan.smem_fwd, a, st))); }
  else { WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, false>, plan, plan.smem_fwd, a, st))); }
  return WT_OK;
}

int resident_backward(const wt_problem* p, const wt_plan& plan, const float* c, const float* b, const float* rho,
                      const int32_t* src_ij,
                      const int32_t* prb_ij, const int32_t* prb_sq, const float* grad_probe, const float* probe_raw,
                      const void* history, float* grad_c, float* grad_b, float* grad_rho, float* grad_x,
                      void* workspace, cudaStream_t st) {
  const size_t plane = (size_t)p->Nx * p->Ny;
  float* a1 = reinterpret_cast<float*>(workspace);
  float* a3 = a1 + plane;
  float* gs = a3 + plane;
  float* Gpart = gs + plane;
  int* status = reinterpret_cast<int*>(Gpart + 2 * (size_t)plan.n_clusters * plane);
  if (!plan.nonlinear)
    k_coeff<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(b, c, (int)plane, p->dt, (p->dt * p->dt) / (p->h * p->h),
                                                              a1, a3, gs);
  WT_CUDA(cudaMemsetAsync(status, 0, sizeof(int), st));
  if (grad_x) WT_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)p->B * p->T * sizeof(float), st));
  ResArgs a = {};
  fill_args(p, plan, &a);
  a.a1 = a1; a.a3 = a3; a.src_ij = src_ij; a.prb_ij = prb_ij; a.prb_sq = prb_sq;
  a.probe_raw = const_cast<float*>(probe_raw); a.grad_probe = grad_probe; a.grad_x = grad_x;
  a.tape = reinterpret_cast<float4*>(const_cast<void*>(history));
  a.Gpart = Gpart; a.status = status;
  a.bpml = b; a.clin = c; a.rho = rho; a.s = make_scalars(p); a.ring = plan.reserved[0];
  const unsigned fg = (unsigned)((plane + 255) / 256);
  if (plan.nonlinear && !plan.reserved[2]) {
    WT_TRY(res_nl_launch_adj(plan, a, st));
    k_finish_grad<<<fg, 256, 0, st>>>(Gpart, nullptr, plan.n_clusters, 2 * plane, plane, grad_c);
    if (grad_rho) k_finish_grad<<<fg, 256, 0, st>>>(Gpart + plane, nullptr, plan.n_clusters, 2 * plane, plane, grad_rho);
    if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));   // needs WT_F_NEED_GRAD_B (streaming path)
    WT_CUDA(cudaGetLastError());
    return WT_OK;
  }
  if (plan.reserved[2]) {   // checkpoint-and-recompute: per segment, newest first: forward from its snapshot with a tape, adjoint
    const int S = plan.reserved[2];
    const CkptLayout lay = ckpt_layout(p, plan.cluster, plan.rows_per_thread, plan.threads, S);
    char* hb = reinterpret_cast<char*>(const_cast<void*>(history));
    float4* snaps = reinterpret_cast<float4*>(hb + lay.x_bytes);
    float4* tape = reinterpret_cast<float4*>(hb + lay.x_bytes + lay.snaps);
    float4* chain = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(status + 1) + 255) & ~(uintptr_t)255);
    for (int k = lay.n_seg - 1; k >= 0; --k) {
      const int t0 = k * S, len = (p->T - t0 < S) ? p->T - t0 : S;
      ResArgs f = a;
      f.T = len; f.Tstride = p->T; f.t_off = t0; f.snap_every = 0; f.snap = nullptr;
      f.snap_in = k > 0 ? snaps + (size_t)(k - 1) * (lay.patch / 16) : nullptr;
      f.x = reinterpret_cast<const float*>(hb);
      f.u1 = f.u2 = nullptr; f.probe_out = nullptr; f.probe_raw = nullptr; f.fields = nullptr; f.grad_x = nullptr;
      f.tape = tape;
      ResArgs g = a;
      g.T = len; g.Tstride = p->T; g.t_off = t0; g.tape = tape; g.chain = chain; g.snap_in = f.snap_in;
      g.chain_in = k < lay.n_seg - 1; g.chain_out = k > 0; g.accumulate = k < lay.n_seg - 1;
      if (plan.nonlinear) {
        WT_TRY(res_nl_launch_fwd(plan, f, st, true));
        WT_TRY(res_nl_launch_adj(plan, g, st, true));
      } else {
        WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_fwd<R, true, 0, 0, false, true>, plan, plan.smem_fwd, f, st)));
        WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_adj<R, 0, 0, 1, 0, true>, plan, plan.smem_bwd, g, st)));
      }
    }
    if (plan.nonlinear) {
      k_finish_grad<<<fg, 256, 0, st>>>(Gpart, nullptr, plan.n_clusters, 2 * plane, plane, grad_c);
      if (grad_rho) k_finish_grad<<<fg, 256, 0, st>>>(Gpart + plane, nullptr, plan.n_clusters, 2 * plane, plane, grad_rho);
      if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));
      WT_CUDA(cudaGetLastError());
      return WT_OK;
    }
    k_finish_grad_p<<<fg, 256, 0, st>>>(Gpart, c, plan.n_clusters, plane, plane, grad_c);
    if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));
    if (grad_rho) WT_CUDA(cudaMemsetAsync(grad_rho, 0, plane * sizeof(float), st));
    WT_CUDA(cudaGetLastError());
    return WT_OK;
  }
  bool launched = false;
  if (a.flags & WT_F_NO_PLAIN_WARPS) {
    WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_adj<R, 0, 0, 1, 0, false, true>, plan, plan.smem_bwd, a, st)));
    launched = true;
  }
  if (!launched && !(a.flags & WT_F_NO_SPECIALIZE)) {
#define WT_SPEC_A(R_, P_, N_, G_, T_)                                                                              \
    if (!launched && plan.rows_per_thread == R_ && a.pitch == P_ && plan.threads == N_ && a.ring == G_) {         \
      if (a.grad_x) WT_TRY(launch_cluster(k_res_adj<R_, P_, N_, 1, G_>, plan, plan.smem_bwd, a, st));             \
      else WT_TRY(launch_cluster(k_res_adj<R_, P_, N_, 0, G_>, plan, plan.smem_bwd, a, st));                      \
      launched = true;                                                                                            \
    }
    WT_SPEC_SHAPES(WT_SPEC_A)
#undef WT_SPEC_A
  }
  if (!launched) { WT_DISPATCH_R(plan.rows_per_thread, WT_TRY(launch_cluster(k_res_adj<R>, plan, plan.smem_bwd, a, st))); }
  k_finish_grad_p<<<fg, 256, 0, st>>>(Gpart, c, plan.n_clusters, plane, plane, grad_c);
  if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));
  if (grad_rho) WT_CUDA(cudaMemsetAsync(grad_rho, 0, plane * sizeof(float), st));
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

}  // namespace wt
