/*
 * wavetorch_b200 -- C ABI of the B200-native wave-RNN hot path.
 *
 * This is the drop-in boundary for the time loop of fancompute/wavetorch (reference paths are relative
 * to the reference repository root):
 *
 *   wt_forward        replaces  WaveRNN.forward's Python loop           wavetorch/rnn.py:36-70
 *                               (WaveCell.forward cell.py:79-107, TimeStep/_time_step cell.py:12-24,
 *                                _laplacian operators.py:5-11, WaveSource.forward source.py:15-22,
 *                                WaveProbe/WaveIntensityProbe.forward probe.py:14-27)
 *   wt_backward       replaces  what autograd replays through the loop: TimeStep.backward cell.py:27-44,
 *                               the nonlinear b(u), c(u) graph of cell.py:94-102, the source add and the
 *                               probe gather/square
 *   wt_step_forward   replaces  TimeStep.forward   (cell.py:22-24)  -- one step, no source/probe
 *   wt_step_backward  replaces  TimeStep.backward  (cell.py:27-44)
 *
 * The reference has no FFI of its own (it is pure Python on ATen); INTEGRATION.md shows the ctypes stub a
 * maintainer would add to wavetorch/rnn.py and wavetorch/cell.py to bind these entry points.
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless stated otherwise
 *   - all fields are float32, contiguous, row-major [.., Nx, Ny] with Ny innermost (the reference layout)
 *   - the caller owns all memory; the library never allocates or frees device memory.  Scratch space is
 *     sized with wt_query_plan() and passed in
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) of device `device`; calls are
 *     asynchronous with respect to the host and may be issued concurrently for different devices/streams
 *   - every entry point returns 0 on success, a negative WT_E* code otherwise; wt_last_error() returns a
 *     thread-local message.  No exceptions cross the boundary.  There is no CPU fallback.
 */
#ifndef WAVETORCH_B200_H
#define WAVETORCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_ABI_VERSION 1

/* status codes */
#define WT_OK 0
#define WT_EINVAL (-1)        /* bad argument (shape, null pointer, misaligned, ...) */
#define WT_ECUDA (-2)         /* a CUDA runtime call failed; see wt_last_error() */
#define WT_EUNSUPPORTED (-3)  /* valid request that this build cannot run (e.g. device is not sm_100) */
#define WT_ENOSPACE (-4)      /* history / workspace buffer too small */

/* wt_problem.flags */
#define WT_F_ZERO_INIT 1u       /* ignore the incoming u1/u2 contents: start from zero fields (rnn.py:39-41) */
#define WT_F_FORCE_STREAM 2u    /* plan: always use the HBM-streaming kernels                             */
#define WT_F_FORCE_RESIDENT 4u  /* plan: fail with WT_EUNSUPPORTED instead of falling back to streaming  */
#define WT_F_NEED_GRAD_B 8u     /* the tape must allow grad w.r.t. the damping field in linear mode      */
#define WT_F_NO_SPECIALIZE 16u  /* on-chip path: run the generic kernels, not a shape-specialised instantiation
                                   (A/B measurements; results are bitwise identical either way)            */
#define WT_F_NO_PLAIN_WARPS 32u /* on-chip path (linear cell): every warp runs the general instantiation of the time step, also
                                   the warps without ghost-row, source, probe or refill duties (A/B; bitwise identical) */

/* wt_plan.path */
#define WT_PATH_STREAM 0    /* one launch per time step, fields live in HBM                     */
#define WT_PATH_RESIDENT 1  /* whole time loop in one launch, fields live in registers + SMEM   */

/* Problem descriptor (host memory).  dt and h are the values of the reference's `cell.dt` / `geom.h`
 * buffers; b0, uth, c_nl those of cell.satdamp_b0 / cell.satdamp_uth / cell.c_nl (cell.py:60-64).
 * Saturable damping is active iff b0 > 0 (cell.py:94); the Kerr term iff c_nl != 0 (cell.py:99). */
typedef struct wt_problem {
  int32_t Nx, Ny;   /* grid */
  int32_t B;        /* independent waveforms (batch) */
  int32_t T;        /* time steps advanced by this call */
  int32_t n_src;    /* source pixel entries; a pixel listed k times receives k*x (rnn.py:56-57); see wt_validate_pixels */
  int32_t n_prb;    /* probe pixels */
  uint32_t flags;
  int32_t device;   /* CUDA device ordinal */
  double dt, h;
  double b0, uth, c_nl;
  /* tuning overrides, 0 = automatic */
  int32_t cluster;  /* CTAs per sample (resident path): 1,2,4,8,16 */
  int32_t rows_per_thread;
  int32_t field_every; /* fields_out keeps every field_every-th field only (0 or 1 = all): snapshot k is the field after
                          step (k+1)*field_every - 1, k < T / field_every.  Forward/inference only. */
  int32_t checkpoint_every; /* > 0, on-chip path (linear or nonlinear cell), WT_F_ZERO_INIT: checkpoint-and-recompute.  wt_forward writes no
                          tape; it stores the field pair every checkpoint_every steps (rounded up to a multiple of 64) into
                          `history`, and wt_backward re-runs each segment with a tape that lives for one segment only.
                          plan.history_bytes shrinks from T to about checkpoint_every + 2*T/checkpoint_every field copies;
                          plan.reserved[2] reports the interval in use (0 = not applicable: the whole tape is kept). */
  int32_t reserved[4];
} wt_problem;

/* What wt_forward/wt_backward will do for a problem, and how much caller-provided scratch they need. */
typedef struct wt_plan {
  int32_t path;            /* WT_PATH_* */
  int32_t cluster;         /* CTAs per sample */
  int32_t rows_per_thread;
  int32_t threads;         /* threads per CTA */
  int32_t rows_per_cta;
  int32_t n_clusters;      /* clusters in the grid (persistent over samples) */
  int32_t smem_fwd, smem_bwd;      /* dynamic shared memory per CTA */
  int32_t nonlinear;       /* bit0 saturable damping, bit1 Kerr */
  int32_t launches_fwd, launches_bwd; /* kernel launches one call performs */
  int32_t reserved[5];
  uint64_t history_bytes;  /* adjoint tape written by wt_forward when `history` is given */
  uint64_t workspace_fwd_bytes;
  uint64_t workspace_bwd_bytes;
} wt_plan;

int wt_abi_version(void);
const char* wt_last_error(void);

/* Fill `plan` for `p` on p->device (queries the device; launches nothing). */
int wt_query_plan(const wt_problem* p, wt_plan* plan);

/*
 * Host-side check of the pixel lists a caller is about to pass to wt_forward / wt_backward.  src_ij_host / prb_ij_host are
 * HOST copies ([n_src,2] / [n_prb,2], (row, col)) of the device arrays.  Returns WT_EINVAL when a coordinate lies outside
 * the p->Nx x p->Ny grid; otherwise WT_OK with *max_listings (nullable) = the largest number of times one source pixel
 * is listed (0 without sources).  wt_forward / wt_backward only see device pointers and cannot check them without a
 * synchronisation, so this is the caller's contract: coordinates must be in range (the kernels redirect an out-of-range
 * entry to cell (0,0) rather than touch foreign memory), and a problem whose *max_listings exceeds WT_MAX_SRC_LISTINGS
 * must set WT_F_FORCE_STREAM -- the on-chip kernels add x[b,t] at most that many times per pixel (rnn.py:56-57 adds it
 * once per listing; the streaming kernels handle any count).
 */
#define WT_MAX_SRC_LISTINGS 2
int wt_validate_pixels(const wt_problem* p, const int32_t* src_ij_host, const int32_t* prb_ij_host, int32_t* max_listings);

/*
 * Advance p->T steps for p->B waveforms.
 *
 *   c, b        [Nx,Ny]   linear wave speed (geom.c) and damping (geom.b)
 *   rho         [Nx,Ny]   density fed to the nonlinear terms (geom.rho); may be NULL when linear
 *   x           [B,T]     input waveforms; x[b,t] is added to every source pixel after step t
 *   src_ij      [n_src,2] int32 (row, col) of the source pixels
 *   prb_ij      [n_prb,2] int32 (row, col) of the probe pixels
 *   prb_square  [n_prb]   int32, nonzero = WaveIntensityProbe (output is the squared field)
 *   u1, u2      [B,Nx,Ny] in: fields at t-1 and t-2 (ignored with WT_F_ZERO_INIT); out: the two latest fields
 *   probe_out   [B,T,n_prb] probe readout (squared where prb_square), the reference's model(x) output
 *   probe_raw   [B,T,n_prb] field at the probes (needed by wt_backward); may be NULL
 *   fields_out  [B,T,Nx,Ny] every field (output_fields=True, rnn.py:65-67), or [B,T/field_every,Nx,Ny] time-decimated
 *               snapshots when p->field_every > 1 (not together with `history`); may be NULL.  Together with `history` it
 *               needs WT_F_FORCE_STREAM (dLoss/dfields is implemented by the streaming adjoint); WT_EUNSUPPORTED otherwise
 *   history     adjoint tape of plan.history_bytes, or NULL for inference
 *   workspace   plan.workspace_fwd_bytes
 */
int wt_forward(const wt_problem* p, const float* c, const float* b, const float* rho, const float* x,
               const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, float* u1, float* u2,
               float* probe_out, float* probe_raw, float* fields_out, void* history, size_t history_bytes,
               void* workspace, size_t workspace_bytes, void* stream);

/*
 * Reverse-time adjoint of wt_forward over the same p->T steps.
 *
 *   grad_probe  [B,T,n_prb]  dLoss/d probe_out
 *   probe_raw   [B,T,n_prb]  as written by wt_forward
 *   grad_fields [B,T,Nx,Ny]  dLoss/d fields_out, or NULL
 *   history     the tape written by wt_forward for the same problem
 *   adj1, adj2  [B,Nx,Ny]    in/out adjoint state, for chaining segments (zero on the first call; on return
 *                            adj1 = dLoss/d u1_in, adj2 = dLoss/d u2_in of the matching wt_forward call)
 *   grad_c, grad_b, grad_rho [Nx,Ny]  OVERWRITTEN with the gradients summed over batch and time
 *                            (grad_b, grad_rho may be NULL; grad_b in linear mode needs WT_F_NEED_GRAD_B
 *                            at forward time)
 *   grad_x      [B,T]        dLoss/dx, or NULL
 */
int wt_backward(const wt_problem* p, const float* c, const float* b, const float* rho,
                const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square,
                const float* grad_probe, const float* probe_raw, const float* grad_fields,
                const void* history, size_t history_bytes, float* adj1, float* adj2, float* grad_c,
                float* grad_b, float* grad_rho, float* grad_x, void* workspace, size_t workspace_bytes,
                void* stream);

/*
 * Row-slab domain decomposition of ONE simulation over the GPUs of a box (BASELINE config 5; the reference has no
 * counterpart).  Every rank runs wt_slab_forward / wt_slab_backward on ITS slab: p->Nx counts the rows it owns plus
 * `halo` ghost rows on each side that borders another rank.  The slab is integrated as an isolated domain for `halo`
 * steps (the error made at its artificial edges moves inwards one row per step, so the owned rows stay exact); then
 * neighbours refresh each other's ghost rows IN THE SAME STREAM, WITHOUT THE HOST: one kernel per exchange stores the
 * `halo` owned rows next to each interior edge, both time levels, straight into the neighbour's ghost rows through
 * NVLink peer-mapped pointers, and synchronises with the neighbours through flag words (st.release.sys / ld.acquire.sys).
 * An exchange also closes every call, so consecutive calls (checkpoint segments) chain without further communication.
 *
 *   halo        steps between exchanges = ghost rows per interior side; a multiple of 8
 *   up, dn      ghost rows above / below my owned rows: 0 at the domain edge, else halo
 *   up_f1/up_f2 address IN THIS PROCESS of the UPPER neighbour's copies of the two state fields the call advances
 *               (u1/u2 for wt_slab_forward, adj1/adj2 for wt_slab_backward), [B, up_Nx, Ny]; dn_*: the lower neighbour's
 *   up_flags, dn_flags   address in this process of the neighbours' flag words (uint32[4], zero before the first call)
 *   flags       my own flag words (same layout: {ready, pushed} written by the upper neighbour, then by the lower one)
 *   state       my own uint32[2], zero before the first call (exchange epoch, block counter)
 * All ranks must issue the same sequence of calls.  The peer mappings come from the caller (cudaIpc / VMM / torch
 * symmetric memory: wavetorch_b200/domain.py); several slabs of one process may also exchange through plain device
 * pointers when each runs on its own stream.  Saturable damping / Kerr terms: forward only (WT_EUNSUPPORTED in
 * wt_slab_backward: the local adjoint coefficients would depend on the inexact ghost fields).
 * The other arguments are those of wt_forward / wt_backward for the local slab (coordinates relative to its first row);
 * u1/u2 resp. adj1/adj2 are required and must be the buffers the neighbours have mapped.
 */
typedef struct wt_slab {
  int32_t halo, up, dn;
  int32_t up_Nx, dn_Nx;     /* rows of the neighbours' slabs */
  int32_t reserved[3];
  uint64_t up_f1, up_f2, dn_f1, dn_f2;
  uint64_t up_flags, dn_flags;
  uint32_t* flags;
  uint32_t* state;
} wt_slab;

int wt_slab_forward(const wt_problem* p, const wt_slab* slab, const float* c, const float* b, const float* rho,
                    const float* x, const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, float* u1,
                    float* u2, float* probe_out, float* probe_raw, void* history, size_t history_bytes, void* workspace,
                    size_t workspace_bytes, void* stream);

int wt_slab_backward(const wt_problem* p, const wt_slab* slab, const float* c, const float* b, const float* rho,
                     const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, const float* grad_probe,
                     const float* probe_raw, const void* history, size_t history_bytes, float* adj1, float* adj2,
                     float* grad_c, float* grad_x, void* workspace, size_t workspace_bytes, void* stream);

/* The exchange on its own (both fields of a [B,Nx,Ny] pair), e.g. to refresh ghost rows after loading a checkpoint. */
int wt_slab_exchange(const wt_slab* slab, int B, int Nx, int Ny, float* f1, float* f2, int device, void* stream);

/*
 * float64 variants of wt_query_plan / wt_forward / wt_backward: the reference's utils.set_dtype('float64') mode
 * (utils.py:14-20).  Same arguments with double fields; one launch per time step on the HBM-streaming kernels (linear and
 * nonlinear cell), tape = every field; no adjoint-state chaining, no dLoss/dfields, no on-chip path.  They also serve as the
 * on-GPU double-precision cross-check of the float32 kernels (tests/test_gpu_parity.py::test_float64_*).
 */
int wt_query_plan_f64(const wt_problem* p, wt_plan* plan);
int wt_forward_f64(const wt_problem* p, const double* c, const double* b, const double* rho, const double* x,
                   const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, double* u1, double* u2,
                   double* probe_out, double* probe_raw, double* fields_out, void* history, size_t history_bytes,
                   void* workspace, size_t workspace_bytes, void* stream);
int wt_backward_f64(const wt_problem* p, const double* c, const double* b, const double* rho, const int32_t* src_ij,
                    const int32_t* prb_ij, const int32_t* prb_square, const double* grad_probe, const double* probe_raw,
                    const void* history, size_t history_bytes, double* grad_c, double* grad_b, double* grad_rho, double* grad_x,
                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * One leapfrog step without sources or probes: y = TimeStep.apply(b, c, y1, y2, dt, h).
 * b and c are [Nx,Ny] (b_batched / c_batched = 0) or [B,Nx,Ny] (= 1).  Uses p->Nx, Ny, B, dt, h, device.
 */
int wt_step_forward(const wt_problem* p, const float* b, int b_batched, const float* c, int c_batched,
                    const float* y1, const float* y2, float* y, void* stream);

/*
 * TimeStep.backward: per-sample gradients, all [B,Nx,Ny]; any output may be NULL (needs_input_grad = False).
 */
int wt_step_backward(const wt_problem* p, const float* b, int b_batched, const float* c, int c_batched,
                     const float* y1, const float* y2, const float* grad_y, float* grad_b, float* grad_c,
                     float* grad_y1, float* grad_y2, void* stream);

/*
 * Geometry parameterisation (the step before and after the loop): c = c0 + (c1-c0) * proj(blur^passes(rho)) and its
 * reverse -- replaces WaveGeometryFreeForm._apply_blur / _apply_projection / .c (geom.py:207-233) and the autograd graph
 * PyTorch builds from them.  taps: [(2*radius+1)^2] normalised disk stencil (the module's blur_kernel buffer);
 * eta, beta, c0, c1: the module's 0-dim float32 device buffers (read on the device, no host synchronisation);
 * blurred: [passes,Nx,Ny] out, every intermediate field (the last one is needed by wt_geom_backward).
 */
int wt_geom_forward(int Nx, int Ny, int radius, int passes, const float* rho, const float* taps, const float* eta,
                    const float* beta, const float* c0, const float* c1, float* blurred, float* c_out, int device,
                    void* stream);

/* grad_rho [Nx,Ny] = d(sum(grad_c * c))/d rho; scratch: [2,Nx,Ny]. */
int wt_geom_backward(int Nx, int Ny, int radius, int passes, const float* blurred_last, const float* grad_c,
                     const float* taps, const float* eta, const float* beta, const float* c0, const float* c1,
                     float* grad_rho, float* scratch, int device, void* stream);

/*
 * Loss head of the classifier (the step after the loop and before its adjoint): replaces
 *     loss = CrossEntropyLoss()(normalize_power(model(x).sum(dim=1)), labels)      train.py:61-62, utils.py:35-36
 * and its autograd graph.
 *   probe_out [B,T,P]  the loop output;  labels [B] int64 class indices in [0,P)
 *   B_total            samples the mean runs over (this call's B of them; 0 = B) -- a batch shard passes the global batch
 *   loss      [1]      sum of this call's cross entropies / B_total;  y_pred [B,P] normalised probe powers (nullable)
 *   dlds      [B,P]    out: dLoss/d(sum_t probe_out[b,:,p]), consumed by wt_loss_backward
 *   scratch   [B]
 * wt_loss_backward: grad_probe[b,t,p] = grad_loss * dlds[b,p]   (grad_loss: device scalar, NULL = 1).
 */
int wt_loss_forward(int B, int T, int P, int B_total, const float* probe_out, const int64_t* labels, float* loss,
                    float* y_pred, float* dlds, float* scratch, int device, void* stream);
int wt_loss_backward(int B, int T, int P, const float* dlds, const float* grad_loss, float* grad_probe, int device,
                     void* stream);

/*
 * Sum-all-reduce of a small float vector across the GPUs of one box through NVLink peer memory, in one kernel launch:
 * the collective of the batch-sharded path (dLoss/dc of every rank's waveforms), replacing ncclAllReduce for this message.
 *   world, rank     ranks on this box (<= 16) and mine
 *   src [n]         my contribution (multiplied by `scale` on the way out);  out [n]: the sum, bitwise identical on all ranks
 *   peer_base       HOST array [world]: device address, valid in THIS process, of rank r's exchange buffer
 *                   (cudaIpc / VMM mapped by the caller).  Layout of each buffer: float gather[2][world][nmax], then
 *                   uint32 flags[world] at byte offset flags_offset_bytes; all zero before the first call
 *   state           my own device uint32[2], zero before the first call (epoch, block counter)
 * Every rank must call it the same number of times in the same order.  The kernel waits on the device for the peers; it
 * is safe to capture in a CUDA graph.
 */
int wt_peer_allreduce(int world, int rank, int n, int nmax, float scale, const float* src, float* out,
                      const uint64_t* peer_base, uint64_t flags_offset_bytes, uint32_t* state, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVETORCH_B200_H */
