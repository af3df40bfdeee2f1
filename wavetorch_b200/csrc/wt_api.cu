// extern "C" entry points (include/wavetorch_b200.h): argument validation, planning, dispatch.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "wt_common.cuh"
#include "wt_resident.h"
#include "wt_slab.h"
#include "wt_stream.h"
#include "wt_tile.h"

namespace wt {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int check_problem(const wt_problem* p) {
  WT_REQUIRE(p != nullptr, "wt_problem is NULL");
  WT_REQUIRE(p->Nx >= 1 && p->Ny >= 1, "bad grid %dx%d", p->Nx, p->Ny);
  WT_REQUIRE(p->B >= 1, "bad batch %d", p->B);
  WT_REQUIRE(p->T >= 0, "bad T %d", p->T);
  WT_REQUIRE(p->n_src >= 0 && p->n_prb >= 0, "negative source/probe count");
  WT_REQUIRE(p->dt > 0 && p->h > 0, "dt and h must be positive (dt=%g, h=%g)", p->dt, p->h);
  WT_REQUIRE(!(p->b0 > 0) || p->uth != 0, "saturable damping needs uth != 0");
  WT_REQUIRE((size_t)p->Nx * p->Ny < (1u << 31), "grid too large for 32-bit cell offsets");
  return WT_OK;
}

static int device_props(int device, cudaDeviceProp* prop) {
  static thread_local int cached_dev = -1;
  static thread_local cudaDeviceProp cached;
  if (cached_dev != device) {
    WT_CUDA(cudaGetDeviceProperties(&cached, device));
    cached_dev = device;
  }
  *prop = cached;
  if (prop->major != 10) {
    set_error("wavetorch_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop->major, prop->minor);
    return WT_EUNSUPPORTED;
  }
  return WT_OK;
}

static int make_plan(const wt_problem* p, bool need_adjoint, bool need_general, wt_plan* plan) {
  WT_TRY(check_problem(p));
  cudaDeviceProp prop;
  WT_TRY(device_props(p->device, &prop));
  memset(plan, 0, sizeof(*plan));
  plan->nonlinear = nonlinear_mask(p);
  const char* env = getenv("WT_FORCE_PATH");
  bool force_stream = (p->flags & WT_F_FORCE_STREAM) || (env && env[0] == 's');
  bool force_res = (p->flags & WT_F_FORCE_RESIDENT) || (env && env[0] == 'r');
  if (!force_stream && !need_general && p->T >= 1 && resident_plan(p, prop, need_adjoint, plan)) return WT_OK;
  if (force_res) {
    int nlm = plan->nonlinear;
    set_error("problem %dx%d B=%d cannot run on the resident path (nonlinear=%d, n_prb=%d)", p->Nx, p->Ny, p->B, nlm,
              p->n_prb);
    return WT_EUNSUPPORTED;
  }
  plan->path = WT_PATH_STREAM;
  plan->cluster = 1;
  plan->threads = 128;
  plan->history_bytes = stream_tape_bytes(p);
  plan->workspace_fwd_bytes = stream_ws_fwd_bytes(p);
  plan->workspace_bwd_bytes = stream_ws_bwd_bytes(p);
  const bool general = plan->nonlinear || (p->flags & WT_F_NEED_GRAD_B);
  plan->launches_fwd = 2 * p->T + 4;
  plan->launches_bwd = (general ? 3 : 2) * p->T + 6;
  if (tile_eligible(p)) {   // temporally blocked kernels (wt_tile.cu); dLoss/dfields requests still go per step
    plan->launches_fwd = tile_launches_fwd(p) + 3;
    plan->launches_bwd = tile_launches_bwd(p) + 3;
  }
  return WT_OK;
}

}  // namespace wt

using namespace wt;

extern "C" {

int wt_abi_version(void) { return WT_ABI_VERSION; }

const char* wt_last_error(void) { return wt::g_err; }

int wt_query_plan(const wt_problem* p, wt_plan* plan) {
  WT_REQUIRE(plan != nullptr, "plan is NULL");
  return make_plan(p, true, false, plan);
}

int wt_validate_pixels(const wt_problem* p, const int32_t* src_ij_host, const int32_t* prb_ij_host, int32_t* max_listings) {
  WT_TRY(check_problem(p));
  WT_REQUIRE(p->n_src == 0 || src_ij_host, "wt_validate_pixels: src_ij_host is NULL");
  WT_REQUIRE(p->n_prb == 0 || prb_ij_host, "wt_validate_pixels: prb_ij_host is NULL");
  for (int k = 0; k < p->n_src; ++k)
    WT_REQUIRE(src_ij_host[2 * k] >= 0 && src_ij_host[2 * k] < p->Nx && src_ij_host[2 * k + 1] >= 0 && src_ij_host[2 * k + 1] < p->Ny,
               "source pixel %d = (%d, %d) lies outside the %dx%d grid", k, src_ij_host[2 * k], src_ij_host[2 * k + 1], p->Nx, p->Ny);
  for (int k = 0; k < p->n_prb; ++k)
    WT_REQUIRE(prb_ij_host[2 * k] >= 0 && prb_ij_host[2 * k] < p->Nx && prb_ij_host[2 * k + 1] >= 0 && prb_ij_host[2 * k + 1] < p->Ny,
               "probe pixel %d = (%d, %d) lies outside the %dx%d grid", k, prb_ij_host[2 * k], prb_ij_host[2 * k + 1], p->Nx, p->Ny);
  if (max_listings) {
    std::vector<long long> cells((size_t)p->n_src);
    for (int k = 0; k < p->n_src; ++k) cells[k] = (long long)src_ij_host[2 * k] * p->Ny + src_ij_host[2 * k + 1];
    std::sort(cells.begin(), cells.end());
    int best = 0, run = 0;
    for (int k = 0; k < p->n_src; ++k) {
      run = (k > 0 && cells[k] == cells[k - 1]) ? run + 1 : 1;
      if (run > best) best = run;
    }
    *max_listings = best;
  }
  return WT_OK;
}

// wt_forward / wt_slab_forward (slab != NULL: streaming kernels on this rank's slab + in-stream ghost-row exchanges)
static int forward_impl(const wt_problem* p_in, const wt_slab* slab, const float* c, const float* b, const float* rho,
                        const float* x, const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, float* u1,
                        float* u2, float* probe_out, float* probe_raw, float* fields_out, void* history,
                        size_t history_bytes, void* workspace, size_t workspace_bytes, void* stream) {
  wt_problem pl;
  const wt_problem* p = p_in;
  if (slab) {
    WT_REQUIRE(p_in != nullptr, "wt_problem is NULL");
    pl = *p_in;
    pl.flags |= WT_F_FORCE_STREAM;
    p = &pl;
    WT_TRY(slab_check(slab, p->B, p->Nx, p->Ny));
  }
  wt_plan plan;
  WT_TRY(make_plan(p, true, false, &plan));
  WT_REQUIRE(c && b && x && u1 && u2, "wt_forward: c, b, x, u1, u2 must not be NULL");
  WT_REQUIRE(p->field_every >= 0 && (p->field_every <= 1 || !history),
             "wt_forward: field_every=%d (time-decimated fields_out) is a forward-only mode", p->field_every);
  WT_REQUIRE(!plan.nonlinear || rho, "wt_forward: rho is required when b0 > 0 or c_nl != 0");
  WT_REQUIRE(p->n_src == 0 || src_ij, "wt_forward: src_ij is NULL");
  WT_REQUIRE(p->n_prb == 0 || (prb_ij && prb_square), "wt_forward: prb_ij / prb_square is NULL");
  WT_REQUIRE(workspace || plan.workspace_fwd_bytes == 0, "wt_forward: workspace is NULL");
  if (workspace_bytes < plan.workspace_fwd_bytes) {
    set_error("wt_forward: workspace %zu < %llu bytes", workspace_bytes, (unsigned long long)plan.workspace_fwd_bytes);
    return WT_ENOSPACE;
  }
  if (history && history_bytes < plan.history_bytes) {
    set_error("wt_forward: history %zu < %llu bytes", history_bytes, (unsigned long long)plan.history_bytes);
    return WT_ENOSPACE;
  }
  WT_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)history & 15) == 0, "workspace/history must be 16-byte aligned");
  WT_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p->T == 0) return WT_OK;
  if (plan.path == WT_PATH_RESIDENT)
    return resident_forward(p, plan, c, b, rho, x, src_ij, prb_ij, prb_square, u1, u2, probe_out, probe_raw, fields_out,
                            history, workspace, st);
  return stream_forward(p, c, b, rho, x, src_ij, prb_ij, prb_square, u1, u2, probe_out, probe_raw, fields_out, history,
                        workspace, st, slab);
}

int wt_forward(const wt_problem* p, const float* c, const float* b, const float* rho, const float* x,
               const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, float* u1, float* u2,
               float* probe_out, float* probe_raw, float* fields_out, void* history, size_t history_bytes,
               void* workspace, size_t workspace_bytes, void* stream) {
  return forward_impl(p, nullptr, c, b, rho, x, src_ij, prb_ij, prb_square, u1, u2, probe_out, probe_raw, fields_out,
                      history, history_bytes, workspace, workspace_bytes, stream);
}

int wt_slab_forward(const wt_problem* p, const wt_slab* slab, const float* c, const float* b, const float* rho,
                    const float* x, const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, float* u1,
                    float* u2, float* probe_out, float* probe_raw, void* history, size_t history_bytes, void* workspace,
                    size_t workspace_bytes, void* stream) {
  WT_REQUIRE(slab != nullptr, "wt_slab_forward: slab is NULL");
  return forward_impl(p, slab, c, b, rho, x, src_ij, prb_ij, prb_square, u1, u2, probe_out, probe_raw, nullptr, history,
                      history_bytes, workspace, workspace_bytes, stream);
}

int wt_slab_exchange(const wt_slab* slab, int B, int Nx, int Ny, float* f1, float* f2, int device, void* stream) {
  WT_REQUIRE(slab && f1 && f2 && B >= 1 && Nx >= 1 && Ny >= 1, "wt_slab_exchange: bad argument");
  WT_TRY(slab_check(slab, B, Nx, Ny));
  WT_CUDA(cudaSetDevice(device));
  return slab_exchange(slab, B, Nx, Ny, f1, f2, reinterpret_cast<cudaStream_t>(stream));
}

static int backward_impl(const wt_problem* p_in, const wt_slab* slab, const float* c, const float* b, const float* rho,
                         const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, const float* grad_probe,
                         const float* probe_raw, const float* grad_fields, const void* history, size_t history_bytes,
                         float* adj1, float* adj2, float* grad_c, float* grad_b, float* grad_rho, float* grad_x,
                         void* workspace, size_t workspace_bytes, void* stream) {
  wt_problem pl;
  const wt_problem* p = p_in;
  if (slab) {
    WT_REQUIRE(p_in != nullptr, "wt_problem is NULL");
    pl = *p_in;
    pl.flags |= WT_F_FORCE_STREAM;
    p = &pl;
    WT_TRY(slab_check(slab, p->B, p->Nx, p->Ny));
    WT_REQUIRE(adj1 && adj2, "wt_slab_backward: adj1/adj2 are required (they are what the neighbours exchange)");
    if (nonlinear_mask(p) || (p->flags & WT_F_NEED_GRAD_B)) {
      set_error("wt_slab_backward: saturable damping / Kerr terms / grad_b are not supported under domain decomposition");
      return WT_EUNSUPPORTED;
    }
  }
  wt_plan plan;
  WT_TRY(make_plan(p, true, false, &plan));
  WT_REQUIRE(c && b && history && grad_c, "wt_backward: c, b, history, grad_c must not be NULL");
  WT_REQUIRE(!plan.nonlinear || rho, "wt_backward: rho is required when b0 > 0 or c_nl != 0");
  WT_REQUIRE(p->n_prb == 0 || (prb_ij && prb_square && grad_probe && probe_raw), "wt_backward: probe arrays missing");
  WT_REQUIRE((adj1 == nullptr) == (adj2 == nullptr), "wt_backward: adj1 and adj2 must both be given or both be NULL");
  if (plan.path == WT_PATH_RESIDENT && (grad_fields || adj1)) {
    set_error("wt_backward: dLoss/dfields and adjoint-state chaining need the streaming path (set WT_F_FORCE_STREAM "
              "for the forward call as well)");
    return WT_EUNSUPPORTED;
  }
  if (workspace_bytes < plan.workspace_bwd_bytes) {
    set_error("wt_backward: workspace %zu < %llu bytes", workspace_bytes, (unsigned long long)plan.workspace_bwd_bytes);
    return WT_ENOSPACE;
  }
  if (history_bytes < plan.history_bytes) {
    set_error("wt_backward: history %zu < %llu bytes", history_bytes, (unsigned long long)plan.history_bytes);
    return WT_ENOSPACE;
  }
  WT_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)history & 15) == 0, "workspace/history must be 16-byte aligned");
  WT_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t plane = (size_t)p->Nx * p->Ny;
  if (p->T == 0) {
    WT_CUDA(cudaMemsetAsync(grad_c, 0, plane * sizeof(float), st));
    if (grad_b) WT_CUDA(cudaMemsetAsync(grad_b, 0, plane * sizeof(float), st));
    if (grad_rho) WT_CUDA(cudaMemsetAsync(grad_rho, 0, plane * sizeof(float), st));
    return WT_OK;
  }
  if (plan.path == WT_PATH_RESIDENT)
    return resident_backward(p, plan, c, b, rho, src_ij, prb_ij, prb_square, grad_probe, probe_raw, history, grad_c, grad_b,
                             grad_rho, grad_x, workspace, st);
  return stream_backward(p, c, b, rho, src_ij, prb_ij, prb_square, grad_probe, probe_raw, grad_fields, history, adj1,
                         adj2, grad_c, grad_b, grad_rho, grad_x, workspace, st, slab);
}

int wt_backward(const wt_problem* p, const float* c, const float* b, const float* rho, const int32_t* src_ij,
                const int32_t* prb_ij, const int32_t* prb_square, const float* grad_probe, const float* probe_raw,
                const float* grad_fields, const void* history, size_t history_bytes, float* adj1, float* adj2,
                float* grad_c, float* grad_b, float* grad_rho, float* grad_x, void* workspace, size_t workspace_bytes,
                void* stream) {
  return backward_impl(p, nullptr, c, b, rho, src_ij, prb_ij, prb_square, grad_probe, probe_raw, grad_fields, history,
                       history_bytes, adj1, adj2, grad_c, grad_b, grad_rho, grad_x, workspace, workspace_bytes, stream);
}

int wt_slab_backward(const wt_problem* p, const wt_slab* slab, const float* c, const float* b, const float* rho,
                     const int32_t* src_ij, const int32_t* prb_ij, const int32_t* prb_square, const float* grad_probe,
                     const float* probe_raw, const void* history, size_t history_bytes, float* adj1, float* adj2,
                     float* grad_c, float* grad_x, void* workspace, size_t workspace_bytes, void* stream) {
  WT_REQUIRE(slab != nullptr, "wt_slab_backward: slab is NULL");
  return backward_impl(p, slab, c, b, rho, src_ij, prb_ij, prb_square, grad_probe, probe_raw, nullptr, history, history_bytes,
                       adj1, adj2, grad_c, nullptr, nullptr, grad_x, workspace, workspace_bytes, stream);
}

int wt_step_forward(const wt_problem* p, const float* b, int b_batched, const float* c, int c_batched,
                    const float* y1, const float* y2, float* y, void* stream) {
  WT_TRY(check_problem(p));
  cudaDeviceProp prop;
  WT_TRY(device_props(p->device, &prop));
  WT_REQUIRE(b && c && y1 && y2 && y, "wt_step_forward: NULL argument");
  WT_CUDA(cudaSetDevice(p->device));
  return step_forward(p, b, b_batched, c, c_batched, y1, y2, y, reinterpret_cast<cudaStream_t>(stream));
}

int wt_step_backward(const wt_problem* p, const float* b, int b_batched, const float* c, int c_batched,
                     const float* y1, const float* y2, const float* grad_y, float* grad_b, float* grad_c,
                     float* grad_y1, float* grad_y2, void* stream) {
  WT_TRY(check_problem(p));
  cudaDeviceProp prop;
  WT_TRY(device_props(p->device, &prop));
  WT_REQUIRE(b && c && y1 && y2 && grad_y, "wt_step_backward: NULL argument");
  WT_CUDA(cudaSetDevice(p->device));
  return step_backward(p, b, b_batched, c, c_batched, y1, y2, grad_y, grad_b, grad_c, grad_y1, grad_y2,
                       reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
