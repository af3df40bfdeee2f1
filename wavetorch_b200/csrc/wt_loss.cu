// Loss head of the vowel classifier, fused: the step right after the time loop and right before its
// adjoint.  Replaces, in one launch each way,
//     yb_pred = normalize_power(model(xb).sum(dim=1));  loss = CrossEntropyLoss()(yb_pred, labels)     train.py:61-62
//     normalize_power(X) = X / sum(X, dim=1, keepdim=True)                                               utils.py:35-36
// and the autograd graph behind them.  dLoss/dprobe_out[b,t,p] does not depend on t, so the backward is a broadcast of a
// [B,P] table computed already in the forward launch.
#include "wt_common.cuh"

namespace wt {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_P = 64;

// One block per sample.  Thread i < used (used = largest multiple of P <= blockDim) walks the contiguous [T*P] slab of its
// sample with stride `used`, so it always sees probe i % P; partials are combined in a fixed order (bitwise reproducible).
__global__ void __launch_bounds__(LOSS_THREADS) k_loss_head(const float* __restrict__ out, const int64_t* __restrict__ labels,
                                                            int T, int P, float inv_B, float* __restrict__ y_pred,
                                                            float* __restrict__ loss_b, float* __restrict__ dlds) {
  __shared__ float part[LOSS_THREADS];
  __shared__ double s[LOSS_MAX_P];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int used = (LOSS_THREADS / P) * P;
  const float* o = out + (size_t)b * T * P;
  float acc = 0.f;
  if (tid < used)
    for (int i = tid; i < T * P; i += used) acc += o[i];
  part[tid] = acc;
  __syncthreads();
  if (tid < P) {
    double a = 0.0;
    for (int i = tid; i < used; i += P) a += (double)part[i];
    s[tid] = a;
  }
  __syncthreads();
  if (tid == 0) {
    double S = 0.0;
    for (int p = 0; p < P; ++p) S += s[p];
    float n[LOSS_MAX_P];
    float mx = -INFINITY;
    for (int p = 0; p < P; ++p) {
      n[p] = (float)s[p] / (float)S;          // normalize_power in float32, like the reference
      mx = fmaxf(mx, n[p]);
    }
    float z = 0.f;
    for (int p = 0; p < P; ++p) z += expf(n[p] - mx);
    const float lse = mx + logf(z);
    const int64_t y = labels[b];
    const int yy = (int)min(max(y, (int64_t)0), (int64_t)(P - 1));
    loss_b[b] = (y < 0 || y >= P) ? NAN : lse - n[yy];   // a label outside [0,P) poisons the loss instead of being ignored
    // dLoss/dn_p = (softmax_p - [p == y]) / B;   dLoss/ds_p = (dLoss/dn_p - sum_q dLoss/dn_q * n_q) / S
    float dn[LOSS_MAX_P];
    float dot = 0.f;
    for (int p = 0; p < P; ++p) {
      dn[p] = (expf(n[p] - lse) - (p == yy ? 1.f : 0.f)) * inv_B;
      dot = fmaf(dn[p], n[p], dot);
    }
    for (int p = 0; p < P; ++p) {
      if (y_pred) y_pred[(size_t)b * P + p] = n[p];
      dlds[(size_t)b * P + p] = (dn[p] - dot) / (float)S;
    }
  }
}

// mean over the batch, fixed order
__global__ void k_loss_mean(const float* __restrict__ loss_b, int B, int B_total, float* __restrict__ loss) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int b = 0; b < B; ++b) a += (double)loss_b[b];
    *loss = (float)(a / B_total);
  }
}

// grad_probe[b,t,p] = grad_loss * dlds[b,p]
__global__ void k_loss_seed(const float* __restrict__ dlds, const float* __restrict__ grad_loss, int T, int P, size_t n,
                            float* __restrict__ grad_probe) {
  const float g = grad_loss ? *grad_loss : 1.f;
  const size_t TP = (size_t)T * P;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / TP;
    const int p = (int)((i - b * TP) % P);
    grad_probe[i] = g * dlds[b * P + p];
  }
}

}  // namespace wt

using namespace wt;

extern "C" {

int wt_loss_forward(int B, int T, int P, int B_total, const float* probe_out, const int64_t* labels, float* loss,
                    float* y_pred, float* dlds, float* scratch, int device, void* stream) {
  WT_REQUIRE(B > 0 && T > 0 && P > 0 && P <= LOSS_MAX_P, "wt_loss_forward: bad shape (B=%d T=%d P=%d, P <= %d)", B, T, P,
             LOSS_MAX_P);
  WT_REQUIRE(probe_out && labels && loss && dlds && scratch, "wt_loss_forward: NULL argument");
  WT_CUDA(cudaSetDevice(device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B_total <= 0) B_total = B;
  k_loss_head<<<B, LOSS_THREADS, 0, st>>>(probe_out, labels, T, P, 1.f / (float)B_total, y_pred, scratch, dlds);
  k_loss_mean<<<1, 32, 0, st>>>(scratch, B, B_total, loss);
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

int wt_loss_backward(int B, int T, int P, const float* dlds, const float* grad_loss, float* grad_probe, int device,
                     void* stream) {
  WT_REQUIRE(B > 0 && T > 0 && P > 0 && P <= LOSS_MAX_P, "wt_loss_backward: bad shape");
  WT_REQUIRE(dlds && grad_probe, "wt_loss_backward: NULL argument");
  WT_CUDA(cudaSetDevice(device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * T * P;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > 148u * 8u) blocks = 148u * 8u;
  k_loss_seed<<<blocks, 256, 0, st>>>(dlds, grad_loss, T, P, n, grad_probe);
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

}  // extern "C"
