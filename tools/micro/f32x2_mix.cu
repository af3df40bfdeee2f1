// Microbenchmark: does FFMA2 free issue slots for other pipes?  Per iteration and warp: NF packed (or 2*NF scalar) FMAs
// mixed with NL shared-memory loads and NA integer ops.
#include <cstdio>
#include <cuda_runtime.h>

template <bool PACKED>
__global__ void k(float* out, int iters, float seed) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = seed * i;
  __syncthreads();
  constexpr int ILP = 8;
  float2 a[ILP];
  unsigned h[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = make_float2(seed + i + threadIdx.x, seed - i); h[i] = threadIdx.x * 7 + i; }
  const float2 m = make_float2(1.0000001f, 0.9999999f);
  const float* p = sm + (threadIdx.x & 31);
  for (int t = 0; t < iters; ++t) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const float l = p[((t + i) & 31) * 32];
      const float2 c = make_float2(l, l);
      if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
      else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
      h[i] = (h[i] ^ (h[i] >> 3)) + 0x9e3779b9u;   // LOP3/SHF + IADD on the alu pipe
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y + (float)h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool PACKED>
void run(int warps, const char* name) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<PACKED><<<148, warps * 32>>>(out, 100, 1.f);
  cudaEventRecord(e0);
  k<PACKED><<<148, warps * 32>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("%-8s warps/SM=%2d : %7.1f cycles per iteration (8 LDS + 16 lane-FMAs + ~24 int ops per warp), %.1f per warp and SMSP\n", name, warps, cyc, cyc / (warps / 4.0));
  cudaFree(out);
}

int main() {
  for (int w : {4, 12, 32}) { run<false>(w, "scalar"); run<true>(w, "packed"); }
  return 0;
}
