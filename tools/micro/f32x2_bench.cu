// Microbenchmark: issue cost and latency of the packed FP32 instructions of sm_100 (FFMA2 / FADD2) against FFMA / FADD.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/f32x2_bench.cu -o gpurun_out/f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool PACKED>
__global__ void k(float* out, int iters, float seed) {
  float2 a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
  const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(seed, -seed);
  for (int t = 0; t < iters; ++t) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
      else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, bool PACKED>
void run(int warps, const char* name) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<ILP, PACKED><<<148, warps * 32>>>(out, 100, 1.f);
  cudaEventRecord(e0);
  k<ILP, PACKED><<<148, warps * 32>>>(out, iters, 1.f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // cycles per loop iteration per SM at 1.965 GHz; FMAs per iteration per warp = 2*ILP (per lane)
  double cyc = ms * 1e-3 * 1.965e9 / iters;
  printf("%-8s ILP=%2d warps/SM=%2d : %7.1f cycles per iteration  -> %.2f lane-FMA pairs per cycle per SMSP\n", name, ILP, warps, cyc,
         (double)ILP * warps / 4 / cyc);
  cudaFree(out);
}

int main() {
  for (int w : {4, 8, 12, 16, 32}) {
    if (w == 4) { run<1, false>(w, "scalar"); run<1, true>(w, "packed"); }
    run<8, false>(w, "scalar"); run<8, true>(w, "packed");
  }
  return 0;
}
