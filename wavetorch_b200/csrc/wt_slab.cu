// Halo exchange of the row-slab domain decomposition: ONE kernel per exchange, peer stores over NVLink, no host, no NCCL.
//
// Rank g integrates rows [r0-halo, r1+halo) of the grid as an isolated domain for `halo` steps (wavetorch_b200/domain.py,
// include/wavetorch_b200.h: wt_slab).  After those steps its ghost rows are stale and its owned rows are exact, so every
// rank stores the `halo` owned rows next to each interior edge -- both time levels, all samples -- straight into the
// neighbour's ghost rows.  The neighbour's fields are peer-mapped into this process; the stores travel through NVSwitch.
//
// Synchronisation (per exchange, epoch e, flags live in the RECEIVER's memory):
//   1. "ready":  I tell each neighbour that my segment is done -- my owned rows are final and I no longer read my ghost
//                rows -- with st.release.sys of e into its flag word.
//   2. every block waits (ld.acquire.sys on my own flags) until the neighbours are ready, then stores its share of the rows.
//   3. "pushed": the last block to finish fences and publishes e on the neighbours, then waits until both neighbours'
//                "pushed" flags have arrived: the kernel, and with it the stream, proceeds only when my ghost rows are fresh.
// The epoch lives in device memory: the whole time loop, exchanges included, is plain stream work (CUDA-graph capturable).
// Traffic per exchange and neighbour: 2 fields x halo rows x Ny x B x 4 bytes in each direction (4.2 MB at config 5, B = 8),
// against ~1 GB of HBM traffic for the 16 time steps in between.
#include <stdlib.h>

#include "wt_slab.h"

namespace wt {

struct XchgArgs {
  int B, Nx, Ny, halo, up, dn, up_Nx, dn_Nx;
  float* f1; float* f2;
  float* up1; float* up2; float* dn1; float* dn2;
  unsigned* up_flags; unsigned* dn_flags;
  unsigned* flags; unsigned* state;
};

__device__ __forceinline__ void slab_st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned slab_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void slab_wait(const unsigned* p, unsigned e) {
  while ((int)(slab_ld_acquire(p) - e) < 0) __nanosleep(32);   // epochs only grow; signed difference tolerates wrap-around
}

// flags layout (mine): [0] ready, [1] pushed -- written by the UPPER neighbour; [2] ready, [3] pushed -- by the LOWER one
template <int VEC>
__global__ void __launch_bounds__(256) k_slab_exchange(XchgArgs a) {
  __shared__ unsigned epoch_s;
  if (threadIdx.x == 0) epoch_s = *reinterpret_cast<volatile unsigned*>(a.state) + 1u;
  __syncthreads();
  const unsigned e = epoch_s;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.up) slab_st_release(a.up_flags + 2, e);   // I am the upper rank's LOWER neighbour
    if (a.dn) slab_st_release(a.dn_flags + 0, e);
  }
  if (threadIdx.x == 0) {
    if (a.up) slab_wait(a.flags + 0, e);
    if (a.dn) slab_wait(a.flags + 2, e);
  }
  __syncthreads();
  // rows to push: to the upper neighbour my first `halo` owned rows -> its last `halo` rows; to the lower one my last
  // `halo` owned rows -> its first `halo` rows
  const int rowv = a.Ny / VEC;                       // vectors per row
  const long long per_dir = (long long)a.B * a.halo * rowv;
  const long long n_up = a.up ? per_dir : 0, n_dn = a.dn ? per_dir : 0;
  const long long total = 2 * (n_up + n_dn);         // two fields
  typedef typename std::conditional<VEC == 4, float4, float>::type V;
  // element i -> (field, direction, sample, row, column vector); consecutive threads take consecutive vectors of a row
  auto locate = [&](long long i, const V*& sp, V*& dp) {
    const int fld = (int)(i / (n_up + n_dn));
    long long k = i - (long long)fld * (n_up + n_dn);
    const bool to_up = k < n_up;
    if (!to_up) k -= n_up;
    const int col = (int)(k % rowv);
    const int row = (int)((k / rowv) % a.halo);
    const int b = (int)(k / ((long long)rowv * a.halo));
    const float* src = fld ? a.f2 : a.f1;
    if (to_up) {
      sp = reinterpret_cast<const V*>(src + ((size_t)b * a.Nx + a.up + row) * a.Ny) + col;
      dp = reinterpret_cast<V*>((fld ? a.up2 : a.up1) + ((size_t)b * a.up_Nx + (a.up_Nx - a.halo) + row) * a.Ny) + col;
    } else {
      sp = reinterpret_cast<const V*>(src + ((size_t)b * a.Nx + (a.Nx - a.dn - a.halo) + row) * a.Ny) + col;
      dp = reinterpret_cast<V*>((fld ? a.dn2 : a.dn1) + ((size_t)b * a.dn_Nx + row) * a.Ny) + col;
    }
  };
  // four independent loads in flight per thread before the first remote store: NVLink needs megabytes in flight
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 4 * stride) {
    const V* sp[4];
    V* dp[4];
    V val[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < total) { locate(i + u * stride, sp[u], dp[u]); val[u] = *sp[u]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u * stride < total) *dp[u] = val[u];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.state + 1, 1u);
    if (prev == gridDim.x - 1) {        // every block's stores are fenced: publish, then wait for my own ghost rows
      __threadfence_system();
      a.state[1] = 0u;
      a.state[0] = e;
      if (a.up) slab_st_release(a.up_flags + 3, e);
      if (a.dn) slab_st_release(a.dn_flags + 1, e);
      if (a.up) slab_wait(a.flags + 1, e);
      if (a.dn) slab_wait(a.flags + 3, e);
    }
  }
}

int slab_check(const wt_slab* s, int B, int Nx, int Ny) {
  WT_REQUIRE(s->halo >= 8 && s->halo % 8 == 0, "wt_slab: halo=%d must be a positive multiple of 8", s->halo);
  WT_REQUIRE((s->up == 0 || s->up == s->halo) && (s->dn == 0 || s->dn == s->halo), "wt_slab: up/dn must be 0 or halo");
  WT_REQUIRE(Nx - s->up - s->dn >= s->halo, "wt_slab: a slab must own at least halo=%d rows (owns %d)", s->halo,
             Nx - s->up - s->dn);
  WT_REQUIRE(s->flags && s->state, "wt_slab: flags/state are NULL");
  if (s->up) WT_REQUIRE(s->up_f1 && s->up_f2 && s->up_flags && s->up_Nx >= 2 * s->halo, "wt_slab: upper neighbour not mapped");
  if (s->dn) WT_REQUIRE(s->dn_f1 && s->dn_f2 && s->dn_flags && s->dn_Nx >= 2 * s->halo, "wt_slab: lower neighbour not mapped");
  return WT_OK;
}

int slab_exchange(const wt_slab* s, int B, int Nx, int Ny, float* f1, float* f2, cudaStream_t st) {
  if (!s || (!s->up && !s->dn)) return WT_OK;
  // WT_SLAB_SKIP=1 (measurement only, results are WRONG): leave the exchange out to see what it costs in situ
  static const bool skip = [] { const char* e = getenv("WT_SLAB_SKIP"); return e && e[0] == '1'; }();
  if (skip) return WT_OK;
  XchgArgs a = {};
  a.B = B; a.Nx = Nx; a.Ny = Ny; a.halo = s->halo; a.up = s->up; a.dn = s->dn; a.up_Nx = s->up_Nx; a.dn_Nx = s->dn_Nx;
  a.f1 = f1; a.f2 = f2;
  a.up1 = reinterpret_cast<float*>(s->up_f1); a.up2 = reinterpret_cast<float*>(s->up_f2);
  a.dn1 = reinterpret_cast<float*>(s->dn_f1); a.dn2 = reinterpret_cast<float*>(s->dn_f2);
  a.up_flags = reinterpret_cast<unsigned*>(s->up_flags); a.dn_flags = reinterpret_cast<unsigned*>(s->dn_flags);
  a.flags = s->flags; a.state = s->state;
  const bool v4 = Ny % 4 == 0 && !(((uintptr_t)f1 | (uintptr_t)f2 | s->up_f1 | s->up_f2 | s->dn_f1 | s->dn_f2) & 15);
  const long long total = 2LL * B * s->halo * (Ny / (v4 ? 4 : 1)) * ((s->up ? 1 : 0) + (s->dn ? 1 : 0));
  int blocks = (int)((total + 256 * 8 - 1) / (256 * 8));   // all blocks wait in step 2: keep the grid co-resident (<= 1 per SM)
  if (blocks > 128) blocks = 128;
  if (blocks < 1) blocks = 1;
  if (v4) k_slab_exchange<4><<<blocks, 256, 0, st>>>(a);
  else k_slab_exchange<1><<<blocks, 256, 0, st>>>(a);
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

}  // namespace wt
