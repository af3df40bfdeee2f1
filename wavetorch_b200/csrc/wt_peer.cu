// Gradient all-reduce over NVLink peer memory, one kernel (the only collective of the batch-sharded path).
//
// After the adjoint every rank holds dLoss/dc (and, with the nonlinear terms, the direct dLoss/drho) of ITS waveforms:
// n = Nx*Ny (or 2*Nx*Ny) floats, 60 KB at BASELINE config 3.  For a message this small an NCCL all-reduce is pure
// latency.  Here each rank
//   1. stores its (scaled) vector straight into EVERY peer's gather buffer (P2P stores over NVLink / NVSwitch),
//   2. publishes a per-rank epoch flag on every peer (st.release.sys) once all its blocks have finished storing,
//   3. waits (ld.acquire.sys) until the flags of all ranks have reached the epoch on its own flag array,
//   4. sums the world's vectors from its LOCAL gather buffer in rank order -- every rank adds the same numbers in the
//      same order, so the result is bitwise identical on all ranks and run-to-run.
// Buffers are double-buffered by epoch parity: a rank can start call e+1 only after every peer has entered call e, i.e.
// finished reading the buffers of call e-1.  The epoch lives in device memory, so the kernel can be replayed from a
// CUDA graph.  The peer-mapped buffers come from the caller (torch symmetric memory: wavetorch_b200/peer.py).
#include "wt_common.cuh"

namespace wt {

constexpr int PEER_MAX_WORLD = 16;

struct PeerArgs {
  int world, rank, n, nmax;
  float scale;
  const float* src;
  float* out;
  float* gather[PEER_MAX_WORLD];      // rank r's buffer [2][world][nmax], mapped into this process
  unsigned* flags[PEER_MAX_WORLD];    // rank r's flags [world]
  unsigned* state;                    // local: [0] epoch of the last completed call, [1] blocks that finished storing
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerArgs a) {
  __shared__ unsigned epoch_s;
  if (threadIdx.x == 0) epoch_s = *reinterpret_cast<volatile unsigned*>(a.state) + 1u;
  __syncthreads();
  const unsigned e = epoch_s;
  const size_t slot = (size_t)(e & 1u) * a.world * a.nmax;
  const int stride = gridDim.x * blockDim.x;
  // 1. my vector into everybody's gather[parity][rank]
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const float v = a.src[i] * a.scale;
    for (int r = 0; r < a.world; ++r) a.gather[r][slot + (size_t)a.rank * a.nmax + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  // 2. the last block to get here has seen every block's stores fenced: publish the epoch on every peer
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(a.state + 1, 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      a.state[1] = 0u;
      a.state[0] = e;     // every block has read the old epoch before it incremented the counter
      for (int r = 0; r < a.world; ++r) st_release_sys(a.flags[r] + a.rank, e);
    }
  }
  // 3. wait for the world (epochs only grow; the signed difference tolerates wrap-around)
  if (threadIdx.x < a.world) {
    const unsigned* f = a.flags[a.rank] + threadIdx.x;
    while ((int)(ld_acquire_sys(f) - e) < 0) __nanosleep(64);
  }
  __syncthreads();
  // 4. ordered sum from my own gather buffer
  const float* g = a.gather[a.rank] + slot;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    float s = 0.f;
    for (int r = 0; r < a.world; ++r) s += __ldcv(g + (size_t)r * a.nmax + i);
    a.out[i] = s;
  }
}

}  // namespace wt

using namespace wt;

extern "C" {

int wt_peer_allreduce(int world, int rank, int n, int nmax, float scale, const float* src, float* out,
                      const uint64_t* peer_base, uint64_t flags_offset_bytes, uint32_t* state, int device, void* stream) {
  WT_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "wt_peer_allreduce: bad world/rank %d/%d",
             world, rank);
  WT_REQUIRE(n > 0 && n <= nmax, "wt_peer_allreduce: n=%d exceeds the gather capacity %d", n, nmax);
  WT_REQUIRE(src && out && peer_base && state, "wt_peer_allreduce: NULL argument");
  WT_REQUIRE(flags_offset_bytes >= (uint64_t)2 * world * nmax * sizeof(float) && flags_offset_bytes % 16 == 0,
             "wt_peer_allreduce: flags overlap the gather buffers");
  WT_CUDA(cudaSetDevice(device));
  PeerArgs a = {};
  a.world = world; a.rank = rank; a.n = n; a.nmax = nmax; a.scale = scale; a.src = src; a.out = out; a.state = state;
  for (int r = 0; r < world; ++r) {
    WT_REQUIRE(peer_base[r] != 0, "wt_peer_allreduce: peer %d is not mapped", r);
    a.gather[r] = reinterpret_cast<float*>(peer_base[r]);
    a.flags[r] = reinterpret_cast<unsigned*>(peer_base[r] + flags_offset_bytes);
  }
  int blocks = (n + 1023) / 1024;     // all blocks spin in step 3, so the grid must be co-resident: keep it small
  if (blocks > 64) blocks = 64;
  if (blocks < 1) blocks = 1;
  k_peer_allreduce<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  WT_CUDA(cudaGetLastError());
  return WT_OK;
}

}  // extern "C"
