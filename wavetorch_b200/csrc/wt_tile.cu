// Temporally blocked forward kernel for large grids (BASELINE config 5): K time steps per HBM round trip.
//
// The grid is cut into tiles; a CTA loads the tile plus a halo of K cells on every side ("extended tile") for one
// sample, advances it K steps entirely on-chip -- registers hold each thread's R x 4 patch of u_t, u_{t-1}, a1, a3;
// shared memory holds the current field for neighbour access, exactly like the on-chip kernels of wt_resident.cu --
// and writes back only the tile itself.  The extended tile is integrated as if it were an isolated domain with a zero
// boundary: the error this makes at its rim moves inwards one cell per step and never reaches the tile in K steps
// (the same argument as the slab decomposition in wavetorch_b200/domain.py).  Cells outside the real domain have
// a1 = a3 = 0 and stay exactly zero, which is the reference's conv2d zero padding (operators.py:11).
//
// HBM traffic per cell update: (2*E/O + 2)*4/K bytes for the fields (E/O = extended/owned cell ratio) instead of 12,
// plus 4 for the tape when a gradient is wanted.  Coefficients are loaded once per tile and reused for every sample.
#include "wt_common.cuh"
#include "wt_stream.h"
#include "wt_tile.h"

namespace wt {

struct TileArgs {
  int Nx, Ny, B, T;
  int K;              // halo depth = max steps per launch
  int steps;          // steps advanced by this launch (<= K)
  int t0;             // index of the first new field
  int TH, TW;         // owned tile (TW multiple of 4)
  int EH, EW;         // extended tile = TH + 2K, TW + 2K
  int P4, pitch, runs, nact;
  int tiles_y;        // tiles along the column direction
  int bchunk;
  int n_src, n_prb;
  const float* a1;
  const float* a3;
  const float* U1;    // [B,Nx,Ny] field at t0-1
  const float* U2;    // [B,Nx,Ny] field at t0-2
  float* V1;          // out: field at t0+steps-1
  float* V2;          // out: field at t0+steps-2
  const float* x;     // [B,T]
  const int32_t* src_ij;
  const int32_t* prb_ij;
  const int32_t* prb_sq;
  float* probe_out;
  float* probe_raw;
  float* tape;        // nullable: [T][B][Nx*Ny] slots of L(u_{t-1}); this launch writes slots t0 .. t0+steps-1
};

constexpr int TILE_MAX_PRB = 32;
constexpr int TILE_MAX_K = 8;

template <int R>
__global__ void __launch_bounds__(R <= 2 ? 512 : 384, R <= 2 ? 2 : 1) k_tile_fwd(TileArgs a) {
  extern __shared__ float4 smem4[];
  const int slab = (a.EH + 2) * a.pitch;
  float* fld = reinterpret_cast<float*>(smem4);       // [2][slab], row 0 / EH+1 and the 4-float row pad stay zero
  float* xs = fld + 2 * slab;                          // [TILE_MAX_K]
  int* poff = reinterpret_cast<int*>(xs + TILE_MAX_K); // [TILE_MAX_PRB] smem offset of an owned probe
  int* pid = poff + TILE_MAX_PRB;                      // [TILE_MAX_PRB] its global index
  __shared__ int n_my_prb;

  const int tid = threadIdx.x, NT = blockDim.x;
  const bool active = tid < a.nact;
  const int run = tid / a.P4;
  const int g = tid - run * a.P4;
  const int lr0 = run * R;
  const int tile = blockIdx.x;
  const int ti0 = (tile / a.tiles_y) * a.TH, tj0 = (tile % a.tiles_y) * a.TW;   // owned tile origin
  const int gi0 = ti0 - a.K + lr0;                     // global row / col of my patch (may be outside the domain)
  const int gj0 = tj0 - a.K + 4 * g;
  const size_t plane = (size_t)a.Nx * a.Ny;

  float k1[R][4], k3[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gi = gi0 + r, gj = gj0 + k;
      const bool ok = active && gi >= 0 && gi < a.Nx && gj >= 0 && gj < a.Ny;
      k1[r][k] = ok ? a.a1[(size_t)gi * a.Ny + gj] : 0.f;
      k3[r][k] = ok ? a.a3[(size_t)gi * a.Ny + gj] : 0.f;
    }
  unsigned m1 = 0, m2 = 0, m3 = 0;   // source listings of my cells: >=1, >=2, >=3 (more: handled by repeated adds below)
  if (active)
    for (int s = 0; s < a.n_src; ++s) {
      const int si = a.src_ij[2 * s] - gi0, sj = a.src_ij[2 * s + 1] - gj0;
      if (si >= 0 && si < R && sj >= 0 && sj < 4) {
        const unsigned bit = 1u << (si * 4 + sj);
        if (m2 & bit) m3 |= bit; else if (m1 & bit) m2 |= bit; else m1 |= bit;
      }
    }
  if (tid == 0) {
    int n = 0;
    for (int p = 0; p < a.n_prb && n < TILE_MAX_PRB; ++p) {
      const int pi = a.prb_ij[2 * p] - ti0, pj = a.prb_ij[2 * p + 1] - tj0;
      if (pi >= 0 && pi < a.TH && pj >= 0 && pj < a.TW) {
        poff[n] = (pi + a.K + 1) * a.pitch + 4 + pj + a.K;
        pid[n] = p;
        ++n;
      }
    }
    n_my_prb = n;
  }
  for (int i = tid; i < 2 * slab; i += NT) fld[i] = 0.f;
  __syncthreads();
  const int own = (lr0 + 1) * a.pitch + 4 + 4 * g;
  const bool row_mine[2] = {true, true};
  (void)row_mine;
  // which of my cells belong to the owned tile (the ones written back)
  const bool col_in = (4 * g >= a.K) && (4 * g < a.K + a.TW) && (gj0 < a.Ny);

  const int b_lo = blockIdx.y * a.bchunk, b_hi = min(a.B, b_lo + a.bchunk);
  for (int b = b_lo; b < b_hi; ++b) {
    const float* u1p = a.U1 + (size_t)b * plane;
    const float* u2p = a.U2 + (size_t)b * plane;
    float v[R][4], w[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int gi = gi0 + r;
      const bool rok = active && gi >= 0 && gi < a.Nx;
      if (rok && gj0 >= 0 && gj0 + 3 < a.Ny) {
        const float4 p = *reinterpret_cast<const float4*>(u1p + (size_t)gi * a.Ny + gj0);
        const float4 q = *reinterpret_cast<const float4*>(u2p + (size_t)gi * a.Ny + gj0);
        v[r][0] = p.x; v[r][1] = p.y; v[r][2] = p.z; v[r][3] = p.w;
        w[r][0] = q.x; w[r][1] = q.y; w[r][2] = q.z; w[r][3] = q.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gj = gj0 + k;
          const bool ok = rok && gj >= 0 && gj < a.Ny;
          v[r][k] = ok ? u1p[(size_t)gi * a.Ny + gj] : 0.f;
          w[r][k] = ok ? u2p[(size_t)gi * a.Ny + gj] : 0.f;
        }
      }
    }
    if (tid < a.steps) xs[tid] = a.x[(size_t)b * a.T + a.t0 + tid];
    if (active) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        *reinterpret_cast<float4*>(fld + (lr0 + r + 1) * a.pitch + 4 + 4 * g) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    }
    __syncthreads();

    auto step = [&](float (&cu)[R][4], float (&pr)[R][4], int j) {
      const float* cur = fld + (j & 1) * slab;
      float* nxt = fld + ((j + 1) & 1) * slab;
      if (active) {
        const float* ownp = cur + own;
        const float4 up = *reinterpret_cast<const float4*>(ownp - a.pitch);
        const float4 dn = *reinterpret_cast<const float4*>(ownp + R * a.pitch);
        const float upv[4] = {up.x, up.y, up.z, up.w};
        const float dnv[4] = {dn.x, dn.y, dn.z, dn.w};
        float lap[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float lf = ownp[r * a.pitch - 1], rt = ownp[r * a.pitch + 4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float n = (r == 0) ? upv[k] : cu[r - 1][k];
            const float s = (r == R - 1) ? dnv[k] : cu[r + 1][k];
            const float wv = (k == 0) ? lf : cu[r][k - 1];
            const float e = (k == 3) ? rt : cu[r][k + 1];
            lap[r][k] = fmaf(-4.f, cu[r][k], (n + s) + (wv + e));
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int k = 0; k < 4; ++k) pr[r][k] = wt_update(k1[r][k], k3[r][k], cu[r][k], pr[r][k], lap[r][k]);
        if (m1) {
          const float xv = xs[j];
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (m1 >> (r * 4 + k) & 1u) pr[r][k] += xv;
              if (m2 >> (r * 4 + k) & 1u) pr[r][k] += xv;
              if (m3 >> (r * 4 + k) & 1u) pr[r][k] += xv;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
          *reinterpret_cast<float4*>(nxt + (lr0 + r + 1) * a.pitch + 4 + 4 * g) = make_float4(pr[r][0], pr[r][1], pr[r][2], pr[r][3]);
        if (a.tape && col_in) {
          float* tp = a.tape + ((size_t)(a.t0 + j) * a.B + b) * plane;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int gi = gi0 + r, li = lr0 + r;
            if (li >= a.K && li < a.K + a.TH && gi < a.Nx)
              *reinterpret_cast<float4*>(tp + (size_t)gi * a.Ny + gj0) = make_float4(lap[r][0], lap[r][1], lap[r][2], lap[r][3]);
          }
        }
      }
      __syncthreads();
      if (tid < n_my_prb) {   // probe.py:15/27 on the field that now sits in `nxt`
        const float val = nxt[poff[tid]];
        const size_t o = ((size_t)b * a.T + a.t0 + j) * a.n_prb + pid[tid];
        if (a.probe_raw) a.probe_raw[o] = val;
        if (a.probe_out) a.probe_out[o] = a.prb_sq[pid[tid]] ? val * val : val;
      }
    };
    int j = 0;
    for (; j + 1 < a.steps; j += 2) {
      step(v, w, j);
      step(w, v, j + 1);
    }
    bool latest_in_v = true;
    if (j < a.steps) { step(v, w, j); latest_in_v = false; }
    // write the owned tile back (latest -> V1, previous -> V2)
    if (active && col_in) {
      float* o1 = a.V1 + (size_t)b * plane;
      float* o2 = a.V2 + (size_t)b * plane;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int gi = gi0 + r, li = lr0 + r;
        if (li >= a.K && li < a.K + a.TH && gi < a.Nx) {
          const float4 hi = latest_in_v ? make_float4(v[r][0], v[r][1], v[r][2], v[r][3]) : make_float4(w[r][0], w[r][1], w[r][2], w[r][3]);
          const float4 lo = latest_in_v ? make_float4(w[r][0], w[r][1], w[r][2], w[r][3]) : make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
          *reinterpret_cast<float4*>(o1 + (size_t)gi * a.Ny + gj0) = hi;
          *reinterpret_cast<float4*>(o2 + (size_t)gi * a.Ny + gj0) = lo;
        }
      }
    }
    __syncthreads();   // the next sample re-initialises the slab buffers
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool tile_eligible(const wt_problem* p) {
  const char* env = getenv("WT_NO_TILE");
  if (env && env[0] == '1') return false;
  if (nonlinear_mask(p) || (p->flags & WT_F_NEED_GRAD_B)) return false;
  if (p->Ny % 4 || p->n_prb > TILE_MAX_PRB) return false;
  // worth it only when the fields do not stay in L2 anyway and there are enough tiles to fill the chip
  const char* emin = getenv("WT_TILE_MIN_CELLS");
  const size_t min_cells = emin ? (size_t)atoll(emin) : ((size_t)1 << 22);
  return (size_t)p->Nx * p->Ny * p->B >= min_cells && p->Nx >= 16 && p->Ny >= 16;
}

size_t tile_extra_ws_bytes(const wt_problem* p) {
  return tile_eligible(p) ? (size_t)2 * p->B * p->Nx * p->Ny * sizeof(float) + 256 : 0;
}

struct TileGeom { int K, R, TH, TW, EH, EW, P4, pitch, runs, nact, threads, tiles_x, tiles_y; size_t smem; };

static TileGeom tile_geom(const wt_problem* p) {
  TileGeom t;
  const char* ek = getenv("WT_TILE_K");
  t.K = ek ? atoi(ek) : 4;
  if (t.K != 8) t.K = 4;
  t.R = 4;
  const char* er = getenv("WT_TILE_R");
  if (er) t.R = atoi(er);
  if (t.R != 2 && t.R != 3 && t.R != 4) t.R = 4;
  // extended tile: EH = runs*R rows, EW = 4*P4 columns with runs*P4 <= threads
  t.P4 = 32;                                   // 128 extended columns -> 128 - 2K owned
  t.EW = 4 * t.P4;
  t.runs = (t.R <= 2 ? 512 : 384) / t.P4;      // 16 or 12 row runs (launch bounds of k_tile_fwd)
  t.EH = t.runs * t.R;
  t.TH = t.EH - 2 * t.K;
  t.TW = t.EW - 2 * t.K;
  t.pitch = t.EW + 4;
  t.nact = t.runs * t.P4;
  t.threads = t.nact;
  t.tiles_x = (p->Nx + t.TH - 1) / t.TH;
  t.tiles_y = (p->Ny + t.TW - 1) / t.TW;
  t.smem = (size_t)2 * (t.EH + 2) * t.pitch * 4 + TILE_MAX_K * 4 + 2 * TILE_MAX_PRB * 4 + 64;
  return t;
}

int tile_forward(const wt_problem* p, const float* a1, const float* a3, const float* x, const int32_t* src_ij,
                 const int32_t* prb_ij, const int32_t* prb_sq, float* u1, float* u2, float* probe_out, float* probe_raw,
                 float* tape, float* extra_ws, cudaStream_t st, int* launches) {
  const TileGeom g = tile_geom(p);
  const size_t field = (size_t)p->B * p->Nx * p->Ny;
  float* A1 = u1;  float* A2 = u2;              // current pair
  float* B1 = extra_ws; float* B2 = extra_ws + field;
  TileArgs a = {};
  a.Nx = p->Nx; a.Ny = p->Ny; a.B = p->B; a.T = p->T; a.K = g.K;
  a.TH = g.TH; a.TW = g.TW; a.EH = g.EH; a.EW = g.EW; a.P4 = g.P4; a.pitch = g.pitch; a.runs = g.runs; a.nact = g.nact;
  a.tiles_y = g.tiles_y; a.n_src = p->n_src; a.n_prb = p->n_prb;
  a.a1 = a1; a.a3 = a3; a.x = x; a.src_ij = src_ij; a.prb_ij = prb_ij; a.prb_sq = prb_sq;
  a.probe_out = probe_out; a.probe_raw = probe_raw; a.tape = tape;
  const int ntiles = g.tiles_x * g.tiles_y;
  // batch chunks: enough CTAs for ~4 waves, but keep samples together so a tile's coefficients are reused
  int nby = (148 * 8 + ntiles - 1) / ntiles;
  if (nby < 1) nby = 1;
  if (nby > p->B) nby = p->B;
  a.bchunk = (p->B + nby - 1) / nby;
  nby = (p->B + a.bchunk - 1) / a.bchunk;
  int n = 0;
  for (int t0 = 0; t0 < p->T; t0 += g.K) {
    a.t0 = t0;
    a.steps = p->T - t0 < g.K ? p->T - t0 : g.K;
    a.U1 = A1; a.U2 = A2; a.V1 = B1; a.V2 = B2;
    dim3 grid(ntiles, nby), block(g.threads);
    switch (g.R) {
      case 2:
        WT_CUDA(cudaFuncSetAttribute(k_tile_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        k_tile_fwd<2><<<grid, block, g.smem, st>>>(a);
        break;
      case 3:
        WT_CUDA(cudaFuncSetAttribute(k_tile_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        k_tile_fwd<3><<<grid, block, g.smem, st>>>(a);
        break;
      default:
        WT_CUDA(cudaFuncSetAttribute(k_tile_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        k_tile_fwd<4><<<grid, block, g.smem, st>>>(a);
        break;
    }
    ++n;
    float* s1 = A1; float* s2 = A2; A1 = B1; A2 = B2; B1 = s1; B2 = s2;
  }
  WT_CUDA(cudaGetLastError());
  if (A1 != u1) {   // odd number of launches: the result sits in the workspace pair
    WT_CUDA(cudaMemcpyAsync(u1, A1, field * sizeof(float), cudaMemcpyDeviceToDevice, st));
    WT_CUDA(cudaMemcpyAsync(u2, A2, field * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  if (launches) *launches = n;
  return WT_OK;
}

}  // namespace wt
