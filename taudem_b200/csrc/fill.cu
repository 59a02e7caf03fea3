// Pit removal (Planchon-Darboux, epsilon = 0) as tile-local relaxation.
//
// reference: flood() src/flood.cpp:50-526 — init :243-271, first scan :292-343,
//            stack ping-pong :357-479.
//
// The reference lowers W from FLT_MAX with W := min(W, max(z, min_nbr W)) on the
// still-wet cells until nothing changes.  That iteration has a unique fixed point
// (the minimax path elevation to the seed cells = priority-flood fill) and only
// compares/copies input floats, so any monotone schedule that reaches the fixed
// point is bit-identical.  Schedule here: an active list of
// 64 x 32 tiles; each CTA loads its tile (+1-cell ring) of W and z into shared
// memory, relaxes it to convergence there (chaotic in-place updates are safe: every
// value ever written is a valid upper bound and values only decrease), writes it
// back and, if anything changed, queues the 3x3 tile neighbourhood for the next
// round.  Rounds end when no tile changes.
#include "ctx.h"
#include "tile_pipe.cuh"

namespace td {
namespace {
constexpr int FW = 64, FH = 32;
constexpr int SW = FW + 2;                 // shared W row stride (with ring)
#define TD_FELNODATA (-3.0e38f)

// Initialisation (src/flood.cpp:243-271): nodata stays nodata, cells of the depression mask, cells on the edge of the
// grid and cells with a nodata neighbour (8 or 4 neighbours) keep their elevation, everything else starts "under water"
// (FLT_MAX).  A 3x3 stencil on the persistent tile pipeline of the flow-direction stencils (tile_pipe.cuh: three stages of
// 2-D TMA tiles per CTA), four cells per thread, one float4 store.  The nodata test is a minimum of |z - nodata| over the
// window, folded column-wise first (three cells of a column are shared by three windows).  Neighbours outside the grid
// need no test: only edge cells have them and edge cells keep their elevation whatever their window holds (what the TMA
// unit zero-fills there is never used for a decision; fminf drops NaNs).  8 B/cell (+2 with a mask).
constexpr int IW = 128, IH = 32, ISTAGES = 3;
using InitRing = TileRing<float, IW, IH, ISTAGES>;
__global__ void __launch_bounds__(256) k_fill_init(const TD_GRID_CONSTANT TileMap tm, const short* __restrict__ mask,
                                                   float* __restrict__ W, Strip s, float nodata, int step) {
  extern __shared__ __align__(128) unsigned char dsm128[];
  using G = InitRing::G;
  InitRing ring;
  ring.init(dsm128, &tm, s);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < ring.ntiles; t += gridDim.x) {
    int r0, c0;
    const float* tile = ring.acquire(t, r0, c0);
#pragma unroll 1
    for (int pass = 0; pass < IH / 8; ++pass) {
      const int tr = warp + 8 * pass;
      const int r = r0 + tr, c = c0 + lane * 4;
      if (r > s.ny || c >= s.pitch) continue;
      const float* pm = tile + tr * G::SW + G::HP + lane * 4;   // row above, column c
      float a[3][6], z[4];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float* p = pm + j * G::SW;
        const float4 v = *reinterpret_cast<const float4*>(p);
        const float nb[6] = {p[-1], v.x, v.y, v.z, v.w, p[4]};
        if (j == 1) { z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w; }
#pragma unroll
        for (int i = 0; i < 6; ++i) a[j][i] = fabsf(nb[i] - nodata);
      }
      unsigned em = ((r == 1 && !s.has_top) || (r == s.ny && !s.has_bot)) ? 0xfu : 0u;     // cells on the edge of the whole grid
      em |= (c == 0) ? 1u : 0u;
      const int klast = s.nx - 1 - c;
      if (klast < 4) em |= (0xfu << max(klast, 0)) & 0xfu;
      short4 mk = make_short4(0, 0, 0, 0);
      if (mask != nullptr) mk = *reinterpret_cast<const short4*>(mask + s.idx(r, c));
      const short m4[4] = {mk.x, mk.y, mk.z, mk.w};
      // 8 neighbours: the minimum over each window column (the centre's own distance may take part: a nodata centre becomes nodata below)
      float col[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) col[i] = fminf(fminf(a[0][i], a[1][i]), a[2][i]);
      float out[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float dmin;
        if (step == 1) dmin = fminf(fminf(col[i], col[i + 1]), col[i + 2]);
        else dmin = fminf(fminf(a[0][i + 1], a[2][i + 1]), fminf(a[1][i], a[1][i + 2]));
        const bool keep = (m4[i] == 1) || ((em >> i) & 1u) || dmin < TD_MINEPS;
        float v = keep ? z[i] : FLT_MAX;
        if (a[1][i + 1] < TD_MINEPS) v = TD_FELNODATA;
        if (c + i >= s.nx) v = TD_FELNODATA;                       // padding columns
        out[i] = v;
      }
      *reinterpret_cast<float4*>(W + s.idx(r, c)) = make_float4(out[0], out[1], out[2], out[3]);
    }
    ring.release(&tm, t);
  }
}

// flag[t] = last round for which tile t has been queued
__device__ __forceinline__ void queue_tile(int tx, int ty, int ntx, int nty, int next_round, int* __restrict__ flag,
                                           int* __restrict__ next, unsigned long long* __restrict__ ctr) {
  if (tx < 0 || ty < 0 || tx >= ntx || ty >= nty) return;
  const int t = ty * ntx + tx;
  if (atomicExch(flag + t, next_round) != next_round) next[atomicAdd(ctr, 1ull)] = t;
}

__global__ void __launch_bounds__(256) k_fill_relax(const float* __restrict__ dem, float* __restrict__ W, Strip s, int step,
                                                    const int* __restrict__ cur, int ntx, int nty, int round,
                                                    int* __restrict__ flag, int* __restrict__ next,
                                                    unsigned long long* __restrict__ ctr) {
  __shared__ float sw[(FH + 2) * SW];
  __shared__ float sz[FH * FW];
  const int tile = cur[blockIdx.x];
  const int tx = tile % ntx, ty = tile / ntx;
  const int c0 = tx * FW, r0 = 1 + ty * FH;
  const int tid = threadIdx.x;
  // load W with its ring; off-grid cells read as FLT_MAX (the reference skips them: hasAccess)
  for (int i = tid; i < (FH + 2) * SW; i += 256) {
    const int lr = i / SW, lc = i - lr * SW;
    const int r = r0 - 1 + lr, c = c0 - 1 + lc;
    sw[i] = s.on_grid(r, c) ? __ldcg(W + s.idx(r, c)) : FLT_MAX;
  }
  for (int i = tid; i < FH * FW; i += 256) {
    const int lr = i / FW, lc = i - lr * FW;
    const int r = r0 + lr, c = c0 + lc;
    sz[i] = s.owned(r, c) ? dem[s.idx(r, c)] : FLT_MAX;   // cells outside the strip: z = FLT_MAX => never wet
  }
  __syncthreads();
  const int lc = tid & 63, lr0 = tid >> 6;       // 4 row phases x 8 rows
  bool any = false;
  for (;;) {
    bool ch = false;
#pragma unroll
    for (int j = 0; j < FH / 4; ++j) {
      const int lr = lr0 + 4 * j;
      float* p = sw + (lr + 1) * SW + lc + 1;
      const float w = *p, z = sz[lr * FW + lc];
      if (w > z && !nd_f(w, TD_FELNODATA)) {
        float m = fminf(fminf(p[1], p[-1]), fminf(p[-SW], p[SW]));
        if (step == 1) m = fminf(m, fminf(fminf(p[-SW + 1], p[-SW - 1]), fminf(p[SW - 1], p[SW + 1])));
        const float nw = fmaxf(z, m);
        if (nw < w) { *p = nw; ch = true; }
      }
    }
    any = any || ch;
    if (!__syncthreads_or(ch)) break;
  }
  any = __syncthreads_or(any);
  if (!any) return;
  for (int i = tid; i < FH * FW; i += 256) {
    const int lr = i / FW, lc2 = i - lr * FW;
    const int r = r0 + lr, c = c0 + lc2;
    if (s.owned(r, c)) W[s.idx(r, c)] = sw[(lr + 1) * SW + lc2 + 1];
  }
  if (tid < 9) queue_tile(tx + tid % 3 - 1, ty + tid / 3 - 1, ntx, nty, round + 1, flag, next, ctr);
}

__global__ void k_fill_all_tiles(int* list, int* flag, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { list[t] = t; flag[t] = 1; }
}
// only the tiles that see a halo row (after an exchange with the neighbour strips nothing else can have changed)
__global__ void k_fill_edge_tiles(int* list, int* flag, int ntx, int nty) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nedge = nty > 1 ? 2 * ntx : ntx;
  if (i >= ntx * nty) return;
  flag[i] = 0;
  if (i < nedge) { const int t = i < ntx ? i : (nty - 1) * ntx + (i - ntx); list[i] = t; }
}
__global__ void k_fill_edge_flags(const int* list, int* flag, int nedge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nedge) flag[list[i]] = 1;
}
}  // namespace

int fill_init(const float* dem, const short* mask, float* W, const Strip& s, float nodata, int four, cudaStream_t st) {
  TileMap tm;
  if (int rc = make_tile_map(&tm, dem, 4, s.pitch, s.ny + 2, InitRing::G::SW, InitRing::G::ROWS)) return rc;
  const long long ntiles = (long long)((s.pitch + IW - 1) / IW) * ((s.ny + IH - 1) / IH);
  int grid = 0;
  if (int rc = stencil_grid((const void*)k_fill_init, 256, InitRing::SMEM, ntiles, &grid)) return rc;
  k_fill_init<<<grid, 256, InitRing::SMEM, st>>>(tm, mask, W, s, nodata, four ? 2 : 1);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

// Relaxes until no tile of the strip changes.  *changed = whether any cell moved.
int fill_relax(td_ctx* ctx, const float* dem, float* W, const Strip& s, int four, int* changed, cudaStream_t st, bool edges_only) {
  const int ntx = (s.nx + FW - 1) / FW, nty = (s.ny + FH - 1) / FH;
  const long long nt = (long long)ntx * nty;
  TD_CUDA(ctx->tileflags.ensure((size_t)nt * 4 * 3));
  int* flag = ctx->tileflags.as<int>();
  int* la = flag + nt;
  int* lb = la + nt;
  unsigned long long n = (unsigned long long)nt;
  if (edges_only) {
    n = (unsigned long long)(nty > 1 ? 2 * ntx : ntx);
    k_fill_edge_tiles<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(la, flag, ntx, nty);
    TD_LAUNCHED();
    k_fill_edge_flags<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(la, flag, (int)n);
    TD_LAUNCHED();
  } else {
    k_fill_all_tiles<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(la, flag, (int)nt);
    TD_LAUNCHED();
  }
  int round = 1;
  *changed = 0;
  while (n > 0) {
    TD_CUDA(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(unsigned long long), st));
    k_fill_relax<<<(unsigned)n, 256, 0, st>>>(dem, W, s, four ? 2 : 1, la, ntx, nty, round, flag, lb, ctx->d_ctr);
    TD_LAUNCHED();
    TD_CUDA(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    TD_CUDA(cudaStreamSynchronize(st));
    n = ctx->h_ctr[0];
    if (n) *changed = 1;
    std::swap(la, lb);
    ++round;
  }
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}
}  // namespace td
