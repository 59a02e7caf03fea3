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

namespace td {
namespace {
constexpr int FW = 64, FH = 32;
constexpr int SW = FW + 2;                 // shared W row stride (with ring)
#define TD_FELNODATA (-3.0e38f)

__global__ void __launch_bounds__(256) k_fill_init(const float* __restrict__ dem, const short* __restrict__ mask,
                                                   float* __restrict__ W, Strip s, float nodata, int step) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x, r = 1 + blockIdx.x;   // rows on grid.x (no 65535 limit)
  if (c >= s.nx) return;
  const long long ci = s.idx(r, c);
  const float z = dem[ci];
  float out;
  if (nd_f(z, nodata)) out = TD_FELNODATA;
  else if (mask != nullptr && mask[ci] == 1) out = z;
  else if (s.global_edge(r, c)) out = z;
  else {
    bool con = false;
    for (int k = 1; k <= 8; k += step) con = con || nd_f(dem[ci + (long long)drow(k) * s.pitch + dcol(k)], nodata);
    out = con ? z : FLT_MAX;
  }
  W[ci] = out;
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
}  // namespace

int fill_init(const float* dem, const short* mask, float* W, const Strip& s, float nodata, int four, cudaStream_t st) {
  dim3 grid(s.ny, (s.nx + 255) / 256);
  k_fill_init<<<grid, 256, 0, st>>>(dem, mask, W, s, nodata, four ? 2 : 1);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

// Relaxes until no tile of the strip changes.  *changed = whether any cell moved.
int fill_relax(td_ctx* ctx, const float* dem, float* W, const Strip& s, int four, int* changed, cudaStream_t st) {
  const int ntx = (s.nx + FW - 1) / FW, nty = (s.ny + FH - 1) / FH;
  const long long nt = (long long)ntx * nty;
  TD_CUDA(ctx->tileflags.ensure((size_t)nt * 4 * 3));
  int* flag = ctx->tileflags.as<int>();
  int* la = flag + nt;
  int* lb = la + nt;
  k_fill_all_tiles<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(la, flag, (int)nt);
  TD_LAUNCHED();
  unsigned long long n = (unsigned long long)nt;
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
