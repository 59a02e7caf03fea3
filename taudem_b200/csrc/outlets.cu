// aread8 / areadinf -o: restrict the evaluation to the cells upstream of the outlets.
//
// reference: the outlet branches of initNeighborD8up / initNeighborDinfup (src/commonLib.cpp:285-385, 137-237):
// starting from the outlet cells, every cell reached by following "neighbour k drains into me" backwards gets a
// dependency count; every other cell keeps nodata, is never evaluated and keeps area nodata (-1).
// Here the dependency stencil has already built the node words (which neighbours drain into each cell) and the
// counts for the whole strip; k_upstream floods the contributor links from the outlets (one cell per lane, the
// warp's stack holds the discovered contributors, overflow spills to a list the host drains with another launch —
// the structure of k_walk) and k_restrict turns everything that was not reached into "not a cell of the flow
// field" (node 0, count 0xFF), after which the ordinary sweep runs.  An outlet on a cell WITHOUT a valid flow direction
// (nodata direction: every D8 edge cell) is handled like the reference does (src/commonLib.cpp:319-343 never looks at
// the outlet's own direction; src/aread8.cpp:274-276 evaluates it with a warning): the cell becomes a node of the flow
// field with the contributors its neighbours' directions give it and no receiver, its contributors are flooded, and the
// sweep evaluates it (contaminated if a neighbour is off the grid or has no direction).
#include <algorithm>
#include <vector>

#include "ctx.h"
#include "dinf_common.cuh"
#include "kernels.h"

namespace td {
namespace {
constexpr unsigned UP_VALID = 0x8000u, IN_SET = 0x4000u;
#ifndef TD_UP_UQ
#define TD_UP_UQ 128
#endif
constexpr int UQ = TD_UP_UQ;   // stack entries per warp

struct UpArgs {
  unsigned short* node;
  Strip s;
  const long long* list;
  unsigned long long nlist;
  unsigned long long* ctr;    // [0] list ticket, [1] spill length, [2] spill list exhausted
  long long* spill;
  unsigned long long spill_cap;
  int* req;                   // row strips: req[c] = 1 asks the strip above to continue at (its last row, c), req[pitch + c] the strip below
  unsigned char* cnt;         // dependency counts (written for outlets on cells without a flow direction)
  int adopt;                  // this launch's list holds outlet cells: cells without a valid direction are adopted as nodes
};

__global__ void __launch_bounds__(256) k_upstream(const UpArgs a) {
  __shared__ long long wq[8][UQ];
  __shared__ int wqn[8];
  const Strip& s = a.s;
  const unsigned lane = threadIdx.x & 31u;
  const int wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  if (lane == 0) wqn[wid] = 0;
  __syncwarp();
  bool list_done = a.nlist == 0;
  for (;;) {
    // ---- one cell per lane: the warp's stack first, then a batch of the list
    long long cur = -1;
    const int nl = wqn[wid];
    if ((int)lane < nl) cur = wq[wid][nl - 1 - (int)lane];
    __syncwarp();
    if (lane == 0) wqn[wid] = max(0, nl - 32);
    __syncwarp();
    const unsigned idle = __ballot_sync(0xffffffffu, cur < 0);
    if (idle && !list_done) {
      const int need = __popc(idle);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.ctr, (unsigned long long)need);
      base = __shfl_sync(0xffffffffu, base, 0);
      const unsigned long long mine = base + (unsigned long long)__popc(idle & lt);
      if (cur < 0 && mine < a.nlist) cur = a.list[mine];
      if (base + (unsigned long long)need >= a.nlist) list_done = true;
    }
    if (__ballot_sync(0xffffffffu, cur >= 0) == 0u) break;
    // ---- mark it; if it is a cell of the flow field seen for the first time, its contributors are next
    if (cur >= 0) {
      unsigned* word = reinterpret_cast<unsigned*>(a.node) + (cur >> 1);
      const unsigned sh = (unsigned)(cur & 1) * 16u;
      unsigned nd = (atomicOr(word, IN_SET << sh) >> sh) & 0xffffu;
      if (a.adopt && !(nd & UP_VALID) && !(nd & IN_SET)) {
        // an outlet on a cell without a flow direction: which neighbours drain into it (their first / second receiver
        // points here), is any of them missing (off the grid / no direction -> the cell is contaminated)
        const int r = (int)(cur / s.pitch), c = (int)(cur - (long long)r * s.pitch);
        unsigned mask = 0, con = 0;
#pragma unroll
        for (int k = 1; k <= 8; ++k) {
          const int rn = r + drow(k), cn = c + dcol(k);
          unsigned nn = 0;
          if (s.on_grid(rn, cn) && rn >= 1 && rn <= s.ny) nn = a.node[s.idx(rn, cn)];
          if (!(nn & UP_VALID)) { con = 1; continue; }
          const int back = k > 4 ? k - 4 : k + 4;
          const int k1 = dinf_node_k1(nn), k2 = dinf_node_k2(nn);   // (D8 words: the direction 0..8, never a second receiver)
          if (k1 == back || k2 == back) mask |= 1u << (k - 1);
        }
        nd = UP_VALID | (con ? 0x1000u : 0u) | mask;            // receiver field 0: nothing downstream of it is decremented
        atomicOr(word, nd << sh);
        a.cnt[cur] = (unsigned char)__popc(mask);
      }
      if ((nd & UP_VALID) && !(nd & IN_SET)) {
#pragma unroll
        for (int k = 1; k <= 8; ++k)
          if ((nd >> (k - 1)) & 1u) {
            const long long ni = cur + (long long)drow(k) * s.pitch + dcol(k);
            if (ni < s.pitch) { a.req[ni] = 1; continue; }                                              // a cell of the strip above
            if (ni >= (long long)(s.ny + 1) * s.pitch) { a.req[s.pitch + (int)(ni - (long long)(s.ny + 1) * s.pitch)] = 1; continue; }
            const int slot = atomicAdd(&wqn[wid], 1);
            if (slot < UQ) wq[wid][slot] = ni;
            else {
              const unsigned long long g = atomicAdd(a.ctr + 1, 1ull);
              if (g < a.spill_cap) a.spill[g] = ni; else a.ctr[2] = 1ull;
            }
          }
      }
    }
    __syncwarp();
    if (lane == 0) wqn[wid] = min(wqn[wid], UQ);
    __syncwarp();
  }
}

// one thread per four cells of a row
__global__ void __launch_bounds__(256) k_restrict(unsigned short* __restrict__ node, unsigned char* __restrict__ cnt, Strip s) {
  const int r = 1 + (int)blockIdx.x, c = ((int)blockIdx.y * 256 + (int)threadIdx.x) * 4;   // rows on grid.x
  if (c >= s.pitch) return;
  const long long o = s.idx(r, c);
  ushort4 nd = *reinterpret_cast<ushort4*>(node + o);
  uchar4 cn = *reinterpret_cast<uchar4*>(cnt + o);
  unsigned short* n4 = reinterpret_cast<unsigned short*>(&nd);
  unsigned char* c4 = reinterpret_cast<unsigned char*>(&cn);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if ((n4[i] & UP_VALID) && (n4[i] & IN_SET)) n4[i] = (unsigned short)(n4[i] & ~IN_SET);
    else { n4[i] = 0; c4[i] = 0xff; }
  }
  *reinterpret_cast<ushort4*>(node + o) = nd;
  *reinterpret_cast<uchar4*>(cnt + o) = cn;
}
}  // namespace

namespace {
// requests received from the neighbour strips -> cells of my first / last row appended to the seed list
__global__ void k_requests(const int* __restrict__ in_top, const int* __restrict__ in_bot, Strip s, long long* __restrict__ list,
                           unsigned long long* __restrict__ ctr) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= s.nx) return;
  if (in_top && in_top[c] > 0) list[atomicAdd(ctr, 1ull)] = s.idx(1, c);
  if (in_bot && in_bot[c] > 0) list[atomicAdd(ctr, 1ull)] = s.idx(s.ny, c);
}
}  // namespace

// One round of the outlet restriction on a strip.  Seeds: the outlets (cols / rows: grid coordinates, row 0 = first owned
// row of the strip, host memory; points outside the strip are ignored like the reference ignores points outside the
// partition; nout < 0 = none this round) and the requests of the neighbour strips (in_top / in_bot: device arrays of
// pitch ints, what the strip above / below wrote into its req_out for my first / last row; NULL = none).  req_out
// (device, 2 x pitch ints, zeroed here) receives this strip's requests to its neighbours.  finish != 0: nobody has
// requests left — everything that was not reached is removed from the flow field (k_restrict).
int sweep_restrict_round(td_ctx* ctx, const Strip& s, const int* cols, const int* rows, int nout, const int* in_top, const int* in_bot,
                         int* req_out, int finish, cudaStream_t st) {
  std::vector<long long> cells;
  for (int i = 0; i < nout; ++i)
    if (cols[i] >= 0 && cols[i] < s.nx && rows[i] >= 0 && rows[i] < s.ny) cells.push_back(s.idx(rows[i] + 1, cols[i]));
  unsigned long long n = cells.size();
  const unsigned long long cap = (unsigned long long)s.nx * s.ny / 16 + 65536;
  TD_CUDA(ctx->listA.ensure(sizeof(long long) * (n + 2 * (unsigned long long)s.nx + 1)));
  TD_CUDA(ctx->listB.ensure(sizeof(long long) * cap));
  TD_CUDA(ctx->listC.ensure(sizeof(long long) * cap));
  if (n) TD_CUDA(cudaMemcpyAsync(ctx->listA.p, cells.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
  UpArgs a;
  a.node = ctx->node.as<unsigned short>(); a.s = s; a.ctr = ctx->d_ctr + 16; a.spill_cap = cap; a.req = req_out;
  a.cnt = ctx->cnt.as<unsigned char>(); a.adopt = 1;
  unsigned long long* hc = ctx->h_ctr + 16;
  if (req_out) TD_CUDA(cudaMemsetAsync(req_out, 0, sizeof(int) * 2 * (size_t)s.pitch, st));
  if (in_top || in_bot) {
    TD_CUDA(cudaMemcpyAsync(a.ctr + 3, &n, sizeof n, cudaMemcpyHostToDevice, st));
    k_requests<<<(s.nx + 255) / 256, 256, 0, st>>>(in_top, in_bot, s, ctx->listA.as<long long>(), a.ctr + 3);
    TD_LAUNCHED();
    TD_CUDA(cudaMemcpyAsync(hc + 3, a.ctr + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    TD_CUDA(cudaStreamSynchronize(st));
    n = hc[3];
  }
  const long long* cur = ctx->listA.as<long long>();
  long long* spill = ctx->listB.as<long long>();
  long long* other = ctx->listC.as<long long>();
  int dev = 0, sms = 1;
  cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  while (n > 0) {
    TD_CUDA(cudaMemsetAsync(a.ctr, 0, 3 * sizeof(unsigned long long), st));
    a.list = cur; a.nlist = n; a.spill = spill;
    const int grid = std::max(1, (int)std::min<unsigned long long>((unsigned long long)sms * 4, (n + 255) / 256));
    k_upstream<<<grid, 256, 0, st>>>(a);
    TD_LAUNCHED();
    TD_CUDA(cudaGetLastError());
    TD_CUDA(cudaMemcpyAsync(hc, a.ctr, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    TD_CUDA(cudaStreamSynchronize(st));
    if (hc[2]) { set_error("outlets: spill list exhausted"); return TD_ERR_ALLOC; }
    n = hc[1];
    cur = spill; std::swap(spill, other);
    a.adopt = 0;
  }
  if (finish) {
    const dim3 rgrid((unsigned)s.ny, (unsigned)(((s.pitch >> 2) + 255) / 256));
    k_restrict<<<rgrid, 256, 0, st>>>(a.node, ctx->cnt.as<unsigned char>(), s);
    TD_LAUNCHED();
  }
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

// the whole restriction on a single strip
int sweep_restrict_upstream(td_ctx* ctx, const Strip& s, const int* cols, const int* rows, int nout, cudaStream_t st) {
  if (s.has_top || s.has_bot) { set_error("sweep_restrict_upstream: row strips go through sweep_restrict_round"); return TD_ERR_ARG; }
  TD_CUDA(ctx->halo.ensure(sizeof(int) * 2 * (size_t)s.pitch));
  return sweep_restrict_round(ctx, s, cols, rows, nout, nullptr, nullptr, ctx->halo.as<int>(), 1, st);
}

}  // namespace td
