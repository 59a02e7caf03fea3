// Contributing-area evaluation without tiles: level passes + chain walking in global memory.
//
// reference: aread8 main loop src/aread8.cpp:216-304, area() main loop src/areadinf.cpp:173-265
// (queue of cells whose dependency count is zero; evaluate, decrement the receivers, push those that
// reach zero).  The gather is k-ordered float32, evaluated once per cell when all contributors are final,
// so the result does not depend on who evaluates a cell or when (same argument as sweep_tiles.cu).
//
// Building blocks (combined by TAUDEM_B200_SWEEP, see capi.cu):
//   k_level : one streaming pass over the count array.  The thread that OWNS a cell evaluates it when its
//             count is zero, decrements the receivers and marks the cell done (0xFE); receivers that reach
//             zero are picked up by their owners in this or the next pass.  A pass costs one byte per cell
//             plus the work of the cells it evaluates, fully coalesced, no queue, no waiting.  Topological
//             levels thin out fast (8192^2 fractal DEM: 91 % of the D8 cells and 76 % of the D-infinity
//             cells lie in the first 16 levels), so a few dozen passes do the bulk of the grid.
//   k_ready : collects the cells whose count is zero but which are not evaluated yet into a list.
//   k_walk  : finishes from such a list by chain walking.  Every lane of a warp owns one chain: it evaluates
//             its cell, releases the area, decrements the receiver(s) with an acq_rel atomic and moves on
//             when it was the last arrival.  Idle lanes refill from the warp's own fork stack (D-infinity:
//             a cell can make two receivers ready) and then from the global ready list in batches (one
//             fetch-add per warp).  No CTA-wide barrier and no inter-warp waiting: a warp that holds one
//             long river does not keep other threads from their work.  A fork stack that overflows spills to
//             a global list that the host drains with another launch.
// Modes: "levels" = K passes of k_level, then k_ready + k_walk;  "hybrid" = one visit of every tile by the
// tile kernel (sweep_tiles.cu, `once`), then k_ready + k_walk;  "walk" = k_ready + k_walk from the sources.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "ctx.h"
#include "dinf_common.cuh"
#include "kernels.h"

namespace td {
namespace {
constexpr unsigned NODE_VALID = 0x8000u, NODE_CON = 0x1000u;
#ifndef TD_WALK_WQ
#define TD_WALK_WQ 96
#endif
constexpr int WQ = TD_WALK_WQ;   // fork stack entries per warp

struct WalkArgs {
  const unsigned short* node;
  unsigned* cntw;
  float* area;
  const float* w;
  const float* ang;
  Strip s;
  int usew, contcheck;
  float w_nodata;
  const double* theta;
  const double* dxc;
  int* halo;
  const long long* list;
  unsigned long long nlist;
  unsigned long long* ctr;    // [0] list ticket, [1] spill length, [2] spill list exhausted, [3] ready cells (collect)
  long long* spill;            // D-infinity: forks that did not fit the warp's stack; D8: parked river heads (k_river)
  unsigned long long spill_cap;
  int river_hops;              // a lane that has followed one chain for this many cells parks it for k_river (0 = never)
  unsigned long long* pass_cells;   // k_level diagnostics (TAUDEM_B200_TIMING=2): cells evaluated by this pass, or NULL
  int diag;                         // k_river diagnostics: ctr[4] += batches, ctr[5] += cells resolved in batches
};

__device__ __forceinline__ unsigned dec_count(unsigned* words, long long cell) {
  unsigned* a = words + (cell >> 2);
  const unsigned sh = (unsigned)(cell & 3) * 8u;
  unsigned old;
#ifdef TD_EMU
  old = atomicAdd(a, 0u - (1u << sh));
#else
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(a), "r"(0u - (1u << sh)) : "memory");
#endif
  return (old >> sh) & 0xffu;
}
// the same decrement without ordering, and the fence that supplies it: "fence; relaxed atomics; fence" releases the area
// once for both receivers of a D-infinity cell and lets the two atomics travel together
__device__ __forceinline__ unsigned dec_count_relaxed(unsigned* words, long long cell) {
  unsigned* a = words + (cell >> 2);
  const unsigned sh = (unsigned)(cell & 3) * 8u;
  unsigned old;
#ifdef TD_EMU
  old = atomicAdd(a, 0u - (1u << sh));
#else
  asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(a), "r"(0u - (1u << sh)) : "memory");
#endif
  return (old >> sh) & 0xffu;
}
__device__ __forceinline__ void fence_acq_rel() {
#ifdef TD_EMU
  __threadfence();
#else
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
#ifdef TD_EMU
  v = __ldcg(p);
#else
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
  return v;
}

// prop(angle, kk) through the full interval search (one copy, out of line: it is the rare path of eval_cell)
__device__ __noinline__ double share_full(float ang, double t, int kk) {
  const Outflow o = dinf_outflow(ang, t);
  return o.k1 == kk ? o.p1 : o.p2;
}

// Evaluates cell (r, c) = ci whose contributors are all final, stores its area and decrements the
// receivers.  Receivers whose count reached zero through this decrement are returned in ready[0..1]
// (with their node words) — the caller decides whether it follows them.
template <bool DINF>
__device__ __forceinline__ int eval_cell(const WalkArgs& a, long long ci, int r, int c, unsigned nd, long long* ready, unsigned* ready_nd,
                                         int* ready_r, int* ready_c) {
  const Strip& s = a.s;
  const unsigned m = nd & 0xffu;
  bool con = (nd & NODE_CON) != 0;
  int nready = 0;
  // contributors' areas (and angles): all loads are issued before any is used
  float an[8], aa[8];
#pragma unroll
  for (int k = 1; k <= 8; ++k) {
    const long long ni = ci + (long long)drow(k) * s.pitch + dcol(k);
    const bool in = (m >> (k - 1)) & 1u;
    an[k - 1] = in ? __ldcg(a.area + ni) : 0.f;
    if (DINF) aa[k - 1] = in ? a.ang[ni] : 0.f;
  }
  float val;
  if (!DINF) {
    // src/aread8.cpp:228-257
    const int d = (int)((nd >> 8) & 0xfu);
    long long cin = -1; unsigned ndn = 0; int rn = 0, cn = 0;
    if (d >= 1 && d <= 8) {
      rn = r + drow(d); cn = c + dcol(d);
      if (s.on_grid(rn, cn)) { cin = s.idx(rn, cn); if (rn != 0 && rn != s.ny + 1) ndn = a.node[cin]; }
    }
    if (a.usew) { const float wv = a.w[ci]; val = nd_f(wv, a.w_nodata) ? -1.0f : wv; }
    else val = 1.0f;
#pragma unroll
    for (int k = 1; k <= 8; ++k)
      if ((m >> (k - 1)) & 1u) { if (nd_f(an[k - 1], -1.0f)) con = true; else val = val + an[k - 1]; }
    if (con && a.contcheck) val = -1.0f;
    a.area[ci] = val;
    // src/aread8.cpp:261-272
    if (cin >= 0) {
      if (rn == 0 || rn == s.ny + 1) { __threadfence(); atomicAdd(a.halo + (rn == 0 ? 0 : s.pitch) + cn, 1); }
      else if ((ndn & NODE_VALID) && dec_count(a.cntw, cin) == 1u) { ready[0] = cin; ready_nd[0] = ndn; ready_r[0] = rn; ready_c[0] = cn; nready = 1; }
    }
  } else {
    // src/areadinf.cpp:187-218.  The share a contributor sends here is prop(angle, direction): for a contributor with
    // two receivers in one of the sectors 1..7 the node word already says which sector (k1 = j, k1 + 1), so the share
    // is one division, (hi - a)/(hi - mid) for its first receiver and (a - mid)/(hi - mid) for the second — exactly
    // the expressions dinf_outflow evaluates; everything else (single receiver, the wrap sector, contributors in a
    // halo row, whose node words belong to the neighbour strip) takes the full interval search.
    unsigned short nn[8];
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      const long long ni = ci + (long long)drow(k) * s.pitch + dcol(k);
      nn[k - 1] = ((m >> (k - 1)) & 1u) ? a.node[ni] : (unsigned short)0;
    }
    // the receivers (from the node word) and their node words, requested together with the contributors' data
    const int k1 = (int)((nd >> 8) & 0xfu);
    int rk[2]; long long rci[2]; unsigned rnd2[2]; int rrn[2], rcn[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = j == 0 ? k1 : ((nd & 0x2000u) ? k1 % 8 + 1 : 0);
      rk[j] = 0; rci[j] = -1; rnd2[j] = 0; rrn[j] = rcn[j] = 0;
      if (k == 0) continue;
      const int rn = r + drow(k), cn = c + dcol(k);
      if (!s.on_grid(rn, cn)) continue;
      rk[j] = k; rrn[j] = rn; rcn[j] = cn; rci[j] = s.idx(rn, cn);
      if (rn != 0 && rn != s.ny + 1) rnd2[j] = a.node[rci[j]];
    }
    val = 0.f;
#pragma unroll
    for (int k = 1; k <= 8; ++k)
      if ((m >> (k - 1)) & 1u) {
        const int kk = k > 4 ? k - 4 : k + 4;              // the direction from that neighbour to this cell
        const int rn = r + drow(k);
        const double t = a.theta[min(max(rn - 1, 0), s.ny - 1)];
        const int k1n = (nn[k - 1] >> 8) & 0xf;
        double p;
        if ((nn[k - 1] & 0x2000u) && k1n <= 7 && rn >= 1 && rn <= s.ny) {
          const double mid = aref(k1n, t), hi = aref(k1n + 1, t);
          const float av = aa[k - 1];
          p = (kk == k1n) ? (hi - av) / (hi - mid) : (av - mid) / (hi - mid);
        } else p = share_full(aa[k - 1], t, kk);
        if (nd_f(an[k - 1], -1.0f)) con = true; else val = (float)((double)val + p * (double)an[k - 1]);
      }
    if (a.usew) val = val + a.w[ci];
    else val = (float)((double)val + a.dxc[r - 1]);
    if (con && a.contcheck) val = -1.0f;
    a.area[ci] = val;
    // src/areadinf.cpp:221-239: every neighbour that receives a share.  One release fence, then both decrements (issued
    // back to back, their round trips overlap), one acquire fence if a receiver became ready.
    unsigned oldc[2] = {0u, 0u};
    const bool any = rk[0] != 0 || rk[1] != 0;
    if (any) fence_acq_rel();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (rk[j] == 0) continue;
      const int rn = rrn[j], cn = rcn[j];
      if (rn == 0 || rn == s.ny + 1) { atomicAdd(a.halo + (rn == 0 ? 0 : s.pitch) + cn, 1); continue; }
      if (!(rnd2[j] & NODE_VALID)) continue;
      oldc[j] = dec_count_relaxed(a.cntw, rci[j]);
    }
    if (oldc[0] == 1u || oldc[1] == 1u) fence_acq_rel();
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (oldc[j] == 1u) {
        if (nready == 0) { ready[0] = rci[j]; ready_nd[0] = rnd2[j]; ready_r[0] = rrn[j]; ready_c[0] = rcn[j]; }
        else { ready[1] = rci[j]; ready_nd[1] = rnd2[j]; ready_r[1] = rrn[j]; ready_c[1] = rcn[j]; }
        ++nready;
      }
  }
  return nready;
}

// ---- level pass: one thread per 16 cells of a row (four count words)
template <bool DINF>
__global__ void __launch_bounds__(256) k_level(const WalkArgs a) {
  const Strip& s = a.s;
  // rows on grid.x (no 65535 limit), 16-cell groups of the row on grid.y x threads (pitch % 32 == 0): no division
  const int r = 1 + (int)blockIdx.x, c0 = ((int)blockIdx.y * 256 + (int)threadIdx.x) * 16;
  const bool inrow = c0 < s.pitch;
  const long long base = s.idx(r, inrow ? c0 : 0);
  uint4 q = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  if (inrow) q = __ldcg(reinterpret_cast<const uint4*>(a.cntw + (base >> 2)));
  const unsigned w4[4] = {q.x, q.y, q.z, q.w};
  unsigned todo = 0;                                                     // bit b: cell base + b is ready
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (((w4[j] - 0x01010101u) & ~w4[j] & 0x80808080u) == 0u) continue;  // no zero byte in this word
    const unsigned word = ld_acquire(a.cntw + ((base + 4 * j) >> 2));    // the contributors' areas are visible after this
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (((word >> (8 * i)) & 0xffu) == 0u && c0 + 4 * j + i < s.nx) todo |= 1u << (4 * j + i);
  }
  if (a.pass_cells && (blockIdx.x & 15u) == 0u) {
    // every 16th row reports how many cells it evaluates (one atomic per CTA): enough to steer the number of passes
    __shared__ unsigned bsum;
    if (threadIdx.x == 0) bsum = 0;
    __syncthreads();
    if (todo) atomicAdd(&bsum, (unsigned)__popc(todo));
    __syncthreads();
    if (threadIdx.x == 0 && bsum) atomicAdd(a.pass_cells, (unsigned long long)bsum);
  }
  unsigned d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma unroll 1
  while (todo) {
    const int b = __ffs(todo) - 1;
    todo &= todo - 1;
    const unsigned nd = a.node[base + b];
    if (!(nd & NODE_VALID)) continue;
    long long ready[2]; unsigned rnd[2]; int rr[2], rc[2];
    eval_cell<DINF>(a, base + b, r, c0 + b, nd, ready, rnd, rr, rc);    // receivers wait for their owners
    const unsigned bit = 0xfeu << (8 * (b & 3));
    if ((b >> 2) == 0) d0 |= bit; else if ((b >> 2) == 1) d1 |= bit; else if ((b >> 2) == 2) d2 |= bit; else d3 |= bit;
  }
  unsigned* cw = a.cntw + (base >> 2);                                   // 0 -> 0xFE (evaluated)
  if (d0) atomicAdd(cw, d0);
  if (d1) atomicAdd(cw + 1, d1);
  if (d2) atomicAdd(cw + 2, d2);
  if (d3) atomicAdd(cw + 3, d3);
}

// ---- ready cells: count byte == 0 on an owned, valid cell.  One thread per count word (4 cells).
template <bool FILL>
__global__ void __launch_bounds__(256) k_ready(const unsigned* __restrict__ cntw, const unsigned short* __restrict__ node, Strip s,
                                               unsigned long long* __restrict__ ctr, long long* __restrict__ list) {
  const int r = 1 + (int)blockIdx.x, c = ((int)blockIdx.y * 256 + (int)threadIdx.x) * 4;   // rows on grid.x, words of the row on grid.y x threads
  int n = 0;
  long long cells[4];
  if (c < s.pitch) {
    const long long ci = s.idx(r, c);
    const unsigned word = cntw[ci >> 2];
    if (((word - 0x01010101u) & ~word & 0x80808080u) != 0u) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (((word >> (8 * i)) & 0xffu) == 0u && c + i < s.nx && (node[ci + i] & NODE_VALID)) cells[n++] = ci + i;
    }
  }
  // block-aggregated reservation: one atomic per CTA (a single address takes every reservation of the grid)
  __shared__ int wtot[8];
  __shared__ unsigned long long bbase;
  const unsigned lane = threadIdx.x & 31u;
  const int wid = threadIdx.x >> 5;
  int incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { if (i < wid) before += wtot[i]; total += wtot[i]; }
  if (total == 0) return;
  if (threadIdx.x == 0) bbase = atomicAdd(ctr + (FILL ? 0 : 3), (unsigned long long)total);
  if (!FILL) return;
  __syncthreads();
  const unsigned long long base = bbase + (unsigned long long)before;
  for (int i = 0; i < n; ++i) list[base + (unsigned long long)(incl - n + i)] = cells[i];
}

template <bool DINF>
__global__ void __launch_bounds__(256) k_walk(const WalkArgs a) {
  __shared__ long long wq[8][WQ];
  __shared__ int wqn[8];
  const Strip& s = a.s;
  const unsigned lane = threadIdx.x & 31u;
  const int wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  if (lane == 0) wqn[wid] = 0;
  __syncwarp();
  long long cur = -1;
  unsigned curnd = 0;
  bool have_nd = false;
  int hops = 0;                         // cells of the current chain (river parking)
  int curr = 0, curc = 0;               // row / column of `cur` when have_nd (a 64-bit division per hop otherwise)
  bool list_done = a.nlist == 0;        // warp-uniform
  for (;;) {
    // ---- refill idle lanes: the warp's fork stack first, then a batch of the global ready list
    unsigned idle = __ballot_sync(0xffffffffu, cur < 0);
    if (idle) {
      const int nl = wqn[wid];
      const int rank = __popc(idle & lt);
      if (cur < 0 && rank < nl) { cur = wq[wid][nl - 1 - rank]; have_nd = false; hops = 0; }
      __syncwarp();
      if (lane == 0) wqn[wid] = max(0, nl - __popc(idle));
      __syncwarp();
      idle = __ballot_sync(0xffffffffu, cur < 0);
      if (idle && !list_done) {
        const int need = __popc(idle);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.ctr, (unsigned long long)need);
        base = __shfl_sync(0xffffffffu, base, 0);
        const unsigned long long mine = base + (unsigned long long)__popc(idle & lt);
        if (cur < 0 && mine < a.nlist) { cur = a.list[mine]; have_nd = false; hops = 0; }
        if (base + (unsigned long long)need >= a.nlist) list_done = true;
      }
    }
    if (__ballot_sync(0xffffffffu, cur >= 0) == 0u) break;   // stack empty and list exhausted

    // ---- one hop per active lane
    long long fork = -1;
    if (cur >= 0) {
      const long long ci = cur;
      int r, c;
      if (have_nd) { r = curr; c = curc; }
      else { r = (int)(ci / s.pitch); c = (int)(ci - (long long)r * s.pitch); }
      const unsigned nd = have_nd ? curnd : (unsigned)a.node[ci];
      long long ready[2]; unsigned rnd[2]; int rr[2], rc[2];
      const int nr = eval_cell<DINF>(a, ci, r, c, nd, ready, rnd, rr, rc);
      atomicAdd(a.cntw + (ci >> 2), 0xfeu << ((unsigned)(ci & 3) * 8u));     // 0 -> 0xFE (evaluated), like k_level and the tile kernel
      if (nr >= 1) { cur = ready[0]; curnd = rnd[0]; curr = rr[0]; curc = rc[0]; have_nd = true; } else cur = -1;
      if (a.river_hops > 0 && cur >= 0 && ++hops >= a.river_hops) {
        // a long chain: most likely a river.  Its next cell (ready, not evaluated) goes to the river list, where a
        // whole warp follows it with look-ahead (k_river) once the short chains are done.
        const unsigned long long g = atomicAdd(a.ctr + 1, 1ull);
        if (g < a.spill_cap) { a.spill[g] = cur; cur = -1; hops = 0; } else atomicAdd(a.ctr + 1, ~0ull);   // list full: keep walking
      }
      if (DINF && nr == 2) fork = ready[1];
    }
    if (DINF) {
      // ---- second ready receivers go to the warp's fork stack (idle lanes take them in the next iteration)
      const unsigned fm = __ballot_sync(0xffffffffu, fork >= 0);
      if (fm) {
        const int nl = wqn[wid];
        const int slot = nl + __popc(fm & lt);
        if (fork >= 0) {
          if (slot < WQ) wq[wid][slot] = fork;
          else {
            const unsigned long long g = atomicAdd(a.ctr + 1, 1ull);
            if (g < a.spill_cap) a.spill[g] = fork; else a.ctr[2] = 1ull;
          }
        }
        __syncwarp();
        if (lane == 0) wqn[wid] = min(WQ, nl + __popc(fm));
        __syncwarp();
      }
    }
  }
}

// Rivers.  A D8 chain never forks and its path is static (the direction codes); a D-infinity chain is the same
// for as long as its cells have a single receiver — 87 % of the cells of the longest dependency chain of the
// 8192^2 test field, three quarters of them in runs of eight or more (flat-resolved valley floors).  So a warp can
// look ahead: it follows the path from a ready head for up to 32 cells, as long as the next cell's only missing
// arrival is the cell before it on the path (count byte == 1: every other contributor has released its area).
// Those cells are resolved together — contributors (and their shares) gathered by all lanes at once, then one
// k-ordered fold per cell handed from lane to lane by shuffle (the exact sequence of operations of
// src/aread8.cpp:228-257 / src/areadinf.cpp:187-218) — with ONE memory round trip per cell (node word + count
// word of the next cell) instead of the four dependent ones of a chain walker.  The warp then arrives at the
// receiver(s) of the last cell with the usual release + decrement and goes on from one that became ready; a
// second one (D-infinity) goes to the spill list for the next launch.
template <bool DINF>
__global__ void __launch_bounds__(256) k_river(const WalkArgs a) {
  const Strip& s = a.s;
  const unsigned FULL = 0xffffffffu;
  const int lane = (int)(threadIdx.x & 31u);
  for (;;) {
    long long head = -1;
    if (lane == 0) { const unsigned long long t = atomicAdd(a.ctr, 1ull); if (t < a.nlist) head = a.list[t]; }
    head = __shfl_sync(FULL, head, 0);
    if (head < 0) return;
    for (;;) {
      const int hr = (int)(head / s.pitch), hc = (int)(head - (long long)hr * s.pitch);
      const unsigned hnd = a.node[head];
      if (lane == 0) (void)ld_acquire(a.cntw + (head >> 2));   // what the head's contributors released ...
      __syncwarp();                                            // ... is visible to every lane of the warp
      // ---- 1. the path: the same loop in every lane (uniform loads), lane i keeps cell i
      long long ci = head; int r = hr, c = hc; unsigned nd = hnd; int kp = 0;
      long long my = -1; int mykp = 0, myr = 0, myc = 0; unsigned mynd = 0;
      long long nxt = -1; int nxr = 0, nxc = 0; bool halo_exit = false;
      int len = 0;
      for (int i = 0; i < 32; ++i) {
        if (lane == i) { my = ci; mynd = nd; mykp = kp; myr = r; myc = c; }
        len = i + 1;
        nxt = -1;
        const int d = (int)((nd >> 8) & 0xfu);               // D8: the direction; D-infinity: the first receiver
        if (d < 1 || d > 8) break;
        if (DINF && (nd & 0x2000u)) break;                   // two receivers: the batch ends with this cell
        const int rn = r + drow(d), cn = c + dcol(d);
        if (!s.on_grid(rn, cn)) break;
        if (rn == 0 || rn == s.ny + 1) { halo_exit = true; nxr = rn; nxc = cn; break; }
        const long long cin = s.idx(rn, cn);
        const unsigned ndn = a.node[cin];
        unsigned cw = 0;                                  // the count word changes under us: one lane reads it for the warp
        if (lane == 0) cw = ld_acquire(a.cntw + (cin >> 2));
        cw = __shfl_sync(FULL, cw, 0);                    // (the shuffle also orders the other lanes' later loads behind the acquire)
        if (!(ndn & NODE_VALID)) break;
        nxt = cin; nxr = rn; nxc = cn;
        if (((cw >> ((unsigned)(cin & 3) * 8u)) & 0xffu) != 1u || i == 31) break;   // not resolvable ahead of time: arrive there
        kp = d > 4 ? d - 4 : d + 4;                      // the direction from the next cell back to this one
        ci = cin; r = rn; c = cn; nd = ndn;
      }
      if (a.diag && lane == 0) { atomicAdd(a.ctr + 4, 1ull); atomicAdd(a.ctr + 5, (unsigned long long)len); }
      // ---- 2. the contributors that are not on the path (all final): areas, and for D-infinity every contributor's share
      const unsigned m = lane < len ? (mynd & 0xffu) : 0u;
      // every lane acquires on its own cell's count word — the location its side contributors released their areas on
      // (lane 0 read these words for the path search; its acquire does not formally order the other lanes' loads)
      if (lane < len) (void)ld_acquire(a.cntw + (my >> 2));
      float an[8];
      double pk[DINF ? 8 : 1];
#pragma unroll
      for (int k = 1; k <= 8; ++k) {
        const bool in = ((m >> (k - 1)) & 1u) != 0u;
        const long long ni = my + (long long)drow(k) * s.pitch + dcol(k);
        an[k - 1] = (in && k != mykp) ? __ldcg(a.area + ni) : 0.f;
        if (DINF) {
          double p = 0.;
          if (in) {
            const float av = a.ang[ni];
            const unsigned nn = a.node[ni];
            const int kk = k > 4 ? k - 4 : k + 4, rn = myr + drow(k), k1n = (int)((nn >> 8) & 0xfu);
            const double t = a.theta[min(max(rn - 1, 0), s.ny - 1)];
            if ((nn & 0x2000u) && k1n <= 7 && rn >= 1 && rn <= s.ny) {
              const double mid = aref(k1n, t), hi = aref(k1n + 1, t);
              p = (kk == k1n) ? (hi - av) / (hi - mid) : (av - mid) / (hi - mid);
            } else p = share_full(av, t, kk);
          }
          pk[k - 1] = p;
        }
      }
      float wv = 1.0f;
      double dxr = 0.;
      if (lane < len) {
        if (a.usew) { const float x = a.w[my]; wv = (!DINF && nd_f(x, a.w_nodata)) ? -1.0f : x; }
        else if (DINF) dxr = a.dxc[myr - 1];
      }
      // ---- 3. the fold, one cell after the other
      float val = 0.f;
      for (int i = 0; i < len; ++i) {
        const float prev = __shfl_sync(FULL, val, (i + 31) & 31);
        if (lane == i) {
          bool con = (mynd & NODE_CON) != 0;
          float v = DINF ? 0.f : wv;
#pragma unroll
          for (int k = 1; k <= 8; ++k)
            if ((m >> (k - 1)) & 1u) {
              const float x = k == mykp ? prev : an[k - 1];
              if (nd_f(x, -1.0f)) con = true;
              else if (DINF) v = (float)((double)v + pk[DINF ? k - 1 : 0] * (double)x);
              else v = v + x;
            }
          if (DINF) { if (a.usew) v = v + wv; else v = (float)((double)v + dxr); }
          if (con && a.contcheck) v = -1.0f;
          val = v;
        }
      }
      // ---- 4. results; the head had count 0, the others 1 (their missing arrival was the path)
      if (lane < len) {
        a.area[my] = val;
        atomicAdd(a.cntw + (my >> 2), (lane == 0 ? 0xfeu : 0xfdu) << ((unsigned)(my & 3) * 8u));
      }
      // ---- 5. the last cell of the batch arrives at its receiver(s): release of its area, decrement
      long long newhead = -1;
      if (lane == len - 1) {
        if (!DINF || !(mynd & 0x2000u)) {                    // one receiver: the cell the path search stopped at
          if (halo_exit) { __threadfence(); atomicAdd(a.halo + (nxr == 0 ? 0 : s.pitch) + nxc, 1); }
          else if (nxt >= 0 && dec_count(a.cntw, nxt) == 1u) newhead = nxt;
        } else {                                             // src/areadinf.cpp:221-239
          const int k1 = (int)((mynd >> 8) & 0xfu);
          long long cin2[2] = {-1, -1};
          unsigned old2[2] = {0u, 0u};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int k = j == 0 ? k1 : k1 % 8 + 1;
            const int rn = myr + drow(k), cn = myc + dcol(k);
            if (!s.on_grid(rn, cn)) continue;
            if (rn == 0 || rn == s.ny + 1) { __threadfence(); atomicAdd(a.halo + (rn == 0 ? 0 : s.pitch) + cn, 1); continue; }
            const long long cin = s.idx(rn, cn);
            if (a.node[cin] & NODE_VALID) cin2[j] = cin;
          }
          fence_acq_rel();                                   // release the area once, both decrements travel together
#pragma unroll
          for (int j = 0; j < 2; ++j) if (cin2[j] >= 0) old2[j] = dec_count_relaxed(a.cntw, cin2[j]);
          if (old2[0] == 1u || old2[1] == 1u) fence_acq_rel();
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (old2[j] == 1u) {
              if (newhead < 0) newhead = cin2[j];
              else {
                const unsigned long long g = atomicAdd(a.ctr + 1, 1ull);
                if (g < a.spill_cap) a.spill[g] = cin2[j]; else a.ctr[2] = 1ull;
              }
            }
        }
      }
      newhead = __shfl_sync(FULL, newhead, len - 1);
      if (newhead < 0) break;
      head = newhead;
    }
  }
}

// Dependency decrements received from the neighbour strips (addBorders, src/linearpart.h:314-328 and
// src/aread8.cpp:283-297): dec_top[c] arrivals for the cell (row 1, c), dec_bot[c] for (row ny, c).  Cells
// that reach zero are found by the next k_ready scan.
__global__ void k_apply_plain(unsigned* __restrict__ cntw, const unsigned short* __restrict__ node, Strip s,
                              const int* __restrict__ dec_top, const int* __restrict__ dec_bot) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= s.nx) return;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int* dec = side == 0 ? dec_top : dec_bot;
    if (dec == nullptr) continue;
    const int d = dec[c];
    if (d <= 0) continue;
    const long long ci = s.idx(side == 0 ? 1 : s.ny, c);
    if (!(node[ci] & NODE_VALID)) continue;
    atomicAdd(cntw + (ci >> 2), 0u - ((unsigned)d << ((unsigned)(ci & 3) * 8u)));
  }
}

template <bool DINF>
int walk_grid(unsigned long long n) {
  static int per_dev = 0;
  if (!per_dev) {
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_walk<DINF>, 256, 0) != cudaSuccess || occ < 1)
      return 0;
    per_dev = sms * occ;
  }
  return (int)std::min<unsigned long long>((unsigned long long)per_dev, (n + 255) / 256);
}

// TAUDEM_B200_TIMING=1: wall-clock milliseconds per phase (with a stream synchronisation at every phase boundary)
struct PhaseTimer {
  bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0;
  PhaseTimer(cudaStream_t s) : st(s) { const char* e = getenv("TAUDEM_B200_TIMING"); on = e && atoi(e) > 0; if (on) { cudaStreamSynchronize(st); t0 = std::chrono::steady_clock::now(); } }
  void lap(double* acc) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const auto t1 = std::chrono::steady_clock::now();
    *acc += std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
  }
};

void walk_args(td_ctx* ctx, WalkArgs& a, float* area, const float* w, const float* ang, const Strip& s, float w_nodata, int usew,
               int contcheck, const double* theta, const double* dxc, int* halo) {
  a.node = ctx->node.as<unsigned short>(); a.cntw = ctx->cnt.as<unsigned>();
  a.area = area; a.w = w; a.ang = ang; a.s = s; a.usew = usew; a.contcheck = contcheck; a.w_nodata = w_nodata;
  a.theta = theta; a.dxc = dxc; a.halo = halo;
  a.list = nullptr; a.nlist = 0; a.spill = nullptr; a.spill_cap = 0; a.river_hops = 0; a.pass_cells = nullptr; a.diag = 0;
  a.ctr = ctx->d_ctr + 16;
}
}  // namespace

int sweep_apply_plain(td_ctx* ctx, const Strip& s, const int* dec_top, const int* dec_bot, cudaStream_t st) {
  k_apply_plain<<<(s.nx + 255) / 256, 256, 0, st>>>(ctx->cnt.as<unsigned>(), ctx->node.as<unsigned short>(), s, dec_top, dec_bot);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

// `passes` streaming level passes over the strip (see k_level).
int sweep_levels(td_ctx* ctx, bool dinf, int passes, float* area, const float* w, const float* ang, const Strip& s, float w_nodata,
                 int usew, int contcheck, const double* theta, const double* dxc, int* halo, cudaStream_t st) {
  WalkArgs a;
  walk_args(ctx, a, area, w, ang, s, w_nodata, usew, contcheck, theta, dxc, halo);
  const dim3 blocks((unsigned)s.ny, (unsigned)(((s.pitch >> 4) + 255) / 256));
  for (double& v : ctx->phase_ms) v = 0.;
  PhaseTimer tm(st);
  // passes < 0: as many passes as pay — a pass costs a scan of the whole count array, so stop when the (sampled: every
  // 16th row) number of cells it evaluated falls below 0.4 % of the strip; the walkers finish such remainders faster
  const bool autostop = passes < 0;
  const int maxp = autostop ? 96 : passes;
  const char* te = getenv("TAUDEM_B200_TIMING");
  const bool diag = te && atoi(te) >= 2 && maxp > 0;             // per-pass cell counts on stderr
  if (diag || autostop) { TD_CUDA(ctx->rows.ensure(sizeof(unsigned long long) * (size_t)maxp)); TD_CUDA(cudaMemsetAsync(ctx->rows.p, 0, sizeof(unsigned long long) * (size_t)maxp, st)); }
  const double sampled = (double)((s.ny + 15) / 16) * s.nx;
  int done = 0;
  for (int p = 0; p < maxp; ++p) {
    a.pass_cells = (diag || autostop) ? ctx->rows.as<unsigned long long>() + p : nullptr;
    if (dinf) k_level<true><<<blocks, 256, 0, st>>>(a); else k_level<false><<<blocks, 256, 0, st>>>(a);
    TD_LAUNCHED();
    ++done;
    if (autostop) {
      unsigned long long* h = ctx->h_ctr + 20;
      TD_CUDA(cudaMemcpyAsync(h, ctx->rows.as<unsigned long long>() + p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      TD_CUDA(cudaStreamSynchronize(st));
      if ((double)*h < 0.004 * sampled) break;
    }
  }
  TD_CUDA(cudaGetLastError());
  tm.lap(&ctx->phase_ms[0]);
  if (diag) {
    std::vector<unsigned long long> pc((size_t)done);
    TD_CUDA(cudaMemcpyAsync(pc.data(), ctx->rows.p, sizeof(unsigned long long) * (size_t)done, cudaMemcpyDeviceToHost, st));
    TD_CUDA(cudaStreamSynchronize(st));
    fprintf(stderr, "[k_level %s] %d passes; cells evaluated per pass in every 16th row (of %.0f):", dinf ? "dinf" : "d8", done, sampled);
    for (int p = 0; p < done; ++p) fprintf(stderr, " %llu", pc[p]);
    fprintf(stderr, "\n");
  }
  return TD_OK;
}

// Finishes a sweep from the current dependency state: every cell whose count is zero is evaluated and the
// chains are followed until nothing is ready any more.
int sweep_walk(td_ctx* ctx, bool dinf, float* area, const float* w, const float* ang, const Strip& s, float w_nodata, int usew,
               int contcheck, const double* theta, const double* dxc, int* halo, cudaStream_t st) {
  WalkArgs a;
  walk_args(ctx, a, area, w, ang, s, w_nodata, usew, contcheck, theta, dxc, halo);
  unsigned long long* hc = ctx->h_ctr + 16;
  const dim3 blocks((unsigned)s.ny, (unsigned)(((s.pitch >> 2) + 255) / 256));
  PhaseTimer tm(st);
  TD_CUDA(cudaMemsetAsync(a.ctr, 0, 4 * sizeof(unsigned long long), st));
  k_ready<false><<<blocks, 256, 0, st>>>(a.cntw, a.node, s, a.ctr, nullptr);
  TD_LAUNCHED();
  TD_CUDA(cudaMemcpyAsync(hc, a.ctr, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  TD_CUDA(cudaStreamSynchronize(st));
  unsigned long long n = hc[3];
  if (n == 0) return TD_OK;
  TD_CUDA(ctx->listA.ensure(sizeof(long long) * n));
  k_ready<true><<<blocks, 256, 0, st>>>(a.cntw, a.node, s, a.ctr, ctx->listA.as<long long>());
  TD_LAUNCHED();
  tm.lap(&ctx->phase_ms[1]);
  const unsigned long long cap = (unsigned long long)s.nx * s.ny / 16 + 65536;
  // TAUDEM_B200_RIVER (aread8) / TAUDEM_B200_RIVER_DINF (areadinf) = number of cells after which a lane parks its chain for
  // k_river (0 / unset = off).  Separate switches: on the 1024^2 test field the look-ahead resolves 27 cells per batch for D8
  // and 2.2 for D-infinity (braided strands feed each other, so the next cell is rarely resolvable ahead of time).
  const char* re = getenv(dinf ? "TAUDEM_B200_RIVER_DINF" : "TAUDEM_B200_RIVER");
  const int river = re ? std::max(0, atoi(re)) : 0;
  const char* te = getenv("TAUDEM_B200_TIMING");
  a.diag = (te && atoi(te) >= 2) ? 1 : 0;
  unsigned long long rbatches = 0, rcells = 0;
  const bool lists = dinf || river;
  if (lists) { TD_CUDA(ctx->listB.ensure(sizeof(long long) * cap)); TD_CUDA(ctx->listC.ensure(sizeof(long long) * cap)); }
  const long long* cur = ctx->listA.as<long long>();
  long long* spill = lists ? ctx->listB.as<long long>() : nullptr;
  long long* other = lists ? ctx->listC.as<long long>() : nullptr;
  bool first = true;
  for (;;) {
    TD_CUDA(cudaMemsetAsync(a.ctr, 0, 6 * sizeof(unsigned long long), st));
    a.list = cur; a.nlist = n; a.spill = spill; a.spill_cap = lists ? cap : 0; a.river_hops = river;
    if (first || !river) {
      // chain walking, one chain per lane (the spill list receives overflowing forks and, with `river`, parked chains)
      const int grid = dinf ? walk_grid<true>(n) : walk_grid<false>(n);
      if (grid < 1) { set_error("walk kernel does not fit on an SM"); return TD_ERR_CUDA; }
      if (dinf) k_walk<true><<<grid, 256, 0, st>>>(a); else k_walk<false><<<grid, 256, 0, st>>>(a);
    } else {
      // what was parked or spilled: one warp each, with look-ahead; rivers that meet continue as one (last arrival)
      const int per_dev = dinf ? walk_grid<true>(1ull << 40) : walk_grid<false>(1ull << 40);
      const int rgrid = (int)std::min<unsigned long long>((unsigned long long)std::max(1, per_dev), (n + 7) / 8);
      if (dinf) k_river<true><<<rgrid, 256, 0, st>>>(a); else k_river<false><<<rgrid, 256, 0, st>>>(a);
    }
    TD_LAUNCHED();
    TD_CUDA(cudaGetLastError());
    tm.lap(&ctx->phase_ms[(first || !river) ? 2 : 3]);
    first = false;
    if (!lists) break;                   // D8 chains never fork: nothing can spill
    TD_CUDA(cudaMemcpyAsync(hc, a.ctr, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    TD_CUDA(cudaStreamSynchronize(st));
    if (hc[2]) { set_error("contributing area: ready-cell spill list exhausted"); return TD_ERR_ALLOC; }
    rbatches += hc[4]; rcells += hc[5];
    n = hc[1];
    if (n == 0) break;
    cur = spill; std::swap(spill, other);
  }
  if (a.diag && rbatches) fprintf(stderr, "[k_river %s] %llu cells in %llu look-ahead batches (%.2f per batch)\n", dinf ? "dinf" : "d8", rcells, rbatches, (double)rcells / (double)rbatches);
  return TD_OK;
}

}  // namespace td
