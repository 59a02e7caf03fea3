// Contributing-area evaluation sweep as a tile-granular dataflow (D8 and D-infinity).
//
// reference: aread8 main loop src/aread8.cpp:216-304, area() main loop src/areadinf.cpp:173-265.
// The value of a cell is a k-ordered gather over the neighbours that drain into it, evaluated
// once when all of them are done; the schedule is free, so:
//
//  * the strip is cut into 64 x 32 (D8) / 64 x 16 (D-infinity) tiles; persistent CTAs take tile ids
//    from a device-side multi-producer/multi-consumer ticket queue;
//  * a CTA loads the tile's dependency counts, node words and areas (with a one-cell ring)
//    into shared memory and runs the wavefront INSIDE shared memory: threads start on cells
//    whose count is zero, evaluate them, decrement the downslope cell with a shared-memory
//    atomic and keep following the chain while they are the last arrival (~100 cycles per hop
//    instead of several L2 round trips);
//  * flow that leaves the tile is delivered after the tile's areas are written back and
//    fenced: one global atomic decrement per crossing; the delivery that brings a count to
//    zero activates the owning tile (per-tile state: idle / queued / running / running+dirty,
//    so a tile is never processed by two CTAs at once);
//  * the kernel ends when no tile is queued or running.
// Flow that crosses the strip boundary (one strip per GPU) is either recorded in `halo` for the
// host-driven exchange rounds (src/aread8.cpp:282-297) or, in peer mode, delivered straight into the
// neighbour GPU's counts / halo buffer / tile queue over NVLink (deliver_peer, sched_activate_peer).
#include <string.h>

#include "ctx.h"
#include "dinf_common.cuh"

namespace td {
namespace {

constexpr int TWX = 64;                              // tile width; the height is a kernel template parameter
constexpr int RW = TWX + 2;                          // ring width
constexpr int TH_D8 = 32, TH_DINF = 16;              // D-infinity keeps two doubles per ring cell: smaller tiles, more CTAs per SM
constexpr int EXTCAP = 512;
constexpr unsigned NODE_VALID = 0x8000u, NODE_CON = 0x1000u;

// what a neighbouring strip exposes to this GPU (device pointers into the peer's memory)
struct PeerStrip {
  unsigned* cntw = nullptr; int* state = nullptr; int* tq = nullptr; unsigned long long* ctr = nullptr; float* halo_in = nullptr;
  unsigned qmask = 0; int ntx = 0, ny = 0, th = 0, valid = 0;
};

struct SweepArgs {
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
  int ntx, nty, th;        // tiles across / down, tile height
  int* state;              // per tile: 0 idle, 1 queued, 2 running, 3 running + re-activated
  unsigned char* visited;  // per tile: area interior has been written at least once
  int* tq;                 // ring of tile ids + 1
  unsigned qmask;
  unsigned long long* ctr; // [0] head, [1] tail, [2] pending (queued + running tiles)
  // peer mode (one strip per GPU, neighbours' buffers mapped over NVLink): no exchange rounds, a tile
  // delivers into the neighbour GPU exactly as it delivers into a neighbour tile
  int peer;
  PeerStrip up, down;      // the strip above (rank-1) / below (rank+1)
  unsigned long long* G;   // queued + running tiles of ALL strips (lives on rank 0)
  const float* halo_in;    // areas of the neighbours' edge cells: [0,pitch) row above, [pitch,2 pitch) row below
  int once;                // every tile is visited exactly once: nothing is re-activated (first phase of the hybrid sweep,
                           // sweep_walk.cu finishes the cells that become ready after their tile's visit)
};

#ifdef TD_EMU   // the CPU emulation switches threads at every volatile load (spin loops must let the writer run)
template <typename T> __device__ __forceinline__ T ldv(const T* p) { emu::yield(); return *((const volatile T*)p); }
#else
template <typename T> __device__ __forceinline__ T ldv(const T* p) { return *((const volatile T*)p); }
#endif

// In peer mode a neighbour GPU operates on this strip's counts and scheduler words with system-scope
// atomics; the owner then uses system scope on the same words too (atomics of different scopes on one
// address are not guaranteed to be atomic with respect to each other).
#define A_ADD(ptr, v) (a.peer ? atomicAdd_system((ptr), (v)) : atomicAdd((ptr), (v)))
#define A_CAS(ptr, c, v) (a.peer ? atomicCAS_system((ptr), (c), (v)) : atomicCAS((ptr), (c), (v)))
#define A_EXCH(ptr, v) (a.peer ? atomicExch_system((ptr), (v)) : atomicExch((ptr), (v)))
#define A_FENCE() do { if (a.peer) __threadfence_system(); else __threadfence(); } while (0)

__device__ __forceinline__ void pending_add(const SweepArgs& a, unsigned long long v) {
  if (a.peer) atomicAdd_system(a.G, v); else atomicAdd(a.ctr + 2, v);
}
__device__ __forceinline__ long long pending_now(const SweepArgs& a) { return (long long)ldv(a.peer ? a.G : a.ctr + 2); }

__device__ void sched_push(const SweepArgs& a, int t) {
  pending_add(a, 1ull);
  const unsigned long long slot = A_ADD(a.ctr + 1, 1ull);
  int* q = a.tq + (slot & a.qmask);
  while (A_CAS(q, 0, t + 1) != 0) {}
}

__device__ void sched_activate(const SweepArgs& a, int t) {
  if (a.once) return;
  for (;;) {
    const int st = ldv(a.state + t);
    if (st == 1 || st == 3) return;
    if (st == 0) { if (A_CAS(a.state + t, 0, 1) == 0) { sched_push(a, t); return; } }
    else if (A_CAS(a.state + t, 2, 3) == 2) return;
  }
}

// the same protocol on a neighbour GPU's scheduler (system-scope atomics over NVLink)
__device__ void sched_activate_peer(const SweepArgs& a, const PeerStrip& P, int t) {
  for (;;) {
    const int st = ldv(P.state + t);
    if (st == 1 || st == 3) return;
    if (st == 0) {
      if (atomicCAS_system(P.state + t, 0, 1) == 0) {
        atomicAdd_system(a.G, 1ull);
        const unsigned long long slot = atomicAdd_system(P.ctr + 1, 1ull);
        int* q = P.tq + (slot & P.qmask);
        while (atomicCAS_system(q, 0, t + 1) != 0) {}
        return;
      }
    } else if (atomicCAS_system(P.state + t, 2, 3) == 2) return;
  }
}

// Flow into the strip above (up = true) / below: the source cell's area goes into the neighbour's halo
// buffer, is fenced system-wide, then the neighbour's count is decremented; zero -> queue its tile.
__device__ void deliver_peer(const SweepArgs& a, bool up, int c_src, float val, int c_dst) {
  const PeerStrip& P = up ? a.up : a.down;
  const int pitch = a.s.pitch;
  P.halo_in[(up ? pitch : 0) + c_src] = val;        // I am the row BELOW the strip above / the row ABOVE the strip below
  __threadfence_system();
  const int r = up ? P.ny : 1;
  const long long ci = (long long)r * pitch + c_dst;
  const unsigned sh = (unsigned)(ci & 3) * 8u;
  const unsigned old = atomicAdd_system(P.cntw + (ci >> 2), 0u - (1u << sh));
  if (((old >> sh) & 0xffu) == 1u) sched_activate_peer(a, P, ((r - 1) / P.th) * P.ntx + c_dst / TWX);
}

// Ticket queue: one fetch-and-add per pop (a CAS loop on the head collapses under the
// contention of ~900 persistent CTAs).  Ticket h is served by the h-th push; a consumer whose
// ticket is never served leaves when no tile is queued or running any more.
__device__ int sched_pop(const SweepArgs& a) {
  const unsigned long long h = A_ADD(a.ctr, 1ull);
  int* q = a.tq + (h & a.qmask);
  for (;;) {
    const int v = ldv(q);
    if (v != 0) {
      A_EXCH(q, 0);
      A_EXCH(a.state + (v - 1), 2);
      atomicAdd(a.ctr + 3, 1ull);          // statistics: tile visits
      if (a.peer) __threadfence_system(); else __threadfence();
      return v - 1;
    }
    if (pending_now(a) <= 0) return -1;
    __nanosleep(64);
  }
}

__device__ void sched_finish(const SweepArgs& a, int t) {
  if (a.peer) __threadfence_system(); else __threadfence();
  if (A_CAS(a.state + t, 2, 0) != 2) { A_EXCH(a.state + t, 1); sched_push(a, t); }
  pending_add(a, ~0ull);   // pending -= 1
}

__global__ void k_sched_init(int* state, unsigned char* visited, int* tq, unsigned qcap, int ntiles, unsigned long long* ctr) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < qcap) tq[i] = (int)i < ntiles ? (int)i + 1 : 0;
  if ((int)i < ntiles) { state[i] = 1; visited[i] = 0; }
  if (i == 0) { ctr[0] = 0; ctr[1] = (unsigned long long)ntiles; ctr[2] = (unsigned long long)ntiles; ctr[3] = 0; ctr[4] = ctr[5] = ctr[6] = ctr[7] = 0; }
}

template <bool DINF, int TWY>
__global__ void __launch_bounds__(256) k_sweep_tiles(const SweepArgs a) {
  constexpr int TC = TWX * TWY, RH = TWY + 2, NWD = TC / 4 / 256;   // cells, ring rows, count words per thread
  // dynamic shared memory carve-up (doubles first)
  extern __shared__ __align__(16) unsigned char dsm[];
  double* sp1 = reinterpret_cast<double*>(dsm);                       // DINF: share of flow to sk1 / sk2 per ring cell
  double* sp2 = sp1 + (DINF ? RH * RW : 0);
  float* sarea = reinterpret_cast<float*>(sp2 + (DINF ? RH * RW : 0));
  float* sang = sarea + RH * RW;                                      // DINF: angles of the ring
  float* sw = sang + (DINF ? RH * RW : 0);                            // weights (only with -wg)
  int* lc = reinterpret_cast<int*>(sw + (a.usew ? TC : 0));
  unsigned short* snode = reinterpret_cast<unsigned short*>(lc + TC);
  unsigned short* ext = snode + TC;
  unsigned short* wq = ext + EXTCAP;                                  // ready cells (local index), 0xFFFF = not yet published
  unsigned char* sk1 = reinterpret_cast<unsigned char*>(wq + TC);     // DINF: the (at most two) receiving directions
  unsigned char* sk2 = sk1 + RH * RW;
  __shared__ int next, cur_tile, self_dirty, wq_head, wq_tail, outstanding;
  __shared__ double saref[DINF ? RH * 10 : 1];                     // prop()'s aref[] table of every ring row
  const Strip& s = a.s;
  const int tid = threadIdx.x;

  for (;;) {
    long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0;
    if (tid == 0) { tk0 = clock64(); cur_tile = sched_pop(a); next = 0; self_dirty = 0; wq_head = 0; wq_tail = 0; outstanding = 0; tk1 = clock64(); }
    __syncthreads();
    const int t = cur_tile;
    if (t < 0) return;
    const int tx = t % a.ntx, ty = t / a.ntx;
    const int c0 = tx * TWX, r0 = 1 + ty * TWY;

    // ---- 1. dependency counts (4 cells per word, 2 words per thread), before anything else
    unsigned g0[NWD];
#pragma unroll
    for (int j = 0; j < NWD; ++j) {
      const int lb = 4 * (tid + 256 * j);
      const int lr = lb / TWX, lx = lb % TWX;
      const int r = r0 + lr, c = c0 + lx;
      unsigned word = 0xffffffffu;
      if (r <= s.ny && c < s.pitch) word = __ldcg(a.cntw + (s.idx(r, c) >> 2));
      g0[j] = word;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const unsigned b = (word >> (8 * i)) & 0xffu; lc[lb + i] = b <= 8u ? (int)b : -1; }
    }
    A_FENCE();
    __syncthreads();
    // ---- 2. node words, areas (+ring), angles (+ring), weights
    const bool first = ldv(a.visited + t) == 0;
#pragma unroll
    for (int j = 0; j < NWD; ++j) {
      const int lb = 4 * (tid + 256 * j);
      const int lr = lb / TWX, lx = lb % TWX;
      const int r = r0 + lr, c = c0 + lx;
      ushort4 nv = make_ushort4(0, 0, 0, 0);
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r <= s.ny && c < s.pitch) {
        nv = *reinterpret_cast<const ushort4*>(a.node + s.idx(r, c));
        if (a.usew) wv = *reinterpret_cast<const float4*>(a.w + s.idx(r, c));
      }
      *reinterpret_cast<ushort4*>(snode + lb) = nv;
      if (a.usew) { sw[lb] = wv.x; sw[lb + 1] = wv.y; sw[lb + 2] = wv.z; sw[lb + 3] = wv.w; }
      *reinterpret_cast<ushort4*>(wq + lb) = make_ushort4(0xffff, 0xffff, 0xffff, 0xffff);
    }
    for (int i = tid; i < RH * RW; i += 256) {
      const int rr = i / RW, rc = i - rr * RW;
      const int r = r0 - 1 + rr, c = c0 - 1 + rc;
      const bool in = r >= 0 && r <= s.ny + 1 && c >= 0 && c < s.nx;
      const bool interior = rr >= 1 && rr <= TWY && rc >= 1 && rc <= TWX && r <= s.ny;   // owned cell of this tile (a halo row can fall inside a partial tile)
      float v = -1.0f;
      if (in && !(first && interior)) {
        if (a.peer && (r == 0 || r == s.ny + 1)) v = __ldcg(a.halo_in + (r == 0 ? 0 : s.pitch) + c);
        else v = __ldcg(a.area + s.idx(r, c));
      }
      sarea[i] = v;
      if (DINF) sang[i] = in ? a.ang[s.idx(r, c)] : TD_MISSINGFLOAT;
    }
    if (DINF) {
      for (int i = tid; i < RH * 10; i += 256) saref[i] = aref(i % 10, a.theta[min(max(r0 - 2 + i / 10, 0), s.ny - 1)]);
      __syncthreads();
      // prop() of every ring cell once, in parallel and off the wavefront's critical path:
      // a cell sends flow to at most two (adjacent) neighbours (src/commonLib.cpp:76-91)
      for (int i = tid; i < RH * RW; i += 256) {
        const int rr = i / RW, rc = i - rr * RW;
        const int r = r0 - 1 + rr, c = c0 - 1 + rc;
        Outflow o; o.k1 = o.k2 = 0; o.p1 = o.p2 = 0.;
        if (r >= 0 && r <= s.ny + 1 && c >= 0 && c < s.nx) o = dinf_outflow(sang[i], saref + rr * 10);
        sk1[i] = (unsigned char)o.k1; sk2[i] = (unsigned char)o.k2; sp1[i] = o.p1; sp2[i] = o.p2;
      }
    }
    __syncthreads();

    // ---- 3. the wavefront inside the tile
    if (tid == 0) tk2 = clock64();
    int npass = 0;
    {
      // D-infinity: every thread publishes the ready cells among its own eight to the shared work queue
      // (a cell can make two neighbours ready; idle threads take the second one at once).
      // D8: a chain never forks, every thread simply walks the chains that start in its own cells
      // (threads spinning on a queue only slow the walkers down).
      int nready = 0;
      unsigned ready = 0;
#pragma unroll
      for (int j = 0; j < NWD; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int l = 4 * (tid + 256 * j) + i;
          if (lc[l] == 0) {
            if (DINF) { wq[atomicAdd(&wq_tail, 1)] = (unsigned short)l; ++nready; }
            else ready |= 1u << (4 * j + i);
          }
        }
      if (nready) atomicAdd(&outstanding, nready);
      __syncthreads();                       // all publications (and `outstanding`) are complete before any count moves
      // work loop: take a ready cell, follow its chain; `outstanding` = live chains + queued cells
      for (;;) {
        int l = -1;
        if (DINF) {
          const int h = atomicAdd(&wq_head, 1);   // ticket order
          for (;;) {
            if (h < ldv(&wq_tail)) { unsigned short v; while ((v = ldv(wq + h)) == 0xffff) {} l = v; break; }
            if (ldv(&outstanding) <= 0) break;
            __nanosleep(100);
          }
        } else if (ready) {
          const int bit = __ffs(ready) - 1;
          ready &= ready - 1;
          l = 4 * (tid + 256 * (bit >> 2)) + (bit & 3);
        }
        if (l < 0) break;
        ++npass;
        __threadfence_block();
        for (;;) {                            // follow the chain while we are the last arrival
          const int lr = l / TWX, lx = l % TWX;
          const int ri = (lr + 1) * RW + lx + 1;
          const unsigned nd = snode[l];
          const unsigned msk = nd & 0xffu;
          bool con = (nd & NODE_CON) != 0;
          float val;
          if (!DINF) {
            // src/aread8.cpp:228-257
            if (a.usew) { const float wv = sw[l]; val = nd_f(wv, a.w_nodata) ? -1.0f : wv; }
            else val = 1.0f;
#pragma unroll
            for (int k = 1; k <= 8; ++k)
              if (msk & (1u << (k - 1))) {
                const float an = sarea[ri + drow(k) * RW + dcol(k)];
                if (nd_f(an, -1.0f)) con = true; else val = val + an;
              }
          } else {
            // src/areadinf.cpp:187-218
            val = 0.f;
#pragma unroll
            for (int k = 1; k <= 8; ++k)
              if (msk & (1u << (k - 1))) {
                const int ni = ri + drow(k) * RW + dcol(k);
                const int kk = k > 4 ? k - 4 : k + 4;            // the direction from that neighbour to this cell
                const double p = sk1[ni] == kk ? sp1[ni] : sp2[ni];
                const float an = sarea[ni];
                if (nd_f(an, -1.0f)) con = true; else val = (float)((double)val + p * (double)an);
              }
            if (a.usew) val = val + sw[l];
            else val = (float)((double)val + a.dxc[r0 + lr - 1]);
          }
          if (con && a.contcheck) val = -1.0f;
          sarea[ri] = val;
          lc[l] = -1;
          __threadfence_block();     // the area is in shared memory before any count says so
          // ---- decrement the downslope cell(s)
          int cont = -1;
          if (!DINF) {
            const int d = (int)((nd >> 8) & 0xfu);
            if (d >= 1 && d <= 8) {
              const int nlr = lr + drow(d), nlx = lx + dcol(d);
              if (nlr >= 0 && nlr < TWY && nlx >= 0 && nlx < TWX && r0 + nlr <= s.ny) {   // owned cell of this tile
                const int l2 = nlr * TWX + nlx;
                if (atomicSub(&lc[l2], 1) == 1) cont = l2;   // invalid / finished cells hold a negative count
              } else if (s.on_grid(r0 + nlr, c0 + nlx)) {
                if (a.peer && (r0 + nlr == 0 || r0 + nlr == s.ny + 1)) deliver_peer(a, r0 + nlr == 0, c0 + lx, val, c0 + nlx);
                else ext[atomicAdd(&next, 1)] = (unsigned short)((nlr + 1) * RW + nlx + 1);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int k = j == 0 ? sk1[ri] : sk2[ri];
              if (k == 0) continue;
              const int nlr = lr + drow(k), nlx = lx + dcol(k);
              if (nlr >= 0 && nlr < TWY && nlx >= 0 && nlx < TWX && r0 + nlr <= s.ny) {   // owned cell of this tile
                const int l2 = nlr * TWX + nlx;
                if (atomicSub(&lc[l2], 1) == 1) {
                  if (cont < 0) cont = l2;
                  else {                        // a second ready neighbour: hand it to an idle thread
                    atomicAdd(&outstanding, 1);
                    __threadfence_block();
                    wq[atomicAdd(&wq_tail, 1)] = (unsigned short)l2;
                  }
                }
              } else if (s.on_grid(r0 + nlr, c0 + nlx)) {
                if (a.peer && (r0 + nlr == 0 || r0 + nlr == s.ny + 1)) deliver_peer(a, r0 + nlr == 0, c0 + lx, val, c0 + nlx);
                else ext[atomicAdd(&next, 1)] = (unsigned short)((nlr + 1) * RW + nlx + 1);
              }
            }
          }
          if (cont < 0) break;
          __threadfence_block();     // counts first, then the areas they announce
          l = cont;
        }
        if (DINF) atomicSub(&outstanding, 1);  // this chain has ended
      }
      __syncthreads();
    }

    // ---- 4. write the areas back, then publish counts and deliver the crossings
    if (tid == 0) tk3 = clock64();
    for (int i = tid; i < TC; i += 256) {
      const int lr = i / TWX, lx = i - lr * TWX;
      const int r = r0 + lr, c = c0 + lx;
      if (r <= s.ny && c < s.nx) a.area[s.idx(r, c)] = sarea[(lr + 1) * RW + lx + 1];
    }
    if (tid == 0) a.visited[t] = 1;
    A_FENCE();
    __syncthreads();
    A_FENCE();
#pragma unroll
    for (int j = 0; j < NWD; ++j) {
      const int lb = 4 * (tid + 256 * j);
      const int lr = lb / TWX, lx = lb % TWX;
      const int r = r0 + lr, c = c0 + lx;
      if (!(r <= s.ny && c < s.pitch)) continue;
      unsigned delta = 0; int dec[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = (int)((g0[j] >> (8 * i)) & 0xffu);
        const int now = lc[lb + i];
        int ci = 0;
        if (b <= 8) ci = now < 0 ? 0xfe - b : now - b;     // evaluated -> 0xFE (done); else minus the local arrivals
        dec[i] = ci;
        delta += (unsigned)ci << (8 * i);
      }
      if (delta != 0) {
        const unsigned old = A_ADD(a.cntw + (s.idx(r, c) >> 2), delta);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (dec[i] < 0 && (int)((old >> (8 * i)) & 0xffu) + dec[i] == 0) self_dirty = 1;   // became ready meanwhile
      }
    }
    const int ne = next;
    for (int e = tid; e < ne; e += 256) {
      const int code = ext[e];
      const int rr = code / RW, rc = code - rr * RW;
      const int r = r0 - 1 + rr, c = c0 - 1 + rc;
      if (r == 0 || r == s.ny + 1) { atomicAdd(a.halo + (r == 0 ? 0 : s.pitch) + c, 1); continue; }
      const long long ci = s.idx(r, c);
      if (!(a.node[ci] & NODE_VALID)) continue;
      const unsigned sh = (unsigned)(ci & 3) * 8u;
      const unsigned old = A_ADD(a.cntw + (ci >> 2), 0u - (1u << sh));
      if (((old >> sh) & 0xffu) == 1u) sched_activate(a, ((r - 1) / TWY) * a.ntx + c / TWX);
    }
    __syncthreads();
    if (tid == 0) {
      if (self_dirty) sched_activate(a, t);
      sched_finish(a, t);
      // statistics (cycles, summed over CTAs): queue wait, load, wavefront, write-back + deliveries
      const long long tk4 = clock64();
      atomicAdd(a.ctr + 4, (unsigned long long)(tk1 - tk0));
      atomicAdd(a.ctr + 5, (unsigned long long)(tk2 - tk1));
      atomicAdd(a.ctr + 6, (unsigned long long)(tk3 - tk2));
      atomicAdd(a.ctr + 7, (unsigned long long)(tk4 - tk3));
      (void)npass;
    }
    __syncthreads();
  }
}
size_t smem_bytes(bool dinf, bool usew) {
  const size_t th = dinf ? TH_DINF : TH_D8, tc = TWX * th, ring = (th + 2) * RW;
  return (dinf ? 2 * ring * sizeof(double) : 0) + (ring + (dinf ? ring : 0) + (usew ? tc : 0)) * sizeof(float) + tc * sizeof(int) +
         (tc + EXTCAP + tc) * sizeof(unsigned short) + 2 * ring;
}
}  // namespace

namespace {
// Applies the dependency decrements received from the neighbour strips (addBorders,
// src/linearpart.h:314-328 and src/aread8.cpp:283-297): dec_top[c] arrivals for the cell (row 1, c),
// dec_bot[c] for (row ny, c).  A count that reaches zero queues the cell's tile.
__global__ void k_apply_halo(SweepArgs a, const int* __restrict__ dec_top, const int* __restrict__ dec_bot) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.s.nx) return;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int* dec = side == 0 ? dec_top : dec_bot;
    if (dec == nullptr) continue;
    const int d = dec[c];
    if (d <= 0) continue;
    const int r = side == 0 ? 1 : a.s.ny;
    const long long ci = a.s.idx(r, c);
    if (!(a.node[ci] & NODE_VALID)) continue;
    const unsigned sh = (unsigned)(ci & 3) * 8u;
    const unsigned old = atomicAdd(a.cntw + (ci >> 2), 0u - ((unsigned)d << sh));
    if ((int)((old >> sh) & 0xffu) == d) sched_activate(a, ((r - 1) / a.th) * a.ntx + c / TWX);
  }
}

__global__ void k_sched_reset(unsigned long long* ctr) { ctr[0] = ctr[1] = ctr[2] = 0; }

int sweep_args(td_ctx* ctx, SweepArgs& a, const Strip& s) {
  a.s = s;
  // fields only the sweep kernel fills; the scheduler helpers (k_apply_halo) read a.peer / a.G
  a.peer = 0; a.G = nullptr; a.halo_in = nullptr; a.up = PeerStrip(); a.down = PeerStrip(); a.once = 0;
  a.area = nullptr; a.w = nullptr; a.ang = nullptr; a.usew = 0; a.contcheck = 1; a.w_nodata = 0.f; a.theta = nullptr; a.dxc = nullptr; a.halo = nullptr;
  a.th = ctx->sweep_dinf ? TH_DINF : TH_D8;
  a.ntx = (s.nx + TWX - 1) / TWX; a.nty = (s.ny + a.th - 1) / a.th;
  const long long nt = (long long)a.ntx * a.nty;
  if (nt > (1ll << 30)) { set_error("strip has too many tiles"); return TD_ERR_ARG; }
  unsigned qcap = 1u << 14;   // always far more slots than persistent CTAs holding tickets
  while (qcap < (unsigned long long)nt) qcap <<= 1;
  TD_CUDA(ctx->tileflags.ensure((size_t)nt * 4 + (size_t)qcap * 4 + (size_t)nt));
  a.state = ctx->tileflags.as<int>();
  a.tq = a.state + nt;
  a.visited = reinterpret_cast<unsigned char*>(a.tq + qcap);
  a.qmask = qcap - 1;
  a.ctr = ctx->d_ctr + 24;
  a.node = ctx->node.as<unsigned short>();
  a.cntw = ctx->cnt.as<unsigned>();
  return TD_OK;
}
}  // namespace

// Queues every tile of the strip (start of a sweep).
int sweep_begin(td_ctx* ctx, const Strip& s, cudaStream_t st) {
  SweepArgs a;
  if (int rc = sweep_args(ctx, a, s)) return rc;
  const int nt = a.ntx * a.nty;
  k_sched_init<<<(a.qmask + 1 + 255) / 256, 256, 0, st>>>(a.state, a.visited, a.tq, a.qmask + 1, nt, a.ctr);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

// Decrements that crossed the strip boundary (from the neighbours' halo records); queues the tiles
// whose cells became ready.  Must be called between two sweep_run calls.
int sweep_apply_halo(td_ctx* ctx, const Strip& s, const int* dec_top, const int* dec_bot, cudaStream_t st) {
  SweepArgs a;
  if (int rc = sweep_args(ctx, a, s)) return rc;
  k_sched_reset<<<1, 1, 0, st>>>(a.ctr);   // tickets abandoned at the end of the previous run are void
  TD_LAUNCHED();
  k_apply_halo<<<(s.nx + 255) / 256, 256, 0, st>>>(a, dec_top, dec_bot);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

namespace {
__global__ void k_add_G(unsigned long long* G, unsigned long long v) { atomicAdd_system(G, v); __threadfence_system(); }
}

// ---- peer mode plumbing (CUDA IPC).  export: make sure every buffer a neighbour touches exists at its
// final size and hand out its IPC handle; connect: open a neighbour's (or rank 0's counter) handles.
int sweep_peer_export(td_ctx* ctx, const Strip& s, int dinf, unsigned char* handles, int* meta, cudaStream_t st) {
  SweepArgs a;
  ctx->sweep_dinf = dinf ? 1 : 0;    // tile geometry of the sweep that is about to run
  const size_t n = (size_t)s.cells();
  TD_CUDA(ctx->node.ensure(n * 2));
  TD_CUDA(ctx->cnt.ensure((n + 3) / 4 * 4));
  if (int rc = sweep_args(ctx, a, s)) return rc;
  TD_CUDA(ctx->peer_halo.ensure(sizeof(float) * 2 * (size_t)s.pitch));
  TD_CUDA(ctx->gbuf.ensure(64));
  TD_CUDA(cudaMemsetAsync(ctx->gbuf.p, 0, 64, st));
  TD_CUDA(cudaStreamSynchronize(st));
  void* ptrs[5] = {ctx->cnt.p, ctx->tileflags.p, ctx->d_ctr, ctx->peer_halo.p, ctx->gbuf.p};
  for (int i = 0; i < 5; ++i) {
    cudaIpcMemHandle_t h;
    TD_CUDA(cudaIpcGetMemHandle(&h, ptrs[i]));
    memcpy(handles + 64 * i, &h, 64);
  }
  meta[0] = (int)a.qmask; meta[1] = a.ntx; meta[2] = s.ny; meta[3] = a.th; meta[4] = a.ntx * a.nty;
  return TD_OK;
}

static void close_peer(td_ctx::PeerInfo& pi) {
  if (pi.cntw) cudaIpcCloseMemHandle(pi.cntw);
  if (pi.tileflags) cudaIpcCloseMemHandle(pi.tileflags);
  if (pi.dctr) cudaIpcCloseMemHandle(pi.dctr);
  if (pi.halo_in) cudaIpcCloseMemHandle(pi.halo_in);
  pi = td_ctx::PeerInfo();
}

// which: 0 = strip above, 1 = strip below, 2 = owner of the global counter (handles == NULL: this rank)
int sweep_peer_connect(td_ctx* ctx, int which, const unsigned char* handles, const int* meta) {
  auto open = [](const unsigned char* h64, void** out) -> cudaError_t {
    cudaIpcMemHandle_t h; memcpy(&h, h64, 64);
    return cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
  };
  if (which == 2) {
    if (ctx->peer_G_opened && ctx->peer_G) cudaIpcCloseMemHandle(ctx->peer_G);
    ctx->peer_G_opened = false;
    if (!handles) { ctx->peer_G = ctx->gbuf.p; return TD_OK; }
    TD_CUDA(open(handles + 64 * 4, &ctx->peer_G));
    ctx->peer_G_opened = true;
    return TD_OK;
  }
  td_ctx::PeerInfo& pi = which == 0 ? ctx->peer_up : ctx->peer_down;
  if (!handles && meta && pi.valid) {       // same buffers, new tile geometry (D8 <-> D-infinity sweep)
    pi.qmask = meta[0]; pi.ntx = meta[1]; pi.ny = meta[2]; pi.th = meta[3]; pi.nt = meta[4];
    return TD_OK;
  }
  close_peer(pi);
  if (!handles) return TD_OK;
  TD_CUDA(open(handles, &pi.cntw));
  TD_CUDA(open(handles + 64, &pi.tileflags));
  TD_CUDA(open(handles + 128, &pi.dctr));
  TD_CUDA(open(handles + 192, &pi.halo_in));
  pi.qmask = meta[0]; pi.ntx = meta[1]; pi.ny = meta[2]; pi.th = meta[3]; pi.nt = meta[4]; pi.valid = 1;
  return TD_OK;
}

// start of a peer-mode sweep: queue all tiles, announce them in the global counter.  The caller must put a
// barrier between this call and sweep_run on every rank (nobody may see G == 0 before everybody announced).
int sweep_peer_begin(td_ctx* ctx, const Strip& s, cudaStream_t st) {
  ctx->peer_on = 1;
  if (int rc = sweep_begin(ctx, s, st)) return rc;
  SweepArgs a;
  if (int rc = sweep_args(ctx, a, s)) return rc;
  TD_CUDA(cudaMemsetAsync(ctx->peer_halo.p, 0, sizeof(float) * 2 * (size_t)s.pitch, st));
  k_add_G<<<1, 1, 0, st>>>((unsigned long long*)ctx->peer_G, (unsigned long long)a.ntx * a.nty);
  TD_LAUNCHED();
  TD_CUDA(cudaStreamSynchronize(st));
  return TD_OK;
}
void sweep_peer_off(td_ctx* ctx) { ctx->peer_on = 0; }

// Runs the evaluation wavefront over the queued tiles until no tile of the strip has a ready cell left.
int sweep_run(td_ctx* ctx, bool dinf, float* area, const float* w, const float* ang, const Strip& s, float w_nodata, int usew,
              int contcheck, const double* theta, const double* dxc, int* halo, cudaStream_t st) {
  SweepArgs a;
  if (int rc = sweep_args(ctx, a, s)) return rc;
  a.area = area; a.w = w; a.ang = ang; a.usew = usew; a.contcheck = contcheck;
  a.w_nodata = w_nodata; a.theta = theta; a.dxc = dxc; a.halo = halo;
  a.peer = ctx->peer_on;
  a.once = ctx->sweep_once;
  if (a.peer && a.once) { set_error("the single-pass tile sweep does not run in peer mode"); return TD_ERR_ARG; }
  if (a.peer) {
    auto fill = [](const td_ctx::PeerInfo& pi, PeerStrip& P) {
      P.valid = pi.valid;
      P.cntw = (unsigned*)pi.cntw; P.state = (int*)pi.tileflags; P.tq = P.state + pi.nt; P.ctr = (unsigned long long*)pi.dctr + 24;
      P.halo_in = (float*)pi.halo_in; P.qmask = (unsigned)pi.qmask; P.ntx = pi.ntx; P.ny = pi.ny; P.th = pi.th;
    };
    fill(ctx->peer_up, a.up); fill(ctx->peer_down, a.down);
    a.G = (unsigned long long*)ctx->peer_G;
    a.halo_in = ctx->peer_halo.as<float>();
    if ((s.has_top && !a.up.valid) || (s.has_bot && !a.down.valid) || !a.G) { set_error("peer mode: neighbours are not connected"); return TD_ERR_ARG; }
  } else { a.G = nullptr; a.halo_in = nullptr; a.up.valid = a.down.valid = 0; }
  const long long nt = (long long)a.ntx * a.nty;
  static int grid_d8 = 0, grid_dinf = 0;
  int& grid = dinf ? grid_dinf : grid_d8;
  if (!grid) {
    int dev = 0, sms = 0, occ = 0;
    TD_CUDA(cudaGetDevice(&dev));
    TD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (dinf) {
      TD_CUDA(cudaFuncSetAttribute(k_sweep_tiles<true, TH_DINF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(true, true)));
      TD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_tiles<true, TH_DINF>, 256, smem_bytes(true, false)));
    } else {
      TD_CUDA(cudaFuncSetAttribute(k_sweep_tiles<false, TH_D8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(false, true)));
      TD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sweep_tiles<false, TH_D8>, 256, smem_bytes(false, false)));
    }
    if (occ < 1) { set_error("sweep kernel does not fit on an SM"); return TD_ERR_CUDA; }
    grid = sms * occ;     // persistent: every CTA is resident, so queue waits cannot deadlock
  }
  const int g = (int)std::min<long long>(grid, nt);
  if (dinf) k_sweep_tiles<true, TH_DINF><<<g, 256, smem_bytes(true, usew != 0), st>>>(a);
  else k_sweep_tiles<false, TH_D8><<<g, 256, smem_bytes(false, usew != 0), st>>>(a);
  TD_LAUNCHED();
  TD_CUDA(cudaGetLastError());
  return TD_OK;
}

}  // namespace td
