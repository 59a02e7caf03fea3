// D-infinity contributing area: the dependency stencil (the evaluation sweep is sweep_warp.cu).
//
// reference: prop()              src/commonLib.cpp:76-91  (share of a cell's flow going to neighbour k)
//            initNeighborDinfup  src/commonLib.cpp:92-136 (in-degree = #neighbours with prop > 0)
//            area() main loop    src/areadinf.cpp:173-265 (k-ordered gather
//                                areares = (float)(areares + p*area_n), + weight or dxc[row],
//                                then decrement every neighbour that receives flow).
// node word of a D-infinity cell: bits 0-7 = which neighbours drain into it, bits 8-11 = its receiver field (dinf_field,
// dinf_common.cuh: the first receiving direction k1 and how its shares are obtained), 0x2000 = it has a second receiver
// (always k1 % 8 + 1), 0x1000 = contaminated, 0x8000 = valid.  Halo rows get the receiver bits alone (no VALID): the sweep's
// gather needs them for contributors that belong to the neighbour strip.
// prop's table aref[] = {-t,0,t,PI/2,PI-t,PI,PI+t,3PI/2,2PI-t,2PI} with t = atan2(dy,dx)
// is rebuilt on the device from t (host glibc atan2, per row) using only +,-: the same
// doubles as the reference.  The per-cell value is a deterministic gather, so any
// topological schedule reproduces it.
#include <stdlib.h>

#include "ctx.h"
#include "dinf_common.cuh"

namespace td {
namespace {
constexpr int TW = 128, TH = 32;

// bit 7 of every byte of the result = (that byte of w == that byte of t); all bytes of w ^ t must be < 0x80
__device__ __forceinline__ unsigned eq_bytes7(unsigned w, unsigned t) { return ~((w ^ t) + 0x7f7f7f7fu) & 0x80808080u; }

__global__ void __launch_bounds__(256) k_deps_dinf(const float* __restrict__ ang, unsigned short* __restrict__ node,
                                                   unsigned char* __restrict__ cnt, float* __restrict__ area, Strip s,
                                                   float nodata, const double* __restrict__ theta, float area_init, int edge_fast) {
  using G = TileGeom<float, TW, TH>;
  __shared__ __align__(128) float tile[G::ELEMS];
  __shared__ __align__(8) uint64_t bar;
  const int c0 = blockIdx.x * TW, r0 = 1 + blockIdx.y * TH;
  load_tile_tma<float, TW, TH>(tile, &bar, ang, s, r0, c0);
  // one byte per staged cell (dinf_node_code): k1 | 0x10 if there is a second receiver (always the next direction, k1 % 8 + 1) |
  // 0x40 / 0x80 how the shares are obtained | 0x20 if the cell is off the grid or nodata; one prop() interval search each
  __shared__ double saref[(TH + 2) * 10];
  __shared__ __align__(16) unsigned char sout[G::ELEMS];
  __shared__ __align__(16) unsigned char sbits[G::ELEMS];     // the receiver bits of the node word's high byte (dinf_node_bits >> 8)
  __shared__ float sarf[(TH + 2) * 10];                          // the same table rounded to float: the pre-screen below
  for (int i = threadIdx.x; i < (TH + 2) * 10; i += 256) { const double v = aref(i % 10, theta_of_row(theta, s.ny, min(r0 - 1 + i / 10, s.ny + 1))); saref[i] = v; sarf[i] = (float)v; }
  __syncthreads();
  for (int i = threadIdx.x; i < G::ELEMS; i += 256) {
    const int t = i / G::SW, sc = i - t * G::SW;
    const int gr = r0 - 1 + t, gc = c0 - G::HP + sc;
    unsigned char code = 0x20;
    if (s.on_grid(gr, gc)) {
      const float av = tile[i];
      if (!nd_f(av, nodata)) {
        // Pre-screen in float: an angle well inside a sector (further than 4e-5 sector widths from both edges — float rounding
        // of the table and of the differences is three orders of magnitude below that) has both shares far above prop()'s
        // 1e-5 threshold: receivers j and j + 1, regular, nothing to decide in double.  Everything near or on an edge
        // (flow exactly along a direction is common) takes the exact path.
        const float* af = sarf + t * 10;
        int j = 0;
#pragma unroll
        for (int e = 1; e <= 9; ++e) j += (av >= af[e]) ? 1 : 0;
        const int jc = min(max(j, 1), 8);
        const float lo = af[jc], hi = af[jc + 1], g = 4e-5f * (hi - lo);
        const float dlo = av - lo, dhi = hi - av;
        if (j >= 1 && j <= 8 && dlo > g && dhi > g) code = (unsigned char)((unsigned)j | 0x10u);
        else {
          // On or next to a direction's own angle (every cell of a resolved flat, every clipped facet: a quarter of a real DEM):
          // closer to the edge e than half of prop()'s 1e-5 band (minus the float error of the table) the share of the sector's
          // other direction is dropped for certain, and which side of the edge the angle lies on — one exact comparison — says
          // how the single receiver e is coded: on or above the edge it is the sector's first direction (e; the wrap sector's
          // direction 8 likewise), below it the upper direction of the sector underneath (e | 0x40; 8 | 0x80 below direction 8).
          // dinf_node_code returns exactly that (its j = 1..7 branch with one share ~1 and one < 1e-5).  Anything else near an
          // edge takes the interval search.
          const bool low = fabsf(dlo) <= fabsf(dhi);
          const int e = low ? jc : jc + 1;
          const float ge = 0.5e-5f * fminf(af[e] - af[e - 1], af[min(e + 1, 9)] - af[e]) - 4e-7f;
          unsigned fast = 0xffu;
          if (edge_fast && j >= 1 && j <= 8 && e <= 8 && fabsf(low ? dlo : dhi) < ge) {
            const bool above = (double)av >= saref[t * 10 + e];
            if (above) fast = (unsigned)e;
            else if (e >= 2) fast = e <= 7 ? ((unsigned)e | 0x40u) : (8u | 0x80u);
          }
          code = fast != 0xffu ? (unsigned char)fast : (unsigned char)dinf_node_code(av, saref + t * 10);
        }
      }
    }
    sout[i] = code;
    const unsigned nb = (code & 0x20u) ? 0u : dinf_node_bits(code);
    sbits[i] = (unsigned char)(nb >> 8);
    // halo rows of the strip (cells of the neighbour strips): receiver bits only, written by the tiles next to them
    if ((gr == 0 || gr == s.ny + 1) && sc >= G::HP && sc < G::HP + TW && gc < s.pitch && (t == 0 || t == s.ny + 2 - r0))
      node[s.idx(gr, gc)] = (unsigned short)nb;
  }
  __syncthreads();
  // four adjacent cells per thread with byte-parallel arithmetic: neighbour k drains into me when one of its
  // receiving directions is kk = (k+4)%8 (src/commonLib.cpp:105-134), i.e. k1 == kk, or k1 == kk-1 with a second receiver
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned* soutw = reinterpret_cast<const unsigned*>(sout);
  constexpr int QW = G::SW / 4;
#pragma unroll 1
  for (int pass = 0; pass < TH / 8; ++pass) {
    const int tr = warp + 8 * pass;
    const int r = r0 + tr, c = c0 + lane * 4;
    if (r > s.ny || c >= s.pitch) continue;
    unsigned W[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const unsigned* q = soutw + (tr + j) * QW + lane;      // word lane + 1 holds the cells c .. c+3 (HP = 4 columns of padding)
      const unsigned wl = q[0], wc = q[1], wr = q[2];
      W[j][0] = __funnelshift_l(wl, wc, 8);
      W[j][1] = wc;
      W[j][2] = __funnelshift_r(wc, wr, 8);
    }
    unsigned mb = 0, all = 0;
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      const unsigned wk = W[1 + drow(k)][1 + dcol(k)];
      const unsigned kk = k > 4 ? k - 4 : k + 4, prev = kk == 1 ? 8u : kk - 1u;
      const unsigned z = eq_bytes7(wk & 0x0f0f0f0fu, kk * 0x01010101u) | eq_bytes7(wk & 0x1f1f1f1fu, (0x10u | prev) * 0x01010101u);
      mb |= z >> (8 - k);
      all |= wk;
    }
    const unsigned wc = W[1][1];
    const unsigned vm = ((~wc >> 5) & 0x01010101u) * 0xffu;              // 0xff per cell of the flow field
    unsigned x = mb - ((mb >> 1) & 0x55555555u);                          // per-byte population count
    x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
    x = (x + (x >> 4)) & 0x0f0f0f0fu;
    const unsigned cw = (x & vm) | ~vm;                                   // count, or 0xff outside the field
    // VALID | CON (a neighbour off the grid or nodata) | the cell's own receivers: k1 in bits 8-11, 0x2000 = a second one (k1 % 8 + 1)
    // VALID | CON (a neighbour off the grid or nodata) | the cell's own receiver bits (dinf_node_bits: field in bits 8-11, 0x2000)
    const unsigned hb = (0x80808080u | ((all & 0x20202020u) >> 1) | reinterpret_cast<const unsigned*>(sbits)[(tr + 1) * QW + lane + 1]) & vm;
    const unsigned mw = mb & vm;
    unsigned short on4[4]; unsigned char oc4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      on4[i] = (unsigned short)(((hb >> (8 * i)) & 0xffu) << 8 | ((mw >> (8 * i)) & 0xffu));
      oc4[i] = (unsigned char)(cw >> (8 * i));
    }
    const long long o = s.idx(r, c);
    *reinterpret_cast<ushort4*>(node + o) = make_ushort4(on4[0], on4[1], on4[2], on4[3]);
    *reinterpret_cast<uchar4*>(cnt + o) = make_uchar4(oc4[0], oc4[1], oc4[2], oc4[3]);
    *reinterpret_cast<float4*>(area + o) = make_float4(area_init, area_init, area_init, area_init);   // nodata everywhere first (src/areadinf.cpp:154)
  }
}

}  // namespace

cudaError_t launch_deps_dinf(const float* ang, unsigned short* node, unsigned char* cnt, float* area, const Strip& s,
                             float nodata, const double* theta, cudaStream_t st, float area_init) {
  dim3 grid((s.pitch + TW - 1) / TW, (s.ny + TH - 1) / TH);
  static const int edge_fast = [] { const char* e = getenv("TAUDEM_B200_DEPS_EDGE"); return e ? atoi(e) : 1; }();   // 0: every near-edge angle takes the interval search
  k_deps_dinf<<<grid, 256, 0, st>>>(ang, node, cnt, area, s, nodata, theta, area_init, edge_fast);
  TD_LAUNCHED();
  return cudaGetLastError();
}


}  // namespace td
