// D8 contributing area: the dependency stencil (the evaluation sweep is sweep_warp.cu).
//
// reference: initNeighborD8up src/commonLib.cpp:240-283 (in-degree per cell),
//            aread8 main loop   src/aread8.cpp:216-304 (pull-gather in k order,
//            decrement the downslope cell, enqueue when its count reaches zero).
//
// The reference's result is a deterministic gather — area(c) = ((w|1) + a_k1) + a_k2 ...
// over the neighbours that drain into c, in k = 1..8 order, float32 — evaluated once
// per cell in ANY topological order, so the GPU schedule below is bit-exact:
//
//   k_deps_d8  : 3x3 stencil over p -> node (u16: inflow mask | dir | flags) and
//                cnt (u8: remaining inflow count).  5 B/cell written-read, streamed
//                through TMA-staged shared-memory tiles.
//   the sweep  : sweep_warp.cu (tile dataflow, one warp per tile visit).
#include "common.cuh"
#include "ctx.h"

namespace td {
namespace {
constexpr int TW = 128, TH = 32;

constexpr unsigned NODE_VALID = 0x8000u, NODE_CON = 0x1000u;

// Generic per-cell evaluation (any int16 codes): the slow path of k_deps_d8 for threads that see a direction
// code outside 0..8 that is not nodata.  pm = staged tile at (row above, column c).
template <int SW>
__device__ __noinline__ void deps_d8_generic(const short* pm, const Strip s, int r, int c, short nodata, unsigned short* on4, unsigned char* oc4) {
  short nb[3][6];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < 6; ++i) nb[j][i] = pm[j * SW + i - 1];
  const bool rowok[3] = {s.on_grid(r - 1, 0), true, s.on_grid(r + 1, 0)};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cc = c + i;
    const int d = nb[1][i + 1];
    const bool colok[3] = {cc - 1 >= 0, true, cc + 1 < s.nx};
    const bool valid = (cc < s.nx) & (d != (int)nodata) & ((unsigned)d <= 8u);
    unsigned mask = 0, con = 0;
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      const int dn = nb[1 + drow(k)][i + 1 + dcol(k)];
      const unsigned miss = (unsigned)(!(rowok[1 + drow(k)] & colok[1 + dcol(k)]) | (dn == (int)nodata));   // off-grid or nodata
      const unsigned toward = (unsigned)((dn - k == 4) | (dn - k == -4));                                    // drains into this cell
      const unsigned inrange = (unsigned)((unsigned)dn <= 8u);
      mask |= ((~miss & 1u) & toward & inrange) << (k - 1);
      // counted by the evaluation loop but never evaluated (code outside 0..8) -> its area stays nodata
      con |= miss | (toward & (inrange ^ 1u));
    }
    on4[i] = valid ? (unsigned short)(NODE_VALID | (con ? NODE_CON : 0u) | ((unsigned)d << 8) | mask) : (unsigned short)0;
    oc4[i] = valid ? (unsigned char)__popc(mask) : (unsigned char)0xff;
  }
}

// bit 7 of every byte of the result = (that byte of w == that byte of t); all bytes of w ^ t must be < 0x80
__device__ __forceinline__ unsigned eq_bytes(unsigned w, unsigned t) { return ~((w ^ t) + 0x7f7f7f7fu) & 0x80808080u; }

// Dependency stencil.  Every staged direction code is first reduced to one byte q (0..8 = the code, 16 = off the
// grid or nodata, 32 = any other value); a thread then handles four adjacent cells at once with byte-parallel
// integer arithmetic: "neighbour k drains into me" is a byte comparison of the four neighbours' q with k+4 / k-4
// (src/commonLib.cpp:262-264; code 0 counts for k = 4 exactly as the reference's "tempShort - k == -4" does).
__global__ void __launch_bounds__(256) k_deps_d8(const short* __restrict__ p, unsigned short* __restrict__ node,
                                                 unsigned char* __restrict__ cnt, float* __restrict__ area, Strip s,
                                                 short nodata, float area_init) {
  using G = TileGeom<short, TW, TH>;
  constexpr int QW = TW / 4 + 2;                       // words per row of q: columns c0-4 .. c0+TW+3
  __shared__ __align__(128) short tile[G::ELEMS];
  __shared__ __align__(8) uint64_t bar;
  __shared__ unsigned qw[(TH + 2) * QW];
  const int c0 = blockIdx.x * TW, r0 = 1 + blockIdx.y * TH;
  load_tile_tma<short, TW, TH>(tile, &bar, p, s, r0, c0);
  // ---- q bytes of the staged tile, four per thread and step
  for (int w = threadIdx.x; w < (TH + 2) * QW; w += 256) {
    const int t = w / QW, x = w - t * QW;              // tile row, word within the row
    const int gr = r0 - 1 + t, gc = c0 - 4 + 4 * x;
    const bool rowon = (gr >= 1 && gr <= s.ny) || (gr == 0 && s.has_top) || (gr == s.ny + 1 && s.has_bot);
    const short* src = tile + t * G::SW + (G::HP - 4) + 4 * x;
    unsigned word = 0;
    if (rowon && gc >= 0 && gc + 3 < s.nx) {           // all four cells on the grid (everything but the tile's rim)
      const short4 v = *reinterpret_cast<const short4*>(src);
      const int d4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const unsigned q = d4[i] == (int)nodata ? 16u : ((unsigned)d4[i] <= 8u ? (unsigned)d4[i] : 32u);
        word |= q << (8 * i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = gc + i;
        const int d = src[i];
        const bool on = rowon && col >= 0 && col < s.nx;
        const unsigned q = (!on || d == (int)nodata) ? 16u : ((unsigned)d <= 8u ? (unsigned)d : 32u);
        word |= q << (8 * i);
      }
    }
    qw[w] = word;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int pass = 0; pass < TH / 8; ++pass) {
    const int tr = warp + 8 * pass;
    const int r = r0 + tr, c = c0 + lane * 4;
    if (r > s.ny || c >= s.pitch) continue;
    // W[j][0..2]: the q bytes of the cells (c-1..c+2), (c..c+3), (c+1..c+4) of tile row tr + j
    unsigned W[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const unsigned* q = qw + (tr + j) * QW + lane;
      const unsigned wl = q[0], wc = q[1], wr = q[2];
      W[j][0] = __funnelshift_l(wl, wc, 8);
      W[j][1] = wc;
      W[j][2] = __funnelshift_r(wc, wr, 8);
    }
    unsigned mb = 0, all = 0;
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      const unsigned wk = W[1 + drow(k)][1 + dcol(k)];
      unsigned z = eq_bytes(wk, (k <= 4 ? (unsigned)(k + 4) : (unsigned)(k - 4)) * 0x01010101u);
      if (k == 4) z |= eq_bytes(wk, 0u);                  // 0 - 4 == -4
      mb |= z >> (8 - k);
      all |= wk;
    }
    const unsigned wc = W[1][1];
    unsigned short on4[4]; unsigned char oc4[4];
    if ((all | wc) & 0x20202020u) {
      // a code outside 0..8 next to these cells: the generic rule decides (it can contaminate, src/aread8.cpp:245-250)
      deps_d8_generic<G::SW>(tile + tr * G::SW + G::HP + lane * 4, s, r, c, nodata, on4, oc4);
    } else {
      const unsigned vbit = ~(wc + 0x77777777u) & 0x80808080u;          // bit 7: q <= 8 (a cell of the flow field)
      const unsigned vm = (vbit >> 7) * 0xffu;                            // 0xff per valid byte
      unsigned x = mb - ((mb >> 1) & 0x55555555u);                        // per-byte population count
      x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
      x = (x + (x >> 4)) & 0x0f0f0f0fu;
      const unsigned cw = (x & vm) | ~vm;                                 // count, or 0xff on cells outside the field
      const unsigned hb = (0x80808080u | ((all & 0x10101010u)) | wc) & vm;   // VALID | CON (bit 12 of the node word) | code
      const unsigned mw = mb & vm;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        on4[i] = (unsigned short)(((hb >> (8 * i)) & 0xffu) << 8 | ((mw >> (8 * i)) & 0xffu));
        oc4[i] = (unsigned char)(cw >> (8 * i));
      }
    }
    const long long o = s.idx(r, c);
    *reinterpret_cast<ushort4*>(node + o) = make_ushort4(on4[0], on4[1], on4[2], on4[3]);
    *reinterpret_cast<uchar4*>(cnt + o) = make_uchar4(oc4[0], oc4[1], oc4[2], oc4[3]);
    // the area partition starts as nodata (-1) everywhere (src/aread8.cpp:193)
    *reinterpret_cast<float4*>(area + o) = make_float4(area_init, area_init, area_init, area_init);
  }
}

}  // namespace

cudaError_t launch_deps_d8(const short* p, unsigned short* node, unsigned char* cnt, float* area, const Strip& s, short nodata,
                           cudaStream_t st, float area_init) {
  dim3 grid((s.pitch + TW - 1) / TW, (s.ny + TH - 1) / TH);
  k_deps_d8<<<grid, 256, 0, st>>>(p, node, cnt, area, s, nodata, area_init);
  TD_LAUNCHED();
  return cudaGetLastError();
}



}  // namespace td
