// D-infinity contributing area: dependency stencil + evaluation sweep.
//
// reference: prop()              src/commonLib.cpp:76-91  (share of a cell's flow going to neighbour k)
//            initNeighborDinfup  src/commonLib.cpp:92-136 (in-degree = #neighbours with prop > 0)
//            area() main loop    src/areadinf.cpp:173-265 (k-ordered gather
//                                areares = (float)(areares + p*area_n), + weight or dxc[row],
//                                then decrement every neighbour that receives flow).
// node word of a D-infinity cell: bits 0-7 = which neighbours drain into it, bits 8-11 = its first receiving
// direction k1 (0 = none), 0x2000 = it has a second receiver (always k1 % 8 + 1), 0x1000 = contaminated, 0x8000 = valid.
// prop's table aref[] = {-t,0,t,PI/2,PI-t,PI,PI+t,3PI/2,2PI-t,2PI} with t = atan2(dy,dx)
// is rebuilt on the device from t (host glibc atan2, per row) using only +,-: the same
// doubles as the reference.  The per-cell value is a deterministic gather, so any
// topological schedule reproduces it.
#include "ctx.h"
#include "dinf_common.cuh"

namespace td {
namespace {
constexpr int TW = 128, TH = 32;
constexpr unsigned NODE_VALID = 0x8000u, NODE_CON = 0x1000u;

// bit 7 of every byte of the result = (that byte of w == that byte of t); all bytes of w ^ t must be < 0x80
__device__ __forceinline__ unsigned eq_bytes7(unsigned w, unsigned t) { return ~((w ^ t) + 0x7f7f7f7fu) & 0x80808080u; }

__global__ void __launch_bounds__(256) k_deps_dinf(const float* __restrict__ ang, unsigned short* __restrict__ node,
                                                   unsigned char* __restrict__ cnt, float* __restrict__ area, Strip s,
                                                   float nodata, const double* __restrict__ theta) {
  using G = TileGeom<float, TW, TH>;
  __shared__ __align__(128) float tile[G::ELEMS];
  __shared__ __align__(8) uint64_t bar;
  const int c0 = blockIdx.x * TW, r0 = 1 + blockIdx.y * TH;
  load_tile_tma<float, TW, TH>(tile, &bar, ang, s, r0, c0);
  // one byte per staged cell: k1 | 0x10 if there is a second receiver (always the next direction, k1 % 8 + 1) |
  // 0x20 if the cell is off the grid or nodata; one prop() interval search each
  __shared__ double saref[(TH + 2) * 10];
  __shared__ __align__(16) unsigned char sout[G::ELEMS];
  for (int i = threadIdx.x; i < (TH + 2) * 10; i += 256) saref[i] = aref(i % 10, theta[min(max(r0 - 2 + i / 10, 0), s.ny - 1)]);
  __syncthreads();
  for (int i = threadIdx.x; i < G::ELEMS; i += 256) {
    const int t = i / G::SW, sc = i - t * G::SW;
    const int gr = r0 - 1 + t, gc = c0 - G::HP + sc;
    unsigned char code = 0x20;
    if (s.on_grid(gr, gc)) {
      const float av = tile[i];
      if (!nd_f(av, nodata)) code = (unsigned char)dinf_receivers(av, saref + t * 10);
    }
    sout[i] = code;
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
    const unsigned hb = (0x80808080u | ((all & 0x20202020u) >> 1) | (wc & 0x0f0f0f0fu) | ((wc & 0x10101010u) << 1)) & vm;
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
    *reinterpret_cast<float4*>(area + o) = make_float4(-1.f, -1.f, -1.f, -1.f);   // src/areadinf.cpp:154
  }
}

__device__ __forceinline__ unsigned atom_dec_byte(unsigned* words, long long cell) {
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

constexpr int STK = 12;

// counters[0] = overflow list length, counters[1] = overflow list overflowed (fatal)
template <int SRC>
__global__ void __launch_bounds__(256) k_sweep_dinf(const unsigned short* __restrict__ node, unsigned* __restrict__ cntw,
                                                    const float* __restrict__ ang, float* __restrict__ area,
                                                    const float* __restrict__ w, Strip s, int usew, int contcheck,
                                                    const double* __restrict__ theta, const double* __restrict__ dxc,
                                                    int* __restrict__ halo, const long long* __restrict__ list,
                                                    unsigned long long nlist, long long* __restrict__ ovf,
                                                    unsigned long long ovf_cap, unsigned long long* __restrict__ counters) {
  int r, c;
  if (SRC == 0) {
    c = blockIdx.x * 64 + (threadIdx.x & 63);
    r = 1 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r > s.ny || c >= s.nx) return;
  } else {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlist) return;
    const long long ci0 = list[t];
    r = (int)(ci0 / s.pitch); c = (int)(ci0 - (long long)r * s.pitch);
  }
  long long ci = s.idx(r, c);
  unsigned nd = node[ci];
  if (!(nd & NODE_VALID)) return;
  if (SRC == 0 && (nd & 0xffu)) return;

  long long stack[STK];
  int sp = 0;
  for (;;) {
    // ---- flow algebra (src/areadinf.cpp:187-218)
    float areares = 0.f;
    bool con = (nd & NODE_CON) != 0;
    const unsigned m = nd & 0xffu;
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      if (m & (1u << (k - 1))) {
        const long long ni = ci + (long long)drow(k) * s.pitch + dcol(k);
        const double p = prop_dev(ang[ni], (k + 4) % 8, theta[min(max(r - 1 + drow(k), 0), s.ny - 1)]);
        const float an = __ldcg(area + ni);
        if (nd_f(an, -1.0f)) con = true;
        else areares = (float)((double)areares + p * (double)an);
      }
    }
    if (usew) areares = areares + w[ci];
    else areares = (float)((double)areares + dxc[r - 1]);
    area[ci] = (con && contcheck) ? -1.0f : areares;

    // ---- decrement every neighbour that receives flow (src/areadinf.cpp:221-239)
    const float a0 = ang[ci];
    const double t0 = theta[r - 1];
    long long next = -1; int nr = 0, nc = 0; unsigned nnd = 0;
#pragma unroll
    for (int k = 1; k <= 8; ++k) {
      if (prop_dev(a0, k, t0) > 0.0) {
        const int rn = r + drow(k), cn = c + dcol(k);
        if (!s.on_grid(rn, cn)) continue;
        const long long cin = s.idx(rn, cn);
        if (rn == 0 || rn == s.ny + 1) { __threadfence(); atomicAdd(halo + (rn == 0 ? 0 : s.pitch) + cn, 1); continue; }
        const unsigned ndn = node[cin];
        if (!(ndn & NODE_VALID)) continue;
        if (atom_dec_byte(cntw, cin) == 1u) {
          if (next < 0) { next = cin; nr = rn; nc = cn; nnd = ndn; }
          else if (sp < STK) stack[sp++] = cin;
          else {
            const unsigned long long slot = atomicAdd(counters, 1ull);
            if (slot < ovf_cap) ovf[slot] = cin; else counters[1] = 1ull;
          }
        }
      }
    }
    if (next >= 0) { ci = next; r = nr; c = nc; nd = nnd; }
    else if (sp > 0) {
      ci = stack[--sp];
      r = (int)(ci / s.pitch); c = (int)(ci - (long long)r * s.pitch);
      nd = node[ci];
    } else return;
  }
}
}  // namespace

cudaError_t launch_deps_dinf(const float* ang, unsigned short* node, unsigned char* cnt, float* area, const Strip& s,
                             float nodata, const double* theta, cudaStream_t st) {
  dim3 grid((s.pitch + TW - 1) / TW, (s.ny + TH - 1) / TH);
  k_deps_dinf<<<grid, 256, 0, st>>>(ang, node, cnt, area, s, nodata, theta);
  TD_LAUNCHED();
  return cudaGetLastError();
}

cudaError_t launch_sweep_dinf(const unsigned short* node, unsigned* cntw, const float* ang, float* area, const float* w,
                              const Strip& s, int usew, int contcheck, const double* theta, const double* dxc, int* halo,
                              const long long* list, unsigned long long nlist, long long* ovf, unsigned long long ovf_cap,
                              unsigned long long* counters, cudaStream_t st) {
  if (list == nullptr) {
    dim3 grid((s.nx + 63) / 64, (s.ny + 3) / 4);
    k_sweep_dinf<0><<<grid, 256, 0, st>>>(node, cntw, ang, area, w, s, usew, contcheck, theta, dxc, halo, nullptr, 0ull, ovf,
                                         ovf_cap, counters);
  } else {
    if (nlist == 0) return cudaSuccess;
    k_sweep_dinf<1><<<(unsigned)((nlist + 255) / 256), 256, 0, st>>>(node, cntw, ang, area, w, s, usew, contcheck, theta, dxc,
                                                                   halo, list, nlist, ovf, ovf_cap, counters);
  }
  TD_LAUNCHED();
  return cudaGetLastError();
}

}  // namespace td
